// b2_ingest_ring: the streaming, shape-agnostic form of the ingest path, callable with HOST pointers only.
//
// A LISTING is what the reference's sync loop holds after its downloads (app/services/webdav_sync.py:273-283,
// :441): n files in host memory, any mix of sizes.  submit() takes per-image pointers and shapes and returns a
// ticket at once; wait(ticket) returns when that listing's digests, dedupe decision + stats
// (webdav_sync.py:311-400), thumbnails and previews are in the caller's host buffers, in LISTING ORDER.  Several
// listings may be in flight, so the hash tail of one hides under the copies of the next.
//
// Why a ring.  SHA-256 is a serial chain per message: one message moves at ~60 MB/s on a warp pair whatever else
// the GPU does, so a 50 MB file needs 0.8 s and PCIe (55 GB/s) is only kept busy when >= ~1 000 messages hash
// concurrently.  The depth of the pipeline is therefore bounded in BYTES, not in batches: one device staging ring
// (tens of GB of the 180 GB of HBM) is carved into chunks of consecutive listing entries
//     [ metadata | images, 16-byte aligned | thumbnails | previews ]
// allocated first-in first-out; a chunk is released when its hash kernel, its resize kernels and the read-back of
// its outputs are done, and submit() blocks only when the ring is full (back-pressure).  Per chunk: one H2D burst
// on the copy stream, ONE hash launch on one of kRingHashStreams streams (warp-pair kernel, lanes ordered by
// decreasing length), one resize launch per shape present in the chunk (cached tap plans, outputs addressed by
// position in the chunk), and one contiguous D2H per output kind straight into the listing-order slots of the
// caller's buffers.  Per listing: the dedupe decision over all its digests on the `fin` stream.
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <deque>
#include <map>
#include <new>
#include <numeric>
#include <utility>
#include <vector>

namespace b2 {

constexpr int kRingHashStreams = 96;         // a chunk's hash runs for the time its LONGEST message needs (up to ~0.8 s for 50 MB)
constexpr int kRingResizeStreams = 4;
constexpr uint32_t kRingMaxChunkImages = 4096;
constexpr uint64_t kRingMinChunkBytes = 32ull << 20;

static inline uint64_t up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct RingChunk {
    uint64_t off = 0, bytes = 0;             // region of the device ring
    uint64_t gen = 0;                        // bumped when the region is released
    cudaEvent_t copied = nullptr, hashed = nullptr, resized = nullptr, flushed = nullptr;
    uint8_t *h_meta = nullptr;               // page-locked metadata staging (lives as long as the chunk record)
    size_t h_meta_cap = 0;
};

struct RingListing {
    bool pending = false;
    uint64_t ticket = 0;
    uint32_t n = 0, cap = 0;
    uint8_t *d_digests = nullptr, *d_is_new = nullptr, *d_valid = nullptr, *d_existing = nullptr;
    int32_t *d_first = nullptr, *d_last = nullptr;
    uint32_t *d_counts = nullptr;
    void *d_ws = nullptr;
    uint64_t ws_bytes = 0, existing_cap = 0;
    cudaEvent_t done = nullptr;
    uint64_t h2d = 0, d2h = 0;
    uint32_t launches = 0;
    struct Part { RingChunk *chunk; uint64_t gen; uint32_t hi; };
    std::vector<Part> parts;                 // chunk -> images [.., hi) of the listing, for progress()
};

}  // namespace b2

struct b2_ingest_ring {
    int device = 0;
    int out_h = 0, out_w = 0;
    bool want_preview = false;
    uint64_t ring_bytes = 0, chunk_bytes = 0;
    uint8_t *d_ring = nullptr;
    uint64_t head = 0;                       // next allocation offset
    std::deque<b2::RingChunk *> live;        // allocation order
    std::vector<b2::RingChunk *> spare;
    std::vector<b2::RingListing> listings;
    uint64_t next_ticket = 1;
    cudaStream_t copy = nullptr, d2h = nullptr, fin = nullptr;
    cudaStream_t hash[b2::kRingHashStreams] = {};
    cudaStream_t resize[b2::kRingResizeStreams] = {};
    uint32_t next_hash = 0, next_resize = 0;
    uint64_t stalls = 0;                     // times submit() had to wait for ring space
};

namespace b2 {

static void ring_free_chunk(RingChunk *c) {
    if (!c) return;
    if (c->copied) cudaEventDestroy(c->copied);
    if (c->hashed) cudaEventDestroy(c->hashed);
    if (c->resized) cudaEventDestroy(c->resized);
    if (c->flushed) cudaEventDestroy(c->flushed);
    if (c->h_meta) cudaFreeHost(c->h_meta);
    delete c;
}

static cudaError_t ring_new_chunk(b2_ingest_ring *r, RingChunk **out) {
    if (!r->spare.empty()) {
        *out = r->spare.back();
        r->spare.pop_back();
        return cudaSuccess;
    }
    RingChunk *c = new (std::nothrow) RingChunk();
    if (!c) return cudaErrorMemoryAllocation;
    cudaError_t e = cudaSuccess;
    for (cudaEvent_t *ev : {&c->copied, &c->hashed, &c->resized, &c->flushed})
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { ring_free_chunk(c); return e; }
    *out = c;
    return cudaSuccess;
}

// Release the oldest live chunk (waiting for it when `block`).  Returns false when nothing could be released.
static bool ring_retire_oldest(b2_ingest_ring *r, bool block, cudaError_t *err) {
    if (r->live.empty()) return false;
    RingChunk *c = r->live.front();
    if (!block) {
        if (cudaEventQuery(c->hashed) != cudaSuccess || cudaEventQuery(c->flushed) != cudaSuccess) {
            cudaGetLastError();                              // cudaErrorNotReady is not an error
            return false;
        }
    } else {
        cudaError_t e = cudaEventSynchronize(c->hashed);
        if (e == cudaSuccess) e = cudaEventSynchronize(c->flushed);
        if (e != cudaSuccess) { *err = e; return false; }
    }
    r->live.pop_front();
    ++c->gen;
    r->spare.push_back(c);
    return true;
}

// First-in first-out allocation of `bytes` (a multiple of 256) in the ring; blocks while the ring is full.
static cudaError_t ring_alloc(b2_ingest_ring *r, uint64_t bytes, uint64_t *off) {
    cudaError_t err = cudaSuccess;
    while (ring_retire_oldest(r, false, &err)) {}
    for (;;) {
        if (r->live.empty()) { r->head = 0; }
        uint64_t cand = r->head + bytes <= r->ring_bytes ? r->head : 0;
        bool clash = false;
        for (const RingChunk *c : r->live)
            if (cand < c->off + c->bytes && c->off < cand + bytes) { clash = true; break; }
        if (!clash) {
            *off = cand;
            r->head = cand + bytes;
            return cudaSuccess;
        }
        ++r->stalls;
        if (!ring_retire_oldest(r, true, &err)) return err != cudaSuccess ? err : cudaErrorUnknown;
    }
}

static void ring_drain(b2_ingest_ring *r) {                  // error path: nothing of a failed submit stays in flight
    cudaStreamSynchronize(r->copy);
    for (auto st : r->hash) if (st) cudaStreamSynchronize(st);
    for (auto st : r->resize) if (st) cudaStreamSynchronize(st);
    cudaStreamSynchronize(r->d2h);
    cudaStreamSynchronize(r->fin);
    cudaGetLastError();
    for (RingChunk *c : r->live) { ++c->gen; r->spare.push_back(c); }
    r->live.clear();
    r->head = 0;
}

int cached_plan(int device, int ih, int iw, int oh, int ow, b2_resize_plan **out);   // host.cu

}  // namespace b2

extern "C" int b2_ingest_ring_destroy(b2_ingest_ring *r) {
    if (!r) return B2_OK;
    cudaSetDevice(r->device);
    if (r->copy) b2::ring_drain(r);
    for (auto *c : r->spare) b2::ring_free_chunk(c);
    for (auto &l : r->listings) {
        cudaFree(l.d_digests); cudaFree(l.d_is_new); cudaFree(l.d_valid); cudaFree(l.d_existing);
        cudaFree(l.d_first); cudaFree(l.d_last); cudaFree(l.d_counts); cudaFree(l.d_ws);
        if (l.done) cudaEventDestroy(l.done);
    }
    cudaFree(r->d_ring);
    if (r->copy) cudaStreamDestroy(r->copy);
    if (r->d2h) cudaStreamDestroy(r->d2h);
    if (r->fin) cudaStreamDestroy(r->fin);
    for (auto st : r->hash) if (st) cudaStreamDestroy(st);
    for (auto st : r->resize) if (st) cudaStreamDestroy(st);
    delete r;
    return B2_OK;
}

extern "C" int b2_ingest_ring_create(int device, uint64_t ring_bytes, uint64_t chunk_bytes, uint32_t max_listings,
                                     int out_h, int out_w, int want_preview, b2_ingest_ring **out) {
    using namespace b2;
    B2_REQUIRE(out != nullptr, "b2_ingest_ring_create: null output");
    *out = nullptr;
    B2_REQUIRE(out_h >= 1 && out_w >= 1, "b2_ingest_ring_create: empty output shape");
    B2_REQUIRE(max_listings >= 1 && max_listings <= 64, "b2_ingest_ring_create: 1 <= max_listings <= 64");
    B2_REQUIRE(ring_bytes >= (64ull << 20), "b2_ingest_ring_create: ring smaller than 64 MiB");
    int rc = b2_init(device);
    if (rc != B2_OK) return rc;
    b2_ingest_ring *r = new (std::nothrow) b2_ingest_ring();
    B2_REQUIRE(r != nullptr, "b2_ingest_ring_create: out of host memory");
    r->device = device;
    r->out_h = out_h; r->out_w = out_w;
    r->want_preview = want_preview != 0;
    r->ring_bytes = ring_bytes & ~uint64_t(255);
    r->chunk_bytes = chunk_bytes ? chunk_bytes : (1ull << 30);
    if (r->chunk_bytes > r->ring_bytes / 4) r->chunk_bytes = r->ring_bytes / 4;
    r->listings.resize(max_listings);
#define B2_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            b2_ingest_ring_destroy(r);                                                                  \
            return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)
    B2_TRY(cudaMalloc(&r->d_ring, size_t(r->ring_bytes)));
    int prio_low = 0, prio_high = 0;
    B2_TRY(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
    B2_TRY(cudaStreamCreateWithPriority(&r->copy, cudaStreamNonBlocking, prio_low));
    B2_TRY(cudaStreamCreateWithPriority(&r->d2h, cudaStreamNonBlocking, prio_low));
    B2_TRY(cudaStreamCreateWithPriority(&r->fin, cudaStreamNonBlocking, prio_low));
    // the hash kernels are the long pole: their CTAs are placed first
    for (auto &st : r->hash) B2_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_high));
    for (auto &st : r->resize) B2_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_low));
    for (auto &l : r->listings) {
        B2_TRY(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        B2_TRY(cudaMalloc(&l.d_counts, 16));
    }
#undef B2_TRY
    *out = r;
    return B2_OK;
}

namespace b2 {

static cudaError_t listing_reserve(RingListing &l, uint32_t n) {
    if (n <= l.cap) return cudaSuccess;
    uint32_t cap = l.cap ? l.cap : 1024;
    while (cap < n) cap *= 2;
    cudaFree(l.d_digests); cudaFree(l.d_is_new); cudaFree(l.d_valid); cudaFree(l.d_first); cudaFree(l.d_last); cudaFree(l.d_ws);
    l.d_digests = l.d_is_new = l.d_valid = nullptr; l.d_first = l.d_last = nullptr; l.d_ws = nullptr; l.cap = 0;
    cudaError_t e = cudaMalloc(&l.d_digests, size_t(cap) * 32);
    if (e == cudaSuccess) e = cudaMalloc(&l.d_is_new, cap);
    if (e == cudaSuccess) e = cudaMalloc(&l.d_valid, cap);
    if (e == cudaSuccess) e = cudaMalloc(&l.d_first, size_t(cap) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&l.d_last, size_t(cap) * 4);
    l.ws_bytes = b2_dedupe_workspace_bytes(cap);
    if (e == cudaSuccess) e = cudaMalloc(&l.d_ws, size_t(l.ws_bytes));
    if (e == cudaSuccess) l.cap = cap;
    return e;
}

}  // namespace b2

extern "C" int b2_ingest_ring_submit(b2_ingest_ring *r, const uint8_t *const *h_pixels, const uint32_t *h_hw,
                                     const uint8_t *const *h_files, const uint64_t *h_file_lens,
                                     const uint8_t *h_valid, uint32_t n,
                                     const uint8_t *h_existing_sorted, uint64_t m,
                                     uint8_t *h_digests, uint8_t *h_is_new, int32_t *h_first_index,
                                     int32_t *h_last_index, uint32_t *h_counts,
                                     uint8_t *h_thumbs, float *h_previews, uint64_t *ticket) {
    using namespace b2;
    B2_REQUIRE(r != nullptr && ticket != nullptr, "b2_ingest_ring_submit: null ring or ticket");
    *ticket = 0;
    B2_REQUIRE(n >= 1 && n < 0x7fffffffu, "b2_ingest_ring_submit: empty listing");
    B2_REQUIRE(h_pixels != nullptr || h_files != nullptr, "b2_ingest_ring_submit: neither pixels nor file bytes given");
    B2_REQUIRE(h_pixels == nullptr || h_hw != nullptr, "b2_ingest_ring_submit: pixels without shapes");
    B2_REQUIRE(h_files == nullptr || h_file_lens != nullptr, "b2_ingest_ring_submit: file bytes without lengths");
    B2_REQUIRE(h_digests && h_is_new && h_counts, "b2_ingest_ring_submit: null output pointer");
    B2_REQUIRE(h_pixels == nullptr || h_thumbs != nullptr, "b2_ingest_ring_submit: pixels given but no thumbnail buffer");
    B2_REQUIRE(!h_previews || r->want_preview, "b2_ingest_ring_submit: previews asked from a ring created without them");
    B2_REQUIRE(m == 0 || h_existing_sorted != nullptr, "b2_ingest_ring_submit: null existing table");
    B2_CUDA_CHECK(cudaSetDevice(r->device));
    RingListing *L = nullptr;
    for (auto &l : r->listings) if (!l.pending) { L = &l; break; }
    B2_REQUIRE(L != nullptr, "b2_ingest_ring_submit: %zu listings already in flight (wait for one first)", r->listings.size());

    const size_t out_px = size_t(r->out_h) * r->out_w * 3;
    auto has_px = [&](uint32_t i) { return h_pixels && h_pixels[i] && h_hw[2 * i] && h_hw[2 * i + 1] && (!h_valid || h_valid[i]); };
    auto px_len = [&](uint32_t i) -> uint64_t { return has_px(i) ? uint64_t(h_hw[2 * i]) * h_hw[2 * i + 1] * 3 : 0; };
    auto msg_len = [&](uint32_t i) -> uint64_t {
        if (h_valid && !h_valid[i]) return 0;
        return h_files ? (h_files[i] ? h_file_lens[i] : 0) : px_len(i);
    };
    // bytes image i occupies in a chunk: pixels (16-byte aligned) + separate file bytes when given
    auto img_bytes = [&](uint32_t i) -> uint64_t { return up(px_len(i), 16) + (h_files ? up(msg_len(i), 16) : 0); };
    uint64_t listing_bytes = 0;
    for (uint32_t i = 0; i < n; ++i) listing_bytes += img_bytes(i);
    uint64_t chunk_bytes = std::min<uint64_t>(r->chunk_bytes, std::max<uint64_t>(kRingMinChunkBytes, listing_bytes / 8));

#define B2_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            ring_drain(r);                                                                              \
            return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)
#define B2_TRY_RC(expr)                                                                                 \
    do {                                                                                                \
        int _rc = (expr);                                                                               \
        if (_rc != B2_OK) { ring_drain(r); return _rc; }                                                \
    } while (0)

    if (n > L->cap) {                                        // grows rarely; cudaFree synchronises the device
        B2_TRY(listing_reserve(*L, n));
    }
    if (m > L->existing_cap) {
        cudaFree(L->d_existing);
        L->d_existing = nullptr;
        L->existing_cap = 0;
        B2_TRY(cudaMalloc(&L->d_existing, size_t(m) * 32));
        L->existing_cap = m;
    }
    L->n = n;
    L->h2d = L->d2h = 0;
    L->launches = 0;
    L->parts.clear();
    if (m) {
        B2_TRY(cudaMemcpyAsync(L->d_existing, h_existing_sorted, size_t(m) * 32, cudaMemcpyHostToDevice, r->copy));
        L->h2d += m * 32;
    }
    if (h_valid) {
        B2_TRY(cudaMemcpyAsync(L->d_valid, h_valid, n, cudaMemcpyHostToDevice, r->copy));
        L->h2d += n;
    }

    std::vector<uint32_t> order, grouped;
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> groups;
    for (uint32_t lo = 0; lo < n;) {
        // ---- chunk = consecutive listing entries up to chunk_bytes
        uint32_t cnt = 0;
        uint64_t data_bytes = 0;
        while (lo + cnt < n && cnt < kRingMaxChunkImages) {
            const uint64_t b = img_bytes(lo + cnt);
            if (cnt > 0 && data_bytes + b > chunk_bytes) break;
            data_bytes += b;
            ++cnt;
        }
        // layout inside the chunk region (offsets relative to its start)
        const uint64_t o_hoff = 0, o_hlen = o_hoff + 8ull * cnt, o_poff = o_hlen + 8ull * cnt,
                       o_order = o_poff + 8ull * cnt, o_slot = o_order + 4ull * cnt, meta_bytes = o_slot + 4ull * cnt;
        const uint64_t o_data = up(meta_bytes, 256);
        const uint64_t o_thumb = up(o_data + data_bytes + 16, 256);
        const uint64_t thumb_bytes = h_pixels ? uint64_t(cnt) * out_px : 0;
        const uint64_t o_prev = up(o_thumb + thumb_bytes, 256);
        const uint64_t prev_bytes = h_pixels && h_previews ? uint64_t(cnt) * out_px * 4 : 0;
        const uint64_t total = up(o_prev + prev_bytes, 256);
        if (total > r->ring_bytes) {
            ring_drain(r);
            return fail(B2_ERR_BAD_ARG, "b2_ingest_ring_submit: image %u needs %llu bytes of staging, the ring has %llu",
                        lo, (unsigned long long)total, (unsigned long long)r->ring_bytes);
        }
        RingChunk *c = nullptr;
        B2_TRY(ring_new_chunk(r, &c));
        if (c->h_meta_cap < meta_bytes) {
            if (c->h_meta) cudaFreeHost(c->h_meta);
            c->h_meta = nullptr;
            c->h_meta_cap = 0;
            const size_t want = size_t(up(meta_bytes, 4096));
            cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&c->h_meta), want, cudaHostAllocDefault);
            if (e != cudaSuccess) { r->spare.push_back(c); B2_TRY(e); }
            c->h_meta_cap = want;
        }
        {
            uint64_t off = 0;
            cudaError_t e = ring_alloc(r, total, &off);
            if (e != cudaSuccess) { r->spare.push_back(c); B2_TRY(e); }
            c->off = off;
            c->bytes = total;
        }
        r->live.push_back(c);
        uint8_t *d_chunk = r->d_ring + c->off;
        uint64_t *m_hoff = reinterpret_cast<uint64_t *>(c->h_meta + o_hoff), *m_hlen = reinterpret_cast<uint64_t *>(c->h_meta + o_hlen),
                 *m_poff = reinterpret_cast<uint64_t *>(c->h_meta + o_poff);
        uint32_t *m_order = reinterpret_cast<uint32_t *>(c->h_meta + o_order), *m_slot = reinterpret_cast<uint32_t *>(c->h_meta + o_slot);

        // ---- placement + copies (adjacent host buffers that land adjacently on the device travel as one copy)
        groups.clear();
        uint64_t pos = c->off + o_data;                      // offsets are relative to the ring base
        const uint8_t *run_src = nullptr;
        uint64_t run_dst = 0, run_len = 0;
        auto flush_run = [&]() -> cudaError_t {
            if (!run_len) return cudaSuccess;
            cudaError_t e = cudaMemcpyAsync(r->d_ring + run_dst, run_src, size_t(run_len), cudaMemcpyHostToDevice, r->copy);
            L->h2d += run_len;
            run_len = 0;
            return e;
        };
        auto put = [&](const uint8_t *src, uint64_t len) -> cudaError_t {
            if (!len) return cudaSuccess;
            if (run_len && src == run_src + run_len && pos == run_dst + run_len) { run_len += len; return cudaSuccess; }
            cudaError_t e = flush_run();
            run_src = src; run_dst = pos; run_len = len;
            return e;
        };
        std::vector<uint64_t> pix_off(cnt, 0);
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t i = lo + j;
            const uint64_t pl = px_len(i), ml = msg_len(i);
            if (pl) {
                pix_off[j] = pos;
                B2_TRY(put(h_pixels[i], pl));
                pos += up(pl, 16);                           // (a padded length ends the run: the next start differs)
                groups[{h_hw[2 * i], h_hw[2 * i + 1]}].push_back(j);
            }
            if (h_files) {
                m_hoff[j] = pos;
                m_hlen[j] = ml;
                if (ml) B2_TRY(put(h_files[i], ml));
                pos += up(ml, 16);
            } else {
                m_hoff[j] = pl ? pix_off[j] : pos;
                m_hlen[j] = ml;
            }
        }
        B2_TRY(flush_run());
        order.resize(cnt);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return m_hlen[a] > m_hlen[b]; });
        memcpy(m_order, order.data(), 4ull * cnt);
        uint32_t g = 0;
        for (auto &kv : groups)
            for (uint32_t j : kv.second) { m_poff[g] = pix_off[j]; m_slot[g] = j; ++g; }
        B2_TRY(cudaMemcpyAsync(d_chunk, c->h_meta, size_t(meta_bytes), cudaMemcpyHostToDevice, r->copy));
        L->h2d += meta_bytes;
        B2_TRY(cudaEventRecord(c->copied, r->copy));

        // ---- hash: one launch for the chunk, digests straight into the listing's slots
        cudaStream_t hs = r->hash[r->next_hash++ % kRingHashStreams];
        B2_TRY(cudaStreamWaitEvent(hs, c->copied, 0));
        B2_TRY_RC(b2_sha256_batch(r->d_ring, reinterpret_cast<const uint64_t *>(d_chunk + o_hoff),
                                  reinterpret_cast<const uint64_t *>(d_chunk + o_hlen),
                                  reinterpret_cast<const uint32_t *>(d_chunk + o_order), cnt, L->d_digests + size_t(lo) * 32, hs));
        B2_TRY(cudaEventRecord(c->hashed, hs));
        B2_TRY(cudaStreamWaitEvent(r->fin, c->hashed, 0));
        ++L->launches;

        // ---- resize: one launch per shape in the chunk, then one read-back per output kind
        if (g) {
            cudaStream_t rs = r->resize[r->next_resize++ % kRingResizeStreams];
            B2_TRY(cudaStreamWaitEvent(rs, c->copied, 0));
            uint8_t *d_thumb = d_chunk + o_thumb;
            float *d_prev = prev_bytes ? reinterpret_cast<float *>(d_chunk + o_prev) : nullptr;
            if (g < cnt) {                                   // entries without pixels: defined (zero) outputs
                B2_TRY(cudaMemsetAsync(d_thumb, 0, size_t(thumb_bytes), rs));
                if (d_prev) B2_TRY(cudaMemsetAsync(d_prev, 0, size_t(prev_bytes), rs));
            }
            uint32_t g0 = 0;
            for (auto &kv : groups) {
                b2_resize_plan *plan = nullptr;
                B2_TRY_RC(cached_plan(r->device, int(kv.first.first), int(kv.first.second), r->out_h, r->out_w, &plan));
                const uint32_t gm = uint32_t(kv.second.size());
                B2_TRY_RC(b2_resize_normalize_batch(plan, r->d_ring, reinterpret_cast<const uint64_t *>(d_chunk + o_poff) + g0,
                                                    reinterpret_cast<const uint32_t *>(d_chunk + o_slot) + g0, gm, d_thumb, d_prev,
                                                    nullptr, nullptr, rs));
                g0 += gm;
                ++L->launches;
            }
            B2_TRY(cudaEventRecord(c->resized, rs));
            B2_TRY(cudaStreamWaitEvent(r->d2h, c->resized, 0));
            B2_TRY(cudaMemcpyAsync(h_thumbs + size_t(lo) * out_px, d_thumb, size_t(thumb_bytes), cudaMemcpyDeviceToHost, r->d2h));
            L->d2h += thumb_bytes;
            if (d_prev) {
                B2_TRY(cudaMemcpyAsync(h_previews + size_t(lo) * out_px, d_prev, size_t(prev_bytes), cudaMemcpyDeviceToHost, r->d2h));
                L->d2h += prev_bytes;
            }
        } else {
            B2_TRY(cudaStreamWaitEvent(r->d2h, c->copied, 0));
        }
        B2_TRY(cudaEventRecord(c->flushed, r->d2h));
        lo += cnt;
        L->parts.push_back({c, c->gen, lo});
    }

    // ---- the listing's dedupe decision, after every chunk's hash (fin already waits for them)
    B2_TRY(cudaEventRecord(L->done, r->copy));               // existing table / validity flags copied
    B2_TRY(cudaStreamWaitEvent(r->fin, L->done, 0));
    B2_TRY_RC(b2_dedupe(L->d_digests, h_valid ? L->d_valid : nullptr, nullptr, n, m ? L->d_existing : nullptr, m, L->d_is_new,
                        L->d_first, L->d_last, L->d_counts, L->d_ws, L->ws_bytes, r->fin));
    L->launches += 2;
    B2_TRY(cudaMemcpyAsync(h_digests, L->d_digests, size_t(n) * 32, cudaMemcpyDeviceToHost, r->fin));
    B2_TRY(cudaMemcpyAsync(h_is_new, L->d_is_new, n, cudaMemcpyDeviceToHost, r->fin));
    B2_TRY(cudaMemcpyAsync(h_counts, L->d_counts, 12, cudaMemcpyDeviceToHost, r->fin));
    L->d2h += uint64_t(n) * 33 + 12;
    if (h_first_index) {
        B2_TRY(cudaMemcpyAsync(h_first_index, L->d_first, size_t(n) * 4, cudaMemcpyDeviceToHost, r->fin));
        L->d2h += uint64_t(n) * 4;
    }
    if (h_last_index) {
        B2_TRY(cudaMemcpyAsync(h_last_index, L->d_last, size_t(n) * 4, cudaMemcpyDeviceToHost, r->fin));
        L->d2h += uint64_t(n) * 4;
    }
    B2_TRY(cudaStreamWaitEvent(r->fin, L->parts.back().chunk->flushed, 0));   // the d2h stream is in order: last chunk = all chunks
    B2_TRY(cudaEventRecord(L->done, r->fin));
#undef B2_TRY
#undef B2_TRY_RC
    L->pending = true;
    L->ticket = r->next_ticket++;
    *ticket = L->ticket;
    return B2_OK;
}

namespace b2 {
static RingListing *ring_find(b2_ingest_ring *r, uint64_t ticket) {
    for (auto &l : r->listings) if (l.pending && l.ticket == ticket) return &l;
    return nullptr;
}
}  // namespace b2

extern "C" int b2_ingest_ring_wait(b2_ingest_ring *r, uint64_t ticket, uint64_t *h2d_bytes, uint64_t *d2h_bytes,
                                   uint32_t *kernel_launches) {
    using namespace b2;
    B2_REQUIRE(r != nullptr, "b2_ingest_ring_wait: null ring");
    RingListing *L = ring_find(r, ticket);
    B2_REQUIRE(L != nullptr, "b2_ingest_ring_wait: ticket %llu is not in flight", (unsigned long long)ticket);
    B2_CUDA_CHECK(cudaSetDevice(r->device));
    cudaError_t e = cudaEventSynchronize(L->done);
    L->pending = false;
    L->parts.clear();
    if (e != cudaSuccess) return fail(B2_ERR_CUDA, "b2_ingest_ring_wait: %s", cudaGetErrorString(e));
    if (h2d_bytes) *h2d_bytes = L->h2d;
    if (d2h_bytes) *d2h_bytes = L->d2h;
    if (kernel_launches) *kernel_launches = L->launches;
    return B2_OK;
}

extern "C" int b2_ingest_ring_poll(b2_ingest_ring *r, uint64_t ticket, int *done, uint32_t *images_flushed) {
    using namespace b2;
    B2_REQUIRE(r != nullptr && done != nullptr, "b2_ingest_ring_poll: null pointer");
    RingListing *L = ring_find(r, ticket);
    B2_REQUIRE(L != nullptr, "b2_ingest_ring_poll: ticket %llu is not in flight", (unsigned long long)ticket);
    B2_CUDA_CHECK(cudaSetDevice(r->device));
    cudaError_t e = cudaEventQuery(L->done);
    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(B2_ERR_CUDA, "b2_ingest_ring_poll: %s", cudaGetErrorString(e));
    *done = e == cudaSuccess;
    if (images_flushed) {                                    // leading entries whose thumbnails / previews are in host memory
        uint32_t hi = 0;
        for (const auto &p : L->parts) {
            const bool gone = p.chunk->gen != p.gen;         // released = finished
            if (!gone && cudaEventQuery(p.chunk->flushed) != cudaSuccess) break;
            hi = p.hi;
        }
        *images_flushed = *done ? L->n : hi;
    }
    cudaGetLastError();
    return B2_OK;
}

extern "C" int b2_ingest_ring_stats(const b2_ingest_ring *r, uint64_t *ring_bytes, uint64_t *bytes_in_flight,
                                    uint32_t *chunks_in_flight, uint64_t *stalls) {
    using namespace b2;
    B2_REQUIRE(r != nullptr, "b2_ingest_ring_stats: null ring");
    uint64_t b = 0;
    for (const RingChunk *c : r->live) b += c->bytes;
    if (ring_bytes) *ring_bytes = r->ring_bytes;
    if (bytes_in_flight) *bytes_in_flight = b;
    if (chunks_in_flight) *chunks_in_flight = uint32_t(r->live.size());
    if (stalls) *stalls = r->stalls;
    return B2_OK;
}
