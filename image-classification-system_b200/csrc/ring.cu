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
// on the copy stream, one resize launch per shape present in the chunk (cached tap plans, outputs addressed by
// position in the chunk), and one contiguous D2H per output kind straight into the listing-order slots of the
// caller's buffers.  Per hash GROUP (consecutive chunks, ~4 GiB): ONE hash launch on one of kRingHashStreams
// streams (warp-pair kernel, lanes ordered by decreasing length over the group).  Per listing: the dedupe decision
// over all its digests on the `fin` stream.
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

constexpr int kRingHashStreams = 32;         // a group's hash runs for the time its LONGEST message needs (up to ~0.8 s for 50 MB)
constexpr int kRingResizeStreams = 4;
constexpr uint32_t kRingMaxChunkImages = 4096;
constexpr uint64_t kRingMinChunkBytes = 32ull << 20;

static inline uint64_t up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
static bool ring_trace() { static const bool t = getenv("B2_RING_TRACE") != nullptr; return t; }   // debug: per-chunk timeline on stderr

struct RingChunk {
    uint64_t off = 0, bytes = 0;             // region of the device ring
    uint64_t gen = 0;                        // bumped when the region is released
    cudaEvent_t copied = nullptr, hashed = nullptr, resized = nullptr, flushed = nullptr;
    uint8_t *h_meta = nullptr;               // page-locked metadata staging (lives as long as the chunk record)
    size_t h_meta_cap = 0;
    bool hash_pending = false;               // data chunk of a hash group that has not been launched yet
};

// A device buffer carved first-in first-out into regions that are released in allocation order.
struct RingArena {
    uint8_t *base = nullptr;
    uint64_t bytes = 0, head = 0;
    std::deque<RingChunk *> live;            // allocation order
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
    struct TraceGroup { RingChunk *meta; uint32_t chunks, msgs; uint64_t bytes; };
    std::vector<TraceGroup> trace_groups;    // trace mode only
};

}  // namespace b2

struct b2_ingest_ring {
    int device = 0;
    int out_h = 0, out_w = 0;
    bool want_preview = false;
    uint64_t ring_bytes = 0, chunk_bytes = 0, hash_group_bytes = 4ull << 30;
    uint8_t *d_ring = nullptr;               // = data.base: images + outputs; hash offsets are relative to it
    b2::RingArena data, meta;                // meta: the hash groups' offset / length / order arrays (tail of d_ring)
    std::vector<b2::RingChunk *> spare;
    std::vector<void *> graveyard;           // outgrown page-locked metadata buffers
    std::vector<b2::RingListing> listings;
    uint64_t next_ticket = 1;
    cudaStream_t copy = nullptr, d2h = nullptr, fin = nullptr;
    cudaStream_t hash[b2::kRingHashStreams] = {};
    cudaStream_t resize[b2::kRingResizeStreams] = {};
    uint32_t next_hash = 0, next_resize = 0;
    uint32_t n_hash = b2::kRingHashStreams;  // streams in use (B2_RING_HASH_STREAMS: experiments)
    uint64_t stalls = 0;                     // times submit() had to wait for ring space
    cudaEvent_t t0 = nullptr;                // trace mode: time origin
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
    if (!r->spare.empty() && !ring_trace()) {               // trace mode never reuses a record: its events are read at wait()
        *out = r->spare.back();
        r->spare.pop_back();
        return cudaSuccess;
    }
    RingChunk *c = new (std::nothrow) RingChunk();
    if (!c) return cudaErrorMemoryAllocation;
    cudaError_t e = cudaSuccess;
    for (cudaEvent_t *ev : {&c->copied, &c->hashed, &c->resized, &c->flushed})
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, ring_trace() ? cudaEventDefault : cudaEventDisableTiming);
    if (e != cudaSuccess) { ring_free_chunk(c); return e; }
    *out = c;
    return cudaSuccess;
}

// Page-locked metadata staging of a chunk record.  Never freed while the ring lives: cudaFreeHost synchronises the
// whole device, i.e. waits for every hash kernel in flight (measured: a record that alternated between the 1 KB chunk
// role and the 6 KB group role was reallocated on every reuse and serialised the stream at one listing per 0.9 s).
static cudaError_t ring_meta_reserve(b2_ingest_ring *r, RingChunk *c, uint64_t bytes) {
    if (c->h_meta_cap >= bytes) return cudaSuccess;
    if (c->h_meta) r->graveyard.push_back(c->h_meta);        // freed by destroy
    c->h_meta = nullptr;
    c->h_meta_cap = 0;
    size_t want = 128 << 10;
    while (want < bytes) want *= 2;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&c->h_meta), want, cudaHostAllocDefault);
    if (e == cudaSuccess) c->h_meta_cap = want;
    return e;
}

// Release the oldest live region of an arena (waiting for it when `block`).  Returns 1 = released, 0 = nothing could
// be released now, 2 = the oldest region belongs to a hash group that has not been launched (the caller must close
// the group first), -1 = CUDA error in *err.
static int ring_retire_oldest(b2_ingest_ring *r, RingArena &a, bool block, cudaError_t *err) {
    if (a.live.empty()) return 0;
    RingChunk *c = a.live.front();
    if (c->hash_pending) return 2;
    if (!block) {
        if (cudaEventQuery(c->hashed) != cudaSuccess || cudaEventQuery(c->flushed) != cudaSuccess) {
            cudaGetLastError();                              // cudaErrorNotReady is not an error
            return 0;
        }
    } else {
        cudaError_t e = cudaEventSynchronize(c->hashed);
        if (e == cudaSuccess) e = cudaEventSynchronize(c->flushed);
        if (e != cudaSuccess) { *err = e; return -1; }
    }
    a.live.pop_front();
    ++c->gen;
    r->spare.push_back(c);
    return 1;
}

// First-in first-out allocation of `bytes` (a multiple of 256) in an arena; blocks while it is full.
// Returns 0 = ok, 2 = blocked by the open hash group (close it and retry), -1 = CUDA error in *err.
static int ring_alloc(b2_ingest_ring *r, RingArena &a, uint64_t bytes, uint64_t *off, cudaError_t *err) {
    while (ring_retire_oldest(r, a, false, err) == 1) {}
    for (;;) {
        if (a.live.empty()) a.head = 0;
        uint64_t cand = a.head + bytes <= a.bytes ? a.head : 0;
        bool clash = false;
        for (const RingChunk *c : a.live)
            if (cand < c->off + c->bytes && c->off < cand + bytes) { clash = true; break; }
        if (!clash) {
            *off = cand;
            a.head = cand + bytes;
            return 0;
        }
        ++r->stalls;
        const int rc = ring_retire_oldest(r, a, true, err);
        if (rc == 2 || rc == -1) return rc;
        if (rc == 0) { *err = cudaErrorUnknown; return -1; }
    }
}

static void ring_drain(b2_ingest_ring *r) {                  // error path: nothing of a failed submit stays in flight
    cudaStreamSynchronize(r->copy);
    for (auto st : r->hash) if (st) cudaStreamSynchronize(st);
    for (auto st : r->resize) if (st) cudaStreamSynchronize(st);
    cudaStreamSynchronize(r->d2h);
    cudaStreamSynchronize(r->fin);
    cudaGetLastError();
    for (RingArena *a : {&r->data, &r->meta}) {
        for (RingChunk *c : a->live) { ++c->gen; c->hash_pending = false; r->spare.push_back(c); }
        a->live.clear();
        a->head = 0;
    }
}

int cached_plan(int device, int ih, int iw, int oh, int ow, b2_resize_plan **out);   // host.cu

}  // namespace b2

extern "C" int b2_ingest_ring_destroy(b2_ingest_ring *r) {
    if (!r) return B2_OK;
    cudaSetDevice(r->device);
    if (r->copy) b2::ring_drain(r);
    for (auto *c : r->spare) b2::ring_free_chunk(c);
    for (void *p : r->graveyard) cudaFreeHost(p);
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
    if (const char *e = getenv("B2_RING_HASH_GROUP_MB"))
        if (atoll(e) >= 1) r->hash_group_bytes = uint64_t(atoll(e)) << 20;
    if (const char *e = getenv("B2_RING_HASH_STREAMS"))
        r->n_hash = uint32_t(atoi(e)) >= 1 && uint32_t(atoi(e)) <= uint32_t(kRingHashStreams) ? uint32_t(atoi(e)) : r->n_hash;
#define B2_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            b2_ingest_ring_destroy(r);                                                                  \
            return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)
    B2_TRY(cudaMalloc(&r->d_ring, size_t(r->ring_bytes)));
    r->meta.bytes = std::min<uint64_t>(64ull << 20, r->ring_bytes / 16) & ~uint64_t(255);     // 20 bytes per message in flight
    r->data.base = r->d_ring;
    r->data.bytes = r->ring_bytes - r->meta.bytes;
    r->meta.base = r->d_ring + r->data.bytes;
    int prio_low = 0, prio_high = 0;
    B2_TRY(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
    B2_TRY(cudaStreamCreateWithPriority(&r->copy, cudaStreamNonBlocking, prio_low));
    B2_TRY(cudaStreamCreateWithPriority(&r->d2h, cudaStreamNonBlocking, prio_low));
    B2_TRY(cudaStreamCreateWithPriority(&r->fin, cudaStreamNonBlocking, prio_low));
    // the hash kernels are the long pole: their CTAs are placed first
    for (auto &st : r->hash) B2_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_high));
    for (auto &st : r->resize) B2_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_low));
    if (ring_trace()) {
        B2_TRY(cudaEventCreate(&r->t0));
        B2_TRY(cudaEventRecord(r->t0, r->copy));
    }
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
    L->trace_groups.clear();
    if (m) {
        B2_TRY(cudaMemcpyAsync(L->d_existing, h_existing_sorted, size_t(m) * 32, cudaMemcpyHostToDevice, r->copy));
        L->h2d += m * 32;
    }
    if (h_valid) {
        B2_TRY(cudaMemcpyAsync(L->d_valid, h_valid, n, cudaMemcpyHostToDevice, r->copy));
        L->h2d += n;
    }

    // Hash GROUPS: the messages of several consecutive chunks are hashed by ONE launch (lanes ordered by decreasing
    // length over the whole group), issued when the group's last chunk has been copied.  A hash kernel lives as long
    // as its longest message (0.8 s for 50 MB) and the GPU runs only a few kernels per hardware connection at once
    // (measured: 1 connection ~8 concurrent kernels, the default 8 connections ~20), so the bytes one hash launch
    // carries — not the copy granularity — decide how many bytes can hash at the same time.
    const uint64_t group_bytes = std::max<uint64_t>(chunk_bytes, std::min<uint64_t>(r->hash_group_bytes, listing_bytes / 2));
    struct Msg { uint64_t off, len; };
    std::vector<Msg> gmsgs;                                  // messages of the open hash group
    std::vector<RingChunk *> gchunks;                        // its data chunks
    uint32_t g_lo = 0;                                       // listing position of its first message
    uint64_t g_bytes = 0;
    std::vector<uint32_t> order;
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> groups;

    auto close_hash_group = [&](cudaError_t *ce) -> int {
        *ce = cudaSuccess;
        const uint32_t cnt = uint32_t(gmsgs.size());
        if (!cnt) return B2_OK;
        // the group's metadata gets its own small region of the ring (released after the hash)
        const uint64_t o_hlen = 8ull * cnt, o_order = 16ull * cnt, meta_bytes = 20ull * cnt, total = up(meta_bytes, 256);
        RingChunk *c = nullptr;
        if ((*ce = ring_new_chunk(r, &c)) != cudaSuccess) return B2_ERR_CUDA;
        if ((*ce = ring_meta_reserve(r, c, meta_bytes)) != cudaSuccess) {
            r->spare.push_back(c);
            return B2_ERR_CUDA;
        }
        if (total > r->meta.bytes) { r->spare.push_back(c); return fail(B2_ERR_BAD_ARG, "b2_ingest_ring_submit: hash group of %u messages exceeds the metadata arena", cnt); }
        uint64_t off = 0;
        if (ring_alloc(r, r->meta, total, &off, ce) != 0) { r->spare.push_back(c); return B2_ERR_CUDA; }   // never blocked by an open group
        c->off = off;
        c->bytes = total;
        c->hash_pending = false;
        r->meta.live.push_back(c);
        uint64_t *m_hoff = reinterpret_cast<uint64_t *>(c->h_meta), *m_hlen = reinterpret_cast<uint64_t *>(c->h_meta + o_hlen);
        uint32_t *m_order = reinterpret_cast<uint32_t *>(c->h_meta + o_order);
        for (uint32_t j = 0; j < cnt; ++j) { m_hoff[j] = gmsgs[j].off; m_hlen[j] = gmsgs[j].len; }
        order.resize(cnt);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return gmsgs[x].len > gmsgs[y].len; });
        memcpy(m_order, order.data(), 4ull * cnt);
        uint8_t *d_meta = r->meta.base + c->off;
        if ((*ce = cudaMemcpyAsync(d_meta, c->h_meta, size_t(meta_bytes), cudaMemcpyHostToDevice, r->copy)) != cudaSuccess) return B2_ERR_CUDA;
        L->h2d += meta_bytes;
        if ((*ce = cudaEventRecord(c->copied, r->copy)) != cudaSuccess) return B2_ERR_CUDA;   // after every data chunk of the group too
        cudaStream_t hs = r->hash[r->next_hash++ % r->n_hash];
        if ((*ce = cudaStreamWaitEvent(hs, c->copied, 0)) != cudaSuccess) return B2_ERR_CUDA;
        int rc = b2_sha256_batch(r->d_ring, reinterpret_cast<const uint64_t *>(d_meta), reinterpret_cast<const uint64_t *>(d_meta + o_hlen),
                                 reinterpret_cast<const uint32_t *>(d_meta + o_order), cnt, L->d_digests + size_t(g_lo) * 32, hs);
        if (rc != B2_OK) return rc;
        ++L->launches;
        if ((*ce = cudaEventRecord(c->hashed, hs)) != cudaSuccess) return B2_ERR_CUDA;
        for (RingChunk *dc : gchunks) {                      // the data chunks are released by the same moment
            if ((*ce = cudaEventRecord(dc->hashed, hs)) != cudaSuccess) return B2_ERR_CUDA;
            dc->hash_pending = false;
        }
        if ((*ce = cudaEventRecord(c->flushed, hs)) != cudaSuccess) return B2_ERR_CUDA;
        if ((*ce = cudaStreamWaitEvent(r->fin, c->hashed, 0)) != cudaSuccess) return B2_ERR_CUDA;
        if (ring_trace()) L->trace_groups.push_back({c, uint32_t(gchunks.size()), cnt, g_bytes});
        gmsgs.clear();
        gchunks.clear();
        g_bytes = 0;
        return B2_OK;
    };
#define B2_CLOSE_GROUP()                                                                                \
    do {                                                                                                \
        cudaError_t _ce = cudaSuccess;                                                                  \
        int _rc = close_hash_group(&_ce);                                                               \
        if (_ce != cudaSuccess) { B2_TRY(_ce); }                                                        \
        if (_rc != B2_OK) { ring_drain(r); return _rc; }                                                \
    } while (0)

    for (uint32_t lo = 0; lo < n;) {
        // ---- chunk = consecutive listing entries up to chunk_bytes: the unit of copy, resize and read-back
        uint32_t cnt = 0;
        uint64_t data_bytes = 0;
        while (lo + cnt < n && cnt < kRingMaxChunkImages) {
            const uint64_t b = img_bytes(lo + cnt);
            if (cnt > 0 && data_bytes + b > chunk_bytes) break;
            data_bytes += b;
            ++cnt;
        }
        // layout inside the chunk region (offsets relative to its start)
        const uint64_t o_poff = 0, o_slot = 8ull * cnt, meta_bytes = 12ull * cnt;
        const uint64_t o_data = up(meta_bytes, 256);
        const uint64_t o_thumb = up(o_data + data_bytes + 16, 256);
        const uint64_t thumb_bytes = h_pixels ? uint64_t(cnt) * out_px : 0;
        const uint64_t o_prev = up(o_thumb + thumb_bytes, 256);
        const uint64_t prev_bytes = h_pixels && h_previews ? uint64_t(cnt) * out_px * 4 : 0;
        const uint64_t total = up(o_prev + prev_bytes, 256);
        if (total > r->data.bytes) {
            ring_drain(r);
            return fail(B2_ERR_BAD_ARG, "b2_ingest_ring_submit: image %u needs %llu bytes of staging, the ring has %llu",
                        lo, (unsigned long long)total, (unsigned long long)r->data.bytes);
        }
        if (g_bytes && g_bytes + data_bytes > group_bytes) B2_CLOSE_GROUP();
        uint64_t region = 0;
        for (;;) {
            cudaError_t e = cudaSuccess;
            const int st = ring_alloc(r, r->data, total, &region, &e);
            if (st == 2) {                                   // the ring is full of the OPEN group's chunks: hash what is there
                B2_CLOSE_GROUP();
                continue;
            }
            if (st != 0) B2_TRY(e != cudaSuccess ? e : cudaErrorUnknown);
            break;
        }
        RingChunk *c = nullptr;
        B2_TRY(ring_new_chunk(r, &c));
        {
            cudaError_t e = ring_meta_reserve(r, c, meta_bytes);
            if (e != cudaSuccess) { r->spare.push_back(c); B2_TRY(e); }
        }
        c->off = region;
        c->bytes = total;
        if (gmsgs.empty()) g_lo = lo;
        c->hash_pending = true;
        r->data.live.push_back(c);
        gchunks.push_back(c);
        uint8_t *d_chunk = r->d_ring + c->off;
        uint64_t *m_poff = reinterpret_cast<uint64_t *>(c->h_meta + o_poff);
        uint32_t *m_slot = reinterpret_cast<uint32_t *>(c->h_meta + o_slot);

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
                gmsgs.push_back({pos, ml});
                if (ml) B2_TRY(put(h_files[i], ml));
                pos += up(ml, 16);
            } else {
                gmsgs.push_back({pl ? pix_off[j] : pos, ml});
            }
        }
        B2_TRY(flush_run());
        g_bytes += data_bytes;
        uint32_t g = 0;
        for (auto &kv : groups)
            for (uint32_t j : kv.second) { m_poff[g] = pix_off[j]; m_slot[g] = j; ++g; }
        if (g) {
            B2_TRY(cudaMemcpyAsync(d_chunk, c->h_meta, size_t(meta_bytes), cudaMemcpyHostToDevice, r->copy));
            L->h2d += meta_bytes;
        }
        B2_TRY(cudaEventRecord(c->copied, r->copy));

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
    RingChunk *last_data = L->parts.back().chunk;
    B2_CLOSE_GROUP();
#undef B2_CLOSE_GROUP

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
    B2_TRY(cudaStreamWaitEvent(r->fin, last_data->flushed, 0));   // the d2h stream is in order: last chunk = all chunks
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
    if (ring_trace() && e == cudaSuccess) {
        auto ms = [&](cudaEvent_t ev) { float t = -1.f; if (cudaEventElapsedTime(&t, r->t0, ev) != cudaSuccess) { cudaGetLastError(); t = -1.f; } return t; };
        for (const auto &p : L->parts)
            fprintf(stderr, "[ring] L%llu chunk hi=%u bytes=%llu copied=%.1f resized=%.1f flushed=%.1f hashed=%.1f\n", (unsigned long long)ticket,
                    p.hi, (unsigned long long)p.chunk->bytes, ms(p.chunk->copied), ms(p.chunk->resized), ms(p.chunk->flushed), ms(p.chunk->hashed));
        for (const auto &g : L->trace_groups)
            fprintf(stderr, "[ring] L%llu group chunks=%u msgs=%u bytes=%llu copied=%.1f hashed=%.1f\n", (unsigned long long)ticket, g.chunks, g.msgs,
                    (unsigned long long)g.bytes, ms(g.meta->copied), ms(g.meta->hashed));
    }
    L->trace_groups.clear();
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
    for (const RingChunk *c : r->data.live) b += c->bytes;
    for (const RingChunk *c : r->meta.live) b += c->bytes;
    if (ring_bytes) *ring_bytes = r->ring_bytes;
    if (bytes_in_flight) *bytes_in_flight = b;
    if (chunks_in_flight) *chunks_in_flight = uint32_t(r->data.live.size());
    if (stalls) *stalls = r->stalls;
    return B2_OK;
}
