// Host-buffer entry points of libb2ingest: the end-to-end form of the path, callable with HOST pointers only.
//
// The reference service is plain Python without a tensor library; what it holds when the path starts is
// bytes in host memory (downloaded files, app/services/webdav_sync.py:441, or rows fetched from the table).
// These entry points own the device side themselves — staging buffers, streams, events — so a caller needs
// nothing but ctypes:
//
//   b2_sha256_host / b2_dedupe_host / b2_thumbnails_host   the service's 50-file batches (blocking)
//   b2_label_tally_host  label rows in host memory -> count matrix (optional) + integer Fleiss partials.
//   b2_host_alloc/free   page-locked host memory (what makes the copies asynchronous and full speed).
#include "common.cuh"

#include <algorithm>
#include <map>
#include <cstring>
#include <mutex>
#include <new>
#include <numeric>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

namespace b2 {

// Page-locked staging buffers for the variable-size host calls, recycled between calls (cudaHostAlloc costs
// milliseconds).  A buffer is owned by one call at a time; the list only grows to what concurrent callers need.
struct PinnedPool {
    struct Buf { void *p; size_t bytes; };
    std::mutex mu;
    std::vector<Buf> free_list;
    void *acquire(size_t bytes, size_t *got) {
        {
            std::lock_guard<std::mutex> lock(mu);
            int best = -1;
            for (int i = 0; i < int(free_list.size()); ++i)
                if (free_list[i].bytes >= bytes && (best < 0 || free_list[i].bytes < free_list[best].bytes)) best = i;
            if (best >= 0) {
                Buf b = free_list[best];
                free_list.erase(free_list.begin() + best);
                *got = b.bytes;
                return b.p;
            }
        }
        size_t want = (bytes + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);        // whole MiB
        void *p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) return nullptr;
        *got = want;
        return p;
    }
    void release(void *p, size_t bytes) {
        std::lock_guard<std::mutex> lock(mu);
        if (free_list.size() >= 8) {                         // keep the largest few
            auto it = std::min_element(free_list.begin(), free_list.end(),
                                       [](const Buf &a, const Buf &b) { return a.bytes < b.bytes; });
            if (it->bytes < bytes) { cudaFreeHost(it->p); *it = Buf{p, bytes}; } else cudaFreeHost(p);
            return;
        }
        free_list.push_back(Buf{p, bytes});
    }
    void clear() {
        std::lock_guard<std::mutex> lock(mu);
        for (auto &b : free_list) cudaFreeHost(b.p);
        free_list.clear();
    }
};
static PinnedPool g_pinned;

static void keep_pool_memory(int device) {                   // freed device blocks stay in the pool between calls
    static std::once_flag once[64];
    std::call_once(once[device & 63], [&] {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    });
}

}  // namespace b2

extern "C" int b2_host_alloc(void **p, uint64_t bytes) {
    using namespace b2;
    B2_REQUIRE(p != nullptr && bytes > 0, "b2_host_alloc: bad argument");
    B2_CUDA_CHECK(cudaHostAlloc(p, size_t(bytes), cudaHostAllocDefault));
    return B2_OK;
}

extern "C" int b2_host_free(void *p) {
    using namespace b2;
    if (p) B2_CUDA_CHECK(cudaFreeHost(p));
    return B2_OK;
}

// Label rows in host memory -> (optional) count matrix and partials in host memory.  Blocking.
extern "C" int b2_label_tally_host(int device, const int32_t *h_image_idx, const uint8_t *h_class_idx,
                                   const uint8_t *h_active, uint64_t rows, uint32_t image_base, uint32_t n_images,
                                   uint32_t k, uint32_t flags, int32_t *h_counts, int64_t *h_partials,
                                   int64_t *h_agree_hist) {
    using namespace b2;
    B2_REQUIRE(h_partials != nullptr, "b2_label_tally_host: null partials");
    B2_REQUIRE(rows == 0 || (h_image_idx && h_class_idx && h_active), "b2_label_tally_host: null row pointer");
    B2_REQUIRE(k >= 1 && k <= 256 && n_images >= 1, "b2_label_tally_host: bad shape");
    int rc = b2_init(device);
    if (rc != B2_OK) return rc;
    cudaStream_t st = nullptr;
    int32_t *d_img = nullptr, *d_counts = nullptr;
    uint8_t *d_cls = nullptr, *d_act = nullptr;
    int64_t *d_part = nullptr, *d_hist = nullptr;
    // stream-ordered allocations: served from the device's memory pool after the first call, no device-wide
    // synchronisation on free
    auto cleanup = [&]() {
        if (st) {
            if (d_hist) cudaFreeAsync(d_hist, st);
            if (d_img) cudaFreeAsync(d_img, st);
            if (d_cls) cudaFreeAsync(d_cls, st);
            if (d_act) cudaFreeAsync(d_act, st);
            if (d_counts) cudaFreeAsync(d_counts, st);
            if (d_part) cudaFreeAsync(d_part, st);
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    };
#define B2_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            cleanup();                                                                                  \
            return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)
    B2_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    keep_pool_memory(device);
    const size_t r = size_t(rows ? rows : 1);
    B2_TRY(cudaMallocAsync(&d_img, r * 4, st));
    B2_TRY(cudaMallocAsync(&d_cls, r, st));
    B2_TRY(cudaMallocAsync(&d_act, r, st));
    B2_TRY(cudaMallocAsync(&d_counts, size_t(n_images) * k * 4, st));
    B2_TRY(cudaMallocAsync(&d_part, (size_t(k) + B2_PARTIALS_EXTRA) * 8, st));
    if (h_agree_hist) B2_TRY(cudaMallocAsync(&d_hist, size_t(B2_AGREE_BINS) * 8, st));
    if (rows) {
        B2_TRY(cudaMemcpyAsync(d_img, h_image_idx, rows * 4, cudaMemcpyHostToDevice, st));
        B2_TRY(cudaMemcpyAsync(d_cls, h_class_idx, rows, cudaMemcpyHostToDevice, st));
        B2_TRY(cudaMemcpyAsync(d_act, h_active, rows, cudaMemcpyHostToDevice, st));
    }
    rc = b2_label_tally(d_img, d_cls, d_act, rows, image_base, n_images, k, flags, d_counts, d_part, d_hist, st);
    if (rc != B2_OK) { cleanup(); return rc; }
    B2_TRY(cudaMemcpyAsync(h_partials, d_part, (size_t(k) + B2_PARTIALS_EXTRA) * 8, cudaMemcpyDeviceToHost, st));
    if (h_agree_hist) B2_TRY(cudaMemcpyAsync(h_agree_hist, d_hist, size_t(B2_AGREE_BINS) * 8, cudaMemcpyDeviceToHost, st));
    if (h_counts) B2_TRY(cudaMemcpyAsync(h_counts, d_counts, size_t(n_images) * k * 4, cudaMemcpyDeviceToHost, st));
    B2_TRY(cudaStreamSynchronize(st));
#undef B2_TRY
    cleanup();
    return b2_label_tally_status(h_partials, k, rows);
}

// Rows sorted by (annotator, image) in host memory -> distinct active images per annotator in host memory
// (the bulk form of app/api/routes/classificacoes.py:224-230).  Blocking.
extern "C" int b2_distinct_images_host(int device, const int32_t *h_annotator_idx, const int32_t *h_image_idx,
                                       const uint8_t *h_active, uint64_t rows, uint32_t n_annotators,
                                       uint32_t *h_distinct) {
    using namespace b2;
    B2_REQUIRE(h_distinct != nullptr && n_annotators >= 1, "b2_distinct_images_host: bad output");
    B2_REQUIRE(rows == 0 || (h_annotator_idx && h_image_idx && h_active), "b2_distinct_images_host: null row pointer");
    int rc = b2_init(device);
    if (rc != B2_OK) return rc;
    keep_pool_memory(device);
    cudaStream_t st = nullptr;
    uint8_t *d = nullptr;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t r = size_t(rows ? rows : 1);
    const size_t o_ann = 0, o_img = up(r * 4), o_act = o_img + up(r * 4), o_out = o_act + up(r), bytes = o_out + up(size_t(n_annotators) * 4);
    auto cleanup = [&]() {
        if (st) {
            if (d) cudaFreeAsync(d, st);
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    };
#define B2_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            cleanup();                                                                                  \
            return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)
    B2_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    B2_TRY(cudaMallocAsync(&d, bytes, st));
    if (rows) {
        B2_TRY(cudaMemcpyAsync(d + o_ann, h_annotator_idx, rows * 4, cudaMemcpyHostToDevice, st));
        B2_TRY(cudaMemcpyAsync(d + o_img, h_image_idx, rows * 4, cudaMemcpyHostToDevice, st));
        B2_TRY(cudaMemcpyAsync(d + o_act, h_active, rows, cudaMemcpyHostToDevice, st));
    }
    rc = b2_distinct_images_per_annotator(reinterpret_cast<const int32_t *>(d + o_ann), reinterpret_cast<const int32_t *>(d + o_img),
                                          d + o_act, rows, n_annotators, reinterpret_cast<uint32_t *>(d + o_out), st);
    if (rc != B2_OK) { cleanup(); return rc; }
    B2_TRY(cudaMemcpyAsync(h_distinct, d + o_out, size_t(n_annotators) * 4, cudaMemcpyDeviceToHost, st));
    B2_TRY(cudaStreamSynchronize(st));
#undef B2_TRY
    cleanup();
    return B2_OK;
}

// ---- variable-size messages in host memory (the service's 50-file batches) -------------------------------
// h_msgs[i] points to h_lens[i] bytes anywhere in host memory (e.g. the buffers of Python bytes objects).  The
// messages are packed, 16-byte aligned, into one recycled page-locked buffer, copied in one transfer, hashed in
// order of decreasing length (a warp's 32 lanes finish together) and the digests (and optionally their lowercase
// hex form, the String(64) primary key) are read back.  Blocking.
extern "C" int b2_sha256_host(int device, const uint8_t *const *h_msgs, const uint64_t *h_lens, uint32_t n,
                              uint8_t *h_digests, char *h_hex) {
    using namespace b2;
    if (n == 0) return B2_OK;
    B2_REQUIRE(h_msgs && h_lens && h_digests, "b2_sha256_host: null pointer");
    int rc = b2_init(device);
    if (rc != B2_OK) return rc;
    keep_pool_memory(device);
    std::vector<uint64_t> meta(size_t(n) * 2);               // offsets then lengths
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i) {
        B2_REQUIRE(h_lens[i] == 0 || h_msgs[i] != nullptr, "b2_sha256_host: message %u is NULL", i);
        meta[i] = total;
        meta[n + i] = h_lens[i];
        total += (h_lens[i] + 15) & ~uint64_t(15);
    }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h_lens[a] > h_lens[b]; });
    const size_t data_bytes = size_t(total ? total : 16);
    const size_t meta_off = (data_bytes + 255) & ~size_t(255);
    const size_t order_off = meta_off + size_t(n) * 16;
    const size_t in_bytes = order_off + size_t(n) * 4;       // staged input: data | offsets | lengths | order
    const size_t out_bytes = size_t(n) * 96;                 // digests | hex
    size_t got = 0;
    uint8_t *stage = static_cast<uint8_t *>(g_pinned.acquire(in_bytes + out_bytes, &got));
    if (!stage) return fail(B2_ERR_CUDA, "b2_sha256_host: cannot allocate %zu bytes of page-locked memory", in_bytes + out_bytes);
    // pack: a single thread copies ~7 GB/s; large batches are split over a few threads by bytes
    {
        const unsigned hw = std::thread::hardware_concurrency();
        unsigned nt = total >= (8u << 20) ? (hw >= 8 ? 8u : (hw ? hw : 1u)) : 1u;
        if (nt > n) nt = n;
        auto pack = [&](uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo; i < hi; ++i)
                if (h_lens[i]) memcpy(stage + meta[i], h_msgs[i], size_t(h_lens[i]));
        };
        if (nt <= 1) {
            pack(0, n);
        } else {
            std::vector<std::thread> workers;
            uint32_t lo = 0;
            for (unsigned t = 0; t < nt; ++t) {                  // cut where the byte offset passes (t+1)/nt of the total
                uint32_t hi = lo;
                const uint64_t target = total / nt * (t + 1);
                while (hi < n && (t + 1 == nt || meta[hi] < target)) ++hi;
                if (hi > lo) workers.emplace_back(pack, lo, hi);
                lo = hi;
            }
            for (auto &w : workers) w.join();
        }
    }
    memcpy(stage + meta_off, meta.data(), size_t(n) * 16);
    memcpy(stage + order_off, order.data(), size_t(n) * 4);
    cudaStream_t st = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    auto cleanup = [&]() {
        if (st) {
            if (d_in) cudaFreeAsync(d_in, st);
            if (d_out) cudaFreeAsync(d_out, st);
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        g_pinned.release(stage, got);
    };
#define B2_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            cleanup();                                                                                  \
            return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)
    B2_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    B2_TRY(cudaMallocAsync(&d_in, in_bytes, st));
    B2_TRY(cudaMallocAsync(&d_out, out_bytes, st));
    B2_TRY(cudaMemcpyAsync(d_in, stage, in_bytes, cudaMemcpyHostToDevice, st));
    rc = b2_sha256_batch(d_in, reinterpret_cast<const uint64_t *>(d_in + meta_off),
                         reinterpret_cast<const uint64_t *>(d_in + meta_off) + n,
                         reinterpret_cast<const uint32_t *>(d_in + order_off), n, d_out, st);
    if (rc == B2_OK && h_hex) rc = b2_digest_hex(d_out, n, reinterpret_cast<char *>(d_out + size_t(n) * 32), st);
    if (rc != B2_OK) { cleanup(); return rc; }
    uint8_t *h_out = stage + in_bytes;
    B2_TRY(cudaMemcpyAsync(h_out, d_out, h_hex ? out_bytes : size_t(n) * 32, cudaMemcpyDeviceToHost, st));
    B2_TRY(cudaStreamSynchronize(st));
#undef B2_TRY
    memcpy(h_digests, h_out, size_t(n) * 32);
    if (h_hex) memcpy(h_hex, h_out + size_t(n) * 32, size_t(n) * 64);
    cleanup();
    return B2_OK;
}

// The dedupe decision of b2_dedupe for digests held in host memory.  Blocking.
extern "C" int b2_dedupe_host(int device, const uint8_t *h_digests, const uint8_t *h_valid, uint32_t n,
                              const uint8_t *h_existing_sorted, uint64_t m, uint8_t *h_is_new,
                              int32_t *h_first_index, int32_t *h_last_index, uint32_t *h_counts) {
    using namespace b2;
    B2_REQUIRE(h_counts != nullptr, "b2_dedupe_host: null counts");
    h_counts[0] = h_counts[1] = h_counts[2] = 0;
    if (n == 0) return B2_OK;
    B2_REQUIRE(h_digests && h_is_new, "b2_dedupe_host: null pointer");
    B2_REQUIRE(m == 0 || h_existing_sorted != nullptr, "b2_dedupe_host: null existing table");
    int rc = b2_init(device);
    if (rc != B2_OK) return rc;
    keep_pool_memory(device);
    const uint64_t ws_bytes = b2_dedupe_workspace_bytes(n);
    // one device block: digests | existing | valid | is_new | first | last | counts | workspace (256-byte aligned parts)
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t o_dig = 0, o_ex = up(size_t(n) * 32), o_val = o_ex + up(size_t(m) * 32), o_new = o_val + up(n),
                 o_first = o_new + up(n), o_last = o_first + up(size_t(n) * 4), o_cnt = o_last + up(size_t(n) * 4),
                 o_ws = o_cnt + 256, bytes = o_ws + up(size_t(ws_bytes ? ws_bytes : 8));
    cudaStream_t st = nullptr;
    uint8_t *d = nullptr;
    auto cleanup = [&]() {
        if (st) {
            if (d) cudaFreeAsync(d, st);
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    };
#define B2_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            cleanup();                                                                                  \
            return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)
    B2_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    B2_TRY(cudaMallocAsync(&d, bytes, st));
    B2_TRY(cudaMemcpyAsync(d + o_dig, h_digests, size_t(n) * 32, cudaMemcpyHostToDevice, st));
    if (m) B2_TRY(cudaMemcpyAsync(d + o_ex, h_existing_sorted, size_t(m) * 32, cudaMemcpyHostToDevice, st));
    if (h_valid) B2_TRY(cudaMemcpyAsync(d + o_val, h_valid, n, cudaMemcpyHostToDevice, st));
    rc = b2_dedupe(d + o_dig, h_valid ? d + o_val : nullptr, nullptr, n, m ? d + o_ex : nullptr, m, d + o_new,
                   reinterpret_cast<int32_t *>(d + o_first), reinterpret_cast<int32_t *>(d + o_last),
                   reinterpret_cast<uint32_t *>(d + o_cnt), d + o_ws, ws_bytes, st);
    if (rc != B2_OK) { cleanup(); return rc; }
    B2_TRY(cudaMemcpyAsync(h_is_new, d + o_new, n, cudaMemcpyDeviceToHost, st));
    if (h_first_index) B2_TRY(cudaMemcpyAsync(h_first_index, d + o_first, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
    if (h_last_index) B2_TRY(cudaMemcpyAsync(h_last_index, d + o_last, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
    B2_TRY(cudaMemcpyAsync(h_counts, d + o_cnt, 12, cudaMemcpyDeviceToHost, st));
    B2_TRY(cudaStreamSynchronize(st));
#undef B2_TRY
    cleanup();
    return B2_OK;
}

// ---- decoded images of any mix of shapes in host memory -> thumbnails / previews in host memory ----------------
// h_rgb[i] points to image i (HWC uint8, row pitch 3*w, no padding), h_hw[2i], h_hw[2i+1] = its height and width.
// Images are grouped by shape; each group is packed into recycled page-locked staging, copied, resized by one
// launch with the group's cached tap plan (outputs land at the images' own positions) and read back.  Blocking.
namespace b2 {
struct PlanCache {
    std::mutex mu;
    std::map<std::tuple<int, int, int, int, int>, b2_resize_plan *> plans;   // (device, in_h, in_w, out_h, out_w)
    int get(int device, int ih, int iw, int oh, int ow, b2_resize_plan **out) {
        std::lock_guard<std::mutex> lock(mu);
        auto key = std::make_tuple(device, ih, iw, oh, ow);
        auto it = plans.find(key);
        if (it != plans.end()) { *out = it->second; return B2_OK; }
        b2_resize_plan *pl = nullptr;
        const int rc = b2_resize_plan_create(ih, iw, oh, ow, &pl);
        if (rc != B2_OK) return rc;
        plans[key] = pl;                                     // plans live until b2_shutdown()
        *out = pl;
        return B2_OK;
    }
    void clear() {
        std::lock_guard<std::mutex> lock(mu);
        for (auto &kv : plans) {
            cudaSetDevice(std::get<0>(kv.first));
            b2_resize_plan_destroy(kv.second);
        }
        plans.clear();
    }
};
static PlanCache g_plans;

// Tap plan for one (device, shape) pair, created on first use and shared by the host-pointer entry points
// (b2_thumbnails_host, b2_ingest_ring_*).  The current device must be `device`.
int cached_plan(int device, int ih, int iw, int oh, int ow, b2_resize_plan **out) {
    return g_plans.get(device, ih, iw, oh, ow, out);
}
}  // namespace b2

// Release what the host-pointer entry points keep between calls: the pool of page-locked staging buffers and the
// cached resize plans (b2_shutdown).  Objects the caller created (plans, rings,
// communicators) are the caller's to destroy first; no other call may be in flight.  The library can be used
// again afterwards (everything is re-created on demand).
extern "C" int b2_shutdown(void) {
    using namespace b2;
    int dev = 0;
    const bool have_dev = cudaGetDevice(&dev) == cudaSuccess;
    g_plans.clear();
    g_pinned.clear();
    if (have_dev) cudaSetDevice(dev);
    cudaGetLastError();
    return B2_OK;
}

extern "C" int b2_thumbnails_host(int device, const uint8_t *const *h_rgb, const uint32_t *h_hw, uint32_t n,
                                  uint32_t out_h, uint32_t out_w, uint8_t *h_thumbs, float *h_previews,
                                  const float mean[3], const float inv_std[3]) {
    using namespace b2;
    if (n == 0) return B2_OK;
    B2_REQUIRE(h_rgb && h_hw && h_thumbs, "b2_thumbnails_host: null pointer");
    B2_REQUIRE(out_h >= 1 && out_w >= 1, "b2_thumbnails_host: empty output shape");
    int rc = b2_init(device);
    if (rc != B2_OK) return rc;
    keep_pool_memory(device);
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> groups;
    for (uint32_t i = 0; i < n; ++i) {
        B2_REQUIRE(h_rgb[i] != nullptr && h_hw[2 * i] >= 1 && h_hw[2 * i + 1] >= 1, "b2_thumbnails_host: image %u is empty", i);
        groups[{h_hw[2 * i], h_hw[2 * i + 1]}].push_back(i);
    }
    const size_t out_px = size_t(out_h) * out_w * 3;
    cudaStream_t st = nullptr;
    uint8_t *d_thumbs = nullptr, *d_in = nullptr;
    float *d_prev = nullptr;
    uint8_t *stage = nullptr;
    size_t got = 0;
    auto cleanup = [&]() {
        if (st) {
            if (d_thumbs) cudaFreeAsync(d_thumbs, st);
            if (d_prev) cudaFreeAsync(d_prev, st);
            if (d_in) cudaFreeAsync(d_in, st);
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        if (stage) g_pinned.release(stage, got);
    };
#define B2_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            cleanup();                                                                                  \
            return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                               \
    } while (0)
    B2_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    B2_TRY(cudaMallocAsync(&d_thumbs, size_t(n) * out_px, st));
    if (h_previews) B2_TRY(cudaMallocAsync(&d_prev, size_t(n) * out_px * 4, st));
    // largest group decides the staging size: images (16-byte aligned starts) | offsets | out slots
    size_t max_bytes = 0;
    for (auto &g : groups) {
        const size_t L = (size_t(g.first.first) * g.first.second * 3 + 15) & ~size_t(15);
        const size_t b = ((L * g.second.size() + 255) & ~size_t(255)) + g.second.size() * 12;
        if (b > max_bytes) max_bytes = b;
    }
    stage = static_cast<uint8_t *>(g_pinned.acquire(max_bytes, &got));
    if (!stage) { cleanup(); return fail(B2_ERR_CUDA, "b2_thumbnails_host: cannot allocate %zu bytes of page-locked memory", max_bytes); }
    B2_TRY(cudaMallocAsync(&d_in, max_bytes, st));
    for (auto &g : groups) {
        const uint32_t ih = g.first.first, iw = g.first.second, m = uint32_t(g.second.size());
        const size_t bytes = size_t(ih) * iw * 3, L = (bytes + 15) & ~size_t(15);
        const size_t off_off = (L * m + 255) & ~size_t(255), slot_off = off_off + size_t(m) * 8;
        b2_resize_plan *plan = nullptr;
        rc = g_plans.get(device, int(ih), int(iw), int(out_h), int(out_w), &plan);
        if (rc != B2_OK) { cleanup(); return rc; }
        uint64_t *offs = reinterpret_cast<uint64_t *>(stage + off_off);
        uint32_t *slots = reinterpret_cast<uint32_t *>(stage + slot_off);
        for (uint32_t j = 0; j < m; ++j) {
            memcpy(stage + L * j, h_rgb[g.second[j]], bytes);
            offs[j] = L * j;
            slots[j] = g.second[j];
        }
        B2_TRY(cudaMemcpyAsync(d_in, stage, slot_off + size_t(m) * 4, cudaMemcpyHostToDevice, st));
        rc = b2_resize_normalize_batch(plan, d_in, reinterpret_cast<const uint64_t *>(d_in + off_off),
                                       reinterpret_cast<const uint32_t *>(d_in + slot_off), m, d_thumbs, d_prev, mean, inv_std, st);
        if (rc != B2_OK) { cleanup(); return rc; }
        B2_TRY(cudaStreamSynchronize(st));                   // the staging buffer is reused by the next group
    }
    B2_TRY(cudaMemcpyAsync(h_thumbs, d_thumbs, size_t(n) * out_px, cudaMemcpyDeviceToHost, st));
    if (h_previews) B2_TRY(cudaMemcpyAsync(h_previews, d_prev, size_t(n) * out_px * 4, cudaMemcpyDeviceToHost, st));
    B2_TRY(cudaStreamSynchronize(st));
#undef B2_TRY
    cleanup();
    return B2_OK;
}
