// Per-image label tally + integer Fleiss partials for sm_100a.
//
// The reference only groups ONE user's active rows by image
// (app/crud/classificacao_crud.py:318-322) and counts a user's distinct images
// (app/api/routes/classificacoes.py:224-230); BASELINE.json configs 1,4,5 additionally require
// the cross-annotator tally and Fleiss' kappa.  Rows are the dictionary-encoded columns of
// table `classificacoes` (app/db/models.py:224-241): image_idx int32 (id_img), class_idx uint8
// (id_opc), active uint8 (ativo); only rows with active != 0 count (classificacao_crud.py:314).
//
// tally_slab_kernel — rows ordered by image_idx (an index scan on id_img).  HBM-bound: 6 B read per row +
//   4*k B written per image, every byte touched once; a thread per row, lanes a slab apart, shared-memory
//   atomics into a sliding window of images (described in full above the kernel).  It replaced, in this order,
//   a shared-atomic tile kernel, two warp-cooperative (ballot / MATCH.ANY) kernels and a thread-per-image
//   kernel, all 3-4x slower (profiles/r1_tally_history.md).
//
// tally_scatter_kernel — any row order: global RED.ADD into a zeroed count matrix (bound by L2
//   atomic throughput, not HBM), followed by the partials pass below.
//
// fleiss_partials_kernel — partials (and the float64 sum of P_i for the general-n kappa) from an
//   existing count matrix, staged through the same shared-memory tile code.
#include "common.cuh"

namespace b2 {

constexpr int kTallyThreads = 512;
constexpr int kRowsPerThread = 16;
constexpr int kTileBudgetBytes = 92 * 1024;                     // two CTAs per SM (beside the 8 KB agreement histogram)
constexpr int kMaxTileImages = 512;

constexpr uint32_t kAgreeBins = B2_AGREE_BINS;

// indices into d_partials after the k class totals
enum { P_S2 = 0, P_R = 1, P_RATED = 2, P_PAIR_IMAGES = 3, P_PAIRS = 4, P_ROWS_SEEN = 5, P_UNSORTED = 6 };

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-level partial sums live in shared memory (s_part[7], 64-bit): threads fold their
// contribution in with one shared atomic per warp, so nothing stays in registers between flushes.
__device__ __forceinline__ void part_add(unsigned long long *s_part, int which, unsigned long long v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_part[which], v);
}

// Agreement histogram (general-n Fleiss kappa without a second pass, exact for any sharding): bin n (2 <= n <
// kAgreeBins) accumulates sum_j n_ij^2 - n_i over the images with n_i = n, so that sum_i P_i =
// sum_n bin[n] / (n (n - 1)) is computed on the host from INTEGERS; bin 0 counts the images with n_i >= kAgreeBins
// (the caller then falls back to b2_fleiss_partials' float64 sum), bin 1 stays 0.
__device__ __forceinline__ void agree_add(unsigned long long *s_hist, uint32_t n, uint32_t s2) {
    if (n >= kAgreeBins) atomicAdd(&s_hist[0], 1ull);
    else if (n >= 2u) atomicAdd(&s_hist[n], (unsigned long long)(s2 - n));
}
__device__ __forceinline__ void commit_hist(const unsigned long long *s_hist, unsigned long long *g_hist) {
    for (uint32_t i = threadIdx.x; i < kAgreeBins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&g_hist[i], s_hist[i]);
}

// Add the CTA's partials to global memory (integer atomics: exact, order-independent).
__device__ void commit_partials(const unsigned long long *s_part, const unsigned long long *s_class_tot, uint32_t k,
                                unsigned long long *g_partials) {
    if (threadIdx.x < 7 && s_part[threadIdx.x]) atomicAdd(&g_partials[k + threadIdx.x], s_part[threadIdx.x]);
    for (uint32_t c = threadIdx.x; c < k; c += blockDim.x)
        if (s_class_tot[c]) atomicAdd(&g_partials[c], s_class_tot[c]);
}

// Write `n_img` images of the shared tile to global memory and fold them into the partials.
// Pass 1 (element order, coalesced stores): class totals + S2.  Pass 2 (one thread per image):
// n_i and the quantities derived from it.  Leaves the tile zeroed.  Caller syncs before and after.
template <bool kStore, bool kSumPi>
__device__ __noinline__ void flush_tile(int32_t *tile, uint32_t n_img, uint32_t k, int32_t *g_counts,
                                        unsigned long long *s_class_tot, unsigned long long *s_part, double *sum_pi,
                                        unsigned long long *s_hist) {
    const uint32_t elems = n_img * k;
    uint32_t c = threadIdx.x % k;
    const uint32_t cstep = blockDim.x % k;
    unsigned long long s2_all = 0;
    for (uint32_t e = threadIdx.x; e < elems; e += blockDim.x) {
        const int32_t v = tile[e];
        if (kStore) g_counts[e] = v;
        if (v) {
            s2_all += (unsigned long long)(uint32_t)v * (uint32_t)v;
            atomicAdd(&s_class_tot[c], (unsigned long long)(uint32_t)v);
        }
        c += cstep;
        if (c >= k) c -= k;
    }
    unsigned long long r = 0, pairs = 0, rated = 0, pair_images = 0;
    for (uint32_t i = threadIdx.x; i < n_img; i += blockDim.x) {
        const int32_t *row = tile + i * k;
        unsigned long long n = 0, s2 = 0;
        for (uint32_t j = 0; j < k; ++j) {
            const uint32_t v = uint32_t(row[j]);
            n += v;
            if (kSumPi) s2 += (unsigned long long)v * v;
        }
        r += n;
        rated += n >= 1;
        pair_images += n >= 2;
        pairs += n * (n - (n > 0));
        if (kSumPi && n >= 2) {
            if (sum_pi) *sum_pi += double(s2 - n) / double(n * (n - 1));
            if (s_hist) agree_add(s_hist, n > 0xffffffffull ? 0xffffffffu : uint32_t(n), uint32_t(s2));
        }
    }
    part_add(s_part, P_S2, s2_all);
    part_add(s_part, P_R, r);
    part_add(s_part, P_RATED, rated);
    part_add(s_part, P_PAIR_IMAGES, pair_images);
    part_add(s_part, P_PAIRS, pairs);
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < elems; e += blockDim.x) tile[e] = 0;
}

struct RowBlock {
    int4 idx[4];
    uint4 cls, act;
};

__device__ __forceinline__ void load_rows(RowBlock &rb, const int32_t *image_idx, const uint8_t *class_idx,
                                          const uint8_t *active, uint64_t row0, uint64_t rows) {
    if (row0 + kRowsPerThread <= rows) {
        const int4 *pi = reinterpret_cast<const int4 *>(image_idx + row0);
#pragma unroll
        for (int j = 0; j < 4; ++j) rb.idx[j] = __ldg(pi + j);
        rb.cls = __ldg(reinterpret_cast<const uint4 *>(class_idx + row0));
        rb.act = __ldg(reinterpret_cast<const uint4 *>(active + row0));
    } else {                                  // ragged end of the table (or past it): scalar loads
        int32_t ii[16];
        uint32_t cw[4] = {0, 0, 0, 0}, aw[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint64_t r = row0 + j;
            ii[j] = INT32_MIN;                // never inside any image range
            if (r < rows) {
                ii[j] = image_idx[r];
                cw[j >> 2] |= uint32_t(class_idx[r]) << (8 * (j & 3));
                aw[j >> 2] |= uint32_t(active[r]) << (8 * (j & 3));
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) rb.idx[j] = make_int4(ii[4 * j], ii[4 * j + 1], ii[4 * j + 2], ii[4 * j + 3]);
        rb.cls = make_uint4(cw[0], cw[1], cw[2], cw[3]);
        rb.act = make_uint4(aw[0], aw[1], aw[2], aw[3]);
    }
}

struct TallySmem {
    int32_t *tile;                        // tile_images * k
    unsigned long long *class_tot;        // k
    unsigned long long *part;             // 8 (7 used)
    double *dred;                         // blockDim (fleiss_partials_kernel only)
    unsigned long long *hist;             // kAgreeBins
};
__device__ __forceinline__ TallySmem carve_smem(uint8_t *raw, uint32_t tile_images, uint32_t k) {
    TallySmem s;
    s.tile = reinterpret_cast<int32_t *>(raw);
    const size_t tile_bytes = (size_t(tile_images) * k * 4 + 7) & ~size_t(7);
    s.class_tot = reinterpret_cast<unsigned long long *>(raw + tile_bytes);
    s.part = s.class_tot + k;
    s.dred = reinterpret_cast<double *>(s.part + 8);
    s.hist = reinterpret_cast<unsigned long long *>(s.dred + kTallyThreads);
    return s;
}

// ----------------------------------------------------------------------------------------
// tally_slab_kernel — the product kernel for rows ordered by image_idx: a thread per ROW, lanes of a warp
// a whole slab (132 rows) apart, shared-memory atomics into a sliding window of images.
//   * CTA b owns the images that START in its nominal row range (same rule as the kernels above) and
//     streams stages of 4224 rows (32 slabs x 132) through a 3-deep shared-memory ring filled by the
//     bulk-copy (TMA) engine: three 1-D copies per stage (image_idx, class_idx, active) issued by one
//     thread, completion on an mbarrier, so up to two stages per CTA are in flight while one is tallied.
//   * Lane l of every warp reads slab l of the stage; warp w takes quads [33w/8, 33(w+1)/8) of each slab.
//     A quad is 4 consecutive rows: one LDS.128 of image indices plus one LDS.32 each of class and active
//     bytes.  The slab pitch is 33 quads = 1 (mod 32) 16-byte units / 4-byte words, so both loads are
//     bank-conflict free without padding the TMA destination.  Lanes that are 132 rows apart sit in
//     different images (BASELINE config 4 has ~100 rows per image), hence no same-address atomics
//     inside a warp instruction; with larger images the atomics serialise but stay correct.
//   * Counters live in a ring of T = 2^t images (slot = image & (T-1)): images below the first image of
//     the current stage are complete (rows are ordered) and are flushed only when the stage would not
//     fit in the window, so a stage is normally tallied exactly once.  A stage spanning more than T
//     images is re-scanned window by window, jumping over image gaps.
//   * Flush: a warp per image, a lane per class — conflict-free LDS, 4k-byte contiguous stores,
//     class totals in registers, n_i with one REDUX.  Images without rows are zero-filled on the way,
//     so d_counts needs no memset.  Partials and checks exactly as in the kernels above.
// ----------------------------------------------------------------------------------------
constexpr int kSlabQuads = 33;                                    // quads per slab: = 1 (mod 32)
constexpr int kSlabRows = 4 * kSlabQuads;                         // 132
constexpr int kSlabStageRows = 32 * kSlabRows;                    // 4224 (multiple of 16: TMA alignment)
constexpr int kSlabStageBytes = kSlabStageRows * 6;               // 25 344
constexpr int kSlabThreads = 256;
constexpr int kSlabWarps = kSlabThreads / 32;

__host__ __device__ inline size_t slab_smem_bytes(uint32_t stages, uint32_t tile_images, uint32_t k, bool hist) {
    return size_t(stages) * kSlabStageBytes + ((size_t(tile_images) * k * 4 + 15) & ~size_t(15)) + 1024 +
           size_t(k) * 8 + 8 * 8 + 16 + size_t(stages) * 8 + (hist ? size_t(kAgreeBins) * 4 : 0);
}

// One row of a quad, branch-free: in the window and class in range => seen++; also active => one shared
// RED.ADD on the counter of (image & tmask, class).  Eleven instructions; written in PTX so that the atomic
// is unconditional: ptxas turns every predicated shared atomic into BSSY / BRA / BSYNC, and a literal 1 into
// ATOMS.POPC.INC, so the increment is a register holding 1 or 0.  The address is always inside the tile plus
// its 1 KB pad (slot < T, class byte < 256), so rows that do not count simply add 0.  No "memory" clobber on purpose: the counters are only read
// after a __syncthreads(), and the clobber would stop the next quad's loads from being hoisted.
template <int J, bool kInc>
__device__ __forceinline__ void slab_row(int32_t img, uint32_t cw, uint32_t aw, int32_t tb, uint32_t span, uint32_t k,
                                         uint32_t tmask, uint32_t k4, uint32_t tile_s, uint32_t one, uint32_t &seen) {
    if (kInc) {
        // Images far longer than a slab: the lanes of a warp sit in the same image and, when one class dominates,
        // on the same counter.  A literal +1 becomes ATOMS.POPC.INC, which the hardware aggregates per address
        // within the warp instruction; it has to be predicated (three more instructions: BSSY / BRA / BSYNC).
        asm volatile(
            "{\n\t"
            ".reg .pred p, q;\n\t"
            ".reg .b32 d, c, a, t;\n\t"
            "sub.u32 d, %1, %2;\n\t"
            "setp.lt.u32 p, d, %3;\n\t"
            "prmt.b32 c, %4, 0, %5;\n\t"
            "setp.lt.and.u32 p, c, %6, p;\n\t"
            "@p add.u32 %0, %0, 1;\n\t"
            "and.b32 a, %7, %8;\n\t"
            "setp.ne.and.u32 q, a, 0, p;\n\t"
            "and.b32 t, %1, %9;\n\t"
            "mad.lo.u32 t, t, %10, %11;\n\t"
            "mad.lo.u32 t, c, 4, t;\n\t"
            "@q red.shared.add.u32 [t], 1;\n\t"
            "}"
            : "+r"(seen)
            : "r"(img), "r"(tb), "r"(span), "r"(cw), "n"(0x4440 + J), "r"(k), "r"(aw), "n"(0xffu << (8 * J)), "r"(tmask),
              "r"(k4), "r"(tile_s));
        return;
    }
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b32 d, c, a, t;\n\t"
        "sub.u32 d, %1, %2;\n\t"
        "setp.lt.u32 p, d, %3;\n\t"
        "prmt.b32 c, %4, 0, %5;\n\t"
        "setp.lt.and.u32 p, c, %6, p;\n\t"
        "@p add.u32 %0, %0, 1;\n\t"
        "and.b32 a, %7, %8;\n\t"
        "setp.ne.and.u32 q, a, 0, p;\n\t"
        "and.b32 t, %1, %9;\n\t"
        "mad.lo.u32 t, t, %10, %11;\n\t"
        "mad.lo.u32 t, c, 4, t;\n\t"
        "selp.u32 a, %12, 0, q;\n\t"
        "red.shared.add.u32 [t], a;\n\t"
        "}"
        : "+r"(seen)
        : "r"(img), "r"(tb), "r"(span), "r"(cw), "n"(0x4440 + J), "r"(k), "r"(aw), "n"(0xffu << (8 * J)), "r"(tmask),
          "r"(k4), "r"(tile_s), "r"(one));
}

// The slab kernel's form: 32-bit counters in shared memory, ONE fire-and-forget shared atomic per lane for the 32
// images parked in a warp's lanes (a warp-aggregated 64-bit form cost +24 % warp instructions, ncu r2).  A numerator
// is < 2^20 (n < kAgreeBins = 2^10), so a bin cannot overflow before kHistFlushImages images have been folded; the CTA
// commits its bins to global memory (64-bit atomics) and zeroes them before that.
constexpr uint32_t kHistFlushImages = 3968;                  // + one window of <= 128 images stays below 2^12
__device__ __forceinline__ void agree_add_lane(uint32_t *s_hist32, uint32_t n, uint32_t s2) {
    const bool big = n >= kAgreeBins;
    if (n >= 2u) atomicAdd(&s_hist32[big ? 0u : n], big ? 1u : s2 - n);
}
__device__ __forceinline__ void commit_hist32(uint32_t *s_hist32, unsigned long long *g_hist) {
    for (uint32_t i = threadIdx.x; i < kAgreeBins; i += blockDim.x) {
        const uint32_t v = s_hist32[i];
        if (v) {
            atomicAdd(&g_hist[i], (unsigned long long)v);
            s_hist32[i] = 0u;
        }
    }
}

// KC = classes per lane (k <= 32 KC), NS = ring depth, kInc = warp-aggregated increments (long images),
// kHist = also build the agreement histogram (g_hist != NULL), kPeer = all-reduce the result in the epilogue
// (a separate instantiation: with the epilogue compiled in, the plain kernel lost 4 %)
template <int KC, int NS, bool kInc, bool kHist, bool kPeer>
__global__ void __launch_bounds__(kSlabThreads, 2)
tally_slab_kernel(const int32_t *__restrict__ image_idx, const uint8_t *__restrict__ class_idx,
                  const uint8_t *__restrict__ active, uint64_t rows, int32_t image_base, uint32_t n_images,
                  uint32_t k, uint32_t tile_log2, int32_t *__restrict__ counts,
                  unsigned long long *__restrict__ g_partials, unsigned long long *__restrict__ g_hist,
                  PeerReduceDesc *__restrict__ peer) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t T = 1u << tile_log2, tmask = T - 1u;
    uint8_t *ring = smem_raw;
    int32_t *tile = reinterpret_cast<int32_t *>(smem_raw + size_t(NS) * kSlabStageBytes);
    unsigned long long *class_tot = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<uint8_t *>(tile) + ((size_t(T) * k * 4 + 15) & ~size_t(15)) + 1024);   // pad: see slab_row
    unsigned long long *part = class_tot + k;                    // 8 (7 used)
    int32_t *s_beyond = reinterpret_cast<int32_t *>(part + 8);   // 1 (+ padding to 16 bytes)
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_beyond + 4);
    uint32_t *hist = reinterpret_cast<uint32_t *>(bars + NS);   // kAgreeBins 32-bit bins (kHist only)
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;

    const uint32_t G = gridDim.x, b = blockIdx.x;
    const int32_t img_end_all = image_base + int32_t(n_images);
    auto nominal = [&](uint64_t i) -> uint64_t {
        return i >= G ? rows : (((rows / G) * i + (rows % G) * i / G) & ~uint64_t(15));
    };
    const uint64_t nom0 = nominal(b), nom1 = nominal(b + 1);
    auto boundary = [&](uint64_t nom) -> int32_t {
        const int32_t v = __ldg(image_idx + nom);
        return v < image_base ? image_base : (v >= img_end_all - 1 ? img_end_all : v + 1);
    };
    int32_t I0, I1;
    if (rows == 0) {
        I0 = image_base + int32_t(uint64_t(n_images) * b / G);
        I1 = image_base + int32_t(uint64_t(n_images) * (b + 1) / G);
    } else {
        I0 = b == 0 ? image_base : boundary(nom0);
        I1 = b + 1 == G ? img_end_all : boundary(nom1);
        if (I1 < I0) I1 = I0;                                     // unsorted input: own nothing
    }
    const bool owns = I1 > I0;

    for (uint32_t e = tid; e < T * k; e += kSlabThreads) tile[e] = 0;
    for (uint32_t c = tid; c < k + 8; c += kSlabThreads) class_tot[c] = 0;      // class totals + partials
    if (kHist)
        for (uint32_t c = tid; c < kAgreeBins; c += kSlabThreads) hist[c] = 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // Per-lane accumulators of the flushes; reduced across the CTA once, at the end of the kernel.
    unsigned long long tot[KC], s2 = 0, pairs = 0;                // tot[cc]: class lane + 32 cc
    uint32_t rated = 0, pair_images = 0, hist_images = 0;
#pragma unroll
    for (int cc = 0; cc < KC; ++cc) tot[cc] = 0;
    auto fold_image = [&](uint32_t n, uint32_t s2i) {             // n_i and sum_j n_ij^2 of one image (0 is harmless)
        rated += n >= 1u;
        pair_images += n >= 2u;
        pairs += (unsigned long long)n * (n - 1u);                // 0 * 0xffffffff = 0
        if (kHist) agree_add_lane(hist, n, s2i);                  // s2i may have wrapped for n >= 65536: unused from kAgreeBins up
    };

    // Write images [a, e) (complete, all mine) to d_counts, fold them into the partials, leave their
    // slots zeroed.  The first min(e - a, T) come from the ring, the rest have no rows.  Uniform.
    // A warp per image, a lane per class; n_i by REDUX, parked in lane (iteration mod 32) so that what
    // derives from it is computed for 32 images at once.
    auto flush = [&](int32_t a, int32_t e) {
        __syncthreads();
        const uint32_t total = uint32_t(e - a);
        const uint32_t n_ring = total < T ? total : T;
        uint32_t n_mine = 0, s2_mine = 0, it = 0;
        for (uint32_t i = wid; i < n_ring; i += kSlabWarps, ++it) {
            const int32_t img = a + int32_t(i);
            int32_t *src = tile + (uint32_t(img) & tmask) * k;
            int32_t *dst = counts + size_t(img - image_base) * k;
            uint32_t v[KC], n = 0, s2i = 0;
#pragma unroll
            for (int cc = 0; cc < KC; ++cc) {
                const uint32_t c = lane + 32u * cc;
                v[cc] = (cc < KC / 2 || c < k) ? uint32_t(src[c]) : 0u;           // the lower half of the lanes' classes always exists
            }
#pragma unroll
            for (int cc = 0; cc < KC; ++cc) {
                const uint32_t c = lane + 32u * cc;
                if (cc < KC / 2 || c < k) {
                    dst[c] = int32_t(v[cc]);
                    src[c] = 0;
                }
                tot[cc] += v[cc];
                s2 += (unsigned long long)v[cc] * v[cc];
                if (kHist) s2i += v[cc] * v[cc];
                n += v[cc];
            }
            n = __reduce_add_sync(0xffffffffu, n);
            if (kHist) s2i = __reduce_add_sync(0xffffffffu, s2i);
            if ((it & 31u) == lane) { n_mine = n; s2_mine = s2i; }
            if ((it & 31u) == 31u) { fold_image(n_mine, s2_mine); n_mine = 0; s2_mine = 0; }
        }
        fold_image(n_mine, s2_mine);
        if (total > n_ring) {                                     // images without rows
            int32_t *z = counts + size_t(a - image_base + int32_t(n_ring)) * k;
            const size_t ne = size_t(total - n_ring) * k;
            for (size_t i = tid; i < ne; i += kSlabThreads) z[i] = 0;
        }
        __syncthreads();
        if (kHist) {
            hist_images += n_ring;
            if (hist_images >= kHistFlushImages) {                // uniform: before a 32-bit bin can overflow
                commit_hist32(hist, g_hist);
                hist_images = 0;
                __syncthreads();
            }
        }
    };

    uint32_t seen = 0, unsorted = 0;
    int32_t tb = I0;                                              // first image of the window [tb, tb + T)

    if (rows > 0 && nom0 < rows) {
        const uint32_t n_stage = uint32_t((rows - nom0 + kSlabStageRows - 1) / kSlabStageRows);   // upper bound
        auto stage_idx = [&](uint32_t s) { return reinterpret_cast<int32_t *>(ring + size_t(s) * kSlabStageBytes); };
        // fill ring slot s with rows [r0, r0 + 4224); all threads call (uniform)
        auto issue = [&](uint64_t r0, uint32_t s) {
            int32_t *d_idx = stage_idx(s);
            uint8_t *d_cls = reinterpret_cast<uint8_t *>(d_idx + kSlabStageRows), *d_act = d_cls + kSlabStageRows;
            const uint64_t left = rows - r0;
            uint32_t bulk_rows = kSlabStageRows;
            if (left < uint64_t(kSlabStageRows)) {                // ragged end of the table
                bulk_rows = uint32_t(left) & ~15u;
                for (uint32_t i = bulk_rows + tid; i < uint32_t(kSlabStageRows); i += kSlabThreads) {
                    const bool in = i < left;
                    d_idx[i] = in ? image_idx[r0 + i] : INT32_MAX;                // sorts last, owned by nobody
                    d_cls[i] = in ? class_idx[r0 + i] : uint8_t(0);
                    d_act[i] = in ? active[r0 + i] : uint8_t(0);
                }
                fence_proxy_async();
            }
            if (tid == 0) {
                if (bulk_rows) {
                    mbar_arrive_expect_tx(&bars[s], bulk_rows * 6u);
                    bulk_g2s(d_idx, image_idx + r0, bulk_rows * 4u, &bars[s]);
                    bulk_g2s(d_cls, class_idx + r0, bulk_rows, &bars[s]);
                    bulk_g2s(d_act, active + r0, bulk_rows, &bars[s]);
                } else {
                    mbar_arrive(&bars[s]);
                }
            }
        };

        int32_t prev_last = nom0 > 0 ? __ldg(image_idx + nom0 - 1) : INT32_MIN;         // row before the stage
        const uint32_t q0 = (uint32_t(kSlabQuads) * wid) / kSlabWarps, q1 = (uint32_t(kSlabQuads) * (wid + 1)) / kSlabWarps;
        const uint32_t Q0 = lane * kSlabQuads + q0;               // my first quad of a stage
        const uint32_t k4 = 4u * k, tile_s = smem_u32(tile);
        const uint32_t one = k < 1u ? k : 1u;                     // = 1 (k >= 1), but not a literal for ptxas

        uint32_t issued = 0, st = 0;                              // stages issued / consumed
        uint32_t is = 0;                                          // ring slot of the next issue
        uint64_t ir0 = nom0;                                      // first row of the next issue
        for (; issued < uint32_t(NS - 1) && issued < n_stage; ++issued) {
            issue(ir0, is);
            ir0 += kSlabStageRows;
            is = is + 1 == uint32_t(NS) ? 0u : is + 1;
        }
        __syncthreads();                                          // ragged-stage plain stores of the prologue
        uint32_t s = 0, parity = 0;                               // ring slot / mbarrier phase of stage st
        uint64_t r0 = nom0;
        for (; st < n_stage; ++st) {
            if (issued < n_stage) {                               // its slot was released by the sync below
                issue(ir0, is);
                ir0 += kSlabStageRows;
                is = is + 1 == uint32_t(NS) ? 0u : is + 1;
                ++issued;
            }
            mbar_wait(&bars[s], parity);
            const uint32_t valid = rows - r0 < uint64_t(kSlabStageRows) ? uint32_t(rows - r0) : uint32_t(kSlabStageRows);
            const int32_t *s_idx = stage_idx(s);
            const int4 *s_quad = reinterpret_cast<const int4 *>(s_idx) + Q0;
            const uint32_t *s_cls = reinterpret_cast<const uint32_t *>(s_idx + kSlabStageRows) + Q0;
            const uint32_t *s_act = s_cls + kSlabStageRows / 4;
            const int32_t first = s_idx[0], last = s_idx[valid - 1];
            // rows of this stage whose pair (r-1, r) this CTA checks: those below nom1 (a multiple of 16)
            const uint32_t chk_rows = r0 >= nom1 ? 0u : (nom1 - r0 < uint64_t(kSlabStageRows) ? uint32_t(nom1 - r0) : uint32_t(kSlabStageRows));

            // one pass over my quads for the window [tb, tb + span)
            auto scan = [&](uint32_t span, bool check_order, bool want_beyond) {
                const int32_t te = tb + int32_t(span);
                int32_t prev = Q0 == 0 ? prev_last : s_idx[4 * Q0 - 1];
                int32_t beyond = INT32_MAX;
                // my quads [0, n_chk) are order-checked in this pass
                const uint32_t n_chk = !check_order || chk_rows / 4 <= Q0 ? 0u : chk_rows / 4 - Q0;
                auto quad = [&](uint32_t q) {
                    const int4 iq = s_quad[q];
                    const uint32_t cw = s_cls[q], aw = s_act[q];
                    // quads holding an out-of-order pair (only zero / non-zero matters to the caller)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.lt.s32 p, %1, %2;\n\t"
                        "setp.lt.or.s32 p, %3, %1, p;\n\t"
                        "setp.lt.or.s32 p, %4, %3, p;\n\t"
                        "setp.lt.or.s32 p, %5, %4, p;\n\t"
                        "setp.lt.and.u32 p, %6, %7, p;\n\t"
                        "@p add.u32 %0, %0, 1;\n\t}"
                        : "+r"(unsorted)
                        : "r"(iq.x), "r"(prev), "r"(iq.y), "r"(iq.z), "r"(iq.w), "r"(q), "r"(n_chk));
                    prev = iq.w;
                    slab_row<0, kInc>(iq.x, cw, aw, tb, span, k, tmask, k4, tile_s, one, seen);
                    slab_row<1, kInc>(iq.y, cw, aw, tb, span, k, tmask, k4, tile_s, one, seen);
                    slab_row<2, kInc>(iq.z, cw, aw, tb, span, k, tmask, k4, tile_s, one, seen);
                    slab_row<3, kInc>(iq.w, cw, aw, tb, span, k, tmask, k4, tile_s, one, seen);
                    if (want_beyond && iq.w >= te) {              // lower bound of the first image past the window
                        const int32_t cand = iq.x > te ? iq.x : te;
                        beyond = cand < beyond ? cand : beyond;
                    }
                };
#pragma unroll
                for (uint32_t q = 0; q < 4; ++q) quad(q);
                if (q1 - q0 > 4) quad(4);                         // 33 = 8 * 4 + 1: the last warp takes the odd quad
                if (want_beyond) {
                    beyond = __reduce_min_sync(0xffffffffu, beyond);
                    if (lane == 0) atomicMin(s_beyond, beyond);
                }
            };

            if (!owns) {
                scan(0u, true, false);                            // order check only
            } else {
                const int32_t lo = first < tb ? tb : (first > I1 ? I1 : first);         // images below are complete
                const int32_t last_c = last < I1 ? last : I1 - 1;
                bool more = int64_t(last_c) - int64_t(tb) >= int64_t(T);                // stage does not fit the window
                if (more && lo > tb) {
                    flush(tb, lo);
                    tb = lo;
                    more = int64_t(last_c) - int64_t(tb) >= int64_t(T);
                }
                if (!more) {                                      // the normal case: one pass, no window change
                    scan(uint32_t(I1 - tb) < T ? uint32_t(I1 - tb) : T, true, false);
                } else {
                    bool first_pass = true;
                    for (;;) {
                        const uint32_t span = uint32_t(I1 - tb) < T ? uint32_t(I1 - tb) : T;
                        if (more) {
                            if (tid == 0) *s_beyond = INT32_MAX;
                            __syncthreads();
                        }
                        scan(span, first_pass, more);
                        first_pass = false;
                        if (!more) break;
                        __syncthreads();
                        int32_t nb = *s_beyond;                   // first image with rows past the window (lower bound)
                        const int32_t te = tb + int32_t(span);
                        nb = nb < te ? te : (nb > I1 ? I1 : nb);
                        flush(tb, nb);                            // the window is complete; [te, nb) has no rows
                        tb = nb;
                        more = int64_t(last_c) - int64_t(tb) >= int64_t(T);
                    }
                }
            }
            prev_last = last;
            __syncthreads();                                      // everybody is done with ring slot s
            // past the nominal end and the stream has left my images (or the table ended)
            const bool done = r0 + kSlabStageRows >= nom1 && (!owns || last >= I1 || r0 + valid >= rows);
            r0 += kSlabStageRows;
            if (++s == uint32_t(NS)) { s = 0; parity ^= 1u; }
            if (done) { ++st; break; }
        }
        // stages issued but not consumed: their copies must land before the shared memory is released
        for (; st < issued; ++st) {
            mbar_wait(&bars[s], parity);
            if (++s == uint32_t(NS)) { s = 0; parity ^= 1u; }
        }
    }
    if (tb < I1) flush(tb, I1);                                   // the open window and any trailing images without rows

    // ---- commit: lanes -> CTA (shared, 64-bit) -> global (one atomic per value and CTA) ----
#pragma unroll
    for (int cc = 0; cc < KC; ++cc)
        if (lane + 32u * cc < k && tot[cc]) atomicAdd(&class_tot[lane + 32u * cc], tot[cc]);
    unsigned long long r = 0;
#pragma unroll
    for (int cc = 0; cc < KC; ++cc) r += tot[cc];
    part_add(part, P_S2, s2);
    part_add(part, P_R, r);
    part_add(part, P_RATED, rated);
    part_add(part, P_PAIR_IMAGES, pair_images);
    part_add(part, P_PAIRS, pairs);
    part_add(part, P_ROWS_SEEN, seen);
    part_add(part, P_UNSORTED, unsorted);
    __syncthreads();
    commit_partials(part, class_tot, k, g_partials);
    if (kHist) commit_hist32(hist, g_hist);
    // Fused collective (b2_label_tally_reduce): the CTA that finishes last all-reduces partials + histogram (one
    // contiguous vector: g_hist == g_partials + k + 7) over NVLink peer memory, in this same kernel.
    if constexpr (kHist && kPeer) {
        __shared__ uint32_t s_last;
        __threadfence();                                          // my atomics are visible before I take a ticket
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&peer->ticket, 1u) == gridDim.x - 1u ? 1u : 0u;
        __syncthreads();
        if (s_last) {
            __threadfence();
            peer_allreduce_cta(peer, g_partials, k + uint32_t(B2_PARTIALS_EXTRA) + kAgreeBins);
        }
    }
}

// Any row order: one RED.ADD per active row into a zeroed count matrix.
__global__ void __launch_bounds__(256)
tally_scatter_kernel(const int32_t *__restrict__ image_idx, const uint8_t *__restrict__ class_idx,
                     const uint8_t *__restrict__ active, uint64_t rows, int64_t image_base, uint32_t n_images,
                     uint32_t k, int32_t *__restrict__ counts, unsigned long long *__restrict__ g_partials) {
    unsigned long long seen = 0;
    const uint64_t groups = (rows + kRowsPerThread - 1) / kRowsPerThread;
    for (uint64_t g = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; g < groups; g += uint64_t(gridDim.x) * blockDim.x) {
        RowBlock rb;
        load_rows(rb, image_idx, class_idx, active, g * kRowsPerThread, rows);
        const int32_t ii[16] = {rb.idx[0].x, rb.idx[0].y, rb.idx[0].z, rb.idx[0].w, rb.idx[1].x, rb.idx[1].y,
                                rb.idx[1].z, rb.idx[1].w, rb.idx[2].x, rb.idx[2].y, rb.idx[2].z, rb.idx[2].w,
                                rb.idx[3].x, rb.idx[3].y, rb.idx[3].z, rb.idx[3].w};
        const uint32_t cw[4] = {rb.cls.x, rb.cls.y, rb.cls.z, rb.cls.w};
        const uint32_t aw[4] = {rb.act.x, rb.act.y, rb.act.z, rb.act.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int64_t rel = int64_t(ii[j]) - image_base;
            const uint32_t c = (cw[j >> 2] >> (8 * (j & 3))) & 0xffu;
            const uint32_t a = (aw[j >> 2] >> (8 * (j & 3))) & 0xffu;
            if (rel >= 0 && rel < int64_t(n_images) && c < k && ii[j] != INT32_MIN) {
                ++seen;
                if (a) atomicAdd(&counts[size_t(rel) * k + c], 1);
            }
        }
    }
    seen = warp_sum(seen);
    if ((threadIdx.x & 31) == 0 && seen) atomicAdd(&g_partials[k + P_ROWS_SEEN], seen);
}

// Partials from a count matrix.  Grid size is a function of n_images only, block sums are
// combined in block order by the last CTA to finish: sum_pi is reproducible for a given shape.
__global__ void __launch_bounds__(kTallyThreads, 2)
fleiss_partials_kernel(const int32_t *__restrict__ counts, uint32_t n_images, uint32_t k, uint32_t tile_images,
                       unsigned long long *__restrict__ g_partials, double *__restrict__ sum_pi_out,
                       double *__restrict__ block_sums, unsigned int *__restrict__ ticket,
                       unsigned long long *__restrict__ g_hist) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const TallySmem sm = carve_smem(smem_raw, tile_images, k);
    int32_t *tile = sm.tile;
    double *s_dred = sm.dred;

    for (uint32_t c = threadIdx.x; c < k + 8; c += blockDim.x) sm.class_tot[c] = 0;
    if (g_hist)
        for (uint32_t c = threadIdx.x; c < kAgreeBins; c += blockDim.x) sm.hist[c] = 0;
    double my_pi = 0.0;
    const uint32_t n_tiles = (n_images + tile_images - 1) / tile_images;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t i0 = t * tile_images;
        const uint32_t n_img = min(tile_images, n_images - i0);
        const int32_t *src = counts + size_t(i0) * k;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < n_img * k; e += blockDim.x) tile[e] = __ldg(src + e);
        __syncthreads();
        if (sum_pi_out || g_hist)
            flush_tile<false, true>(tile, n_img, k, nullptr, sm.class_tot, sm.part, sum_pi_out ? &my_pi : nullptr,
                                    g_hist ? sm.hist : nullptr);
        else flush_tile<false, false>(tile, n_img, k, nullptr, sm.class_tot, sm.part, nullptr, nullptr);
    }
    __syncthreads();
    commit_partials(sm.part, sm.class_tot, k, g_partials);
    if (g_hist) commit_hist(sm.hist, g_hist);
    if (sum_pi_out) {
        s_dred[threadIdx.x] = my_pi;
        __syncthreads();
        for (int s = blockDim.x >> 1; s > 0; s >>= 1) {                          // fixed-order tree
            if (int(threadIdx.x) < s) s_dred[threadIdx.x] += s_dred[threadIdx.x + s];
            __syncthreads();
        }
        __shared__ bool last;
        if (threadIdx.x == 0) {
            block_sums[blockIdx.x] = s_dred[0];
            __threadfence();
            last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (last && threadIdx.x == 0) {
            __threadfence();
            double s = 0.0;
            for (uint32_t i = 0; i < gridDim.x; ++i) s += reinterpret_cast<volatile double *>(block_sums)[i];
            *sum_pi_out = s;
        }
    }
}

static uint32_t pick_tile_images(uint32_t k) {
    uint32_t t = kTileBudgetBytes / (4u * k);
    if (t > uint32_t(kMaxTileImages)) t = kMaxTileImages;
    if (t >= 32) t &= ~31u;
    return t < 1 ? 1 : t;
}
static size_t tally_smem_bytes(uint32_t tile_images, uint32_t k) {
    size_t tile = (size_t(tile_images) * k * 4 + 7) & ~size_t(7);
    return tile + size_t(k) * 8 + 8 * 8 + size_t(kTallyThreads) * 8 + size_t(kAgreeBins) * 8;
}
constexpr uint32_t kFleissGridMax = 592;

static cudaError_t ensure_smem(const void *fn, size_t bytes) {
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

}  // namespace b2

extern "C" uint64_t b2_fleiss_workspace_bytes(uint32_t n_images) {
    (void)n_images;
    return 16 + 8ull * b2::kFleissGridMax;
}

namespace b2 {
int label_tally_impl(const int32_t *d_image_idx, const uint8_t *d_class_idx, const uint8_t *d_active,
                     uint64_t rows, uint32_t image_base, uint32_t n_images, uint32_t k, uint32_t flags,
                     int32_t *d_counts, int64_t *d_partials, int64_t *d_agree_hist, PeerReduceDesc *peer, void *stream);
}

extern "C" int b2_label_tally(const int32_t *d_image_idx, const uint8_t *d_class_idx, const uint8_t *d_active,
                              uint64_t rows, uint32_t image_base, uint32_t n_images, uint32_t k, uint32_t flags,
                              int32_t *d_counts, int64_t *d_partials, int64_t *d_agree_hist, void *stream) {
    return b2::label_tally_impl(d_image_idx, d_class_idx, d_active, rows, image_base, n_images, k, flags, d_counts, d_partials,
                                d_agree_hist, nullptr, stream);
}

// `peer` != NULL (sorted mode, histogram contiguous behind the partials): the slab kernel all-reduces its own result.
int b2::label_tally_impl(const int32_t *d_image_idx, const uint8_t *d_class_idx, const uint8_t *d_active,
                         uint64_t rows, uint32_t image_base, uint32_t n_images, uint32_t k, uint32_t flags,
                         int32_t *d_counts, int64_t *d_partials, int64_t *d_agree_hist, PeerReduceDesc *peer, void *stream) {
    using namespace b2;
    B2_REQUIRE(peer == nullptr || ((flags & B2_TALLY_SORTED) && d_agree_hist == d_partials + k + B2_PARTIALS_EXTRA),
               "b2_label_tally_reduce: needs B2_TALLY_SORTED and the histogram right behind the partials");
    B2_REQUIRE(d_counts && d_partials, "b2_label_tally: null output pointer");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(d_agree_hist) & 7) == 0, "b2_label_tally: misaligned histogram");
    B2_REQUIRE(k >= 1 && k <= 256, "b2_label_tally: k must be in 1..256 (class_idx is uint8)");
    B2_REQUIRE(n_images >= 1 && uint64_t(n_images) * k < (1ull << 40), "b2_label_tally: n_images out of range");
    B2_REQUIRE(uint64_t(image_base) + n_images <= 0x7fffffffull, "b2_label_tally: image range exceeds int32");
    B2_REQUIRE(rows == 0 || (d_image_idx && d_class_idx && d_active), "b2_label_tally: null row pointer");
    B2_REQUIRE(rows == 0 || ((reinterpret_cast<uintptr_t>(d_image_idx) | reinterpret_cast<uintptr_t>(d_class_idx) |
                 reinterpret_cast<uintptr_t>(d_active)) & 15) == 0, "b2_label_tally: row arrays must be 16-byte aligned");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(d_partials) & 7) == 0 && (reinterpret_cast<uintptr_t>(d_counts) & 3) == 0,
               "b2_label_tally: misaligned output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *partials = reinterpret_cast<unsigned long long *>(d_partials);
    unsigned long long *hist = reinterpret_cast<unsigned long long *>(d_agree_hist);
    B2_CUDA_CHECK(cudaMemsetAsync(partials, 0, (size_t(k) + B2_PARTIALS_EXTRA) * 8, st));
    if (hist) B2_CUDA_CHECK(cudaMemsetAsync(hist, 0, size_t(kAgreeBins) * 8, st));
    const uint32_t tile_images = pick_tile_images(k);
    const size_t smem = tally_smem_bytes(tile_images, k);
    if (flags & B2_TALLY_SORTED) {                           // thread-per-row slab kernel
        uint32_t stages = 3, per_sm = 2;
        if (const char *e = getenv("B2_TALLY_STAGES")) stages = atoi(e) == 2 ? 2u : 3u;
        if (const char *e = getenv("B2_TALLY_CTAS")) per_sm = uint32_t(atoi(e)) < 1 ? 1u : uint32_t(atoi(e));
        const size_t budget = (227u * 1024u) / per_sm - 1024u;               // per CTA, incl. the 1 KB the driver reserves
        uint32_t t = 0;
        while (t < 13 && slab_smem_bytes(stages, 2u << t, k, hist != nullptr) <= budget) ++t;
        if (const char *e = getenv("B2_TALLY_TILE_LOG2")) t = uint32_t(atoi(e));
        const size_t smem_slab = slab_smem_bytes(stages, 1u << t, k, hist != nullptr);
        uint64_t want = rows ? (rows + 2ull * kSlabStageRows - 1) / (2ull * kSlabStageRows)
                             : (uint64_t(n_images) + 1023) / 1024;
        const uint64_t cap_ctas = uint64_t(per_sm) * uint64_t(sm_count());
        if (want > cap_ctas) want = cap_ctas;
        if (want < 1) want = 1;
        auto launch = [&](auto kern) -> int {
            B2_CUDA_CHECK(ensure_smem(reinterpret_cast<const void *>(kern), smem_slab));
            if (getenv("B2_TALLY_DEBUG")) {
                int occ = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSlabThreads, smem_slab);
                fprintf(stderr, "[b2] tally_slab_kernel: %u CTAs, %zu B smem, T=2^%u, hist=%d, occupancy %d CTAs/SM\n",
                        uint32_t(want), smem_slab, t, hist != nullptr, occ);
            }
            kern<<<uint32_t(want), kSlabThreads, smem_slab, st>>>(d_image_idx, d_class_idx, d_active, rows,
                                                                 int32_t(image_base), n_images, k, t, d_counts, partials, hist, peer);
            B2_LAUNCH_CHECK("tally_slab_kernel");
            return B2_OK;
        };
        // Several slabs of rows per image on average: the lanes of a warp share images -> aggregated increments.
        // Measured (tools/tally_long_images.py, 100 M rows, k = 50, 70 % of an image's rows in one class): 300 rows
        // per image 0.125 ms plain / 0.129 aggregated; 1 000: 0.145 / 0.116; 10 000: 0.278 / 0.114.
        bool inc = rows / n_images >= 512;
        if (const char *e = getenv("B2_TALLY_INC")) inc = atoi(e) != 0;
#define B2_SLAB_DISPATCH_H(NS, INC, HIST, PEER)                                       \
    do {                                                                              \
        if (k <= 32) return launch(tally_slab_kernel<1, NS, INC, HIST, PEER>);        \
        if (k <= 64) return launch(tally_slab_kernel<2, NS, INC, HIST, PEER>);        \
        if (k <= 128) return launch(tally_slab_kernel<4, NS, INC, HIST, PEER>);       \
        return launch(tally_slab_kernel<8, NS, INC, HIST, PEER>);                     \
    } while (0)
#define B2_SLAB_DISPATCH(NS, INC)                                                     \
    do {                                                                              \
        if (hist && peer) B2_SLAB_DISPATCH_H(NS, INC, true, true);                    \
        if (hist) B2_SLAB_DISPATCH_H(NS, INC, true, false);                           \
        B2_SLAB_DISPATCH_H(NS, INC, false, false);                                    \
    } while (0)
        if (stages == 2) {
            if (inc) B2_SLAB_DISPATCH(2, true);
            B2_SLAB_DISPATCH(2, false);
        }
        if (inc) B2_SLAB_DISPATCH(3, true);
        B2_SLAB_DISPATCH(3, false);
#undef B2_SLAB_DISPATCH
#undef B2_SLAB_DISPATCH_H
    }
    B2_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, size_t(n_images) * k * 4, st));
    if (rows) {
        const uint64_t groups = (rows + kRowsPerThread - 1) / kRowsPerThread;
        uint64_t grid = (groups + 255) / 256;
        const uint64_t cap = 8ull * uint64_t(sm_count());
        if (grid > cap) grid = cap;
        tally_scatter_kernel<<<unsigned(grid), 256, 0, st>>>(d_image_idx, d_class_idx, d_active, rows,
                                                            int64_t(image_base), n_images, k, d_counts, partials);
        B2_LAUNCH_CHECK("tally_scatter_kernel");
    }
    B2_CUDA_CHECK(ensure_smem(reinterpret_cast<const void *>(fleiss_partials_kernel), smem));
    const uint32_t n_tiles = (n_images + tile_images - 1) / tile_images;
    const uint32_t grid = n_tiles < kFleissGridMax ? n_tiles : kFleissGridMax;
    fleiss_partials_kernel<<<grid, kTallyThreads, smem, st>>>(d_counts, n_images, k, tile_images, partials,
                                                             nullptr, nullptr, nullptr, hist);
    B2_LAUNCH_CHECK("fleiss_partials_kernel");
    return B2_OK;
}

extern "C" int b2_fleiss_partials(const int32_t *d_counts, uint32_t n_images, uint32_t k, int64_t *d_partials,
                                  double *d_sum_pi, int64_t *d_agree_hist, void *d_workspace, uint64_t workspace_bytes,
                                  void *stream) {
    using namespace b2;
    B2_REQUIRE(d_counts && d_partials, "b2_fleiss_partials: null pointer");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(d_agree_hist) & 7) == 0, "b2_fleiss_partials: misaligned histogram");
    B2_REQUIRE(k >= 1 && k <= 256 && n_images >= 1, "b2_fleiss_partials: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *partials = reinterpret_cast<unsigned long long *>(d_partials);
    unsigned long long *hist = reinterpret_cast<unsigned long long *>(d_agree_hist);
    B2_CUDA_CHECK(cudaMemsetAsync(partials, 0, (size_t(k) + B2_PARTIALS_EXTRA) * 8, st));
    if (hist) B2_CUDA_CHECK(cudaMemsetAsync(hist, 0, size_t(kAgreeBins) * 8, st));
    const uint32_t tile_images = pick_tile_images(k);
    const size_t smem = tally_smem_bytes(tile_images, k);
    B2_CUDA_CHECK(ensure_smem(reinterpret_cast<const void *>(fleiss_partials_kernel), smem));
    const uint32_t n_tiles = (n_images + tile_images - 1) / tile_images;
    const uint32_t grid = n_tiles < kFleissGridMax ? n_tiles : kFleissGridMax;
    double *block_sums = nullptr;
    unsigned int *ticket = nullptr;
    if (d_sum_pi) {
        B2_REQUIRE(d_workspace && (reinterpret_cast<uintptr_t>(d_workspace) & 7) == 0,
                   "b2_fleiss_partials: d_sum_pi needs an 8-byte aligned workspace");
        if (workspace_bytes < b2_fleiss_workspace_bytes(n_images))
            return fail(B2_ERR_WORKSPACE, "b2_fleiss_partials: workspace %llu < required %llu bytes",
                        (unsigned long long)workspace_bytes, (unsigned long long)b2_fleiss_workspace_bytes(n_images));
        ticket = static_cast<unsigned int *>(d_workspace);
        block_sums = reinterpret_cast<double *>(static_cast<uint8_t *>(d_workspace) + 16);
        B2_CUDA_CHECK(cudaMemsetAsync(ticket, 0, 16, st));
    }
    fleiss_partials_kernel<<<grid, kTallyThreads, smem, st>>>(d_counts, n_images, k, tile_images, partials,
                                                             d_sum_pi, block_sums, ticket, hist);
    B2_LAUNCH_CHECK("fleiss_partials_kernel");
    return B2_OK;
}

// Host-side verdict on a tally's partials (copied to the host by the caller).
extern "C" int b2_label_tally_status(const int64_t *h_partials, uint32_t k, uint64_t rows) {
    using namespace b2;
    B2_REQUIRE(h_partials != nullptr, "b2_label_tally_status: null pointer");
    if (h_partials[k + P_UNSORTED] != 0)
        return fail(B2_ERR_NOT_SORTED, "label tally: %lld adjacent row pairs are out of image order (B2_TALLY_SORTED)",
                    (long long)h_partials[k + P_UNSORTED]);
    if (uint64_t(h_partials[k + P_ROWS_SEEN]) != rows)
        return fail(B2_ERR_BAD_ARG, "label tally: %llu of %llu rows tallied; the rest have image_idx or class_idx out of range",
                    (unsigned long long)h_partials[k + P_ROWS_SEEN], (unsigned long long)rows);
    return B2_OK;
}

// Bulk COUNT(DISTINCT id_img) WHERE id_con = ? AND ativo (app/api/routes/classificacoes.py:224-230)
// for every annotator at once.  Rows sorted by (annotator_idx, image_idx): a run of equal (annotator, image)
// counts once iff it holds an active row.
namespace b2 {
__global__ void __launch_bounds__(256)
distinct_images_kernel(const int32_t *__restrict__ annotator_idx, const int32_t *__restrict__ image_idx,
                       const uint8_t *__restrict__ active, uint64_t rows, uint32_t n_annotators,
                       uint32_t *__restrict__ distinct) {
    // A warp takes 32 consecutive rows per step.  The row that STARTS a run (its key differs from the row before:
    // one shuffle, lane 0 loads its predecessor) scans the run forward until it meets an active row, so every row
    // is visited by at most one scanning thread — linear in the table whatever the run lengths (a backward scan
    // from every row was quadratic on long runs of inactive rows).  A warp sees one or two annotators: lanes with
    // the same annotator are grouped with MATCH.ANY and the group's first lane adds the group's count in one
    // atomic — no two lanes of a warp instruction on the same address.
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
    const uint64_t n_warps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t base = warp * 32; base < rows; base += n_warps * 32) {
        const uint64_t r = base + lane;
        int32_t a = -1, img = -1;
        if (r < rows) { a = annotator_idx[r]; img = image_idx[r]; }
        int32_t pa = __shfl_up_sync(0xffffffffu, a, 1), pi = __shfl_up_sync(0xffffffffu, img, 1);
        if (lane == 0) {
            pa = base > 0 ? annotator_idx[base - 1] : -1;
            pi = base > 0 ? image_idx[base - 1] : -1;
        }
        bool counted = false;
        if (r < rows && a >= 0 && uint32_t(a) < n_annotators && (r == 0 || pa != a || pi != img)) {
            for (uint64_t q = r; q < rows; ++q) {
                if (q != r && (annotator_idx[q] != a || image_idx[q] != img)) break;
                if (active[q]) { counted = true; break; }
            }
        }
        const uint32_t flags = __ballot_sync(0xffffffffu, counted);
        const uint32_t same = __match_any_sync(0xffffffffu, a);
        const uint32_t mine = flags & same;
        if (counted && (mine & ((1u << lane) - 1u)) == 0) atomicAdd(&distinct[a], uint32_t(__popc(mine)));
    }
}
}  // namespace b2

extern "C" int b2_distinct_images_per_annotator(const int32_t *d_annotator_idx, const int32_t *d_image_idx,
                                                const uint8_t *d_active, uint64_t rows, uint32_t n_annotators,
                                                uint32_t *d_distinct, void *stream) {
    using namespace b2;
    B2_REQUIRE(d_distinct != nullptr && n_annotators >= 1, "b2_distinct_images_per_annotator: bad output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2_CUDA_CHECK(cudaMemsetAsync(d_distinct, 0, size_t(n_annotators) * 4, st));
    if (rows == 0) return B2_OK;
    B2_REQUIRE(d_annotator_idx && d_image_idx && d_active, "b2_distinct_images_per_annotator: null pointer");
    uint64_t grid = (rows + 255) / 256;
    const uint64_t cap = 16ull * uint64_t(sm_count());
    if (grid > cap) grid = cap;
    distinct_images_kernel<<<unsigned(grid), 256, 0, st>>>(d_annotator_idx, d_image_idx, d_active, rows,
                                                          n_annotators, d_distinct);
    B2_LAUNCH_CHECK("distinct_images_kernel");
    return B2_OK;
}
