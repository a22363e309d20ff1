// Per-image label tally + integer Fleiss partials for sm_100a.
//
// The reference only groups ONE user's active rows by image
// (app/crud/classificacao_crud.py:318-322) and counts a user's distinct images
// (app/api/routes/classificacoes.py:224-230); BASELINE.json configs 1,4,5 additionally require
// the cross-annotator tally and Fleiss' kappa.  Rows are the dictionary-encoded columns of
// table `classificacoes` (app/db/models.py:224-241): image_idx int32 (id_img), class_idx uint8
// (id_opc), active uint8 (ativo); only rows with active != 0 count (classificacao_crud.py:314).
//
// tally_sorted_kernel — rows ordered by image_idx (an index scan on id_img).  HBM-bound:
//   6 B read per row + 4*k B written per image, every byte touched once.
//   * Persistent CTAs.  CTA b nominally owns rows [b*R/G, (b+1)*R/G); so that every image is
//     written by exactly one CTA with plain stores, the boundary is moved to the next image
//     change: CTA b owns images [I_b, I_{b+1}), I_b = image_idx[b*R/G] + 1.  It starts
//     streaming at its nominal row and simply ignores rows of foreign images, so no search
//     and no pre-pass is needed.
//   * A thread loads 16 consecutive rows with six 128-bit loads (4 x int4 image_idx, 1 x uint4
//     class, 1 x uint4 active); a warp covers 512 consecutive rows, a CTA 8192.  The next
//     block is prefetched into registers while the current one is tallied.
//   * The CTA keeps a TILE x k int32 count tile in shared memory and tallies with shared-memory
//     atomics.  Lanes are 16 rows apart, which spreads a warp over several images and keeps
//     same-address conflicts low even when one class dominates an image.
//   * When the stream leaves the tile, the tile is written to d_counts with coalesced stores
//     (d_counts need not be zeroed) and reduced on the fly into the integer partials:
//     class totals (64-bit shared accumulators), S2 = sum n_ij^2, R = sum n_i, images with
//     n_i >= 1 / >= 2, sum n_i (n_i - 1).  Partials leave the CTA as 64-bit integer atomics, so
//     they are exact and independent of scheduling and of the GPU count.
//   * Every adjacent row pair is checked for order and every row for range; violations are
//     reported in the partials (unsorted pairs; rows_seen != rows) — the host wrapper turns them
//     into B2_ERR_NOT_SORTED / B2_ERR_BAD_ARG without the library having to synchronise.
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
constexpr int kBlockRows = kTallyThreads * kRowsPerThread;      // 8192
constexpr int kTileBudgetBytes = 100 * 1024;                    // two CTAs per SM
constexpr int kMaxTileImages = 512;

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
                                        unsigned long long *s_class_tot, unsigned long long *s_part, double *sum_pi) {
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
        if (kSumPi && n >= 2) *sum_pi += double(s2 - n) / double(n * (n - 1));
    }
    part_add(s_part, P_S2, s2_all);
    part_add(s_part, P_R, r);
    part_add(s_part, P_RATED, rated);
    part_add(s_part, P_PAIR_IMAGES, pair_images);
    part_add(s_part, P_PAIRS, pairs);
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < elems; e += blockDim.x) tile[e] = 0;
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
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
};
__device__ __forceinline__ TallySmem carve_smem(uint8_t *raw, uint32_t tile_images, uint32_t k) {
    TallySmem s;
    s.tile = reinterpret_cast<int32_t *>(raw);
    const size_t tile_bytes = (size_t(tile_images) * k * 4 + 7) & ~size_t(7);
    s.class_tot = reinterpret_cast<unsigned long long *>(raw + tile_bytes);
    s.part = s.class_tot + k;
    s.dred = reinterpret_cast<double *>(s.part + 8);
    return s;
}

__global__ void __launch_bounds__(kTallyThreads, 2)
tally_sorted_kernel(const int32_t *__restrict__ image_idx, const uint8_t *__restrict__ class_idx,
                    const uint8_t *__restrict__ active, uint64_t rows, int32_t image_base, uint32_t n_images,
                    uint32_t k, uint32_t tile_images, int32_t *__restrict__ counts,
                    unsigned long long *__restrict__ g_partials) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const TallySmem sm = carve_smem(smem_raw, tile_images, k);
    int32_t *tile = sm.tile;

    const uint32_t G = gridDim.x, b = blockIdx.x;
    const int32_t img_end_all = image_base + int32_t(n_images);                  // host checked: fits int32
    // nominal row range, aligned to the 16-row vector granule
    const uint64_t nom0 = ((rows / G) * b + (rows % G) * b / G) & ~uint64_t(kRowsPerThread - 1);
    const uint64_t nom1 = b + 1 == G ? rows : (((rows / G) * (b + 1) + (rows % G) * (b + 1) / G) & ~uint64_t(kRowsPerThread - 1));
    // owned image range [I0, I1): the image under the nominal boundary row belongs to the CTA before
    auto boundary = [&](uint64_t nom) -> int32_t {
        const int32_t v = image_idx[nom];
        return v < image_base ? image_base : (v >= img_end_all - 1 ? img_end_all : v + 1);
    };
    int32_t I0, I1;
    if (rows == 0) {                                                             // nothing to stream: split the zero fill
        I0 = image_base + int32_t(uint64_t(n_images) * b / G);
        I1 = image_base + int32_t(uint64_t(n_images) * (b + 1) / G);
    } else {
        I0 = b == 0 ? image_base : boundary(nom0);
        I1 = b + 1 == G ? img_end_all : boundary(nom1);
        if (I1 < I0) I1 = I0;                                                    // unsorted input: own nothing
    }

    for (uint32_t e = threadIdx.x; e < tile_images * k; e += blockDim.x) tile[e] = 0;
    for (uint32_t c = threadIdx.x; c < k + 8; c += blockDim.x) sm.class_tot[c] = 0;   // class totals + partials
    __syncthreads();

    int32_t base = I0;                                                           // first image of the tile
    int32_t tile_end = I1 - base > int32_t(tile_images) ? base + int32_t(tile_images) : I1;
    uint32_t rows_seen = 0, unsorted = 0;

    auto flush = [&]() {                                                         // uniform
        __syncthreads();
        flush_tile<true, false>(tile, uint32_t(tile_end - base), k,
                                counts + size_t(base - image_base) * k, sm.class_tot, sm.part, nullptr);
        base = tile_end;
        tile_end = I1 - base > int32_t(tile_images) ? base + int32_t(tile_images) : I1;
        __syncthreads();
    };

    RowBlock cur;
    uint64_t blk = nom0;
    while (blk < rows) {
        const uint64_t row0 = blk + uint64_t(threadIdx.x) * kRowsPerThread;
        const uint64_t nblk = blk + kBlockRows;
        load_rows(cur, image_idx, class_idx, active, row0, rows);
        {   // pull the next block into L2 while this one is tallied (no registers held)
            const uint64_t nrow0 = row0 + kBlockRows;
            if (nrow0 < rows) {
                prefetch_l2(image_idx + nrow0);
                if ((threadIdx.x & 7) == 0) {                                     // one 128-byte line per 8 threads
                    prefetch_l2(class_idx + nrow0);
                    prefetch_l2(active + nrow0);
                }
            }
        }

        const int32_t ii[16] = {cur.idx[0].x, cur.idx[0].y, cur.idx[0].z, cur.idx[0].w, cur.idx[1].x, cur.idx[1].y,
                                cur.idx[1].z, cur.idx[1].w, cur.idx[2].x, cur.idx[2].y, cur.idx[2].z, cur.idx[2].w,
                                cur.idx[3].x, cur.idx[3].y, cur.idx[3].z, cur.idx[3].w};
        const uint32_t cw[4] = {cur.cls.x, cur.cls.y, cur.cls.z, cur.cls.w};
        const uint32_t aw[4] = {cur.act.x, cur.act.y, cur.act.z, cur.act.w};

        // order check of every adjacent pair this CTA is responsible for: pairs (r-1, r) with
        // nom0 <= r < nom1 (nom1 is a multiple of 16 or the table end, so a thread is all in or out)
        if (row0 < nom1) {
            int32_t prev = row0 > 0 ? __ldg(image_idx + row0 - 1) : INT32_MIN;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                unsorted += (ii[j] < prev) & (ii[j] != INT32_MIN);               // INT32_MIN = past the table
                prev = ii[j];
            }
        }

        bool any_mine = false;
#pragma unroll
        for (int j = 0; j < 16; ++j) any_mine |= (ii[j] < I1) & (ii[j] != INT32_MIN);
        for (;;) {
            // Tile-relative counter index of each of my rows (kNoKey: not in this tile / bad class /
            // inactive).  A thread's 16 rows are consecutive, so most share the image and its dominant
            // class: rows equal to the first key are merged into ONE shared atomic, which removes most
            // same-address conflicts between the lanes of a warp.
            constexpr uint32_t kNoKey = 0xffffffffu;
            bool beyond_tile = false;
            auto key_of = [&](int j) -> uint32_t {
                const int32_t img = ii[j];
                const uint32_t c = (cw[j >> 2] >> (8 * (j & 3))) & 0xffu;
                const uint32_t a = (aw[j >> 2] >> (8 * (j & 3))) & 0xffu;
                const bool in_tile = (img >= base) & (img < tile_end) & (c < k);
                return (in_tile && a) ? uint32_t(img - base) * k + c : kNoKey;
            };
            const uint32_t k0 = key_of(0);
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int32_t img = ii[j];
                const uint32_t c = (cw[j >> 2] >> (8 * (j & 3))) & 0xffu;
                rows_seen += (img >= base) & (img < tile_end) & (c < k);
                beyond_tile |= (img >= tile_end) & (img < I1);
                m += key_of(j) == k0;
            }
            if (k0 != kNoKey) atomicAdd(&tile[k0], int32_t(m));
#pragma unroll
            for (int j = 1; j < 16; ++j) {
                const uint32_t kj = key_of(j);
                if (kj != k0 && kj != kNoKey) atomicAdd(&tile[kj], 1);
            }
            if (!__syncthreads_or(beyond_tile)) break;
            flush();                                                             // tile complete: write it, open the next
        }
        // the stream has left this CTA's images once a whole block past the nominal end is foreign
        const bool stream_live = __syncthreads_or(any_mine) || nblk < nom1;
        if (!stream_live) break;
        blk = nblk;
    }
    // remaining tiles (the open one and any image range without rows)
    while (base < I1) flush();
    part_add(sm.part, P_ROWS_SEEN, rows_seen);
    part_add(sm.part, P_UNSORTED, unsorted);
    __syncthreads();
    commit_partials(sm.part, sm.class_tot, k, g_partials);
}

// ----------------------------------------------------------------------------------------
// tally_warp_kernel — the product kernel for rows ordered by image_idx.  No atomics on the row
// path at all (measured: shared-memory atomics cost ~2 cycles per lane and bound the tile
// kernel above at 22 % of HBM peak):
//   * A WARP streams a contiguous row range and owns the images that START in it (same
//     ownership rule as above, per warp instead of per CTA).
//   * Rows reach the warp through its own ring of two shared-memory stages of 256 rows filled by
//     the bulk-copy (TMA) engine (cp.async.bulk + mbarrier, three 1-D copies per stage issued by
//     one lane): the next stage is in flight while this one is tallied, without holding registers.
//   * A step is 32 consecutive rows, one per lane.  MATCH.ANY on the class byte gives every lane
//     the set of lanes with its class; ANDed with the ballot of "active, in range, current
//     image" it is the group whose size is the increment.  The lowest lane of each group adds
//     it to the warp's PRIVATE k-entry counter array in shared memory with a plain
//     read-modify-write: groups have distinct classes, so there are no conflicts and no atomics.
//   * When the image changes the lanes copy the k counters to d_counts — one coalesced k*4-byte
//     store per image, no tile, no flush — zero them and fold them into per-warp class totals
//     and the sum of squares.  n_i is the population count of the row mask, uniform across the
//     warp, so R, rated images and pairs need no reduction.
//   * Partials are combined per CTA in shared memory, then one 64-bit atomic per value and CTA.
// ----------------------------------------------------------------------------------------
constexpr int kWarpKernelThreads = 256;
constexpr int kWarpsPerCta = kWarpKernelThreads / 32;
constexpr int kStageRows = 256;
constexpr int kGroupRows = kStageRows;                        // nominal starts are aligned to one stage
constexpr int kStageBytes = kStageRows * 6;                   // int32 image + uint8 class + uint8 active
constexpr int kStages = 2;
constexpr int kWarpSmemBytes = kStages * kStageBytes + 256 * 4 + 256 * 8;            // ring + counters + class totals
constexpr int kWarpKernelSmem = kWarpsPerCta * kWarpSmemBytes;                       // 48 KB per CTA, four CTAs per SM

__global__ void __launch_bounds__(kWarpKernelThreads, 4)
tally_warp_kernel(const int32_t *__restrict__ image_idx, const uint8_t *__restrict__ class_idx,
                  const uint8_t *__restrict__ active, uint64_t rows, int32_t image_base, uint32_t n_images,
                  uint32_t k, uint64_t n_workers, int32_t *__restrict__ counts,
                  unsigned long long *__restrict__ g_partials) {
    __shared__ unsigned long long s_tot[256 + 8];
    extern __shared__ __align__(128) uint8_t warp_smem[];
    __shared__ __align__(8) uint64_t ring_bars[kWarpsPerCta * kStages];
    for (uint32_t i = threadIdx.x; i < 256 + 8; i += blockDim.x) s_tot[i] = 0;

    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t *my = warp_smem + size_t(wid) * kWarpSmemBytes;
    uint32_t *cnt_s = reinterpret_cast<uint32_t *>(my + kStages * kStageBytes);             // k counters of image `cur`
    unsigned long long *tot_s = reinterpret_cast<unsigned long long *>(my + kStages * kStageBytes + 256 * 4);
    for (uint32_t c = lane; c < 256; c += 32) { cnt_s[c] = 0; tot_s[c] = 0; }
    __syncthreads();

    // n_workers warps share the rows; the host keeps it <= rows/512 so that nominal starts are distinct
    const uint64_t W = n_workers;
    const uint64_t w = uint64_t(blockIdx.x) * kWarpsPerCta + wid;
    const bool worker = w < W;
    const int32_t img_end_all = image_base + int32_t(n_images);
    auto nominal = [&](uint64_t i) -> uint64_t {
        return i >= W ? rows : (((rows / W) * i + (rows % W) * i / W) & ~uint64_t(kGroupRows - 1));
    };
    const uint64_t nom0 = nominal(w), nom1 = nominal(w + 1);
    auto boundary = [&](uint64_t nom) -> int32_t {
        const int32_t v = __ldg(image_idx + nom);
        return v < image_base ? image_base : (v >= img_end_all - 1 ? img_end_all : v + 1);
    };
    int32_t I0, I1;
    if (!worker) {
        I0 = I1 = img_end_all;                                // spare warp of the last CTA: owns nothing
    } else if (rows == 0) {
        I0 = image_base + int32_t(uint64_t(n_images) * w / W);
        I1 = image_base + int32_t(uint64_t(n_images) * (w + 1) / W);
    } else {
        I0 = w == 0 ? image_base : boundary(nom0);
        I1 = w + 1 == W ? img_end_all : boundary(nom1);
        if (I1 < I0) I1 = I0;                                 // unsorted input
    }
    const int32_t span = I1 - I0;                             // "mine" <=> unsigned(img - I0) < span
    const uint32_t lanes_below = (1u << lane) - 1u;

    unsigned long long s2 = 0, sum_r = 0, pairs = 0;
    uint32_t rated = 0, pair_images = 0, seen = 0, unsorted = 0, n_cur = 0;
    int32_t cur = I0;

    // store the finished image, fold it into the partials, zero-fill images without rows up to `next`
    auto finish_image = [&](int32_t next) {
        __syncwarp();
        if (cur < I1) {
            int32_t *dst = counts + size_t(cur - image_base) * k;
            for (uint32_t c = lane; c < k; c += 32) {
                const uint32_t v = cnt_s[c];
                dst[c] = int32_t(v);
                if (v) {
                    cnt_s[c] = 0;
                    tot_s[c] += v;
                    s2 += (unsigned long long)v * v;
                }
            }
            sum_r += n_cur;
            rated += n_cur >= 1;
            pair_images += n_cur >= 2;
            pairs += (unsigned long long)n_cur * (n_cur - (n_cur > 0));
            n_cur = 0;
            const int32_t stop = next < I1 ? next : I1;
            for (int32_t img = cur + 1; img < stop; ++img) {
                int32_t *z = counts + size_t(img - image_base) * k;
                for (uint32_t c = lane; c < k; c += 32) z[c] = 0;
            }
        }
        cur = next;
        __syncwarp();
    };

    // add the rows of `vm` (a ballot: rows of image `cur` that count) to the warp's counters
    auto accumulate = [&](uint32_t same_class, uint32_t vm, uint32_t c) {
        const uint32_t grp = same_class & vm;                 // rows with my class that count
        if (((vm >> lane) & 1u) && (grp & lanes_below) == 0) cnt_s[c] += __popc(grp);   // lowest lane of the group
        n_cur += __popc(vm);
    };

    // one step of 32 consecutive rows; returns false once the stream has left this warp's images
    auto step = [&](int32_t img, uint32_t c, uint32_t act, int32_t prev_img, bool check_order) -> bool {
        if (check_order) {
            // lane l > 0 compares with lane l-1 of this step, lane 0 with lane 31 of the previous one
            const int32_t z = lane == 31 ? prev_img : img;
            const int32_t before = __shfl_sync(0xffffffffu, z, (lane + 31) & 31);
            unsorted += (img < before) & (img != INT32_MAX);
        }
        if (span <= 0) return false;                          // this warp owns no image: order check only
        const uint32_t same_class = __match_any_sync(0xffffffffu, c);
        const bool good = (c < k) & (act != 0);
        if (__all_sync(0xffffffffu, img == cur)) {            // common case: one image, the current one
            seen += c < k;
            accumulate(same_class, __ballot_sync(0xffffffffu, good), c);
            __syncwarp();
            return true;
        }
        const bool mine = uint32_t(img - I0) < uint32_t(span);
        seen += mine & (c < k);
        uint32_t rem = __ballot_sync(0xffffffffu, mine);
        while (rem) {
            const int first = __ffs(rem) - 1;
            const int32_t nxt = __shfl_sync(0xffffffffu, img, first);
            if (nxt != cur) finish_image(nxt);
            const uint32_t same = __ballot_sync(0xffffffffu, img == nxt) & rem;
            accumulate(same_class, __ballot_sync(0xffffffffu, mine & good & (img == nxt)), c);
            __syncwarp();
            rem &= ~same;
        }
        return __any_sync(0xffffffffu, img < I1) != 0;
    };

    if (worker && nom0 < rows && span >= 0) {
        int32_t *s_idx = reinterpret_cast<int32_t *>(my);
        uint64_t *bars = ring_bars + wid * kStages;
        auto stage_idx = [&](int b) { return s_idx + b * (kStageBytes / 4); };
        auto stage_cls = [&](int b) { return reinterpret_cast<uint8_t *>(stage_idx(b)) + kStageRows * 4; };
        auto stage_act = [&](int b) { return stage_cls(b) + kStageRows; };
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < kStages; ++b) mbar_init(&bars[b], 1);
            fence_mbar_init();
        }
        __syncwarp();
        const uint64_t n_stage = (rows - nom0 + kStageRows - 1) / kStageRows;    // upper bound; the stream usually ends earlier
        auto issue = [&](uint64_t st) {                       // whole warp calls; lane 0 issues
            const uint64_t r0 = nom0 + st * kStageRows;
            const int b = int(st % kStages);
            if (r0 + kStageRows <= rows) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&bars[b], kStageRows * 6);
                    bulk_g2s(stage_idx(b), image_idx + r0, kStageRows * 4, &bars[b]);
                    bulk_g2s(stage_cls(b), class_idx + r0, kStageRows, &bars[b]);
                    bulk_g2s(stage_act(b), active + r0, kStageRows, &bars[b]);
                }
            } else {                                          // ragged end of the table: plain loads, sentinel fill
                for (int i = lane; i < kStageRows; i += 32) {
                    const bool in = r0 + i < rows;
                    stage_idx(b)[i] = in ? image_idx[r0 + i] : INT32_MAX;        // sorts last, owned by nobody
                    stage_cls(b)[i] = in ? class_idx[r0 + i] : uint8_t(0);
                    stage_act(b)[i] = in ? active[r0 + i] : uint8_t(0);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[b]);
            }
        };
        int32_t prev_img = nom0 > 0 ? __ldg(image_idx + nom0 - 1) : INT32_MIN;   // row before the first
        const bool owns = span > 0;
        uint64_t issued = 0, st = 0;
        for (; issued < kStages - 1 && issued < n_stage; ++issued) issue(issued);
        for (; st < n_stage; ++st) {
            if (issued < n_stage) { issue(issued); ++issued; }
            const int b = int(st % kStages);
            mbar_wait(&bars[b], uint32_t((st / kStages) & 1));
            const int32_t *si = stage_idx(b) + lane;
            const uint8_t *sc = stage_cls(b) + lane, *sa = stage_act(b) + lane;
            const uint64_t r0 = nom0 + st * kStageRows;
            const bool check_order = r0 < nom1;               // nom1 is stage aligned (or the table end)
            bool any_mine = false;
#pragma unroll 4
            for (int t = 0; t < kStageRows / 32; ++t) {
                const int32_t img = si[32 * t];
                any_mine |= step(img, sc[32 * t], sa[32 * t], prev_img, check_order);
                prev_img = img;
            }
            __syncwarp();                                     // every lane is done with this stage before it is refilled
            // past the nominal end with nothing below I1 in this stage: the stream has left my images
            if (r0 + kStageRows >= nom1 && !(any_mine && owns)) { ++st; break; }
        }
        // stages issued but not consumed: their copies must land before this CTA's shared memory is released
        for (; st < issued; ++st) mbar_wait(&bars[st % kStages], uint32_t((st / kStages) & 1));
    }
    finish_image(I1);

    // ---- commit: warp -> CTA (shared, 64-bit) -> global (one atomic per value and CTA) ----
    for (uint32_t c = lane; c < k; c += 32)
        if (tot_s[c]) atomicAdd(&s_tot[c], tot_s[c]);
    {
        const unsigned long long v_s2 = warp_sum(s2), v_seen = warp_sum(seen), v_uns = warp_sum(unsorted);
        if (lane == 0) {
            if (v_s2) atomicAdd(&s_tot[256 + P_S2], v_s2);
            if (sum_r) atomicAdd(&s_tot[256 + P_R], sum_r);
            if (rated) atomicAdd(&s_tot[256 + P_RATED], (unsigned long long)rated);
            if (pair_images) atomicAdd(&s_tot[256 + P_PAIR_IMAGES], (unsigned long long)pair_images);
            if (pairs) atomicAdd(&s_tot[256 + P_PAIRS], pairs);
            if (v_seen) atomicAdd(&s_tot[256 + P_ROWS_SEEN], v_seen);
            if (v_uns) atomicAdd(&s_tot[256 + P_UNSORTED], v_uns);
        }
    }
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < k; c += blockDim.x)
        if (s_tot[c]) atomicAdd(&g_partials[c], s_tot[c]);
    if (threadIdx.x < 7 && s_tot[256 + threadIdx.x]) atomicAdd(&g_partials[k + threadIdx.x], s_tot[256 + threadIdx.x]);
}

// ----------------------------------------------------------------------------------------
// tally_image_kernel — thread-per-image formulation for rows ordered by image_idx.
// The warp kernel above spends ~108 warp-instructions per 32 rows because the whole warp cooperates on
// each row group.  Here every THREAD tallies a whole image by itself, so the SIMT width is used on
// independent rows: ~10 thread-instructions per row.
//   * CTA b owns the images that START in its nominal row range (same rule as above) and walks them in
//     tiles of `tile_images` images whose counters (int32, row pitch k|1 so a column walk is
//     conflict-free) live in shared memory.
//   * Phase A (coalesced): blocks of 4096 rows, 16 consecutive rows per thread (six 128-bit loads).
//     Each row becomes ONE byte in a shared row buffer — its class, or 0xFF when inactive / foreign /
//     out of range — and every image change records where the image's rows start and end in that
//     buffer.  Order and range checks happen here.  A phase ends when the buffer is full or the stream
//     leaves the tile.
//   * Phase B: thread i walks the rows [start_i, end_i) of image tile_base+i in the row buffer and
//     increments its own counters with plain read-modify-writes: no atomics, no warp collectives.
//   * A finished tile is written to d_counts with coalesced stores and folded into the integer
//     partials exactly like the other kernels.
// ----------------------------------------------------------------------------------------
constexpr int kImgThreads = 256;
constexpr int kImgBlockRows = kImgThreads * kRowsPerThread;       // 4096
constexpr uint32_t kNoRow = 0xffffffffu;

struct ImgSmem {
    int32_t *cnt;                 // tile_images * pitch
    uint8_t *rowbuf;              // cap_rows
    uint32_t *seg_start, *seg_end;// tile_images each
    unsigned long long *class_tot;// k
    unsigned long long *part;     // 8
    uint32_t *first_beyond;       // 1 (row offset in the buffer of the first row past the tile)
    uint32_t *tile_tot;           // k (class totals of the open tile, 32-bit)
};
__host__ __device__ inline size_t img_smem_layout(uint32_t tile_images, uint32_t pitch, uint32_t cap_rows, uint32_t k,
                                                  size_t *o_rowbuf, size_t *o_seg, size_t *o_tot) {
    size_t off = (size_t(tile_images) * pitch * 4 + 15) & ~size_t(15);
    *o_rowbuf = off; off += (size_t(cap_rows) + 15) & ~size_t(15);
    *o_seg = off; off += size_t(tile_images) * 8;
    off = (off + 7) & ~size_t(7);
    *o_tot = off; off += size_t(k) * 8 + 8 * 8 + 16 + size_t(k) * 4;
    return off;
}

__global__ void __launch_bounds__(kImgThreads, 3)
tally_image_kernel(const int32_t *__restrict__ image_idx, const uint8_t *__restrict__ class_idx,
                   const uint8_t *__restrict__ active, uint64_t rows, int32_t image_base, uint32_t n_images,
                   uint32_t k, uint32_t tile_images, uint32_t cap_rows, int32_t *__restrict__ counts,
                   unsigned long long *__restrict__ g_partials) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t pitch = k | 1u;
    size_t o_rowbuf, o_seg, o_tot;
    img_smem_layout(tile_images, pitch, cap_rows, k, &o_rowbuf, &o_seg, &o_tot);
    ImgSmem sm;
    sm.cnt = reinterpret_cast<int32_t *>(smem_raw);
    sm.rowbuf = smem_raw + o_rowbuf;
    sm.seg_start = reinterpret_cast<uint32_t *>(smem_raw + o_seg);
    sm.seg_end = sm.seg_start + tile_images;
    sm.class_tot = reinterpret_cast<unsigned long long *>(smem_raw + o_tot);
    sm.part = sm.class_tot + k;
    sm.first_beyond = reinterpret_cast<uint32_t *>(sm.part + 8);
    sm.tile_tot = sm.first_beyond + 4;
    const uint32_t tid = threadIdx.x;

    const uint32_t G = gridDim.x, b = blockIdx.x;
    const int32_t img_end_all = image_base + int32_t(n_images);
    auto nominal = [&](uint64_t i) -> uint64_t {
        return i >= G ? rows : (((rows / G) * i + (rows % G) * i / G) & ~uint64_t(kRowsPerThread - 1));
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
        if (I1 < I0) I1 = I0;
    }

    for (uint32_t e = tid; e < tile_images * pitch; e += blockDim.x) sm.cnt[e] = 0;
    for (uint32_t i = tid; i < tile_images; i += blockDim.x) { sm.seg_start[i] = 0; sm.seg_end[i] = 0; }
    for (uint32_t c = tid; c < k + 8; c += blockDim.x) sm.class_tot[c] = 0;
    for (uint32_t c = tid; c < k; c += blockDim.x) sm.tile_tot[c] = 0;
    if (tid == 0) *sm.first_beyond = kNoRow;
    __syncthreads();

    uint32_t seen = 0, unsorted = 0;
    int32_t tb = I0;                                                         // first image of the open tile
    int32_t tile_end = I1 - tb > int32_t(tile_images) ? tb + int32_t(tile_images) : I1;

    // write the open tile, fold it into the partials, zero it, open the next one (uniform)
    auto flush = [&]() {
        __syncthreads();
        const uint32_t n_img = uint32_t(tile_end - tb);
        int32_t *dst = counts + size_t(tb - image_base) * k;
        unsigned long long s2 = 0;
        // element order: coalesced stores, class totals, sum of squares
        uint32_t i = tid / k, c = tid - i * k;
        const uint32_t di = blockDim.x / k, dc = blockDim.x - di * k;
        for (uint32_t e = tid; e < n_img * k; e += blockDim.x) {
            const uint32_t v = uint32_t(sm.cnt[i * pitch + c]);
            dst[e] = int32_t(v);
            if (v) {
                s2 += (unsigned long long)v * v;
                atomicAdd(&sm.tile_tot[c], v);                              // native 32-bit shared atomic; a tile holds < 2^32 ratings
            }
            i += di; c += dc;
            if (c >= k) { c -= k; ++i; }
        }
        // one thread per image: n_i and what derives from it (pitch is odd: conflict-free)
        unsigned long long r = 0, pairs = 0, rated = 0, pair_images = 0;
        for (uint32_t im = tid; im < n_img; im += blockDim.x) {
            const int32_t *row = sm.cnt + im * pitch;
            unsigned long long n = 0;
            for (uint32_t j = 0; j < k; ++j) n += uint32_t(row[j]);
            r += n; rated += n >= 1; pair_images += n >= 2; pairs += n * (n - (n > 0));
        }
        part_add(sm.part, P_S2, s2);
        part_add(sm.part, P_R, r);
        part_add(sm.part, P_RATED, rated);
        part_add(sm.part, P_PAIR_IMAGES, pair_images);
        part_add(sm.part, P_PAIRS, pairs);
        __syncthreads();
        for (uint32_t c2 = tid; c2 < k; c2 += blockDim.x) {                  // fold the tile's class totals into 64 bits
            sm.class_tot[c2] += sm.tile_tot[c2];
            sm.tile_tot[c2] = 0;
        }
        for (uint32_t e = tid; e < n_img * pitch; e += blockDim.x) sm.cnt[e] = 0;
        tb = tile_end;
        tile_end = I1 - tb > int32_t(tile_images) ? tb + int32_t(tile_images) : I1;
        __syncthreads();
    };

    uint64_t consume_from = nom0;                                            // first row not yet tallied (uniform)
    bool stream_done = rows == 0 || nom0 >= rows;
    while (!stream_done) {
        // ---------------- phase A: fill the row buffer ----------------
        const uint64_t buf_row0 = consume_from & ~uint64_t(kRowsPerThread - 1);
        // number of 4096-row blocks this phase may buffer (uniform), and where they end
        uint32_t n_blk = cap_rows / kImgBlockRows;
        {
            const uint64_t left = (rows - buf_row0 + kImgBlockRows - 1) / kImgBlockRows;
            if (left < n_blk) n_blk = uint32_t(left);
        }
        const uint64_t loaded_end = buf_row0 + uint64_t(n_blk) * kImgBlockRows < rows
                                        ? buf_row0 + uint64_t(n_blk) * kImgBlockRows : rows;
        auto process = [&](const RowBlock &rbk, uint64_t row0) {
            const int32_t ii[16] = {rbk.idx[0].x, rbk.idx[0].y, rbk.idx[0].z, rbk.idx[0].w, rbk.idx[1].x, rbk.idx[1].y,
                                    rbk.idx[1].z, rbk.idx[1].w, rbk.idx[2].x, rbk.idx[2].y, rbk.idx[2].z, rbk.idx[2].w,
                                    rbk.idx[3].x, rbk.idx[3].y, rbk.idx[3].z, rbk.idx[3].w};
            const uint32_t cw[4] = {rbk.cls.x, rbk.cls.y, rbk.cls.z, rbk.cls.w};
            const uint32_t aw[4] = {rbk.act.x, rbk.act.y, rbk.act.z, rbk.act.w};
            int32_t prev = (row0 > 0 && row0 <= rows) ? __ldg(image_idx + row0 - 1) : INT32_MIN;
            uint32_t packed[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            const uint32_t off0 = uint32_t(row0 - buf_row0);
            // Common path: all 16 rows exist, are new to this phase and lie in the open tile (rows are ordered,
            // so the first and the last decide).  The 16 class bytes are merged with the active / range masks
            // four at a time; image changes are rare (one per ~100 rows), so each row costs one compare and a
            // short, rarely taken block that records where the images start and end in the row buffer.  Only
            // threads at the edges of a phase or of a tile take the general path below: divergence is
            // confined to the one or two warps that hold such an edge.
            if (row0 > consume_from && row0 + kRowsPerThread <= rows && ii[0] >= tb && ii[15] < tile_end &&
                prev <= ii[0]) {
                const uint32_t n_tile = uint32_t(tile_end - tb);
                const uint32_t k4 = k >= 256 ? 0u : k * 0x01010101u;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t in_range = k >= 256 ? 0xffffffffu : __vcmpltu4(cw[w], k4);
                    const uint32_t keep = in_range & __vcmpne4(aw[w], 0u);
                    packed[w] = (cw[w] & keep) | ~keep;
                    seen += __popc(in_range) >> 3;
                }
                *reinterpret_cast<uint4 *>(sm.rowbuf + off0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                const bool count_order = row0 + kRowsPerThread <= nom1;      // nom1 is 16-aligned: all rows or none
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int32_t img = ii[j];
                    if (img != prev) {
                        if (count_order) unsorted += img < prev;
                        const uint32_t rel = uint32_t(img - tb), relp = uint32_t(prev - tb);
                        if (rel < n_tile) sm.seg_start[rel] = off0 + j;
                        if (relp < n_tile) sm.seg_end[relp] = off0 + j;
                    }
                    prev = img;
                }
                return;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint64_t row = row0 + j;
                const int32_t img = ii[j];
                const bool valid = row < rows && row >= consume_from;        // exists and not tallied in an earlier phase
                const uint32_t c = (cw[j >> 2] >> (8 * (j & 3))) & 0xffu;
                const uint32_t a = (aw[j >> 2] >> (8 * (j & 3))) & 0xffu;
                if (valid) {
                    if (row < nom1) unsorted += img < prev;                  // pairs (r-1, r), nom0 <= r < nom1
                    const bool in_tile = (img >= tb) & (img < tile_end);
                    const bool first_here = (img != prev) | (row == consume_from);
                    if (in_tile) {
                        seen += c < k;
                        if (a && c < k) packed[j >> 2] = (packed[j >> 2] & ~(0xffu << (8 * (j & 3)))) | (c << (8 * (j & 3)));
                        if (first_here) sm.seg_start[img - tb] = off0 + j;
                    } else if (img >= tile_end && first_here && (prev < tile_end || row == consume_from)) {
                        atomicMin(sm.first_beyond, off0 + j);                // the stream leaves the tile here
                    }
                    if (img != prev && row > consume_from && prev >= tb && prev < tile_end)
                        sm.seg_end[prev - tb] = off0 + j;                    // the previous image's rows end here
                }
                prev = img;
            }
            *reinterpret_cast<uint4 *>(sm.rowbuf + off0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        };
        // Software pipeline: the loads of block nb+1 are in flight while block nb is scanned; no barrier
        // between blocks.  Rows are ordered, so the image of a block's LAST row tells every thread alike
        // whether the stream leaves the tile inside that block: if it does, nothing further is loaded.
        {
            RowBlock q[2];
            const uint64_t trow = uint64_t(tid) * kRowsPerThread;
            auto last_img_of = [&](uint32_t cb) -> int32_t {
                const uint64_t end = buf_row0 + uint64_t(cb + 1) * kImgBlockRows;
                return __ldg(image_idx + (end < rows ? end : rows) - 1);
            };
            int32_t last_img = 0, next_last = 0;
            if (n_blk) { load_rows(q[0], image_idx, class_idx, active, buf_row0 + trow, rows); last_img = last_img_of(0); }
            for (uint32_t nb = 0; nb < n_blk; nb += 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t cb = nb + u;
                    if (cb < n_blk) {
                        const uint64_t blk = buf_row0 + uint64_t(cb) * kImgBlockRows;
                        const bool more = cb + 1 < n_blk && last_img < tile_end;
                        if (more) {
                            load_rows(q[u ^ 1], image_idx, class_idx, active, blk + kImgBlockRows + trow, rows);
                            next_last = last_img_of(cb + 1);
                        }
                        process(q[u], blk + trow);
                        last_img = next_last;
                        if (!more) nb = n_blk;                               // leave both loops after this block
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t fb = *sm.first_beyond;
        const uint64_t stop = fb != kNoRow ? buf_row0 + fb : loaded_end;     // first row NOT tallied by this phase
        if (tid == 0 && stop > consume_from) {                               // close the last image of the buffer
            const int32_t li = __ldg(image_idx + stop - 1);
            if (li >= tb && li < tile_end) sm.seg_end[li - tb] = uint32_t(stop - buf_row0);
        }
        __syncthreads();
        // ---------------- phase B: one thread per image ----------------
        const uint32_t n_tile = uint32_t(tile_end - tb);
        for (uint32_t im = tid; im < n_tile; im += blockDim.x) {
            const uint32_t s = sm.seg_start[im], e = sm.seg_end[im];
            int32_t *my = sm.cnt + im * pitch;
            uint32_t r = s;
            for (; r + 4 <= e; r += 4) {          // four row bytes first, then the counter updates: the loads
                const uint32_t m0 = sm.rowbuf[r], m1 = sm.rowbuf[r + 1], m2 = sm.rowbuf[r + 2], m3 = sm.rowbuf[r + 3];
                if (m0 != 0xffu) my[m0] += 1;     // do not wait behind the stores
                if (m1 != 0xffu) my[m1] += 1;
                if (m2 != 0xffu) my[m2] += 1;
                if (m3 != 0xffu) my[m3] += 1;
            }
            for (; r < e; ++r) {
                const uint32_t m = sm.rowbuf[r];
                if (m != 0xffu) my[m] += 1;
            }
            sm.seg_start[im] = 0;
            sm.seg_end[im] = 0;
        }
        if (tid == 0) *sm.first_beyond = kNoRow;
        consume_from = stop;
        // the tile is complete when the stream left it (or ended); the stream is over for this CTA when it
        // ended or when the next row belongs to somebody else's images
        const bool left_tile = fb != kNoRow || stop >= rows;
        bool next_foreign = stop >= rows;
        if (!next_foreign && fb != kNoRow) next_foreign = __ldg(image_idx + stop) >= I1;
        if (left_tile) {
            if (tb < I1) flush(); else __syncthreads();
            if (!next_foreign) {                                             // tiles without any row: write zeros, no reload
                const int32_t ni = __ldg(image_idx + stop);
                while (tb < I1 && ni >= tile_end) flush();
            }
            if (next_foreign) {
                // order check still has to reach nom1 when this CTA's images end early (unsorted input only)
                stream_done = true;
            } else if (tb >= I1) {
                stream_done = true;
            }
        } else {
            __syncthreads();
        }
    }
    // remaining tiles (image ranges without rows), then the order check of rows this CTA did not stream
    while (tb < I1) flush();
    for (uint64_t r = (consume_from > nom0 ? consume_from : nom0) + tid; r < nom1 && r < rows; r += blockDim.x) {
        const int32_t prev = r > 0 ? __ldg(image_idx + r - 1) : INT32_MIN;
        unsorted += __ldg(image_idx + r) < prev;
    }
    part_add(sm.part, P_ROWS_SEEN, seen);
    part_add(sm.part, P_UNSORTED, unsorted);
    __syncthreads();
    commit_partials(sm.part, sm.class_tot, k, g_partials);
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

__host__ __device__ inline size_t slab_smem_bytes(uint32_t stages, uint32_t tile_images, uint32_t k) {
    return size_t(stages) * kSlabStageBytes + ((size_t(tile_images) * k * 4 + 15) & ~size_t(15)) + 1024 +
           size_t(k) * 8 + 8 * 8 + 16 + size_t(stages) * 8;
}

// One row of a quad, branch-free: in the window and class in range => seen++; also active => one shared
// RED.ADD on the counter of (image & tmask, class).  Eleven instructions; written in PTX so that the atomic
// is unconditional: ptxas turns every predicated shared atomic into BSSY / BRA / BSYNC, and a literal 1 into
// ATOMS.POPC.INC, so the increment is a register holding 1 or 0.  The address is always inside the tile plus
// its 1 KB pad (slot < T, class byte < 256), so rows that do not count simply add 0.  No "memory" clobber on purpose: the counters are only read
// after a __syncthreads(), and the clobber would stop the next quad's loads from being hoisted.
template <int J>
__device__ __forceinline__ void slab_row(int32_t img, uint32_t cw, uint32_t aw, int32_t tb, uint32_t span, uint32_t k,
                                         uint32_t tmask, uint32_t k4, uint32_t tile_s, uint32_t one, uint32_t &seen) {
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

template <int KC, int NS>                                         // KC = classes per lane (k <= 32 KC), NS = ring depth
__global__ void __launch_bounds__(kSlabThreads, 2)
tally_slab_kernel(const int32_t *__restrict__ image_idx, const uint8_t *__restrict__ class_idx,
                  const uint8_t *__restrict__ active, uint64_t rows, int32_t image_base, uint32_t n_images,
                  uint32_t k, uint32_t tile_log2, int32_t *__restrict__ counts,
                  unsigned long long *__restrict__ g_partials) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t T = 1u << tile_log2, tmask = T - 1u;
    uint8_t *ring = smem_raw;
    int32_t *tile = reinterpret_cast<int32_t *>(smem_raw + size_t(NS) * kSlabStageBytes);
    unsigned long long *class_tot = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<uint8_t *>(tile) + ((size_t(T) * k * 4 + 15) & ~size_t(15)) + 1024);   // pad: see slab_row
    unsigned long long *part = class_tot + k;                    // 8 (7 used)
    int32_t *s_beyond = reinterpret_cast<int32_t *>(part + 8);   // 1 (+ padding to 16 bytes)
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_beyond + 4);
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
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // Per-lane accumulators of the flushes; reduced across the CTA once, at the end of the kernel.
    unsigned long long tot[KC], s2 = 0, pairs = 0;                // tot[cc]: class lane + 32 cc
    uint32_t rated = 0, pair_images = 0;
#pragma unroll
    for (int cc = 0; cc < KC; ++cc) tot[cc] = 0;
    auto fold_image = [&](uint32_t n) {                           // n_i of one image (0 is harmless)
        rated += n >= 1u;
        pair_images += n >= 2u;
        pairs += (unsigned long long)n * (n - 1u);                // 0 * 0xffffffff = 0
    };

    // Write images [a, e) (complete, all mine) to d_counts, fold them into the partials, leave their
    // slots zeroed.  The first min(e - a, T) come from the ring, the rest have no rows.  Uniform.
    // A warp per image, a lane per class; n_i by REDUX, parked in lane (iteration mod 32) so that what
    // derives from it is computed for 32 images at once.
    auto flush = [&](int32_t a, int32_t e) {
        __syncthreads();
        const uint32_t total = uint32_t(e - a);
        const uint32_t n_ring = total < T ? total : T;
        uint32_t n_mine = 0, it = 0;
        for (uint32_t i = wid; i < n_ring; i += kSlabWarps, ++it) {
            const int32_t img = a + int32_t(i);
            int32_t *src = tile + (uint32_t(img) & tmask) * k;
            int32_t *dst = counts + size_t(img - image_base) * k;
            uint32_t v[KC], n = 0;
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
                n += v[cc];
            }
            n = __reduce_add_sync(0xffffffffu, n);
            if ((it & 31u) == lane) n_mine = n;
            if ((it & 31u) == 31u) { fold_image(n_mine); n_mine = 0; }
        }
        fold_image(n_mine);
        if (total > n_ring) {                                     // images without rows
            int32_t *z = counts + size_t(a - image_base + int32_t(n_ring)) * k;
            const size_t ne = size_t(total - n_ring) * k;
            for (size_t i = tid; i < ne; i += kSlabThreads) z[i] = 0;
        }
        __syncthreads();
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
                    slab_row<0>(iq.x, cw, aw, tb, span, k, tmask, k4, tile_s, one, seen);
                    slab_row<1>(iq.y, cw, aw, tb, span, k, tmask, k4, tile_s, one, seen);
                    slab_row<2>(iq.z, cw, aw, tb, span, k, tmask, k4, tile_s, one, seen);
                    slab_row<3>(iq.w, cw, aw, tb, span, k, tmask, k4, tile_s, one, seen);
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
                       double *__restrict__ block_sums, unsigned int *__restrict__ ticket) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const TallySmem sm = carve_smem(smem_raw, tile_images, k);
    int32_t *tile = sm.tile;
    double *s_dred = sm.dred;

    for (uint32_t c = threadIdx.x; c < k + 8; c += blockDim.x) sm.class_tot[c] = 0;
    double my_pi = 0.0;
    const uint32_t n_tiles = (n_images + tile_images - 1) / tile_images;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t i0 = t * tile_images;
        const uint32_t n_img = min(tile_images, n_images - i0);
        const int32_t *src = counts + size_t(i0) * k;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < n_img * k; e += blockDim.x) tile[e] = __ldg(src + e);
        __syncthreads();
        if (sum_pi_out) flush_tile<false, true>(tile, n_img, k, nullptr, sm.class_tot, sm.part, &my_pi);
        else flush_tile<false, false>(tile, n_img, k, nullptr, sm.class_tot, sm.part, nullptr);
    }
    __syncthreads();
    commit_partials(sm.part, sm.class_tot, k, g_partials);
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
    return tile + size_t(k) * 8 + 8 * 8 + size_t(kTallyThreads) * 8;
}
constexpr uint32_t kFleissGridMax = 592;

static cudaError_t ensure_smem(const void *fn, size_t bytes) {
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

}  // namespace b2

// B2_TALLY_PATH: unset = auto (slab kernel; warp kernel for >= 512 rows per image), 0 = thread-per-image kernel,
// 1 = MATCH.ANY warp kernel, 2 = shared-atomic tile kernel, 3 = thread-per-row slab kernel.
static int tally_path_override() {
    const char *e = getenv("B2_TALLY_PATH");
    return e ? atoi(e) : -1;
}

extern "C" uint64_t b2_label_tally_workspace_bytes(uint32_t n_images) {
    (void)n_images;
    return 0;
}

extern "C" uint64_t b2_fleiss_workspace_bytes(uint32_t n_images) {
    (void)n_images;
    return 16 + 8ull * b2::kFleissGridMax;
}

extern "C" int b2_label_tally(const int32_t *d_image_idx, const uint8_t *d_class_idx, const uint8_t *d_active,
                              uint64_t rows, uint32_t image_base, uint32_t n_images, uint32_t k, uint32_t flags,
                              int32_t *d_counts, int64_t *d_partials, void *d_workspace,
                              uint64_t workspace_bytes, void *stream) {
    using namespace b2;
    (void)d_workspace; (void)workspace_bytes;
    B2_REQUIRE(d_counts && d_partials, "b2_label_tally: null output pointer");
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
    B2_CUDA_CHECK(cudaMemsetAsync(partials, 0, (size_t(k) + B2_PARTIALS_EXTRA) * 8, st));
    const uint32_t tile_images = pick_tile_images(k);
    const size_t smem = tally_smem_bytes(tile_images, k);
    if (flags & B2_TALLY_SORTED) {
        // Auto: the slab kernel keeps the lanes of a warp in different images as long as an image has fewer rows
        // than a slab; images with very many rows go to the warp kernel (whose common case is "32 rows of the
        // same image").  The other two kernels stay selectable for comparison (profiles/r1_tally_history.md).
        int path = tally_path_override();
        if (path < 0) path = (rows / n_images >= 512) ? 1 : 3;
        if (path == 0) {                                     // thread-per-image kernel
            const uint32_t pitch = k | 1u;
            uint32_t ti = 192, cap = 20480;                 // 39 KB counters (k = 50) + 20 KB rows: three CTAs per SM
            if (const char *e = getenv("B2_TALLY_TILE")) ti = uint32_t(atoi(e));
            if (const char *e = getenv("B2_TALLY_CAP")) cap = uint32_t(atoi(e));
            const uint32_t by_smem = (56u * 1024u) / (pitch * 4u);           // counters <= 56 KB
            if (ti > by_smem) ti = by_smem;
            if (ti < 1) ti = 1;
            cap = cap / kImgBlockRows * kImgBlockRows;
            if (cap < uint32_t(kImgBlockRows)) cap = kImgBlockRows;
            size_t o1, o2, o3;
            const size_t smem_img = img_smem_layout(ti, pitch, cap, k, &o1, &o2, &o3);
            B2_CUDA_CHECK(ensure_smem(reinterpret_cast<const void *>(tally_image_kernel), smem_img));
            int per_sm = int((220u * 1024u) / (smem_img + 1024));
            if (per_sm > 4) per_sm = 4;
            if (per_sm < 1) per_sm = 1;
            uint64_t want = rows ? (rows + 4ull * kImgBlockRows - 1) / (4ull * kImgBlockRows)
                                 : (uint64_t(n_images) + ti - 1) / ti;
            const uint64_t cap_ctas = uint64_t(per_sm) * uint64_t(sm_count());
            if (want > cap_ctas) want = cap_ctas;
            if (want < 1) want = 1;
            tally_image_kernel<<<uint32_t(want), kImgThreads, smem_img, st>>>(d_image_idx, d_class_idx, d_active, rows,
                                                                             int32_t(image_base), n_images, k, ti, cap,
                                                                             d_counts, partials);
            B2_LAUNCH_CHECK("tally_image_kernel");
            return B2_OK;
        }
        if (path == 3) {                                     // thread-per-row slab kernel
            uint32_t stages = 3, per_sm = 2;
            if (const char *e = getenv("B2_TALLY_STAGES")) stages = atoi(e) == 2 ? 2u : 3u;
            if (const char *e = getenv("B2_TALLY_CTAS")) per_sm = uint32_t(atoi(e)) < 1 ? 1u : uint32_t(atoi(e));
            const size_t budget = (227u * 1024u) / per_sm - 1024u;               // per CTA, incl. the 1 KB the driver reserves
            uint32_t t = 0;
            while (t < 13 && slab_smem_bytes(stages, 2u << t, k) <= budget) ++t;
            if (const char *e = getenv("B2_TALLY_TILE_LOG2")) t = uint32_t(atoi(e));
            const size_t smem_slab = slab_smem_bytes(stages, 1u << t, k);
            uint64_t want = rows ? (rows + 2ull * kSlabStageRows - 1) / (2ull * kSlabStageRows)
                                 : (uint64_t(n_images) + 1023) / 1024;
            const uint64_t cap_ctas = uint64_t(per_sm) * uint64_t(sm_count());
            if (want > cap_ctas) want = cap_ctas;
            if (want < 1) want = 1;
            auto launch = [&](auto kern) -> int {
                B2_CUDA_CHECK(ensure_smem(reinterpret_cast<const void *>(kern), smem_slab));
                kern<<<uint32_t(want), kSlabThreads, smem_slab, st>>>(d_image_idx, d_class_idx, d_active, rows,
                                                                     int32_t(image_base), n_images, k, t, d_counts, partials);
                B2_LAUNCH_CHECK("tally_slab_kernel");
                return B2_OK;
            };
            if (stages == 2) {
                if (k <= 32) return launch(tally_slab_kernel<1, 2>);
                if (k <= 64) return launch(tally_slab_kernel<2, 2>);
                if (k <= 128) return launch(tally_slab_kernel<4, 2>);
                return launch(tally_slab_kernel<8, 2>);
            }
            if (k <= 32) return launch(tally_slab_kernel<1, 3>);
            if (k <= 64) return launch(tally_slab_kernel<2, 3>);
            if (k <= 128) return launch(tally_slab_kernel<4, 3>);
            return launch(tally_slab_kernel<8, 3>);
        }
        if (path == 2) {                                     // shared-memory-atomic tile kernel (comparison only)
            B2_CUDA_CHECK(ensure_smem(reinterpret_cast<const void *>(tally_sorted_kernel), smem));
            uint64_t want = (rows + 4ull * kBlockRows - 1) / (4ull * kBlockRows);
            const uint64_t by_images = (uint64_t(n_images) + tile_images - 1) / tile_images;
            if (want < by_images) want = by_images;
            const uint64_t cap = 2ull * uint64_t(sm_count());
            const uint32_t grid = uint32_t(want < 1 ? 1 : (want > cap ? cap : want));
            tally_sorted_kernel<<<grid, kTallyThreads, smem, st>>>(d_image_idx, d_class_idx, d_active, rows,
                                                                   int32_t(image_base), n_images, k, tile_images,
                                                                   d_counts, partials);
            B2_LAUNCH_CHECK("tally_sorted_kernel");
            return B2_OK;
        }
        // one warp per ~4096 rows (or per 64 images when there are no rows), at most one resident wave
        const uint64_t warps_per_cta = kWarpKernelThreads / 32;
        uint64_t want_warps = (rows + 4095) / 4096;
        const uint64_t by_images = (uint64_t(n_images) + 63) / 64;
        if (rows == 0 && want_warps < by_images) want_warps = by_images;
        if (want_warps < 1) want_warps = 1;
        const uint64_t cap = 4ull * uint64_t(sm_count()) * warps_per_cta;   // 4 CTAs of 8 warps per SM (48 KB each): one wave
        if (want_warps > cap) want_warps = cap;
        if (rows > 0 && want_warps > rows / kGroupRows) want_warps = rows / kGroupRows ? rows / kGroupRows : 1;
        const uint64_t nw = want_warps;                       // distinct, 128-row aligned nominal starts
        const uint32_t grid = uint32_t((nw + warps_per_cta - 1) / warps_per_cta);
        B2_CUDA_CHECK(ensure_smem(reinterpret_cast<const void *>(tally_warp_kernel), kWarpKernelSmem));
        tally_warp_kernel<<<grid, kWarpKernelThreads, kWarpKernelSmem, st>>>(d_image_idx, d_class_idx, d_active, rows,
                                                                            int32_t(image_base), n_images, k, nw,
                                                                            d_counts, partials);
        B2_LAUNCH_CHECK("tally_warp_kernel");
        return B2_OK;
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
                                                             nullptr, nullptr, nullptr);
    B2_LAUNCH_CHECK("fleiss_partials_kernel");
    return B2_OK;
}

extern "C" int b2_fleiss_partials(const int32_t *d_counts, uint32_t n_images, uint32_t k, int64_t *d_partials,
                                  double *d_sum_pi, void *d_workspace, uint64_t workspace_bytes, void *stream) {
    using namespace b2;
    B2_REQUIRE(d_counts && d_partials, "b2_fleiss_partials: null pointer");
    B2_REQUIRE(k >= 1 && k <= 256 && n_images >= 1, "b2_fleiss_partials: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *partials = reinterpret_cast<unsigned long long *>(d_partials);
    B2_CUDA_CHECK(cudaMemsetAsync(partials, 0, (size_t(k) + B2_PARTIALS_EXTRA) * 8, st));
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
                                                             d_sum_pi, block_sums, ticket);
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
// for every annotator at once.  Rows sorted by (annotator_idx, image_idx): an active row opens a
// new distinct image iff no earlier active row of the same annotator has the same image.
namespace b2 {
__global__ void __launch_bounds__(256)
distinct_images_kernel(const int32_t *__restrict__ annotator_idx, const int32_t *__restrict__ image_idx,
                       const uint8_t *__restrict__ active, uint64_t rows, uint32_t n_annotators,
                       uint32_t *__restrict__ distinct) {
    for (uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; r < rows; r += uint64_t(gridDim.x) * blockDim.x) {
        if (!active[r]) continue;
        const int32_t a = annotator_idx[r], img = image_idx[r];
        if (a < 0 || uint32_t(a) >= n_annotators) continue;
        // walk back over the rows of the same (annotator, image) run looking for an earlier active one
        bool first = true;
        for (uint64_t q = r; q-- > 0;) {
            if (annotator_idx[q] != a || image_idx[q] != img) break;
            if (active[q]) { first = false; break; }
        }
        if (first) atomicAdd(&distinct[a], 1u);
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
