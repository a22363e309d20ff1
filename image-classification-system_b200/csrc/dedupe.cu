// Dedupe decision on 256-bit content hashes for sm_100a.
//
// Replaces the sequential lookup-then-insert loop of WebDAVSync._process_image_batch
// (app/services/webdav_sync.py:311-400): `SELECT ... WHERE content_hash = ?` (:324), insert +
// flush when absent (:329-354, created += 1), update when present — including a copy inserted
// earlier in the same batch (:371-398, updated += 1) — processed += 1 (:400).
//
// Parallel restatement: an open-addressing table keyed by the digest holds, per distinct
// digest, the (seq, index) of its earliest and latest valid occurrence (64-bit atomicMin /
// atomicMax, so the answer does not depend on thread scheduling).  Pass 2 looks every digest
// up again: is_new[i] = (first occurrence == i) && digest not in the sorted `existing` table
// (binary search, memcmp order).  Counts are block-reduced and added with integer atomics.
#include "common.cuh"

namespace b2 {

constexpr unsigned long long kEmpty = ~0ull;

struct Digest {
    uint4 lo, hi;
};
__device__ __forceinline__ Digest load_digest(const uint8_t *digests, uint64_t i) {
    const uint4 *p = reinterpret_cast<const uint4 *>(digests + 32ull * i);
    Digest d;
    d.lo = p[0];
    d.hi = p[1];
    return d;
}
__device__ __forceinline__ bool digest_eq(const Digest &a, const Digest &b) {
    return a.lo.x == b.lo.x && a.lo.y == b.lo.y && a.lo.z == b.lo.z && a.lo.w == b.lo.w &&
           a.hi.x == b.hi.x && a.hi.y == b.hi.y && a.hi.z == b.hi.z && a.hi.w == b.hi.w;
}
// memcmp order: compare bytes ascending = compare big-endian words.
__device__ __forceinline__ int digest_cmp(const Digest &a, const Digest &b) {
    const uint32_t aw[8] = {a.lo.x, a.lo.y, a.lo.z, a.lo.w, a.hi.x, a.hi.y, a.hi.z, a.hi.w};
    const uint32_t bw[8] = {b.lo.x, b.lo.y, b.lo.z, b.lo.w, b.hi.x, b.hi.y, b.hi.z, b.hi.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t x = __byte_perm(aw[j], 0, 0x0123), y = __byte_perm(bw[j], 0, 0x0123);
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}
__device__ __forceinline__ uint32_t digest_slot(const Digest &d, uint32_t mask) {
    // digests are uniformly distributed already; fold two words so that a truncated or
    // structured test digest still spreads
    uint32_t h = d.lo.x ^ (d.hi.w * 0x9E3779B1u) ^ (d.lo.z >> 7);
    return h & mask;
}

// -1 if absent, else position in the sorted table.
__device__ __forceinline__ int64_t find_sorted(const uint8_t *existing, uint64_t m, const Digest &key) {
    uint64_t lo = 0, hi = m;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        const int c = digest_cmp(load_digest(existing, mid), key);
        if (c == 0) return int64_t(mid);
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    return -1;
}

__global__ void __launch_bounds__(256)
dedupe_insert_kernel(const uint8_t *__restrict__ digests, const uint8_t *__restrict__ valid,
                     const uint32_t *__restrict__ seq, uint32_t n,
                     unsigned long long *__restrict__ first_tab, unsigned long long *__restrict__ last_tab,
                     uint32_t mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (valid && !valid[i]) return;
    const Digest key = load_digest(digests, i);
    const unsigned long long s = seq ? seq[i] : i;
    const unsigned long long vfirst = (s << 32) | i;
    const unsigned long long vlast = ((s + 1) << 32) | i;     // 0 = empty for the max table
    uint32_t slot = digest_slot(key, mask);
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        unsigned long long cur = atomicCAS(&first_tab[slot], kEmpty, vfirst);
        bool mine = (cur == kEmpty);
        if (!mine) {
            const uint32_t j = uint32_t(cur & 0xffffffffu);
            mine = digest_eq(load_digest(digests, j), key);
            if (mine) atomicMin(&first_tab[slot], vfirst);
        }
        if (mine) {
            atomicMax(&last_tab[slot], vlast);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(256)
dedupe_resolve_kernel(const uint8_t *__restrict__ digests, const uint8_t *__restrict__ valid, uint32_t n,
                      const uint8_t *__restrict__ existing, uint64_t m,
                      const unsigned long long *__restrict__ first_tab,
                      const unsigned long long *__restrict__ last_tab, uint32_t mask,
                      uint8_t *__restrict__ is_new, int32_t *__restrict__ first_index,
                      int32_t *__restrict__ last_index, uint32_t *__restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t processed = 0, created = 0;
    if (i < n) {
        int32_t first = -1, last = -1;
        uint8_t fresh = 0;
        if (!valid || valid[i]) {
            processed = 1;
            const Digest key = load_digest(digests, i);
            uint32_t slot = digest_slot(key, mask);
            for (uint32_t probe = 0; probe <= mask; ++probe) {
                const unsigned long long cur = first_tab[slot];
                if (cur == kEmpty) break;                       // cannot happen after insert
                const uint32_t j = uint32_t(cur & 0xffffffffu);
                if (j == i || digest_eq(load_digest(digests, j), key)) {
                    first = int32_t(j);
                    last = int32_t(last_tab[slot] & 0xffffffffu);
                    break;
                }
                slot = (slot + 1) & mask;
            }
            if (first == int32_t(i)) {
                fresh = (m == 0 || find_sorted(existing, m, key) < 0) ? 1 : 0;
                created = fresh;
            }
        }
        is_new[i] = fresh;
        if (first_index) first_index[i] = first;
        if (last_index) last_index[i] = last;
    }
    // block reduction of the two counters
    processed = __reduce_add_sync(0xffffffffu, processed);
    created = __reduce_add_sync(0xffffffffu, created);
    __shared__ uint32_t sp[8], sc[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sp[warp] = processed; sc[warp] = created; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t p = 0, c = 0;
        for (int w = 0; w < int(blockDim.x >> 5); ++w) { p += sp[w]; c += sc[w]; }
        if (p) atomicAdd(&counts[0], p);
        if (c) atomicAdd(&counts[1], c);
        if (p - c) atomicAdd(&counts[2], p - c);
    }
}

__global__ void __launch_bounds__(256)
lookup_sorted_kernel(const uint8_t *__restrict__ digests, uint32_t n, const uint8_t *__restrict__ existing,
                     uint64_t m, int64_t *__restrict__ found) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    found[i] = m ? find_sorted(existing, m, load_digest(digests, i)) : -1;
}

// ---- dictionary-encode rows of table `classificacoes` -----------------------------
// id_img arrives as the 64 lowercase hex characters of the content hash (String(64) foreign key,
// app/db/models.py:229), id_opc as a 16-byte UUID (:231), ativo as a byte (:233).  The image dictionary is the
// sorted digest table of b2_dedupe (image index = position), the option dictionary the environment's sorted
// option UUIDs (class index = position).  A thread per row: decode the hex key (rejecting anything that is
// not [0-9a-f]), binary-search both dictionaries (the image table of 1 M keys = 32 MB stays in L2).
__device__ __forceinline__ bool hex_word(uint32_t c4lo, uint32_t c4hi, uint32_t &out) {
    // eight ASCII hex characters (c4lo = chars 0..3, c4hi = chars 4..7) -> four bytes in memory order
    uint32_t v = 0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t c = ((i < 4 ? c4lo : c4hi) >> (8 * (i & 3))) & 0xffu;
        const bool digit = c - 0x30u <= 9u, alpha = c - 0x61u <= 5u;
        ok &= digit | alpha;
        const uint32_t nib = digit ? c - 0x30u : c - 0x57u;
        v |= (nib & 0xfu) << (8 * (i >> 1) + ((i & 1) ? 0 : 4));       // first character of a pair = high nibble
    }
    out = v;
    return ok;
}

__device__ __forceinline__ int cmp16(const uint4 &a, const uint4 &b) {
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t x = __byte_perm(aw[j], 0, 0x0123), y = __byte_perm(bw[j], 0, 0x0123);
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}

__global__ void __launch_bounds__(256)
encode_label_rows_kernel(const char *__restrict__ img_hex, const uint8_t *__restrict__ opc_uuid,
                         const uint8_t *__restrict__ ativo, uint64_t rows, const uint8_t *__restrict__ image_keys,
                         uint64_t n_images, const uint8_t *__restrict__ option_keys, uint32_t k,
                         int32_t *__restrict__ image_idx, uint8_t *__restrict__ class_idx, uint8_t *__restrict__ active,
                         unsigned long long *__restrict__ unknown) {
    uint32_t bad_img = 0, bad_opc = 0;
    for (uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; r < rows; r += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 *hx = reinterpret_cast<const uint4 *>(img_hex + 64 * r);
        const uint4 h0 = __ldg(hx), h1 = __ldg(hx + 1), h2 = __ldg(hx + 2), h3 = __ldg(hx + 3);
        Digest key;
        bool ok = hex_word(h0.x, h0.y, key.lo.x);
        ok &= hex_word(h0.z, h0.w, key.lo.y);
        ok &= hex_word(h1.x, h1.y, key.lo.z);
        ok &= hex_word(h1.z, h1.w, key.lo.w);
        ok &= hex_word(h2.x, h2.y, key.hi.x);
        ok &= hex_word(h2.z, h2.w, key.hi.y);
        ok &= hex_word(h3.x, h3.y, key.hi.z);
        ok &= hex_word(h3.z, h3.w, key.hi.w);
        const int64_t pos = ok && n_images ? find_sorted(image_keys, n_images, key) : -1;
        image_idx[r] = int32_t(pos);
        bad_img += pos < 0;
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(opc_uuid + 16 * r));
        uint32_t lo = 0, hi = k, cls = 0xffu;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const int c = cmp16(__ldg(reinterpret_cast<const uint4 *>(option_keys + 16 * size_t(mid))), u);
            if (c == 0) { cls = mid; break; }
            if (c < 0) lo = mid + 1; else hi = mid;
        }
        class_idx[r] = uint8_t(cls);
        bad_opc += cls == 0xffu;
        active[r] = ativo[r] ? 1 : 0;
    }
    bad_img = __reduce_add_sync(0xffffffffu, bad_img);
    bad_opc = __reduce_add_sync(0xffffffffu, bad_opc);
    if ((threadIdx.x & 31) == 0) {
        if (bad_img) atomicAdd(&unknown[0], (unsigned long long)bad_img);
        if (bad_opc) atomicAdd(&unknown[1], (unsigned long long)bad_opc);
    }
}

static uint32_t table_capacity(uint32_t n) {
    uint64_t cap = 64;
    while (cap < 2ull * n) cap <<= 1;
    return uint32_t(cap);
}

}  // namespace b2

extern "C" uint64_t b2_dedupe_workspace_bytes(uint32_t n) {
    return 16ull * b2::table_capacity(n);
}

extern "C" int b2_dedupe(const uint8_t *d_digests, const uint8_t *d_valid, const uint32_t *d_seq, uint32_t n,
                         const uint8_t *d_existing, uint64_t m, uint8_t *d_is_new, int32_t *d_first_index,
                         int32_t *d_last_index, uint32_t *d_counts, void *d_workspace,
                         uint64_t workspace_bytes, void *stream) {
    using namespace b2;
    B2_REQUIRE(d_counts != nullptr, "b2_dedupe: d_counts is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, 3 * sizeof(uint32_t), st));
    if (n == 0) return B2_OK;
    B2_REQUIRE(d_digests && d_is_new && d_workspace, "b2_dedupe: null pointer");
    B2_REQUIRE(n < 0x7fffffffu, "b2_dedupe: n too large");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(d_digests) & 15) == 0, "b2_dedupe: d_digests must be 16-byte aligned");
    B2_REQUIRE(m == 0 || (d_existing && (reinterpret_cast<uintptr_t>(d_existing) & 15) == 0),
               "b2_dedupe: d_existing must be non-null and 16-byte aligned when m > 0");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 7) == 0, "b2_dedupe: workspace must be 8-byte aligned");
    const uint32_t cap = table_capacity(n);
    if (workspace_bytes < 16ull * cap)
        return fail(B2_ERR_WORKSPACE, "b2_dedupe: workspace %llu < required %llu bytes",
                    (unsigned long long)workspace_bytes, 16ull * cap);
    unsigned long long *first_tab = static_cast<unsigned long long *>(d_workspace);
    unsigned long long *last_tab = first_tab + cap;
    B2_CUDA_CHECK(cudaMemsetAsync(first_tab, 0xff, 8ull * cap, st));
    B2_CUDA_CHECK(cudaMemsetAsync(last_tab, 0x00, 8ull * cap, st));
    const uint32_t grid = (n + 255) / 256;
    dedupe_insert_kernel<<<grid, 256, 0, st>>>(d_digests, d_valid, d_seq, n, first_tab, last_tab, cap - 1);
    B2_LAUNCH_CHECK("dedupe_insert_kernel");
    dedupe_resolve_kernel<<<grid, 256, 0, st>>>(d_digests, d_valid, n, d_existing, m, first_tab, last_tab,
                                                cap - 1, d_is_new, d_first_index, d_last_index, d_counts);
    B2_LAUNCH_CHECK("dedupe_resolve_kernel");
    return B2_OK;
}

extern "C" int b2_lookup_sorted(const uint8_t *d_digests, uint32_t n, const uint8_t *d_existing, uint64_t m,
                                int64_t *d_found_index, void *stream) {
    using namespace b2;
    if (n == 0) return B2_OK;
    B2_REQUIRE(d_digests && d_found_index, "b2_lookup_sorted: null pointer");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(d_digests) & 15) == 0, "b2_lookup_sorted: d_digests must be 16-byte aligned");
    B2_REQUIRE(m == 0 || (d_existing && (reinterpret_cast<uintptr_t>(d_existing) & 15) == 0),
               "b2_lookup_sorted: d_existing must be non-null and 16-byte aligned when m > 0");
    lookup_sorted_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_digests, n, d_existing, m, d_found_index);
    B2_LAUNCH_CHECK("lookup_sorted_kernel");
    return B2_OK;
}

extern "C" int b2_encode_label_rows(const char *d_img_hex, const uint8_t *d_opc_uuid, const uint8_t *d_ativo,
                                    uint64_t rows, const uint8_t *d_image_keys, uint64_t n_images,
                                    const uint8_t *d_option_keys, uint32_t k, int32_t *d_image_idx,
                                    uint8_t *d_class_idx, uint8_t *d_active, uint64_t *d_unknown, void *stream) {
    using namespace b2;
    B2_REQUIRE(d_unknown != nullptr, "b2_encode_label_rows: null d_unknown");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B2_CUDA_CHECK(cudaMemsetAsync(d_unknown, 0, 16, st));
    if (rows == 0) return B2_OK;
    B2_REQUIRE(d_img_hex && d_opc_uuid && d_ativo && d_image_idx && d_class_idx && d_active, "b2_encode_label_rows: null pointer");
    B2_REQUIRE(k <= 255, "b2_encode_label_rows: at most 255 options (class 255 marks an unknown option)");
    B2_REQUIRE(n_images < 0x7fffffffull, "b2_encode_label_rows: image dictionary too large for int32 indices");
    B2_REQUIRE(n_images == 0 || d_image_keys, "b2_encode_label_rows: null image dictionary");
    B2_REQUIRE(k == 0 || d_option_keys, "b2_encode_label_rows: null option dictionary");
    B2_REQUIRE(((reinterpret_cast<uintptr_t>(d_img_hex) | reinterpret_cast<uintptr_t>(d_opc_uuid) |
                 reinterpret_cast<uintptr_t>(d_image_keys) | reinterpret_cast<uintptr_t>(d_option_keys)) & 15) == 0,
               "b2_encode_label_rows: keys and dictionaries must be 16-byte aligned");
    uint64_t grid = (rows + 255) / 256;
    const uint64_t cap = 32ull * uint64_t(sm_count());
    if (grid > cap) grid = cap;
    encode_label_rows_kernel<<<unsigned(grid), 256, 0, st>>>(d_img_hex, d_opc_uuid, d_ativo, rows, d_image_keys, n_images,
                                                            d_option_keys, k, d_image_idx, d_class_idx, d_active,
                                                            reinterpret_cast<unsigned long long *>(d_unknown));
    B2_LAUNCH_CHECK("encode_label_rows_kernel");
    return B2_OK;
}
