// Batched multi-message SHA-256 (FIPS 180-4) for sm_100a.
//
// Replaces hashlib.sha256(data).hexdigest() of the reference
// (app/services/webdav_sync.py:59, app/services/activity_api_sync.py:798,
//  app/api/routes/images.py:62) for a whole batch of file buffers.
//
// SHA-256 is a serial chain per message (Merkle-Damgard), so the only parallelism is ACROSS
// messages: one lane owns one message and runs the 64-round compression in registers; a
// warp owns 32 messages (sorted by length by the caller so lanes finish together).  The
// kernel is bound by the INT32 ALU pipe (~1400 LOP3/SHF/IADD3 per 64-byte block), not by
// HBM — see DESIGN.md "sha256_lanes".  Batches that leave SM sub-partitions idle run as warp
// pairs instead (sha256_pair_kernel, Path 2 below).
//
// sha256_lanes_kernel: each lane streams its own message with 128-bit loads (four per 64-byte block,
// prefetched one block ahead).  Every 32-byte sector fetched is fully used, so DRAM traffic = message bytes
// (ncu: 19.87 GB read for 19.86 GB of messages).  A variant that staged the warp's 32 messages through shared
// memory with per-lane cp.async.bulk copies measured 760 GB/s against 825 and was removed.
#include "common.cuh"

#include <mutex>

namespace b2 {

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

#define B2_BSIG0(x) (rotr32(x, 2) ^ rotr32(x, 13) ^ rotr32(x, 22))
#define B2_BSIG1(x) (rotr32(x, 6) ^ rotr32(x, 11) ^ rotr32(x, 25))
#define B2_SSIG0(x) (rotr32(x, 7) ^ rotr32(x, 18) ^ ((x) >> 3))
#define B2_SSIG1(x) (rotr32(x, 17) ^ rotr32(x, 19) ^ ((x) >> 10))
#define B2_CH(e, f, g) (((e) & (f)) ^ (~(e) & (g)))
#define B2_MAJ(a, b, c) (((a) & (b)) ^ ((a) & (c)) ^ ((b) & (c)))

// Round constants: with the 64 rounds fully unrolled these become instruction immediates.
__device__ constexpr uint32_t kK[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

struct Sha256State {
    uint32_t h[8];
    __device__ __forceinline__ void init() {
        h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
        h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
    }
};

// Integer add, optionally forced onto the FMA pipe.  ncu on this kernel (profiles/r1_*): the ALU
// pipe (SHF/LOP3/IADD3/PRMT) is 87 % busy and the FMA pipe 4 %.  Every rotate and boolean has to
// stay on the ALU pipe, an add does not.  V = 0 leaves the choice to ptxas (mostly IADD3); V = 2 sends
// the two-input adds to the FMA pipe as IMADs (see sha_variant_override for when that pays).
template <int V>
__device__ __forceinline__ uint32_t addp(uint32_t a, uint32_t b, uint32_t one) {
    if (V == 0) return a + b;
    // `one` is a kernel argument equal to 1: ptxas cannot fold the multiply away (it turns a literal
    // `mad x, 1, y` back into IADD3), so this stays an IMAD on the FMA pipe.
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}
// Variant 2: only the adds that are two-input anyway go to the FMA pipe (one IMAD replaces one IADD, the
// instruction count does not grow); three-input sums stay IADD3 on the ALU pipe.
template <int V>
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b, uint32_t one) {
    return V == 0 ? a + b : addp<1>(a, b, one);
}
template <int V>
__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b, uint32_t c, uint32_t one) {
    if (V == 2) {
        uint32_t d;                                   // kept as one IADD3: ptxas must not re-associate it with the IMADs
        asm("{\n\t.reg .u32 t;\n\tadd.u32 t, %1, %2;\n\tadd.u32 %0, t, %3;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        return d;
    }
    return a + b + c;
}

// One 64-byte block.  w[] holds the 16 big-endian message words and is used as the rolling
// 16-word window of the message schedule (destroyed).
template <int V>
__device__ __forceinline__ void sha256_compress_v(Sha256State &s, uint32_t (&w)[16], uint32_t one) {
    uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3];
    uint32_t e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
#pragma unroll
    for (int t = 0; t < 64; ++t) {
        if (t >= 16) {
            w[t & 15] = add2<V>(add3<V>(w[t & 15], B2_SSIG0(w[(t + 1) & 15]), w[(t + 9) & 15], one),
                                B2_SSIG1(w[(t + 14) & 15]), one);
        }
        // off the critical path: h + K + W (+ d); on it: Sigma1(e) + Ch(e,f,g)
        const uint32_t hkw = add3<V>(h, kK[t], w[t & 15], one);
        const uint32_t t1 = add3<V>(B2_BSIG1(e), B2_CH(e, f, g), hkw, one);
        const uint32_t t2 = add2<V>(B2_BSIG0(a), B2_MAJ(a, b, c), one);
        h = g; g = f; f = e; e = add2<V>(d, t1, one);
        d = c; c = b; b = a; a = add2<V>(t1, t2, one);
    }
    s.h[0] = add2<V>(s.h[0], a, one); s.h[1] = add2<V>(s.h[1], b, one);
    s.h[2] = add2<V>(s.h[2], c, one); s.h[3] = add2<V>(s.h[3], d, one);
    s.h[4] = add2<V>(s.h[4], e, one); s.h[5] = add2<V>(s.h[5], f, one);
    s.h[6] = add2<V>(s.h[6], g, one); s.h[7] = add2<V>(s.h[7], h, one);
}
__device__ __forceinline__ void sha256_compress(Sha256State &s, uint32_t (&w)[16]) { sha256_compress_v<0>(s, w, 1u); }

// Final 1-2 blocks: `rem` (< 64) trailing message bytes at `tail`, then 0x80, zeros and the
// 64-bit big-endian bit length.  Static indexing only (no local memory).
template <typename ByteLoader>
__device__ __forceinline__ void sha256_finish(Sha256State &s, ByteLoader load_byte, uint32_t rem,
                                              uint64_t total_len) {
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t idx = 4 * j + b;
            uint32_t byte = 0;
            if (idx < rem) byte = load_byte(idx);
            else if (idx == rem) byte = 0x80u;
            v = (v << 8) | byte;
        }
        w[j] = v;
    }
    const uint64_t bits = total_len << 3;
    if (rem >= 56) {
        sha256_compress(s, w);
#pragma unroll
        for (int j = 0; j < 14; ++j) w[j] = 0;
    }
    w[14] = static_cast<uint32_t>(bits >> 32);
    w[15] = static_cast<uint32_t>(bits);
    sha256_compress(s, w);
}

__device__ __forceinline__ void sha256_store_digest(const Sha256State &s, uint8_t *out) {
    uint4 lo = make_uint4(bswap32(s.h[0]), bswap32(s.h[1]), bswap32(s.h[2]), bswap32(s.h[3]));
    uint4 hi = make_uint4(bswap32(s.h[4]), bswap32(s.h[5]), bswap32(s.h[6]), bswap32(s.h[7]));
    if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        reinterpret_cast<uint4 *>(out)[0] = lo;
        reinterpret_cast<uint4 *>(out)[1] = hi;
    } else {
        const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) out[4 * j + b] = static_cast<uint8_t>(v[j] >> (8 * b));
    }
}

__device__ __forceinline__ void words_from_v4(uint32_t (&w)[16], const uint4 (&q)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[4 * i + 0] = bswap32(q[i].x); w[4 * i + 1] = bswap32(q[i].y);
        w[4 * i + 2] = bswap32(q[i].z); w[4 * i + 3] = bswap32(q[i].w);
    }
}

// ----------------------------------------------------------------------------------------
// Path 1: lane streams its own message straight from global memory.
// ----------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(128)
sha256_lanes_kernel(const uint8_t *__restrict__ data, const uint64_t *__restrict__ offsets,
                    const uint64_t *__restrict__ lengths, const uint32_t *__restrict__ order,
                    uint32_t n, uint8_t *__restrict__ digests, uint32_t one) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const uint32_t msg = order ? order[slot] : slot;
    const uint8_t *p = data + offsets[msg];
    const uint64_t len = lengths[msg];
    const uint64_t nfull = len >> 6;

    Sha256State s;
    s.init();
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
        uint4 cur[4], nxt[4];
        if (nfull > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[i] = __ldg(q + i);
        }
        for (uint64_t blk = 0; blk < nfull; ++blk) {
            if (blk + 1 < nfull) {        // prefetch the next block while this one compresses
#pragma unroll
                for (int i = 0; i < 4; ++i) nxt[i] = __ldg(q + 4 * (blk + 1) + i);
            }
            uint32_t w[16];
            words_from_v4(w, cur);
            sha256_compress_v<V>(s, w, one);
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
        }
    } else {                               // unaligned start: byte loads (correct, slower)
        for (uint64_t blk = 0; blk < nfull; ++blk) {
            const uint8_t *b = p + (blk << 6);
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                w[j] = (uint32_t(b[4 * j]) << 24) | (uint32_t(b[4 * j + 1]) << 16) |
                       (uint32_t(b[4 * j + 2]) << 8) | uint32_t(b[4 * j + 3]);
            sha256_compress(s, w);
        }
    }
    const uint8_t *tail = p + (nfull << 6);
    sha256_finish(s, [&](uint32_t i) -> uint32_t { return tail[i]; },
                  static_cast<uint32_t>(len & 63), len);
    sha256_store_digest(s, digests + 32ull * msg);
}

// ----------------------------------------------------------------------------------------
// Path 2: a warp PAIR per 32 messages, for batches that leave SM sub-partitions idle.
// One lane hashes ~45 MB/s whatever the batch holds (a lone warp issues one instruction every two clocks and a
// block costs 1 416 of them), so a few large messages take as long as a GPU full of them.  Here the block's work
// is split across two warps that the hardware places on different sub-partitions: warp 0 loads the message,
// byte-swaps it, expands the message schedule and writes W[t] + K[t] for the 64 rounds into shared memory
// (16 STS.128 per block, ~600 instructions); warp 1 reads them (16 LDS.128) and runs the rounds (~990
// instructions) — the longer half, so a message moves ~1.4x faster.  Two block buffers, handed over with the
// named-barrier producer / consumer pattern (bar.arrive by the side that is done, bar.sync by the side that
// waits; 64 participants).  Lanes of a warp can have different lengths: both warps loop to the longest one.
// Only worth it while message warps x 2 fit the sub-partitions (b2_sha256_batch decides).
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" :: "r"(id) : "memory"); }

__global__ void __launch_bounds__(64)
sha256_pair_kernel(const uint8_t *__restrict__ data, const uint64_t *__restrict__ offsets,
                   const uint64_t *__restrict__ lengths, const uint32_t *__restrict__ order,
                   uint32_t n, uint8_t *__restrict__ digests) {
    __shared__ __align__(16) uint4 wk[2][16 * 32];           // [buffer][quad of rounds][lane]: W[t] + K[t], t = 4q .. 4q+3
    const uint32_t lane = threadIdx.x & 31;
    const bool rounds_warp = threadIdx.x >= 32;
    const uint32_t slot = blockIdx.x * 32 + lane;
    const bool live = slot < n;
    const uint32_t msg = live ? (order ? order[slot] : slot) : 0u;
    const uint8_t *p = data + (live ? offsets[msg] : 0ull);
    const uint64_t len = live ? lengths[msg] : 0ull;
    const uint64_t nfull64 = len >> 6;
    const uint32_t nfull = nfull64 > 0xffffffffull ? 0xffffffffu : uint32_t(nfull64);   // < 256 GiB per message
    const uint32_t nmax = __reduce_max_sync(0xffffffffu, nfull);
    constexpr int kFull = 1, kEmpty = 3;                      // named barriers 1, 2 (full) and 3, 4 (empty); 0 = __syncthreads

    if (!rounds_warp) {
        // ---- schedule warp
        const bool al = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
        uint4 cur[4] = {}, nxt[4] = {};
        if (al && nfull > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[i] = __ldg(q + i);
        }
        for (uint32_t blk = 0; blk < nmax; ++blk) {
            const int buf = blk & 1;
            if (blk >= 2) named_bar_sync(kEmpty + buf);       // the rounds warp has read this buffer's previous block
            if (blk < nfull) {
                uint32_t w[16];
                if (al) {
                    // This warp needs ~1 200 clocks per block, less than a trip to HBM: besides the register prefetch of
                    // the next block, pull the line 8 blocks ahead into L2 (one prefetch per 128-byte line).
                    if ((blk & 1) == 0 && blk + 8 < nfull)
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(q + 4 * (uint64_t(blk) + 8)));
                    if (blk + 1 < nfull) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) nxt[i] = __ldg(q + 4 * (uint64_t(blk) + 1) + i);
                    }
                    words_from_v4(w, cur);
#pragma unroll
                    for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
                } else {                                      // unaligned start: byte loads (correct, slower)
                    const uint8_t *b = p + (uint64_t(blk) << 6);
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        w[j] = (uint32_t(b[4 * j]) << 24) | (uint32_t(b[4 * j + 1]) << 16) |
                               (uint32_t(b[4 * j + 2]) << 8) | uint32_t(b[4 * j + 3]);
                }
                uint4 *dst = &wk[buf][lane];
#pragma unroll
                for (int t4 = 0; t4 < 16; ++t4) {
                    uint32_t o[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = 4 * t4 + u;
                        if (t >= 16)
                            w[t & 15] = w[t & 15] + B2_SSIG0(w[(t + 1) & 15]) + w[(t + 9) & 15] + B2_SSIG1(w[(t + 14) & 15]);
                        o[u] = w[t & 15] + kK[t];
                    }
                    dst[t4 * 32] = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            named_bar_arrive(kFull + buf);
        }
        return;
    }

    // ---- rounds warp
    Sha256State s;
    s.init();
    for (uint32_t blk = 0; blk < nmax; ++blk) {
        const int buf = blk & 1;
        named_bar_sync(kFull + buf);
        if (blk < nfull) {
            const uint4 *src = &wk[buf][lane];
            uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3];
            uint32_t e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
#pragma unroll
            for (int t4 = 0; t4 < 16; ++t4) {
                const uint4 x = src[t4 * 32];
                const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t hkw = h + xs[u];
                    const uint32_t t1 = B2_BSIG1(e) + B2_CH(e, f, g) + hkw;
                    const uint32_t t2 = B2_BSIG0(a) + B2_MAJ(a, b, c);
                    h = g; g = f; f = e; e = d + t1;
                    d = c; c = b; b = a; a = t1 + t2;
                }
            }
            s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d;
            s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
        }
        if (blk + 2 < nmax) named_bar_arrive(kEmpty + buf);   // only when the schedule warp will wait for it
    }
    if (!live) return;
    const uint8_t *tail = p + (nfull64 << 6);
    sha256_finish(s, [&](uint32_t i) -> uint32_t { return tail[i]; }, static_cast<uint32_t>(len & 63), len);
    sha256_store_digest(s, digests + 32ull * msg);
}

__global__ void digest_hex_kernel(const uint8_t *__restrict__ digests, uint32_t n, char *__restrict__ hex) {
    // one thread per digest byte-pair word: thread t handles 4 digest bytes -> 8 hex chars
    const uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (t >= uint64_t(n) * 8) return;
    const uint32_t v = reinterpret_cast<const uint32_t *>(digests)[t];
    uint32_t out[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t packed = 0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const uint32_t byte = (v >> (8 * (2 * half + b))) & 0xff;
            const uint32_t hi = byte >> 4, lo = byte & 15;
            const uint32_t chi = hi < 10 ? '0' + hi : 'a' + hi - 10;
            const uint32_t clo = lo < 10 ? '0' + lo : 'a' + lo - 10;
            packed |= (chi | (clo << 8)) << (16 * b);
        }
        out[half] = packed;
    }
    reinterpret_cast<uint2 *>(hex)[t] = make_uint2(out[0], out[1]);
}

}  // namespace b2

// B2_SHA_VARIANT: unset = auto, 0 = adds left to ptxas (IADD3), 2 = two-input adds forced onto the FMA pipe.
// Measured on B200 (GB/s, 256 KiB messages; 1 / 2 / 4 warps per SM sub-partition): variant 0 819 / 850 / 861,
// variant 2 794 / 884 / 899, and a variant with EVERY add as IMAD 734 / 852 / 870.  A lone warp issues one
// instruction every ~2 cycles whatever the pipe, so with one warp per sub-partition (all that 1080p images
// leave room for in HBM) only the instruction count matters; from two warps up the IMADs pay.
static int sha_variant_override() {
    const char *e = getenv("B2_SHA_VARIANT");
    return e ? atoi(e) : -1;
}

extern "C" int b2_sha256_batch(const uint8_t *d_data, const uint64_t *d_offsets,
                               const uint64_t *d_lengths, const uint32_t *d_order, uint32_t n,
                               uint8_t *d_digests, void *stream) {
    using namespace b2;
    if (n == 0) return B2_OK;
    B2_REQUIRE(d_data && d_offsets && d_lengths && d_digests, "b2_sha256_batch: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Spread warps over SM sub-partitions: small batches use one warp per CTA so the block
    // scheduler places consecutive warps on different SMs.
    const uint32_t warps = (n + 31) / 32;
    const int block = warps <= uint32_t(sm_count()) * 16u ? 32 : 128;
    const uint32_t grid = (n + block - 1) / block;
    // Ask for the shared-memory-heavy L1 split, like the resize kernel: an SM only changes its carve-out
    // when it is idle, so kernels with different preferences never share an SM and a hash launched beside
    // a resize on another stream would simply wait for it (measured: no overlap at all without this).
    static std::once_flag carve_once[64], pair_once[64];     // function attributes are per device
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    std::call_once(carve_once[cur_dev & 63], [] {
        cudaFuncSetAttribute(sha256_lanes_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(sha256_lanes_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    });
    // Few message warps: the warp-pair kernel (two sub-partitions per 32 messages, ~1.4x the per-message speed).
    // B2_SHA_PAIR = 0 / 1 forces the choice (tests, comparisons).
    bool pair = warps * 2u <= uint32_t(sm_count()) * 4u;
    if (const char *e = getenv("B2_SHA_PAIR")) pair = atoi(e) != 0;
    if (pair) {
        std::call_once(pair_once[cur_dev & 63], [] {
            cudaFuncSetAttribute(sha256_pair_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        });
        sha256_pair_kernel<<<warps, 64, 0, st>>>(d_data, d_offsets, d_lengths, d_order, n, d_digests);
        B2_LAUNCH_CHECK("sha256_pair_kernel");
        return B2_OK;
    }
    int variant = sha_variant_override();
    if (variant < 0) variant = warps > uint32_t(sm_count()) * 4u ? 2 : 0;    // more than one warp per sub-partition
    if (variant == 0)
        sha256_lanes_kernel<0><<<grid, block, 0, st>>>(d_data, d_offsets, d_lengths, d_order, n, d_digests, 1u);
    else
        sha256_lanes_kernel<2><<<grid, block, 0, st>>>(d_data, d_offsets, d_lengths, d_order, n, d_digests, 1u);
    B2_LAUNCH_CHECK("sha256_lanes_kernel");
    return B2_OK;
}

extern "C" int b2_digest_hex(const uint8_t *d_digests, uint32_t n, char *d_hex, void *stream) {
    using namespace b2;
    if (n == 0) return B2_OK;
    B2_REQUIRE(d_digests && d_hex, "b2_digest_hex: null pointer");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(d_digests) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_hex) & 7) == 0,
               "b2_digest_hex: digests must be 4-byte and hex 8-byte aligned");
    const uint64_t threads = uint64_t(n) * 8;
    digest_hex_kernel<<<unsigned((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_digests, n, d_hex);
    B2_LAUNCH_CHECK("digest_hex_kernel");
    return B2_OK;
}
