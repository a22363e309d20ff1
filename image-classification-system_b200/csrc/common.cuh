// Shared helpers for libb2ingest (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/b2ingest.h"

namespace b2 {

// Thread-local last-error text (b2_last_error()).
char *error_buffer();
int fail(int code, const char *fmt, ...);

#define B2_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return ::b2::fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                              cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

#define B2_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) return ::b2::fail(B2_ERR_BAD_ARG, __VA_ARGS__);                     \
    } while (0)

// Launch-error check that does not synchronise.
#define B2_LAUNCH_CHECK(name)                                                            \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess)                                                           \
            return ::b2::fail(B2_ERR_CUDA, "launch of %s failed: %s", name,              \
                              cudaGetErrorString(_e));                                   \
    } while (0)

int sm_count();   // SMs of the current device (cached per device)

// ---- device-side helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) --------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, size multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- all-reduce (sum, int64) of a small vector over NVLink peer memory, executed by ONE CTA ------------------------
// The tally's partial vector is 8.6 KB; an NCCL all-reduce of that size is pure latency (~30 us at eight ranks, as
// long as the 12.5 M-row tally of config 4 itself).  Every rank owns a "mailbox" in its own HBM — two epochs x world
// slots of `cap` int64 values plus two epochs x world flags — that its peers map through CUDA IPC.  A step:
//   1. write my vector into slot [epoch parity][my rank] of EVERY rank's mailbox (plain stores over NVLink),
//   2. __threadfence_system(), then publish flag [parity][my rank] = epoch on every rank (st.release.sys),
//   3. wait until all `world` flags of my own mailbox show this epoch (ld.acquire.sys; bounded spin),
//   4. sum the `world` slots in rank order — the same integers in the same order on every rank — into the vector.
// Two parities make a second barrier unnecessary: a rank can only start epoch e+2 (and overwrite parity e) after it
// has seen every peer's flag of epoch e+1, which a peer publishes after it finished reading epoch e.
constexpr int kPeerMaxWorld = 16;
struct PeerReduceDesc {
    int world, rank;
    uint32_t cap;                     // int64 values per slot
    uint32_t epoch;                   // last completed epoch (device-resident, advanced by the kernel)
    uint32_t ticket;                  // last-CTA detection of the fused tally; reset by the CTA that takes it
    uint32_t error;                   // 1 = a peer did not arrive within the spin budget
    unsigned long long *mail[kPeerMaxWorld];   // rank q's slots, as mapped in this process
    uint32_t *flags[kPeerMaxWorld];            // rank q's flags
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All threads of ONE CTA call this with the same arguments; `vec` (n <= cap values, global memory, complete and
// visible: the caller fenced) is replaced by the sum over ranks.
static __device__ __noinline__ void peer_allreduce_cta(PeerReduceDesc *d, unsigned long long *vec, uint32_t n) {
    const int world = d->world, rank = d->rank;
    const uint32_t cap = d->cap, e = d->epoch + 1u, par = e & 1u;
    __syncthreads();                                          // everybody has read the epoch before thread 0 advances it
    // my values first (independent loads in flight together), then one store per value and peer
    constexpr int kU = 4;
    for (uint32_t base = threadIdx.x; base < n; base += kU * blockDim.x) {
        unsigned long long v[kU];
#pragma unroll
        for (int j = 0; j < kU; ++j) {
            const uint32_t i = base + j * blockDim.x;
            v[j] = i < n ? ld_relaxed_sys_u64(vec + i) : 0ull;
        }
        for (int q = 0; q < world; ++q) {
            unsigned long long *dst = d->mail[q] + (size_t(par) * world + rank) * cap;
#pragma unroll
            for (int j = 0; j < kU; ++j) {
                const uint32_t i = base + j * blockDim.x;
                if (i < n) dst[i] = v[j];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (int(threadIdx.x) < world) {
        st_release_sys_u32(d->flags[threadIdx.x] + par * world + rank, e);
        const uint32_t *mine = d->flags[rank] + par * world + threadIdx.x;
        const long long t0 = clock64();
        const bool broken = *reinterpret_cast<volatile uint32_t *>(&d->error) != 0u;   // an earlier step timed out: do not wait again
        while (!broken && ld_acquire_sys_u32(mine) != e) {
            if (clock64() - t0 > 6000000000ll) { d->error = 1u; break; }     // ~3 s: a peer died; never hang the GPU
        }
    }
    __syncthreads();
    const unsigned long long *src = d->mail[rank] + size_t(par) * world * cap;
    for (uint32_t base = threadIdx.x; base < n; base += kU * blockDim.x) {
        unsigned long long acc[kU] = {};
        for (int q = 0; q < world; ++q) {                         // rank order: the same sum on every rank
            unsigned long long v[kU];
#pragma unroll
            for (int j = 0; j < kU; ++j) {
                const uint32_t i = base + j * blockDim.x;
                v[j] = i < n ? ld_relaxed_sys_u64(src + size_t(q) * cap + i) : 0ull;
            }
#pragma unroll
            for (int j = 0; j < kU; ++j) acc[j] += v[j];
        }
#pragma unroll
        for (int j = 0; j < kU; ++j) {
            const uint32_t i = base + j * blockDim.x;
            if (i < n) vec[i] = acc[j];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { d->epoch = e; d->ticket = 0u; }
}

}  // namespace b2
