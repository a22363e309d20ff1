// Instruction-throughput microbenchmarks for sm_100a (B200): which pipes the hot kernels' integer
// ops run on and at what rate.  Prints thread-ops per clock per SM for each op.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
enum Op { IADD3, LOP3, SHF, PRMT, IMAD_RRR, IMAD_CONST, IMAD_IMM1, IMAD_WIDE, DP4A, VOTE, POPC, SHFL, ATOMS_SPREAD,
          ATOMS_SAME, MIX_SHF_IMAD, MIX_LOP_DP4A, MIX_PRMT_IMAD, MATCH, LDS32, NOPS };
const char *names[] = {"IADD3", "LOP3", "SHF(funnel rot)", "PRMT", "IMAD r,r,r", "IMAD r,c[],r (x*one+y)", "IMAD.IADD (x*1+y imm)",
                       "IMAD.WIDE.U32", "IDP.4A u8.u8", "VOTE.ballot", "POPC", "SHFL.idx", "ATOMS.ADD spread", "ATOMS.ADD same addr",
                       "mix 1 SHF : 1 IMAD(c)", "mix 1 LOP3 : 1 DP4A", "mix 1 PRMT : 1 IMAD", "MATCH.ANY", "LDS.32 conflict-free"};

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t one, uint32_t seed, long long *cycles) {
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed * (threadIdx.x + 1) + j * 77u;
    uint32_t b = seed ^ 0x9e3779b9u, c = seed + 12345u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[j]) : "r"(b), "r"(c));   // fuses to IADD3
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c));
            if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[j]));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(a[j]) : "r"(b));
            if (OP == IMAD_RRR) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
            if (OP == IMAD_CONST) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(one), "r"(c));
            if (OP == IMAD_IMM1) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(a[j]) : "r"(c));
            if (OP == IMAD_WIDE) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[j]), "r"(b)); a[j] = uint32_t(w) ^ uint32_t(w >> 32); }
            if (OP == DP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[j]) : "r"(b), "r"(c));
            if (OP == VOTE) { uint32_t v; asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "=r"(v) : "r"(a[j] & 1)); a[j] += v; }
            if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(a[j]));
            if (OP == SHFL) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(a[j]) : "r"(b & 31));
            if (OP == ATOMS_SPREAD) atomicAdd(&sm[(threadIdx.x * 1 + j * 256) & 2047], 1u);
            if (OP == ATOMS_SAME) atomicAdd(&sm[(threadIdx.x >> 5) * 32 + j], 1u);
            if (OP == MIX_SHF_IMAD) { asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[j])); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[(j + 4) & 7]) : "r"(one), "r"(c)); }
            if (OP == MIX_LOP_DP4A) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c)); asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[(j + 4) & 7]) : "r"(b), "r"(c)); }
            if (OP == MIX_PRMT_IMAD) { asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(a[j]) : "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[(j + 4) & 7]) : "r"(b), "r"(c)); }
            if (OP == MATCH) { uint32_t v; asm volatile("match.any.sync.b32 %0, %1, 0xffffffff;" : "=r"(v) : "r"(a[j] & 7)); a[j] += v; }
            if (OP == LDS32) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(uint32_t(__cvta_generic_to_shared(&sm[(threadIdx.x + (a[j] & 1) * 32 + j * 256) & 2047])))); a[j] += v; }
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r ^= a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + sm[threadIdx.x];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, int ctas_per_sm, uint32_t *out, long long *cyc) {
    const int grid = sms * ctas_per_sm;
    k<OP><<<grid, 256>>>(out, 1u, 12345u, cyc);
    k<OP><<<grid, 256>>>(out, 1u, 12345u, cyc);
    cudaDeviceSynchronize();
    long long *h = new long long[grid];
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += double(h[i]);
    avg /= grid;
    int per_iter = (OP == MIX_SHF_IMAD || OP == MIX_LOP_DP4A || OP == MIX_PRMT_IMAD) ? 16 : 8;
    if (OP == IADD3) per_iter = 8;
    const double ops = double(ITER) * per_iter * 256 * ctas_per_sm;    // thread-ops per SM
    printf("%-28s ctas/SM=%d  %8.1f thread-ops/clk/SM  (%.2f cyc per warp-instr per SMSP)\n", names[OP], ctas_per_sm, ops / avg,
           avg / (double(ITER) * per_iter * (256 / 32) * ctas_per_sm / 4.0));
    delete[] h;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, sms * 8 * 256 * 4);
    cudaMalloc(&cyc, sms * 8 * 8);
    for (int c : {1, 4}) {
        run<IADD3>(sms, c, out, cyc); run<LOP3>(sms, c, out, cyc); run<SHF>(sms, c, out, cyc); run<PRMT>(sms, c, out, cyc);
        run<IMAD_RRR>(sms, c, out, cyc); run<IMAD_CONST>(sms, c, out, cyc); run<IMAD_IMM1>(sms, c, out, cyc);
        run<IMAD_WIDE>(sms, c, out, cyc); run<DP4A>(sms, c, out, cyc); run<VOTE>(sms, c, out, cyc); run<POPC>(sms, c, out, cyc);
        run<SHFL>(sms, c, out, cyc); run<ATOMS_SPREAD>(sms, c, out, cyc); run<ATOMS_SAME>(sms, c, out, cyc);
        run<MIX_SHF_IMAD>(sms, c, out, cyc); run<MIX_LOP_DP4A>(sms, c, out, cyc); run<MIX_PRMT_IMAD>(sms, c, out, cyc);
        run<MATCH>(sms, c, out, cyc); run<LDS32>(sms, c, out, cyc);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
