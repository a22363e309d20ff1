// Separable antialiased BILINEAR resize + normalise (thumbnail / preview tensor), sm_100a.
//
// The reference has no resize; BASELINE.json configs 1,2,3,5 require
// "hash + 256x256 thumbnail".  Semantics = Pillow's 8-bit two-pass resampler (the reference's
// pinned image library, requirements.txt:9), reproduced bit-exactly:
//   per axis: scale = in/out; support = max(scale,1); ksize = 2*ceil(support)+1;
//   triangle weights in float64 normalised by their sequential sum, quantised to 22-bit
//   fixed point (int(0.5 + w * 2^22)); out = clip8((2^21 + sum px*k) >> 22);
//   HORIZONTAL pass first into a uint8 intermediate, then the VERTICAL pass.
//
// resize_bands_kernel<KQ, kVMode> — the product kernel.  One CTA = one band of output rows of one
// image (the whole image when the batch fills the GPU).  The band's input rows are a single contiguous byte
// range of the HWC image; the bulk-copy (TMA) engine streams it through a ring of shared-memory stages
// (cp.async.bulk + an mbarrier per stage), so global memory is read once, in large aligned bursts, with no
// register staging.
//   Horizontal pass ("quads"): thread x owns output column x.  Per input row it pulls its byte-aligned window
// from shared memory as 32-bit words (funnel-shifted), de-interleaves it IN REGISTERS — four pixels = three
// words -> one R, one G and one B word, two PRMT each — and every plane word meets three coefficient-limb
// words (a 22-bit tap = three 8-bit limbs) in three IDP.4A: 6 PRMT + 9 IDP.4A per four taps and three
// channels.  KQ is the tap capacity, a multiple of 4 chosen from the widest window the plan has (16 for
// 1080p -> 256).  Any input width and alignment; taps are non-negative (BILINEAR).
//   Vertical pass, scatter form (kVMode 2, every downscale): never more than three output rows have tap
// windows over one input row, and they finish in order.  The thread keeps three rows of accumulators; the
// horizontal result of an input row, still in registers, is multiplied into them with that row's three taps
// (one uniform 16-byte load from a per-plan table); when the table says a row is complete it is written out
// (uint8 HWC thumbnail + float32 CHW preview).  No intermediate in shared memory: three CTAs per SM.
//   Vertical pass, gather form (kVMode 0 / 1, upscales): the horizontal results go to a rolling,
// thread-private shared-memory intermediate and finished output rows gather their taps from it; the tap
// table travels in the kernel parameters when it fits (kVMode 1), else it is staged in shared memory.
// No tensor cores: 2 MAC per input byte, nothing to contract.
//
// resize_generic_kernel — any shape / any alignment, one thread per output pixel, straight
// from global memory.  Used when the fast path's limits do not hold, and by the tests as an
// independent device implementation.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace b2 {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kRound = 1 << (kPrecisionBits - 1);

// ---------------------------------------------------------------------------------------
// Host: tap tables, identical arithmetic (float64, same operation order) to Pillow.
// ---------------------------------------------------------------------------------------
struct AxisTaps {
    int ksize = 0;
    std::vector<int32_t> bounds;   // out x {first, count}
    std::vector<int32_t> coeffs;   // out x ksize
};

static AxisTaps make_taps(int in_size, int out_size) {
    AxisTaps t;
    const double scale = double(in_size) / double(out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    t.ksize = int(std::ceil(support)) * 2 + 1;
    t.bounds.assign(size_t(out_size) * 2, 0);
    t.coeffs.assign(size_t(out_size) * t.ksize, 0);
    std::vector<double> w(t.ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = int(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = int(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        const int n = xmax - xmin;
        double ww = 0.0;
        for (int x = 0; x < n; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            const double v = a < 1.0 ? 1.0 - a : 0.0;
            w[x] = v;
            ww += v;
        }
        for (int x = 0; x < n; ++x) {
            double v = w[x];
            if (ww != 0.0) v /= ww;
            t.coeffs[size_t(xx) * t.ksize + x] =
                v < 0 ? int(-0.5 + v * (1 << kPrecisionBits)) : int(0.5 + v * (1 << kPrecisionBits));
        }
        t.bounds[2 * xx] = xmin;
        t.bounds[2 * xx + 1] = n;
    }
    return t;
}

}  // namespace b2

struct b2_resize_plan {
    int in_h, in_w, out_h, out_w;
    b2::AxisTaps h, v;
    int device;
    int32_t *d_hbounds, *d_hcoeffs, *d_vbounds, *d_vcoeffs;
    // fast-path geometry
    int threads;
    int q_bucket;         // tap capacity (multiple of 4) covering the widest horizontal window; 0 = generic kernel only
    int n_stages;
    // gather-form vertical pass
    int vparam;           // 1 = the vertical tap table fits the kernel parameters (out_h * (2 + ksize_v) <= kVtabInts)
    int tmp_ring_rows;    // rows of the rolling intermediate (power of two)
    int tmp_pitch;        // bytes per intermediate row (multiple of 16)
    int rows_per_stage;
    int stage_bytes;      // bytes of one ring stage (incl. alignment + over-read padding)
    size_t smem_fixed;    // ring + intermediate (+ slack); staged vertical tap tables add band_rows*(2+ksize_v)*4
    size_t smem_max;      // with band_rows = out_h
    // scatter-form vertical pass: per INPUT row the taps of the <= 3 output rows whose windows cover it
    int scat;             // 1 = available (downscale: never more than 3 output rows per input row, one finishing at a time)
    int4 *d_vscat;        // in_h x {tap of the row in accumulator 0, 1, 2 (row oy lives in accumulator oy % 3),
                          //         (oy << 2 | accumulator + 1) of the output row that ENDS with this input row, else 0}
    int rps_scat, stage_bytes_scat;
    size_t smem_scat;
    // pair kernel (two output columns per thread): quad offset / quad count of the odd column inside the union window
    int pair;             // 1 = an instantiation of resize_pairs_kernel fits this plan
    int pair_dq, pair_nqb, pair_threads, rps_pair, stage_bytes_pair;
    size_t smem_pair;
};

namespace b2 {

constexpr int kMaxStages = 4;
constexpr int kVtabInts = 3456;              // 256 output rows x (2 + 11 taps) = 3328 fit; 13.5 KB of the 32 KB parameter space

struct ResizeParams {
    const uint8_t *rgb;
    const uint64_t *offsets;
    const uint32_t *out_slot;
    uint8_t *thumb;
    float *preview;
    const int32_t *hbounds, *hcoeffs, *vbounds, *vcoeffs;
    const int4 *vscat;
    int in_h, in_w, out_h, out_w;
    int ksize_h, ksize_v;
    int band_rows, n_bands;
    int rows_per_stage, stage_bytes, n_stages, tmp_pitch, tmp_ring_rows;
    float mean[3], inv_std[3];
    // kVMode 1: the vertical taps of every output row, {first, count, k_0 .. k_{ksize_v-1}} per row, travel
    // in the kernel parameters (constant bank, uniform loads: no shared-memory traffic, no staging at CTA start).
    int32_t vtab[kVtabInts];
};

__device__ __forceinline__ uint32_t clip8(int32_t acc) {
    const int32_t v = acc >> kPrecisionBits;
    return uint32_t(min(max(v, 0), 255));
}

// (u8 * (1/255) - mean) * inv_std as three separately rounded float32 operations (no FMA
// contraction), so the preview is bit-identical to the NumPy oracle.
__device__ __forceinline__ float normalise(uint32_t u8, float mean, float inv_std) {
    return __fmul_rn(__fsub_rn(__fmul_rn(float(u8), 1.0f / 255.0f), mean), inv_std);
}

// BILINEAR taps are non-negative and sum to 2^22 +- ksize/2, so (2^21 + sum(px*k)) >> 22 is already inside
// 0..255: the band kernel needs no clip (the generic kernel keeps Pillow's).
__device__ __forceinline__ uint32_t to_u8(uint32_t acc) { return acc >> kPrecisionBits; }

template <int KQ, int kVMode>
__global__ void __launch_bounds__(256)
resize_bands_kernel(const __grid_constant__ ResizeParams p) {
    constexpr bool kVScat = kVMode == 2;
    constexpr bool kVParam = kVMode == 1;
    constexpr int NQ = KQ / 4;                  // 4-pixel groups of the window
    static_assert(KQ % 4 == 0, "tap capacity is a whole number of quads");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];

    const int tid = threadIdx.x;
    const int img = blockIdx.x / p.n_bands;
    const int band = blockIdx.x - img * p.n_bands;
    const int oy0 = band * p.band_rows;
    const int oy1 = min(oy0 + p.band_rows, p.out_h);
    const int r0 = p.vbounds[2 * oy0];
    const int r1 = p.vbounds[2 * (oy1 - 1)] + p.vbounds[2 * (oy1 - 1) + 1];
    const int n_rows = r1 - r0;
    const int pitch = p.in_w * 3;
    const uint64_t img_off = p.offsets[img];
    const uint8_t *img_ptr = p.rgb + img_off;
    const uint64_t img_bytes = uint64_t(p.in_h) * pitch;
    const bool aligned = ((reinterpret_cast<uintptr_t>(img_ptr)) & 15) == 0;

    uint8_t *ring = smem;                                            // n_stages * stage_bytes
    // gather form: rolling intermediate of tmp_ring_rows rows, one packed RGBX word per output column; column x is
    // written and read by thread x only
    uint32_t *tmp = reinterpret_cast<uint32_t *>(smem + size_t(p.n_stages) * p.stage_bytes);
    const int tmp_pitch_w = p.tmp_pitch >> 2;
    int32_t *vb_s = reinterpret_cast<int32_t *>(smem + size_t(p.n_stages) * p.stage_bytes +
                                                size_t(p.tmp_ring_rows) * p.tmp_pitch);            // band_rows*2 (kVMode 0)
    int32_t *vk_s = vb_s + 2 * p.band_rows;                          // band_rows * ksize_v
    const int tmp_mask = p.tmp_ring_rows - 1;                        // power of two

    if (tid == 0) {
        for (int s = 0; s < p.n_stages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    if (kVMode == 0) {                 // vertical taps of this band -> shared memory
        for (int i = tid; i < (oy1 - oy0) * 2; i += blockDim.x) vb_s[i] = p.vbounds[2 * oy0 + i];
        for (int i = tid; i < (oy1 - oy0) * p.ksize_v; i += blockDim.x) vk_s[i] = p.vcoeffs[oy0 * p.ksize_v + i];
    }

    // horizontal taps of my column -> registers: limb i of the taps under pixels 4g .. 4g+3 of my window
    const bool col_active = tid < p.out_w;
    const int xmin = col_active ? p.hbounds[2 * tid] : 0;
    uint32_t kq[3][NQ];
#pragma unroll
    for (int g = 0; g < NQ; ++g) {
        uint32_t l0 = 0, l1 = 0, l2 = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int t = 4 * g + b;
            const uint32_t k = (col_active && t < p.ksize_h) ? uint32_t(p.hcoeffs[tid * p.ksize_h + t]) : 0u;
            l0 |= (k & 0xffu) << (8 * b);
            l1 |= ((k >> 8) & 0xffu) << (8 * b);
            l2 |= ((k >> 16) & 0xffu) << (8 * b);
        }
        kq[0][g] = l0; kq[1][g] = l1; kq[2][g] = l2;
    }
    __syncthreads();

    const int total_stages = (n_rows + p.rows_per_stage - 1) / p.rows_per_stage;

    // Stage s covers input rows [r0 + s*RPS, min(r0 + (s+1)*RPS, r1)) = bytes [b0, b1) of the
    // image.  The copy engine moves the enclosing 16-byte aligned range [a0, a1), clamped to
    // the last whole 16-byte unit of the image; a ragged tail (< 16 B, last rows of the image
    // only) is patched with ordinary loads.
    auto stage_range = [&](int s, uint64_t &b0, uint64_t &b1) {
        const int ra = r0 + s * p.rows_per_stage;
        const int rb = min(ra + p.rows_per_stage, r1);
        b0 = uint64_t(ra) * pitch;
        b1 = uint64_t(rb) * pitch;
    };
    auto issue = [&](int s, int buf) { // thread 0 only, aligned images only
        uint64_t b0, b1;
        stage_range(s, b0, b1);
        const uint64_t a0 = b0 & ~uint64_t(15);
        uint64_t a1 = (b1 + 15) & ~uint64_t(15);
        const uint64_t lim = img_bytes & ~uint64_t(15);
        if (a1 > lim) a1 = lim;
        const uint32_t bytes = a1 > a0 ? uint32_t(a1 - a0) : 0u;
        if (bytes) {
            mbar_arrive_expect_tx(&full_bar[buf], bytes);
            bulk_g2s(ring + size_t(buf) * p.stage_bytes, img_ptr + a0, bytes, &full_bar[buf]);
        } else {
            mbar_arrive(&full_bar[buf]);
        }
    };

    if (aligned && tid == 0) {
        for (int s = 0; s < p.n_stages - 1 && s < total_stages; ++s) issue(s, s);
    }

    const int row_bytes = p.out_w * 3;
    const uint32_t slot = p.out_slot ? p.out_slot[img] : uint32_t(img);
    uint8_t *thumb_px = p.thumb + uint64_t(slot) * p.out_h * row_bytes + 3 * tid;                     // my column, row 0
    float *prev_px = p.preview ? p.preview + uint64_t(slot) * 3 * p.out_h * p.out_w + tid : nullptr;
    const uint32_t plane = uint32_t(p.out_h) * uint32_t(p.out_w);
    const float mean0 = p.mean[0], mean1 = p.mean[1], mean2 = p.mean[2];
    const float istd0 = p.inv_std[0], istd1 = p.inv_std[1], istd2 = p.inv_std[2];
    // One thread per output PIXEL: the float32 CHW preview is written with fully coalesced 128-byte stores.
    auto write_pixel = [&](int oy, uint32_t o0, uint32_t o1, uint32_t o2) {
        uint8_t *tpx = thumb_px + uint32_t(oy) * uint32_t(row_bytes);
        tpx[0] = uint8_t(o0); tpx[1] = uint8_t(o1); tpx[2] = uint8_t(o2);
        if (prev_px) {
            float *pp = prev_px + uint32_t(oy) * uint32_t(p.out_w);
            pp[0] = normalise(o0, mean0, istd0);
            pp[plane] = normalise(o1, mean1, istd1);
            pp[2 * plane] = normalise(o2, mean2, istd2);
        }
    };

    // ---- gather-form vertical pass, run incrementally: after every stage the output rows whose taps are all in
    // the rolling intermediate are finished and written.  The intermediate is a ring of tmp_ring_rows
    // (>= rows_per_stage + ksize_v + 2) rows, so a band can be the whole image: no input row is read or
    // filtered twice.
    int next_oy = oy0;                                               // first output row not yet written
    const int vstride = 2 + p.ksize_v;
    auto v_first = [&](int oy) { return kVParam ? p.vtab[oy * vstride] : vb_s[2 * (oy - oy0)]; };
    auto v_count = [&](int oy) { return kVParam ? p.vtab[oy * vstride + 1] : vb_s[2 * (oy - oy0) + 1]; };
    auto v_coeff = [&](int oy, int t) { return kVParam ? p.vtab[oy * vstride + 2 + t] : vk_s[(oy - oy0) * p.ksize_v + t]; };
    auto emit_ready = [&](int rows_done) {
        int e1 = next_oy;
        while (e1 < oy1 && v_first(e1) + v_count(e1) <= rows_done) ++e1;
        if (col_active) {
            for (int oy = next_oy; oy < e1; ++oy) {
                const int cnt = v_count(oy);
                uint32_t a0 = kRound, a1 = kRound, a2 = kRound;
                const uint32_t *col = tmp + tid;
                int rr = v_first(oy) - r0;
                for (int t = 0; t < cnt; ++t, ++rr) {
                    const uint32_t px = col[size_t(rr & tmp_mask) * tmp_pitch_w];     // one RGBX word per tap
                    const uint32_t k = uint32_t(v_coeff(oy, t));                       // uniform: constant bank or broadcast
                    a0 += (px & 0xffu) * k;
                    a1 += __byte_perm(px, 0, 0x4441) * k;
                    a2 += __byte_perm(px, 0, 0x4442) * k;
                }
                write_pixel(oy, to_u8(a0), to_u8(a1), to_u8(a2));
            }
        }
        next_oy = e1;
    };

    // ---- scatter-form vertical pass: three output rows in flight.  The accumulators are never reset (a reset on the
    // rarely taken "row complete" path costs nine register moves on the hot one): they run on modulo 2^32 and
    // vb remembers where the row in that accumulator started.
    uint32_t va[3][3], vb[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) { va[j][c] = uint32_t(kRound); vb[j][c] = 0u; }

    // ring position of stage s (buf), of the stage issued this iteration (ibuf) and the parities of their
    // barriers are carried along instead of being recomputed with % and / every stage
    int buf = 0, ibuf = p.n_stages - 1;
    uint32_t parity = 0;
    for (int s = 0; s < total_stages; ++s) {
        uint8_t *sbuf = ring + size_t(buf) * p.stage_bytes;
        uint64_t b0, b1;
        stage_range(s, b0, b1);
        const uint64_t a0 = b0 & ~uint64_t(15);
        if (aligned) {
            // slot ibuf held stage s - 1, which every warp left behind at the barrier that ended the last iteration
            if (tid == 0 && s + p.n_stages - 1 < total_stages) issue(s + p.n_stages - 1, ibuf);
            mbar_wait(&full_bar[buf], parity);
            const uint64_t lim = img_bytes & ~uint64_t(15);
            if (b1 > lim) {            // ragged image tail: uniform branch, last stage of last band
                for (uint64_t b = max(lim, a0) + tid; b < b1; b += blockDim.x) sbuf[b - a0] = img_ptr[b];
                __syncthreads();
            }
        } else {                       // misaligned image: cooperative copy, same smem layout
            for (uint64_t b = b0 + tid; b < b1; b += blockDim.x) sbuf[b - a0] = img_ptr[b];
            __syncthreads();
        }

        const int ra = r0 + s * p.rows_per_stage;
        const int rb = min(ra + p.rows_per_stage, r1);
        if (col_active) {
            [[maybe_unused]] const int4 *vsp = p.vscat + ra;
            // byte address (in shared memory) of my first source byte of row ra; rows are `pitch` apart
            uint32_t src = smem_u32(sbuf) + uint32_t(uint64_t(ra) * pitch - a0) + 3u * xmin;
            for (int r = ra; r < rb; ++r, src += uint32_t(pitch)) {
                const uint32_t base = src & ~3u;
                const uint32_t sh = src << 3;                          // the funnel shift takes it modulo 32
                uint32_t w[3 * NQ + 1];
#pragma unroll
                for (int j = 0; j < 3 * NQ + 1; ++j)
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[j]) : "r"(base + 4u * j));
                uint32_t acc[3][3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { acc[c][0] = uint32_t(kRound); acc[c][1] = 0; acc[c][2] = 0; }
#pragma unroll
                for (int g = 0; g < NQ; ++g) {
                    const uint32_t w0 = __funnelshift_r(w[3 * g], w[3 * g + 1], sh);
                    const uint32_t w1 = __funnelshift_r(w[3 * g + 1], w[3 * g + 2], sh);
                    const uint32_t w2 = __funnelshift_r(w[3 * g + 2], w[3 * g + 3], sh);
                    uint32_t px[3];
                    px[0] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);     // R of pixels 4g .. 4g+3
                    px[1] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);     // G
                    px[2] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);     // B
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        acc[c][0] = __dp4a(px[c], kq[0][g], acc[c][0]);
                        acc[c][1] = __dp4a(px[c], kq[1][g], acc[c][1]);
                        acc[c][2] = __dp4a(px[c], kq[2][g], acc[c][2]);
                    }
                }
                uint32_t out[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) out[c] = to_u8(acc[c][0] + (acc[c][1] << 8) + (acc[c][2] << 16));
                if constexpr (kVScat) {
                    const int4 e = __ldg(vsp++);                       // same address in every thread
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        va[0][c] += out[c] * uint32_t(e.x);
                        va[1][c] += out[c] * uint32_t(e.y);
                        va[2][c] += out[c] * uint32_t(e.z);
                    }
                    if (e.w) {                                         // uniform: an output row ends with this input row
                        const int oy = e.w >> 2;
                        const bool mine = oy >= oy0 && oy < oy1;       // rows of a neighbouring band pass through unwritten
                        auto finish = [&](uint32_t (&a)[3], uint32_t (&b)[3]) {
                            if (mine) write_pixel(oy, to_u8(a[0] - b[0]), to_u8(a[1] - b[1]), to_u8(a[2] - b[2]));
#pragma unroll
                            for (int c = 0; c < 3; ++c) b[c] = a[c] - uint32_t(kRound);
                        };
                        const int sl = e.w & 3;                        // which accumulator row: (oy % 3) + 1
                        if (sl == 1) finish(va[0], vb[0]);
                        else if (sl == 2) finish(va[1], vb[1]);
                        else finish(va[2], vb[2]);
                    }
                } else {
                    tmp[size_t((r - r0) & tmp_mask) * tmp_pitch_w + tid] = out[0] | (out[1] << 8) | (out[2] << 16);
                }
            }
        }
        // Stage consumed: the ring slot may be refilled.  (Per-warp releases on an "empty" mbarrier with a dedicated
        // producer warp were measured instead of this barrier: 3.55 ms against 3.34 ms, profiles/r1_resize_notes.md.)
        __syncthreads();
        if constexpr (!kVScat) emit_ready(rb);
        if (++buf == p.n_stages) { buf = 0; parity ^= 1u; }
        if (++ibuf == p.n_stages) ibuf = 0;
    }
}

// ---------------------------------------------------------------------------------------
// resize_pairs_kernel<KQ, DQ, NQB> — scatter form with TWO adjacent output columns per thread.
// The windows of neighbouring columns overlap by half (window = 2 x scale, step = scale), so a thread that owns
// columns 2t and 2t+1 loads, aligns and de-interleaves the UNION of their windows once: for 1080p -> 256 (KQ = 16
// taps, columns 7 or 8 pixels apart) 19 LDS + 18 SHF + 36 PRMT instead of 2 x (13 + 12 + 24).  Column A's taps sit
// under quads 0 .. KQ/4-1 of the union, column B's under quads DQ .. DQ+NQB-1 (its coefficient limbs are placed at
// (xmin_B - xmin_A) - 4 DQ inside that range, zero elsewhere): 9 IDP.4A per quad and column.  Everything else —
// stage ring, scatter-form vertical pass, never-reset accumulators — is the band kernel's.  Half the threads per CTA
// (out_w / 2), so stages are shorter and more CTAs share an SM.
// ---------------------------------------------------------------------------------------
template <int KQ, int DQ, int NQB>
__global__ void __launch_bounds__(128, (KQ <= 20 ? 4 : 2))
resize_pairs_kernel(const __grid_constant__ ResizeParams p) {
    constexpr int NQ = KQ / 4;
    constexpr int NU = (DQ + NQB > NQ) ? DQ + NQB : NQ;              // quads of the union window
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];

    const int tid = threadIdx.x;
    const int img = blockIdx.x / p.n_bands;
    const int band = blockIdx.x - img * p.n_bands;
    const int oy0 = band * p.band_rows;
    const int oy1 = min(oy0 + p.band_rows, p.out_h);
    const int r0 = p.vbounds[2 * oy0];
    const int r1 = p.vbounds[2 * (oy1 - 1)] + p.vbounds[2 * (oy1 - 1) + 1];
    const int n_rows = r1 - r0;
    const int pitch = p.in_w * 3;
    const uint8_t *img_ptr = p.rgb + p.offsets[img];
    const uint64_t img_bytes = uint64_t(p.in_h) * pitch;
    const bool aligned = ((reinterpret_cast<uintptr_t>(img_ptr)) & 15) == 0;
    uint8_t *ring = smem;

    if (tid == 0) {
        for (int s = 0; s < p.n_stages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }

    // my two columns (2 tid, 2 tid + 1); a thread past the last pair computes a copy of the last pair, unwritten
    const int n_pairs = p.out_w >> 1;
    const bool active = tid < n_pairs;
    const int ca = 2 * (active ? tid : n_pairs - 1), cb = ca + 1;
    const int xmin = p.hbounds[2 * ca];
    const int ob = p.hbounds[2 * cb] - xmin - 4 * DQ;                // first tap of B inside its quad range (>= 0: the plan checked)
    uint32_t ka[3][NQ], kb[3][NQB];
#pragma unroll
    for (int g = 0; g < NQ; ++g) {
        uint32_t l0 = 0, l1 = 0, l2 = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int t = 4 * g + b;
            const uint32_t k = t < p.ksize_h ? uint32_t(p.hcoeffs[ca * p.ksize_h + t]) : 0u;
            l0 |= (k & 0xffu) << (8 * b); l1 |= ((k >> 8) & 0xffu) << (8 * b); l2 |= ((k >> 16) & 0xffu) << (8 * b);
        }
        ka[0][g] = l0; ka[1][g] = l1; ka[2][g] = l2;
    }
#pragma unroll
    for (int g = 0; g < NQB; ++g) {
        uint32_t l0 = 0, l1 = 0, l2 = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int t = 4 * g + b - ob;
            const uint32_t k = (t >= 0 && t < p.ksize_h) ? uint32_t(p.hcoeffs[cb * p.ksize_h + t]) : 0u;
            l0 |= (k & 0xffu) << (8 * b); l1 |= ((k >> 8) & 0xffu) << (8 * b); l2 |= ((k >> 16) & 0xffu) << (8 * b);
        }
        kb[0][g] = l0; kb[1][g] = l1; kb[2][g] = l2;
    }
    __syncthreads();

    const int total_stages = (n_rows + p.rows_per_stage - 1) / p.rows_per_stage;
    auto stage_range = [&](int s, uint64_t &b0, uint64_t &b1) {
        const int ra = r0 + s * p.rows_per_stage;
        const int rb = min(ra + p.rows_per_stage, r1);
        b0 = uint64_t(ra) * pitch;
        b1 = uint64_t(rb) * pitch;
    };
    auto issue = [&](int s, int buf) { // thread 0 only, aligned images only
        uint64_t b0, b1;
        stage_range(s, b0, b1);
        const uint64_t a0 = b0 & ~uint64_t(15);
        uint64_t a1 = (b1 + 15) & ~uint64_t(15);
        const uint64_t lim = img_bytes & ~uint64_t(15);
        if (a1 > lim) a1 = lim;
        const uint32_t bytes = a1 > a0 ? uint32_t(a1 - a0) : 0u;
        if (bytes) {
            mbar_arrive_expect_tx(&full_bar[buf], bytes);
            bulk_g2s(ring + size_t(buf) * p.stage_bytes, img_ptr + a0, bytes, &full_bar[buf]);
        } else {
            mbar_arrive(&full_bar[buf]);
        }
    };
    if (aligned && tid == 0) {
        for (int s = 0; s < p.n_stages - 1 && s < total_stages; ++s) issue(s, s);
    }

    const int row_bytes = p.out_w * 3;
    const uint32_t slot = p.out_slot ? p.out_slot[img] : uint32_t(img);
    const uint32_t plane = uint32_t(p.out_h) * uint32_t(p.out_w);
    const float mean0 = p.mean[0], mean1 = p.mean[1], mean2 = p.mean[2];
    const float istd0 = p.inv_std[0], istd1 = p.inv_std[1], istd2 = p.inv_std[2];
    // two adjacent pixels: 6 thumbnail bytes (2-byte aligned) and, per plane, two floats (8-byte aligned)
    uint8_t *thumb_px = p.thumb + uint64_t(slot) * p.out_h * row_bytes + 3 * ca;
    float *prev_px = p.preview ? p.preview + uint64_t(slot) * 3 * p.out_h * p.out_w + ca : nullptr;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(thumb_px) | uint32_t(row_bytes)) & 1u) == 0u &&
                        (!prev_px || ((reinterpret_cast<uintptr_t>(prev_px) | (uint32_t(p.out_w) * 4u) | (plane * 4u)) & 7u) == 0u);
    auto write_pair = [&](int oy, const uint32_t (&a)[3], const uint32_t (&b)[3]) {
        uint8_t *tpx = thumb_px + uint32_t(oy) * uint32_t(row_bytes);
        float *pp = prev_px ? prev_px + uint32_t(oy) * uint32_t(p.out_w) : nullptr;
        if (vec_ok) {
            uint16_t *t16 = reinterpret_cast<uint16_t *>(tpx);
            t16[0] = uint16_t(a[0] | (a[1] << 8)); t16[1] = uint16_t(a[2] | (b[0] << 8)); t16[2] = uint16_t(b[1] | (b[2] << 8));
            if (pp) {
                *reinterpret_cast<float2 *>(pp) = make_float2(normalise(a[0], mean0, istd0), normalise(b[0], mean0, istd0));
                *reinterpret_cast<float2 *>(pp + plane) = make_float2(normalise(a[1], mean1, istd1), normalise(b[1], mean1, istd1));
                *reinterpret_cast<float2 *>(pp + 2 * plane) = make_float2(normalise(a[2], mean2, istd2), normalise(b[2], mean2, istd2));
            }
        } else {
            tpx[0] = uint8_t(a[0]); tpx[1] = uint8_t(a[1]); tpx[2] = uint8_t(a[2]);
            tpx[3] = uint8_t(b[0]); tpx[4] = uint8_t(b[1]); tpx[5] = uint8_t(b[2]);
            if (pp) {
                pp[0] = normalise(a[0], mean0, istd0); pp[1] = normalise(b[0], mean0, istd0);
                pp[plane] = normalise(a[1], mean1, istd1); pp[plane + 1] = normalise(b[1], mean1, istd1);
                pp[2 * plane] = normalise(a[2], mean2, istd2); pp[2 * plane + 1] = normalise(b[2], mean2, istd2);
            }
        }
    };

    // scatter-form vertical pass: three output rows in flight per column, accumulators never reset (see the band kernel)
    uint32_t vaa[3][3], vba[3][3], vab[3][3], vbb[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) { vaa[j][c] = vab[j][c] = uint32_t(kRound); vba[j][c] = vbb[j][c] = 0u; }

    int buf = 0, ibuf = p.n_stages - 1;
    uint32_t parity = 0;
    for (int s = 0; s < total_stages; ++s) {
        uint8_t *sbuf = ring + size_t(buf) * p.stage_bytes;
        uint64_t b0, b1;
        stage_range(s, b0, b1);
        const uint64_t a0 = b0 & ~uint64_t(15);
        if (aligned) {
            if (tid == 0 && s + p.n_stages - 1 < total_stages) issue(s + p.n_stages - 1, ibuf);
            mbar_wait(&full_bar[buf], parity);
            const uint64_t lim = img_bytes & ~uint64_t(15);
            if (b1 > lim) {
                for (uint64_t b = max(lim, a0) + tid; b < b1; b += blockDim.x) sbuf[b - a0] = img_ptr[b];
                __syncthreads();
            }
        } else {
            for (uint64_t b = b0 + tid; b < b1; b += blockDim.x) sbuf[b - a0] = img_ptr[b];
            __syncthreads();
        }

        const int ra = r0 + s * p.rows_per_stage;
        const int rb = min(ra + p.rows_per_stage, r1);
        {
            const int4 *vsp = p.vscat + ra;
            uint32_t src = smem_u32(sbuf) + uint32_t(uint64_t(ra) * pitch - a0) + 3u * xmin;
            for (int r = ra; r < rb; ++r, src += uint32_t(pitch)) {
                const uint32_t base = src & ~3u;
                const uint32_t sh = src << 3;
                uint32_t w[3 * NU + 1];
#pragma unroll
                for (int j = 0; j < 3 * NU + 1; ++j)
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[j]) : "r"(base + 4u * j));
                uint32_t aa[3][3], ab[3][3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { aa[c][0] = ab[c][0] = uint32_t(kRound); aa[c][1] = aa[c][2] = ab[c][1] = ab[c][2] = 0; }
#pragma unroll
                for (int g = 0; g < NU; ++g) {
                    const uint32_t w0 = __funnelshift_r(w[3 * g], w[3 * g + 1], sh);
                    const uint32_t w1 = __funnelshift_r(w[3 * g + 1], w[3 * g + 2], sh);
                    const uint32_t w2 = __funnelshift_r(w[3 * g + 2], w[3 * g + 3], sh);
                    uint32_t px[3];
                    px[0] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
                    px[1] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
                    px[2] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
                    if (g < NQ) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            aa[c][0] = __dp4a(px[c], ka[0][g], aa[c][0]);
                            aa[c][1] = __dp4a(px[c], ka[1][g], aa[c][1]);
                            aa[c][2] = __dp4a(px[c], ka[2][g], aa[c][2]);
                        }
                    }
                    if (g >= DQ && g < DQ + NQB) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            ab[c][0] = __dp4a(px[c], kb[0][g - DQ], ab[c][0]);
                            ab[c][1] = __dp4a(px[c], kb[1][g - DQ], ab[c][1]);
                            ab[c][2] = __dp4a(px[c], kb[2][g - DQ], ab[c][2]);
                        }
                    }
                }
                uint32_t oa[3], ob_[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    oa[c] = to_u8(aa[c][0] + (aa[c][1] << 8) + (aa[c][2] << 16));
                    ob_[c] = to_u8(ab[c][0] + (ab[c][1] << 8) + (ab[c][2] << 16));
                }
                const int4 e = __ldg(vsp++);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    vaa[0][c] += oa[c] * uint32_t(e.x); vaa[1][c] += oa[c] * uint32_t(e.y); vaa[2][c] += oa[c] * uint32_t(e.z);
                    vab[0][c] += ob_[c] * uint32_t(e.x); vab[1][c] += ob_[c] * uint32_t(e.y); vab[2][c] += ob_[c] * uint32_t(e.z);
                }
                if (e.w) {
                    const int oy = e.w >> 2;
                    const bool mine = oy >= oy0 && oy < oy1 && active;
                    auto finish = [&](uint32_t (&xa)[3], uint32_t (&ya)[3], uint32_t (&xb)[3], uint32_t (&yb)[3]) {
                        if (mine) {
                            const uint32_t pa[3] = {to_u8(xa[0] - ya[0]), to_u8(xa[1] - ya[1]), to_u8(xa[2] - ya[2])};
                            const uint32_t pb[3] = {to_u8(xb[0] - yb[0]), to_u8(xb[1] - yb[1]), to_u8(xb[2] - yb[2])};
                            write_pair(oy, pa, pb);
                        }
#pragma unroll
                        for (int c = 0; c < 3; ++c) { ya[c] = xa[c] - uint32_t(kRound); yb[c] = xb[c] - uint32_t(kRound); }
                    };
                    const int sl = e.w & 3;
                    if (sl == 1) finish(vaa[0], vba[0], vab[0], vbb[0]);
                    else if (sl == 2) finish(vaa[1], vba[1], vab[1], vbb[1]);
                    else finish(vaa[2], vba[2], vab[2], vbb[2]);
                }
            }
        }
        __syncthreads();
        if (++buf == p.n_stages) { buf = 0; parity ^= 1u; }
        if (++ibuf == p.n_stages) ibuf = 0;
    }
}

// One thread per output pixel (all three channels), any shape, any alignment.
__global__ void __launch_bounds__(256)
resize_generic_kernel(const ResizeParams p, uint32_t n) {
    const uint64_t gid = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    const uint64_t per_img = uint64_t(p.out_h) * p.out_w;
    if (gid >= per_img * n) return;
    const uint32_t img = uint32_t(gid / per_img);
    const uint32_t rem = uint32_t(gid - img * per_img);
    const int oy = rem / p.out_w, ox = rem - oy * p.out_w;
    const uint8_t *src = p.rgb + p.offsets[img];
    const int pitch = p.in_w * 3;
    const int xmin = p.hbounds[2 * ox], xcnt = p.hbounds[2 * ox + 1];
    const int ymin = p.vbounds[2 * oy], ycnt = p.vbounds[2 * oy + 1];
    const int32_t *kh = p.hcoeffs + size_t(ox) * p.ksize_h;
    const int32_t *kv = p.vcoeffs + size_t(oy) * p.ksize_v;
    int32_t v0 = kRound, v1 = kRound, v2 = kRound;
    for (int ty = 0; ty < ycnt; ++ty) {
        const uint8_t *row = src + size_t(ymin + ty) * pitch + 3 * xmin;
        int32_t h0 = kRound, h1 = kRound, h2 = kRound;
        for (int tx = 0; tx < xcnt; ++tx) {
            const int32_t k = kh[tx];
            h0 += int32_t(row[3 * tx]) * k;
            h1 += int32_t(row[3 * tx + 1]) * k;
            h2 += int32_t(row[3 * tx + 2]) * k;
        }
        const int32_t k = kv[ty];
        v0 += int32_t(clip8(h0)) * k;        // uint8 intermediate, as Pillow stores it
        v1 += int32_t(clip8(h1)) * k;
        v2 += int32_t(clip8(h2)) * k;
    }
    const uint32_t o[3] = {clip8(v0), clip8(v1), clip8(v2)};
    const uint32_t slot = p.out_slot ? p.out_slot[img] : img;
    uint8_t *t = p.thumb + (uint64_t(slot) * per_img + rem) * 3;
    t[0] = uint8_t(o[0]); t[1] = uint8_t(o[1]); t[2] = uint8_t(o[2]);
    if (p.preview) {
        float *pv = p.preview + uint64_t(slot) * 3 * per_img + rem;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            pv[size_t(c) * per_img] = normalise(o[c], p.mean[c], p.inv_std[c]);
    }
}

static int pick_quads_bucket(int max_taps) {
    const int buckets[] = {4, 8, 12, 16, 20, 28, 36};
    for (int b : buckets)
        if (max_taps <= b) return b;
    return 0;
}

// The (tap capacity, quad offset, quad count) triples of resize_pairs_kernel that are built.  (16, 1, 5) is 1080p ->
// 256 wide, BASELINE config 2: columns alternately 7 and 8 pixels apart, so the odd column genuinely needs five quads.
// (36, 2, 10) is 4K (3840 wide) -> 256, BASELINE config 5; (4, 0, 2) and (8, 0, 3) are 512 and 1024 wide (configs 1, 3:
// 2.71 against 2.87 ms and 1.61 against 1.72 ms per ~6 GB of images).  Measured against the band kernel in one call
// (tools/resize_ab.sh, tools/resize_shapes.py): 1080p 3.12 ms against 3.38 ms (-7.5 %), 4K 1.40 against 1.63 ms (-14 %);
// 2048^2, which maps to the first triple, was 4 % SLOWER because its lanes are 12 words apart and collide 4-way in the
// shared-memory banks: the plan checks the bank spread of the pair mapping before taking this kernel.  Other plans
// use the band kernel.
#define B2_PAIRS_VARIANTS(X) X(0, 16, 1, 5) X(1, 36, 2, 10) X(2, 4, 0, 2) X(3, 8, 0, 3)
static int pairs_variant(int kq, int dq, int nqb) {
#define B2_X(I, KQ, DQ, NQB) if (kq == KQ && dq == DQ && nqb == NQB) return I;
    B2_PAIRS_VARIANTS(B2_X)
#undef B2_X
    return -1;
}

template <int KQ>
static cudaError_t set_smem_attr(size_t bytes) {
    const void *fns[3] = {reinterpret_cast<const void *>(resize_bands_kernel<KQ, 0>),
                          reinterpret_cast<const void *>(resize_bands_kernel<KQ, 1>),
                          reinterpret_cast<const void *>(resize_bands_kernel<KQ, 2>)};
    for (const void *fn : fns) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int KQ>
static void launch_bands(const ResizeParams &p, uint32_t n, int threads, size_t smem, cudaStream_t st, int vmode) {
    const uint32_t grid = n * uint32_t(p.n_bands);
    if (vmode == 2) resize_bands_kernel<KQ, 2><<<grid, threads, smem, st>>>(p);
    else if (vmode == 1) resize_bands_kernel<KQ, 1><<<grid, threads, smem, st>>>(p);
    else resize_bands_kernel<KQ, 0><<<grid, threads, smem, st>>>(p);
}

// switch over the tap-capacity buckets
#define B2_QUADS_DISPATCH(qb, CALL)                                                      \
    switch (qb) {                                                                        \
        case 4: CALL(4); break;   case 8: CALL(8); break;   case 12: CALL(12); break;    \
        case 16: CALL(16); break; case 20: CALL(20); break; case 28: CALL(28); break;    \
        default: CALL(36); break;                                                        \
    }

}  // namespace b2

extern "C" int b2_resize_plan_create(int in_h, int in_w, int out_h, int out_w, b2_resize_plan **plan_out) {
    using namespace b2;
    B2_REQUIRE(plan_out != nullptr, "b2_resize_plan_create: null output");
    B2_REQUIRE(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "b2_resize_plan_create: non-positive dimension");
    B2_REQUIRE(in_h <= 65536 && in_w <= 65536 && out_h <= 16384 && out_w <= 16384, "b2_resize_plan_create: dimension too large");
    b2_resize_plan *pl = new b2_resize_plan();
    pl->in_h = in_h; pl->in_w = in_w; pl->out_h = out_h; pl->out_w = out_w;
    pl->h = make_taps(in_w, out_w);
    pl->v = make_taps(in_h, out_h);
    B2_CUDA_CHECK(cudaGetDevice(&pl->device));
    auto upload = [](const std::vector<int32_t> &v, int32_t **d) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, v.size() * sizeof(int32_t));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*d, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    };
    B2_CUDA_CHECK(upload(pl->h.bounds, &pl->d_hbounds));
    B2_CUDA_CHECK(upload(pl->h.coeffs, &pl->d_hcoeffs));
    B2_CUDA_CHECK(upload(pl->v.bounds, &pl->d_vbounds));
    B2_CUDA_CHECK(upload(pl->v.coeffs, &pl->d_vcoeffs));

    // ---- fast-path geometry -----------------------------------------------------------
    pl->threads = ((out_w + 31) / 32) * 32;
    {
        int max_taps = 0;
        for (int x = 0; x < out_w; ++x) max_taps = std::max(max_taps, pl->h.bounds[2 * x + 1]);
        pl->q_bucket = pl->threads <= 256 ? pick_quads_bucket(max_taps) : 0;    // thread-per-column layout: out_w <= 256
    }
    const int pitch = in_w * 3;
    pl->tmp_pitch = ((out_w * 4 + 15) / 16) * 16;
    pl->vparam = out_h * (2 + pl->v.ksize) <= kVtabInts;
    if (const char *e = getenv("B2_RESIZE_VPARAM")) pl->vparam = pl->vparam && atoi(e) != 0;
    // Stage geometry: two stages per CTA (the other CTAs of the SM cover the refill latency), as many rows per stage
    // as fit the shared-memory budget.  74 KB = three CTAs (24 warps) per SM: with the scatter-form vertical pass,
    // 6-row stages at 1080p, measured 3.51 ms per 2 368 images against 3.94 ms for two CTAs with 9-row stages and
    // 3.51 ms for four CTAs with 4-row stages.
    // Over-read padding per stage = the widest register window + alignment slack.
    const int overread = 3 * 36 + 64;
    pl->n_stages = 2;
    if (const char *e = getenv("B2_RESIZE_STAGES")) pl->n_stages = atoi(e) >= 2 && atoi(e) <= kMaxStages ? atoi(e) : 2;
    size_t budget = 74 * 1024;
    if (const char *e = getenv("B2_RESIZE_SMEM_KB")) budget = size_t(atoi(e)) * 1024;
    int rps_cap = 32;
    if (const char *e = getenv("B2_RESIZE_RPS")) rps_cap = atoi(e) >= 1 && atoi(e) <= 32 ? atoi(e) : rps_cap;
    auto layout = [&](int rps) {
        pl->rows_per_stage = rps;
        pl->stage_bytes = ((rps * pitch + 32 + overread + 127) / 128) * 128;
        int ring_rows = 8;                                   // rolling intermediate (thread-private columns): power of
        while (ring_rows < rps + pl->v.ksize + 2) ring_rows <<= 1;   // two covering one stage plus one tap window
        pl->tmp_ring_rows = ring_rows;
        pl->smem_fixed = size_t(pl->n_stages) * pl->stage_bytes + size_t(ring_rows) * pl->tmp_pitch + 64;
        pl->smem_max = pl->smem_fixed + (pl->vparam ? 0 : size_t(out_h) * (2 + pl->v.ksize) * 4);
    };
    int rps = rps_cap;
    for (; rps > 1; --rps) {
        layout(rps);
        if (pl->smem_max <= budget) break;
    }
    layout(rps);
    if (pl->smem_max > 220 * 1024) pl->q_bucket = 0;         // does not fit at all: generic kernel

    // ---- scatter-form vertical pass: per input row, the taps of output rows oy_lo .. oy_lo + 2, where oy_lo is the
    // first output row whose window has not ended before this input row.
    pl->scat = 0; pl->d_vscat = nullptr; pl->rps_scat = 0; pl->stage_bytes_scat = 0; pl->smem_scat = 0;
    if (pl->q_bucket != 0) {
        std::vector<int4> tab(size_t(in_h), make_int4(0, 0, 0, 0));
        bool ok = true;
        int lo = 0;
        auto first = [&](int oy) { return pl->v.bounds[2 * oy]; };
        auto last = [&](int oy) { return pl->v.bounds[2 * oy] + pl->v.bounds[2 * oy + 1] - 1; };
        for (int oy = 0; oy + 1 < out_h && ok; ++oy) ok = first(oy) <= first(oy + 1) && last(oy) < last(oy + 1);
        for (int r = 0; r < in_h && ok; ++r) {
            while (lo < out_h && last(lo) < r) ++lo;
            int k[3] = {0, 0, 0};                                            // by accumulator row = oy % 3
            for (int j = 0; j < 3; ++j) {
                const int oy = lo + j;
                if (oy < out_h && first(oy) <= r && r <= last(oy)) {
                    const int c = pl->v.coeffs[size_t(oy) * pl->v.ksize + (r - first(oy))];
                    if (c < 0) ok = false;
                    k[oy % 3] = c;
                }
            }
            if (lo + 3 < out_h && first(lo + 3) <= r) ok = false;            // a fourth row would be in flight
            const bool done = lo < out_h && last(lo) == r;
            tab[r] = make_int4(k[0], k[1], k[2], done ? (lo << 2) | (lo % 3 + 1) : 0);
        }
        if (ok) {
            int rs = rps_cap;
            for (; rs >= 1; --rs) {
                pl->rps_scat = rs;
                pl->stage_bytes_scat = ((rs * pitch + 32 + overread + 127) / 128) * 128;
                pl->smem_scat = size_t(pl->n_stages) * pl->stage_bytes_scat + 64;
                if (pl->smem_scat <= budget) break;
            }
            if (rs >= 1 && pl->smem_scat <= 220 * 1024) {
                B2_CUDA_CHECK(cudaMalloc(&pl->d_vscat, tab.size() * sizeof(int4)));
                B2_CUDA_CHECK(cudaMemcpy(pl->d_vscat, tab.data(), tab.size() * sizeof(int4), cudaMemcpyHostToDevice));
                pl->scat = 1;
            }
        }
    }
    // ---- pair kernel eligibility: even out_w, scatter form available, and (KQ, DQ, NQB) among the built instantiations
    pl->pair = 0;
    if (pl->scat && out_w % 2 == 0 && out_w >= 64) {
        int dq = 1 << 30, hi = 0;
        for (int x = 0; x + 1 < out_w; x += 2) {
            const int ob = pl->h.bounds[2 * (x + 1)] - pl->h.bounds[2 * x];
            dq = std::min(dq, ob / 4);
        }
        for (int x = 0; x + 1 < out_w; x += 2) {
            const int ob = pl->h.bounds[2 * (x + 1)] - pl->h.bounds[2 * x];
            hi = std::max(hi, ob - 4 * dq + pl->h.bounds[2 * (x + 1) + 1]);
        }
        const int nqb = (hi + 3) / 4;
        // Shared-memory banks: the 32 lanes of a warp read their union windows 2 x scale x 3 bytes apart.  1080p (45 bytes) and
        // 4K (90 bytes) spread over the banks; 2048^2 (48 bytes = 12 words) lands on 8 banks, 4096^2 (24 words) on 4, and
        // the pair kernel then LOSES to the band kernel (measured: 2048^2 1.428 against 1.374 ms).  Worst number of
        // distinct words per bank over the warps of a row decides.
        int worst = 1;
        for (int w0 = 0; w0 < out_w / 2; w0 += 32) {
            int words[32][32], cnt[32] = {};
            for (int l = 0; l < 32 && w0 + l < out_w / 2; ++l) {
                const int word = (3 * pl->h.bounds[2 * 2 * (w0 + l)]) >> 2, bank = word & 31;
                bool seen = false;
                for (int i = 0; i < cnt[bank]; ++i) seen = seen || words[bank][i] == word;
                if (!seen) words[bank][cnt[bank]++] = word;
            }
            for (int b = 0; b < 32; ++b) worst = std::max(worst, cnt[b]);
        }
        const bool built = pairs_variant(pl->q_bucket, dq, nqb) >= 0 && worst <= 3;
        if (built && !getenv("B2_RESIZE_NO_PAIRS")) {
            pl->pair_dq = dq; pl->pair_nqb = nqb;
            pl->pair_threads = ((out_w / 2 + 31) / 32) * 32;
            const int nu = std::max(pl->q_bucket / 4, dq + nqb);
            const int over = 12 * nu + 4 + 64;
            size_t pair_budget = (pl->q_bucket <= 20 ? 52 : 72) * 1024;      // four (three: wide taps need > 128 registers) CTAs of four warps per SM
            if (const char *e = getenv("B2_RESIZE_PAIR_SMEM_KB")) pair_budget = size_t(atoi(e)) * 1024;
            int rs = rps_cap;
            for (; rs >= 1; --rs) {
                pl->rps_pair = rs;
                pl->stage_bytes_pair = ((rs * pitch + 32 + over + 127) / 128) * 128;
                pl->smem_pair = size_t(pl->n_stages) * pl->stage_bytes_pair + 64;
                if (pl->smem_pair <= pair_budget) break;
            }
            pl->pair = rs >= 1 && pl->smem_pair <= 220 * 1024;
        }
    }
    *plan_out = pl;
    return B2_OK;
}

extern "C" int b2_resize_plan_destroy(b2_resize_plan *pl) {
    if (!pl) return B2_OK;
    cudaFree(pl->d_hbounds); cudaFree(pl->d_hcoeffs); cudaFree(pl->d_vbounds); cudaFree(pl->d_vcoeffs);
    if (pl->d_vscat) cudaFree(pl->d_vscat);
    delete pl;
    return B2_OK;
}

extern "C" int b2_resize_plan_taps(const b2_resize_plan *pl, int axis, int *ksize, int32_t *h_bounds,
                                   int32_t *h_coeffs, uint64_t coeffs_capacity) {
    using namespace b2;
    B2_REQUIRE(pl && ksize, "b2_resize_plan_taps: null pointer");
    B2_REQUIRE(axis == 0 || axis == 1, "b2_resize_plan_taps: axis must be 0 or 1");
    const AxisTaps &t = axis == 0 ? pl->h : pl->v;
    *ksize = t.ksize;
    if (h_bounds) memcpy(h_bounds, t.bounds.data(), t.bounds.size() * sizeof(int32_t));
    if (h_coeffs) {
        B2_REQUIRE(coeffs_capacity >= t.coeffs.size(), "b2_resize_plan_taps: coeffs buffer too small");
        memcpy(h_coeffs, t.coeffs.data(), t.coeffs.size() * sizeof(int32_t));
    }
    return B2_OK;
}

// B2_RESIZE_PATH: 0 = auto, 1 = force generic kernel (tests / comparison).
static int resize_path_override() {
    const char *e = getenv("B2_RESIZE_PATH");
    return e ? atoi(e) : 0;
}

extern "C" int b2_resize_normalize_batch(const b2_resize_plan *pl, const uint8_t *d_rgb,
                                         const uint64_t *d_offsets, const uint32_t *d_out_slot, uint32_t n,
                                         uint8_t *d_thumb, float *d_preview, const float mean[3],
                                         const float inv_std[3], void *stream) {
    using namespace b2;
    if (n == 0) return B2_OK;
    B2_REQUIRE(pl && d_rgb && d_offsets && d_thumb, "b2_resize_normalize_batch: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ResizeParams p;
    p.rgb = d_rgb; p.offsets = d_offsets; p.out_slot = d_out_slot; p.thumb = d_thumb; p.preview = d_preview;
    p.hbounds = pl->d_hbounds; p.hcoeffs = pl->d_hcoeffs; p.vbounds = pl->d_vbounds; p.vcoeffs = pl->d_vcoeffs;
    p.vscat = nullptr;
    p.in_h = pl->in_h; p.in_w = pl->in_w; p.out_h = pl->out_h; p.out_w = pl->out_w;
    p.ksize_h = pl->h.ksize; p.ksize_v = pl->v.ksize;
    // Bands: the whole image per CTA when the batch alone fills the GPU (no input row is read twice);
    // small batches are cut into bands of >= 16 output rows so that every SM gets work.
    int n_bands = 1;
    {
        const int want_ctas = 4 * sm_count();
        if (int64_t(n) < want_ctas) n_bands = int((want_ctas + n - 1) / n);
        const int max_bands = pl->out_h / 16 > 0 ? pl->out_h / 16 : 1;
        if (n_bands > max_bands) n_bands = max_bands;
    }
    p.band_rows = (pl->out_h + n_bands - 1) / n_bands;
    p.n_bands = (pl->out_h + p.band_rows - 1) / p.band_rows;
    p.rows_per_stage = pl->rows_per_stage; p.stage_bytes = pl->stage_bytes; p.n_stages = pl->n_stages;
    p.tmp_pitch = pl->tmp_pitch; p.tmp_ring_rows = pl->tmp_ring_rows;
    const size_t smem_launch = pl->smem_fixed + (pl->vparam ? 0 : size_t(p.band_rows) * (2 + pl->v.ksize) * 4);
    if (pl->vparam) {
        const int vs = 2 + pl->v.ksize;
        for (int oy = 0; oy < pl->out_h; ++oy) {
            p.vtab[oy * vs] = pl->v.bounds[2 * oy];
            p.vtab[oy * vs + 1] = pl->v.bounds[2 * oy + 1];
            memcpy(&p.vtab[oy * vs + 2], &pl->v.coeffs[size_t(oy) * pl->v.ksize], size_t(pl->v.ksize) * 4);
        }
    }
    for (int c = 0; c < 3; ++c) {
        p.mean[c] = mean ? mean[c] : 0.0f;
        p.inv_std[c] = inv_std ? inv_std[c] : 1.0f;
    }
    const bool fast = pl->q_bucket != 0 && resize_path_override() != 1 &&
                      uint64_t(n) * uint64_t(p.n_bands) < 0x7fffffffull;
    if (!fast) {
        const uint64_t threads = uint64_t(n) * pl->out_h * pl->out_w;
        B2_REQUIRE((threads + 255) / 256 < 0x7fffffffull, "b2_resize_normalize_batch: batch too large");
        resize_generic_kernel<<<unsigned((threads + 255) / 256), 256, 0, st>>>(p, n);
        B2_LAUNCH_CHECK("resize_generic_kernel");
        return B2_OK;
    }
    // Vertical pass: scatter form for every downscale (B2_RESIZE_VSCAT=0 forces the gather form: tests, comparisons)
    bool vscat = pl->scat != 0;
    if (const char *ev = getenv("B2_RESIZE_VSCAT")) vscat = vscat && atoi(ev) != 0;
    const int vmode = vscat ? 2 : (pl->vparam ? 1 : 0);
    size_t smem = smem_launch;
    if (vscat) {
        p.vscat = pl->d_vscat;
        p.rows_per_stage = pl->rps_scat; p.stage_bytes = pl->stage_bytes_scat;
        smem = pl->smem_scat;
    }
    if (vscat && pl->pair) {
        p.rows_per_stage = pl->rps_pair; p.stage_bytes = pl->stage_bytes_pair;
        const int variant = pairs_variant(pl->q_bucket, pl->pair_dq, pl->pair_nqb);
        static std::mutex pmu;
        static size_t pair_attr[64][8];
        cudaError_t pe = cudaSuccess;
        {
            std::lock_guard<std::mutex> lock(pmu);
            if (pair_attr[pl->device & 63][variant] < pl->smem_pair) {
#define B2_X(I, KQ, DQ, NQB)                                                                                                    \
                if (variant == I) {                                                                                             \
                    pe = cudaFuncSetAttribute(resize_pairs_kernel<KQ, DQ, NQB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem_pair)); \
                    if (pe == cudaSuccess)                                                                                      \
                        pe = cudaFuncSetAttribute(resize_pairs_kernel<KQ, DQ, NQB>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                                  cudaSharedmemCarveoutMaxShared);                                             \
                }
                B2_PAIRS_VARIANTS(B2_X)
#undef B2_X
                if (pe == cudaSuccess) pair_attr[pl->device & 63][variant] = pl->smem_pair;
            }
        }
        B2_CUDA_CHECK(pe);
        const uint32_t grid = n * uint32_t(p.n_bands);
#define B2_X(I, KQ, DQ, NQB) if (variant == I) resize_pairs_kernel<KQ, DQ, NQB><<<grid, pl->pair_threads, pl->smem_pair, st>>>(p);
        B2_PAIRS_VARIANTS(B2_X)
#undef B2_X
        B2_LAUNCH_CHECK("resize_pairs_kernel");
        return B2_OK;
    }
    const size_t smem_attr = std::max(pl->smem_max, pl->smem_scat);
    static std::mutex mu;
    static size_t attr_bytes[64][8];     // [device][bucket index]: largest smem opt-in done so far
    cudaError_t e = cudaSuccess;
    {
        std::lock_guard<std::mutex> lock(mu);
        const int qb = pl->q_bucket;
        const int bi = qb <= 20 ? qb / 4 - 1 : (qb == 28 ? 5 : 6);
        const int dev = pl->device & 63;
        if (attr_bytes[dev][bi] < smem_attr) {
#define B2_CALL(KQ) e = set_smem_attr<KQ>(smem_attr)
            B2_QUADS_DISPATCH(qb, B2_CALL)
#undef B2_CALL
            if (e == cudaSuccess) attr_bytes[dev][bi] = smem_attr;
        }
    }
    B2_CUDA_CHECK(e);
#define B2_CALL(KQ) launch_bands<KQ>(p, n, pl->threads, smem, st, vmode)
    B2_QUADS_DISPATCH(pl->q_bucket, B2_CALL)
#undef B2_CALL
    B2_LAUNCH_CHECK("resize_bands_kernel");
    return B2_OK;
}
