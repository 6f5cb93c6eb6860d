// acq_common.cuh -- plans and device helpers shared by the acquisition kernels.
#pragma once
#include "acq_kernels.cuh"
#include "fft_smem.cuh"

namespace gb {

// ------------------------------------------------------------------ plans
//                 N      T  MINB PAD  radices (forward DIF order; odd radices last => no padding needed)
using P1024 = Plan<1024, 64, 8, 4, 4, 16, 16>;
using P2048 = Plan<2048, 128, 4, 4, 8, 16, 16>;
using P4092 = PfaPlan<4092, 160, 4, 0, 12, 11, 31>;
// stage geometry of acq_inverse_lw_kernel: the same radices on 128 working threads (+ one leftover warp = P4092::T)
using P4092W = PfaPlan<4092, 128, 4, 0, 12, 11, 31>;
using P4092W3 = PfaPlan<4092, 128, 3, 0, 12, 11, 31>;   // the default: 3 CTAs per SM, 128 registers, no accumulator spills
using P4092W2 = PfaPlan<4092, 128, 2, 0, 12, 11, 31>;
using P4092W5 = PfaPlan<4092, 128, 5, 0, 12, 11, 31>;   // A/B: five CTAs per SM with the accumulators in tensor memory (80 registers)
using P4096 = Plan<4096, 256, 2, 4, 16, 16, 16>;
// 3 x 11 is one Good-Thomas radix-33 butterfly in registers (no internal twiddles): three shared-memory stages, not four
using P8184 = PfaPlan<8184, 288, 1, 0, 8, 33, 31>;
using P16368 = PfaPlan<16368, 544, 1, 0, 16, 33, 31>;
#if defined(GB_P20000_ALT) && GB_P20000_ALT == 1     // A/B builds (tools/build_alt.sh): three shared-memory stages instead of four
using P20000 = Plan<20000, 320, 1, 0, 32, 25, 25>;
#elif defined(GB_P20000_ALT) && GB_P20000_ALT == 2
using P20000 = Plan<20000, 400, 1, 0, 32, 25, 25>;
#elif defined(GB_P20000_ALT) && GB_P20000_ALT == 3
using P20000 = Plan<20000, 640, 1, 0, 32, 25, 25>;
#else
using P20000 = Plan<20000, 512, 1, 0, 8, 4, 25, 25>;
#endif

// tuning variants of the headline plan (GB_TUNING builds only; selected with gb_tuning_set("acq_variant", 1..))
using P4092v1 = Plan<4092, 160, 4, 0, 12, 11, 31>;   // the Cooley-Tukey form of the default plan (A/B)
using P4092v2 = PfaPlan<4092, 160, 3, 0, 12, 11, 31>;
using P4092v3 = Plan<4092, 192, 2, 0, 12, 11, 31>;
using P4092v4 = Plan<4092, 384, 1, 0, 12, 11, 31>;
using P4092v5 = PfaStreamPlan<3, 4092, 160, 4, 0, 12, 11, 31>;
using P4092v6 = PfaStreamPlan<5, 4092, 160, 4, 0, 12, 11, 31>;
using P4092v7 = PfaStreamPlan<3, 4092, 160, 3, 0, 12, 11, 31>;
using P4092v8 = PfaStreamPlan<5, 4092, 160, 3, 0, 12, 11, 31>;
using P16368v1 = PfaPlan<16368, 544, 1, 0, 16, 3, 11, 31>;   // the four-stage form (A/B)
using P16368v2 = PfaPlan<16368, 512, 1, 0, 16, 33, 31>;

// reference arithmetic of multiply_simd_block (doppler_shift.rs:43-58): separate roundings, no FMA
__device__ __forceinline__ float2 wipe(float2 x, float2 w)
{
    return make_float2(__fadd_rn(__fmul_rn(x.x, w.x), -__fmul_rn(x.y, w.y)),
                       __fadd_rn(__fmul_rn(x.x, w.y), __fmul_rn(x.y, w.x)));
}

__device__ __forceinline__ float2 ld_iq(const AcqArgs& a, unsigned long long idx)
{
    return __ldg(&a.iq[(a.iq_start + idx) & a.iq_mask]);
}

// satellite_detection_two_peaks' second-peak window (acquisition_bk.rs:371-390), slice bounds verbatim: with
// left = cp - spc and right = cp + spc the second peak is searched in
//   left < 1    : [right-1, N+left)          (code phase within spc of the start)
//   right >= N  : [right-N-1, left)          (within spc of the end; the lower bound underflows for right == N, clamped)
//   otherwise   : [0, left) u [right, N)     -- asymmetric: index cp+spc is searched, cp-spc is not
__device__ __forceinline__ bool two_peak_searched(int n, int cp, int spc, int N)
{
    const int left = cp - spc, right = cp + spc;
    if (left < 1) return n >= right - 1 && n < N + left;
    if (right >= N) return n >= max(right - N - 1, 0) && n < left;
    return n < left || n >= right;
}

struct PeakIdx {
    float v;
    unsigned idx;
};
// "first index of the strict maximum": larger value wins, ties go to the smaller index; NaN never wins
__device__ __forceinline__ PeakIdx peak_merge(PeakIdx a, PeakIdx b)
{
    const bool take_b = (b.v > a.v) || (b.v == a.v && b.idx < a.idx);
    return take_b ? b : a;
}
__device__ __forceinline__ PeakIdx warp_peak(PeakIdx p)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        PeakIdx q;
        q.v = __shfl_xor_sync(0xffffffffu, p.v, o);
        q.idx = __shfl_xor_sync(0xffffffffu, p.idx, o);
        p = peak_merge(p, q);
    }
    return p;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Row reduction shared by the fused and the shared-forward kernels: peak / first argmax / 8-lane sum
// (Q2: only the first 8*floor(N/8) bins) / second peak outside +-spc of the first.
// For prime-factor plans the code-phase index of line position l is npos[l] (the Ruritanian map).
// BAR = 0: the CTA is exactly P::T threads (__syncthreads); BAR > 0: only the first P::T threads of a larger CTA take
// part and synchronise on named barrier BAR (acq_inverse_lw_kernel's working warps).
template <class P, int BAR = 0>
__device__ __forceinline__ void reduce_row_to_cell(float (&acc)[StageGeo<P, 0>::ITERS][StageGeo<P, 0>::R], float2* line,
                                                   int spc, gb_acq_cell* out, const int* __restrict__ npos)
{
    auto cta_sync = [] {
        if constexpr (BAR == 0) __syncthreads();
        else named_bar_sync(BAR, P::T);
    };
    using G0 = StageGeo<P, 0>;
    constexpr int N = P::N;
    constexpr int NSUM = (N / 8) * 8;
    constexpr int NW = P::T / 32;
    PeakIdx pk;
    pk.v = 0.f;
    pk.idx = 0u;
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) {
        const int i = threadIdx.x + it * P::T;
        if (G0::NB % P::T == 0 || i < G0::NB) {
#pragma unroll
            for (int j = 0; j < G0::R; j++) {
                const int n = P::PFA ? __ldg(&npos[i + j * G0::SUB]) : i + j * G0::SUB;
                const float v = acc[it][j];
                PeakIdx c;
                c.v = v;
                c.idx = (unsigned)n;
                if (v > 0.f) pk = peak_merge(pk, c);
                if (NSUM == N || n < NSUM) sum += v;
            }
        }
    }
    pk = warp_peak(pk);
    sum = warp_sum(sum);
    float* red_v = reinterpret_cast<float*>(line);
    unsigned* red_i = reinterpret_cast<unsigned*>(line) + 64;
    float* red_s = reinterpret_cast<float*>(line) + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red_v[warp] = pk.v;
        red_i[warp] = pk.idx;
        red_s[warp] = sum;
    }
    cta_sync();
    if (warp == 0) {
        PeakIdx q;
        q.v = lane < NW ? red_v[lane] : 0.f;
        q.idx = lane < NW ? red_i[lane] : 0u;
        float s = lane < NW ? red_s[lane] : 0.f;
        q = warp_peak(q);
        s = warp_sum(s);
        if (lane == 0) {
            red_v[32] = q.v;
            red_i[32] = q.idx;
            red_s[32] = s;
        }
    }
    cta_sync();
    const float peak = red_v[32];
    const unsigned arg = red_i[32];
    const float total = red_s[32];
    float p2 = 0.f;
    if (spc > 0) {
#pragma unroll
        for (int it = 0; it < G0::ITERS; it++) {
            const int i = threadIdx.x + it * P::T;
            if (G0::NB % P::T == 0 || i < G0::NB) {
#pragma unroll
                for (int j = 0; j < G0::R; j++) {
                    const int n = P::PFA ? __ldg(&npos[i + j * G0::SUB]) : i + j * G0::SUB;
                    if (two_peak_searched(n, (int)arg, spc, N)) p2 = fmaxf(p2, acc[it][j]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p2 = fmaxf(p2, __shfl_xor_sync(0xffffffffu, p2, o));
        cta_sync();
        if (lane == 0) red_v[warp] = p2;
        cta_sync();
        if (warp == 0) {
            float v = lane < NW ? red_v[lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            p2 = v;
        }
    }
    if (threadIdx.x == 0) {
        gb_acq_cell c;
        c.peak = peak;
        c.argmax = arg;
        c.sum8 = total;
        c.peak2 = p2;
        *out = c;
    }
}

// Final inverse stage (DIT, L = N) fused with |.|^2 accumulate; outputs are in natural order.
template <class P>
__device__ __forceinline__ void final_stage_accumulate(const float2* __restrict__ line, const float2* __restrict__ tw,
                                                       float (&acc)[StageGeo<P, 0>::ITERS][StageGeo<P, 0>::R])
{
    using G0 = StageGeo<P, 0>;
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) {
        const int i = threadIdx.x + it * P::T;
        if (G0::NB % P::T == 0 || i < G0::NB) {
            float2 v[G0::R];
            v[0] = line[P::phys(i)];
#pragma unroll
            for (int q = 1; q < G0::R; q++) {
                const float2 u = line[P::phys(i + q * G0::SUB)];
                v[q] = P::PFA ? u : cmul_conj(u, __ldg(&tw[(q - 1) * G0::SUB + i]));
            }
            Dft<G0::R, true>::run(v);
#pragma unroll
            // two FMAs per point (acc + im^2, then + re^2): one instruction fewer than |.|^2 followed by an add
            for (int j = 0; j < G0::R; j++) acc[it][j] = __fmaf_rn(v[j].x, v[j].x, __fmaf_rn(v[j].y, v[j].y, acc[it][j]));
        }
    }
}

}  // namespace gb
