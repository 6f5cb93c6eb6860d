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
using P4096 = Plan<4096, 256, 2, 4, 16, 16, 16>;
// 3 x 11 is one Good-Thomas radix-33 butterfly in registers (no internal twiddles): three shared-memory stages, not four
using P8184 = PfaPlan<8184, 288, 1, 0, 8, 33, 31>;
using P16368 = PfaPlan<16368, 544, 1, 0, 16, 33, 31>;
using P20000 = Plan<20000, 512, 1, 0, 8, 4, 25, 25>;

// tuning variants of the headline plan (selected with the environment variable GB_ACQ_VARIANT=1..4)
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

}  // namespace gb
