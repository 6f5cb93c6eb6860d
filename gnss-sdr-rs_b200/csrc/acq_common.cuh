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
#elif defined(GB_P20000_ALT) && GB_P20000_ALT == 4
using P20000 = Plan<20000, 512, 1, 0, 32, 25, 25>;   // with the accumulators in tensor memory (acq_tmem 1): no 64-register accumulator block
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

// ------------------------------------------------------------------ power accumulators in tensor memory
// Tensor memory (256 KB per SM: 128 lanes x 512 columns x 32 bit) is an accumulator store with its own datapath
// (tcgen05.ld / tcgen05.st, SASS LDTM / STTM): no register, no L1 data-pipe traffic.  A warp may touch the 32 lanes of its
// quarter (warp id % 4) only, so thread (warp w, lane l) owns TMEM lane 32 (w % 4) + l; warps that share a quarter take
// different column ranges.  The accumulate is the same two-FMA form in the same order as final_stage_accumulate.
template <uint32_t COLS> __device__ __forceinline__ uint32_t tmem_alloc_cta(uint32_t* base_smem)
{
    static_assert(COLS >= 32 && COLS <= 512 && (COLS & (COLS - 1)) == 0, "power of two, 32 .. 512 columns");
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(base_smem)), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *base_smem;
}
// by the warp that allocated (warp 0), after every warp's last TMEM access
template <uint32_t COLS> __device__ __forceinline__ void tmem_dealloc_warp(uint32_t base)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(COLS) : "memory");
}
template <int R> __device__ __forceinline__ void tmem_ld(float (&a)[R], uint32_t taddr)
{
    static_assert(R == 4 || R == 8 || R == 12 || R == 16 || R == 32, "radix of the last inverse stage");
    if constexpr (R == 32) {
        tmem_ld<16>(*reinterpret_cast<float(*)[16]>(&a[0]), taddr);
        tmem_ld<16>(*reinterpret_cast<float(*)[16]>(&a[16]), taddr + 16);
    } else if constexpr (R == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]), "=f"(a[8]), "=f"(a[9]),
                       "=f"(a[10]), "=f"(a[11]), "=f"(a[12]), "=f"(a[13]), "=f"(a[14]), "=f"(a[15])
                     : "r"(taddr));
    } else {
        if constexpr (R >= 8)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "r"(taddr));
        if constexpr (R != 8)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(a[R - 4]), "=f"(a[R - 3]), "=f"(a[R - 2]), "=f"(a[R - 1]) : "r"(taddr + (R - 4)));
    }
}
// the loaded registers are operands of the wait so that no use of them can be scheduled ahead of it
template <int R> __device__ __forceinline__ void tmem_wait_ld(float (&a)[R])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < R; j++) asm volatile("" : "+f"(a[j]));
}
template <int R> __device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&a)[R])
{
    static_assert(R == 4 || R == 8 || R == 12 || R == 16 || R == 32, "radix of the last inverse stage");
    if constexpr (R == 32) {
        tmem_st<16>(taddr, *reinterpret_cast<const float(*)[16]>(&a[0]));
        tmem_st<16>(taddr + 16, *reinterpret_cast<const float(*)[16]>(&a[16]));
    } else if constexpr (R == 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     :: "r"(taddr), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]), "f"(a[8]), "f"(a[9]),
                        "f"(a[10]), "f"(a[11]), "f"(a[12]), "f"(a[13]), "f"(a[14]), "f"(a[15])
                     : "memory");
    } else {
        if constexpr (R >= 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         :: "r"(taddr), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
        if constexpr (R != 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                         :: "r"(taddr + (R - 4)), "f"(a[R - 4]), "f"(a[R - 3]), "f"(a[R - 2]), "f"(a[R - 1]) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld8_nm(float (&a)[8], uint32_t taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "r"(taddr));
}
// no "memory" clobber: the spectrum loads of the group may be scheduled across it
__device__ __forceinline__ void tmem_wait_ld8_nm(float (&a)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]));
}

// TMEM columns one thread needs, and the CTA's allocation (warps that share a lane quarter stack their column ranges)
// MODE bit 0: the power accumulators (ITERS0 x R0 columns), bit 1: the thread's code-spectrum values of the first inverse
// stage (per iteration ceil(RM / 4) chunks of four complex values = eight columns), stacked in that order
template <class P> __host__ __device__ constexpr uint32_t tmem_acc_cols() { return StageGeo<P, 0>::ITERS * StageGeo<P, 0>::R; }
template <class P> __host__ __device__ constexpr uint32_t tmem_code_chunks() { return (StageGeo<P, P::NSTAGE - 1>::R + 3) / 4; }
template <class P> __host__ __device__ constexpr uint32_t tmem_code_cols() { return StageGeo<P, P::NSTAGE - 1>::ITERS * tmem_code_chunks<P>() * 8; }
template <class P, int MODE = 1> __host__ __device__ constexpr uint32_t tmem_cols_per_thread()
{
    return ((MODE & 1) ? tmem_acc_cols<P>() : 0) + ((MODE & 2) ? tmem_code_cols<P>() : 0);
}
template <class P, int MODE = 1, int T_ = P::T> __host__ __device__ constexpr uint32_t tmem_cols_cta()
{
    uint32_t need = ((T_ + 127) / 128) * tmem_cols_per_thread<P, MODE>(), c = 32;
    while (c < need) c *= 2;
    return c;
}
// this thread's first column: base + (lane quarter << 16) + column range of its warp
template <class P, int MODE = 1> __device__ __forceinline__ uint32_t tmem_thread_addr(uint32_t base)
{
    const uint32_t w = threadIdx.x >> 5;
    return base + (((w & 3u) * 32u) << 16) + (w >> 2) * tmem_cols_per_thread<P, MODE>();
}
template <class P> __device__ __forceinline__ void tmem_zero_accumulators(uint32_t taddr)
{
    using G0 = StageGeo<P, 0>;
    float z[G0::R];
#pragma unroll
    for (int j = 0; j < G0::R; j++) z[j] = 0.f;
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) tmem_st<G0::R>(taddr + it * G0::R, z);
}
template <class P> __device__ __forceinline__ void tmem_load_accumulators(uint32_t taddr, float (&acc)[StageGeo<P, 0>::ITERS][StageGeo<P, 0>::R])
{
    using G0 = StageGeo<P, 0>;
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) tmem_ld<G0::R>(acc[it], taddr + it * G0::R);
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) tmem_wait_ld<G0::R>(acc[it]);
}
// final inverse stage fused with |.|^2 accumulate, accumulators in tensor memory (see final_stage_accumulate)
template <class P>
__device__ __forceinline__ void final_stage_accumulate_tmem(const float2* __restrict__ line, const float2* __restrict__ tw, uint32_t taddr)
{
    using G0 = StageGeo<P, 0>;
    const int warp0 = threadIdx.x & ~31;
    // the previous group's stores (or the zero fill) must have landed before these loads: by now they long have
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) {
        if (warp0 + it * P::T >= G0::NB) continue;   // the whole warp is past the last butterfly (warp-uniform)
        const int i = threadIdx.x + it * P::T;
        const bool active = G0::NB % P::T == 0 || i < G0::NB;
        float a[G0::R];
        tmem_ld<G0::R>(a, taddr + it * G0::R);
        float2 v[G0::R];
        if (active) {
            v[0] = line[P::phys(i)];
#pragma unroll
            for (int q = 1; q < G0::R; q++) {
                const float2 u = line[P::phys(i + q * G0::SUB)];
                v[q] = P::PFA ? u : cmul_conj(u, __ldg(&tw[(q - 1) * G0::SUB + i]));
            }
            Dft<G0::R, true>::run(v);
        }
        tmem_wait_ld<G0::R>(a);
        if (active) {
#pragma unroll
            for (int j = 0; j < G0::R; j++) a[j] = __fmaf_rn(v[j].x, v[j].x, __fmaf_rn(v[j].y, v[j].y, a[j]));
        }
        tmem_st<G0::R>(taddr + it * G0::R, a);
    }
}

}  // namespace gb
