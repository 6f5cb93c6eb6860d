// fft_smem.cuh -- in-place mixed-radix FFT stages on a shared-memory line, sm_100a.
//
// Replaces the rustfft plans the reference builds at do_acquisition.rs:132-142 (forward at :182,
// inverse at :188, both unnormalised).  Design (not a translation of rustfft):
//   * forward = decimation-in-frequency, radices r1..rm, output left in mixed-radix digit-reversed
//     order; inverse = decimation-in-time over the same stages in reverse, consuming that order and
//     producing natural order.  The pointwise multiply by the conjugated code spectrum happens in
//     the scrambled domain (the code spectra are produced by the same forward stages), so no
//     permutation pass exists anywhere on the hot path.
//   * one thread owns one radix-R butterfly entirely in registers; radix-R DFTs of odd primes use
//     the conjugate-symmetric half form with compile-time constants (FFMA-immediate).
//   * stage twiddles W_L^(i*q) come from per-stage tables laid out [q][i] (f64-evaluated, f32-rounded)
//     so a warp's loads are contiguous; read through the read-only path.
//   * power-of-two plans use a padded line (one complex of padding per 16) so the short-stride
//     stages are bank-conflict free; plans whose last radix is odd need no padding.
#pragma once
#include <cuda_runtime.h>

#include "radix_tables.cuh"

namespace gb {

// ---- packed FP32 pairs (Blackwell FFMA2 / FADD2: one instruction, two IEEE-rn f32 operations) ----
// A complex value lives in an aligned register pair anyway (it is loaded and stored as 64 bits), so complex
// add/sub and "complex x real scalar + complex" map 1:1 onto the sm_100a packed instructions.  Measured on B200
// (tools/ubench/ffma2.cu): FFMA2 sustains the same 128 FMA/clk/SM as FFMA with half the issue slots -- these kernels
// are issue-bound, not FMA-pipe-bound, so halving the FP instruction count is the lever.  Each lane of a packed op
// rounds exactly like the scalar instruction: results are bit-identical to the scalar formulation.
typedef unsigned long long pk64;
__device__ __forceinline__ pk64 pk(float x, float y)
{
    pk64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ pk64 pk(float2 v) { return pk(v.x, v.y); }
__device__ __forceinline__ float2 upk(pk64 v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ pk64 add2(pk64 a, pk64 b)
{
    pk64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk64 sub2(pk64 a, pk64 b)
{
    pk64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// a * (s, s) + c : the scalar becomes an FFMA2 immediate / scalar operand
__device__ __forceinline__ pk64 fma2s(pk64 a, float s, pk64 c)
{
    pk64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(pk(s, s)), "l"(c));
    return r;
}

__device__ __forceinline__ pk64 mul2(pk64 a, pk64 b)
{
    pk64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk64 fma2(pk64 a, pk64 b, pk64 c)
{
    pk64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// Complex multiplies in TWO packed instructions (build with -DGB_PACKED_CMUL; -DGB_PACKED_ROT does the same for the
// +-i rotations of the odd-prime combines).  The packed operands take a lane swizzle and a per-lane sign for free
// (SASS: R.F32x2.LO_HI.NP) and a broadcast scalar (R.F32), so
//   a * b       = a (.) (b.x, b.x) + swap(a) (.) (-b.y,  b.y)     FMUL2 + FFMA2
//   a * conj(b) = a (.) (b.x, b.x) + swap(a) (.) ( b.y, -b.y)
// and ptxas folds the pk(a.y, a.x) / pk(-b.y, b.y) packings below into those operand modifiers (no MOV is emitted).
// Measured on B200 (config 2 / config 1, samples resident): scalar 1.776 / 0.734 ms, ROT 1.786 / 0.732, CMUL 1.778 /
// 0.736, both 1.781 / 0.732 -- 9 % fewer instructions, identical time: the inverse kernel is bound by FMA-pipe CYCLES
// (a packed instruction holds the pipe for two), not by issue slots, so the scalar forms stay the default.
#ifdef GB_PACKED_CMUL
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return upk(fma2(pk(a.y, a.x), pk(-b.y, b.y), mul2(pk(a), pk(b.x, b.x))));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b)  // a * conj(b)
{
    return upk(fma2(pk(a.y, a.x), pk(b.y, -b.y), mul2(pk(a), pk(b.x, b.x))));
}
#else
// Scalar forms with the contraction written out (one FMUL + one FFMA per component, the same roundings as the packed
// forms above): left to the compiler, `a*b + c*d` is contracted differently in different inlining contexts, and the
// kernels that must agree bit for bit (fused / shared / leftover-warp chains) would drift apart in the last ulp.
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(__fmaf_rn(-a.y, b.y, __fmul_rn(a.x, b.x)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b)  // a * conj(b)
{
    return make_float2(__fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x)), __fmaf_rn(-a.x, b.y, __fmul_rn(a.y, b.x)));
}
#endif
// c + i*s and c - i*s for packed complex c, s (i*s = (-s.y, s.x)): one FADD2 each, the rotation is an operand modifier
__device__ __forceinline__ float2 add_i(pk64 c, pk64 s)
{
#ifdef GB_PACKED_ROT
    const float2 t = upk(s);
    return upk(add2(c, pk(-t.y, t.x)));
#else
    const float2 cc = upk(c), ss = upk(s);
    return make_float2(cc.x - ss.y, cc.y + ss.x);
#endif
}
__device__ __forceinline__ float2 sub_i(pk64 c, pk64 s)
{
#ifdef GB_PACKED_ROT
    const float2 t = upk(s);
    return upk(add2(c, pk(t.y, -t.x)));
#else
    const float2 cc = upk(c), ss = upk(s);
    return make_float2(cc.x + ss.y, cc.y - ss.x);
#endif
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return upk(add2(pk(a), pk(b))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return upk(sub2(pk(a), pk(b))); }
// multiply by -i (forward W_4) or +i (inverse)
template <bool INV> __device__ __forceinline__ float2 rot90(float2 a)
{
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// ------------------------------------------------------------------ in-register DFTs
// Dft<R,INV>::run(v): v[q] <- sum_j v[j] * exp(-/+ 2 pi i j q / R), unnormalised.
template <int R, bool INV> struct Dft;

template <bool INV> struct Dft<1, INV> {
    static __device__ __forceinline__ void run(float2 (&)[1]) {}
};

template <bool INV> struct Dft<2, INV> {
    static __device__ __forceinline__ void run(float2 (&v)[2])
    {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <bool INV> struct Dft<4, INV> {
    static __device__ __forceinline__ void run(float2 (&v)[4])
    {
        const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        const float2 c = cadd(v[1], v[3]), d = rot90<INV>(csub(v[1], v[3]));
        v[0] = cadd(a, c);
        v[1] = cadd(b, d);
        v[2] = csub(a, c);
        v[3] = csub(b, d);
    }
};

// multiply by W_8^k (forward) / conj (inverse), k = 1, 3
template <bool INV> __device__ __forceinline__ float2 mulw8_1(float2 a)
{
    const float h = 0.70710678118654752440f;
    return INV ? make_float2((a.x - a.y) * h, (a.x + a.y) * h) : make_float2((a.x + a.y) * h, (a.y - a.x) * h);
}
template <bool INV> __device__ __forceinline__ float2 mulw8_3(float2 a)
{
    const float h = 0.70710678118654752440f;
    return INV ? make_float2(-(a.x + a.y) * h, (a.x - a.y) * h) : make_float2((a.y - a.x) * h, -(a.x + a.y) * h);
}

template <bool INV> struct Dft<8, INV> {
    static __device__ __forceinline__ void run(float2 (&v)[8])
    {
        // 2 x radix-4 over even / odd inputs, then radix-2 combine with W_8^q
        float2 e[4] = {v[0], v[2], v[4], v[6]};
        float2 o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4, INV>::run(e);
        Dft<4, INV>::run(o);
        o[1] = mulw8_1<INV>(o[1]);
        o[2] = rot90<INV>(o[2]);
        o[3] = mulw8_3<INV>(o[3]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            v[q] = cadd(e[q], o[q]);
            v[q + 4] = csub(e[q], o[q]);
        }
    }
};

// multiply by exp(-/+ 2 pi i k / R) with compile-time k (generic constant twiddle)
template <int R, bool INV> __device__ __forceinline__ float2 mulw(float2 a, int k)
{
    const float c = RT<R>::c(k), s = INV ? RT<R>::s(k) : -RT<R>::s(k);
#ifdef GB_PACKED_CMUL
    return upk(fma2(pk(a.y, a.x), pk(-s, s), mul2(pk(a), pk(c, c))));
#else
    return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
#endif
}

template <bool INV> struct Dft<16, INV> {
    static __device__ __forceinline__ void run(float2 (&v)[16])
    {
        // 4 x 4: j = 4*j1 + j2, q = q1 + 4*q2
        float2 t[4][4];
#pragma unroll
        for (int j2 = 0; j2 < 4; j2++) {
            float2 c[4] = {v[j2], v[4 + j2], v[8 + j2], v[12 + j2]};
            Dft<4, INV>::run(c);
#pragma unroll
            for (int q1 = 0; q1 < 4; q1++) t[q1][j2] = c[q1];
        }
#pragma unroll
        for (int q1 = 0; q1 < 4; q1++) {
            float2 c[4];
            c[0] = t[q1][0];
#pragma unroll
            for (int j2 = 1; j2 < 4; j2++) {
                const int k = j2 * q1;
                if (k == 0) c[j2] = t[q1][j2];
                else if (k == 4) c[j2] = rot90<INV>(t[q1][j2]);
                else if (k == 2) c[j2] = mulw8_1<INV>(t[q1][j2]);
                else if (k == 6) c[j2] = mulw8_3<INV>(t[q1][j2]);
                else c[j2] = mulw<16, INV>(t[q1][j2], k);
            }
            Dft<4, INV>::run(c);
#pragma unroll
            for (int q2 = 0; q2 < 4; q2++) v[q1 + 4 * q2] = c[q2];
        }
    }
};

// Odd prime radix, conjugate-symmetric half form:
//   a_j = v_j + v_{R-j}, b_j = v_j - v_{R-j}
//   X_q, X_{R-q} = v_0 + sum_j a_j cos(2 pi j q/R)  +/-  (-/+ i) sum_j b_j sin(2 pi j q/R)
template <int R, bool INV> struct DftOddPrime {
    static __device__ __forceinline__ void run(float2 (&v)[R])
    {
        constexpr int H = (R - 1) / 2;
        pk64 a[H + 1], b[H + 1];
        const pk64 v0 = pk(v[0]);
        pk64 x0 = v0;
#pragma unroll
        for (int j = 1; j <= H; j++) {
            const pk64 p = pk(v[j]), m = pk(v[R - j]);
            a[j] = add2(p, m);
            b[j] = sub2(p, m);
            x0 = add2(x0, a[j]);
        }
        v[0] = upk(x0);
#pragma unroll
        for (int q = 1; q <= H; q++) {
            // c2 = v0 + sum_j a_j cos;  s2 = sum_j b_j sin  (packed re/im);  i*s2 = (-s2.y, s2.x)
            pk64 c2 = v0, s2 = pk(0.f, 0.f);
#pragma unroll
            for (int j = 1; j <= H; j++) {
                const int k = (j * q) % R;
                const float c = RT<R>::c(k);
                const float s = INV ? RT<R>::s(k) : -RT<R>::s(k);  // Im of exp(-/+ i theta)
                c2 = fma2s(a[j], c, c2);
                s2 = fma2s(b[j], s, s2);
            }
            v[q] = add_i(c2, s2);
            v[R - q] = sub_i(c2, s2);
        }
    }
};
// Same arithmetic as DftOddPrime::run, but every output is handed to `emit(q, X_q)` the moment it is
// final instead of being written back into v[]: the a/b half-sums are the only long-lived registers, so
// a radix-31 butterfly needs ~75 registers instead of ~130.  450 FFMA2 + 45 FADD2 + 60 FADD for R = 31
// (900 FFMA + 150 FADD in scalar form).
template <int R, bool INV, class Emit> __device__ __forceinline__ void dft_odd_prime_emit(float2 (&v)[R], Emit emit)
{
    constexpr int H = (R - 1) / 2;
    pk64 a[H + 1], b[H + 1];
    const pk64 v0 = pk(v[0]);
    pk64 x0 = v0;
#pragma unroll
    for (int j = 1; j <= H; j++) {
        const pk64 p = pk(v[j]), m = pk(v[R - j]);
        a[j] = add2(p, m);
        b[j] = sub2(p, m);
        x0 = add2(x0, a[j]);
    }
    emit(0, upk(x0));
#pragma unroll
    for (int q = 1; q <= H; q++) {
        pk64 c2 = v0, s2 = pk(0.f, 0.f);
#pragma unroll
        for (int j = 1; j <= H; j++) {
            const int k = (j * q) % R;
            const float c = RT<R>::c(k);
            const float s = INV ? RT<R>::s(k) : -RT<R>::s(k);
            c2 = fma2s(a[j], c, c2);
            s2 = fma2s(b[j], s, s2);
        }
        emit(q, add_i(c2, s2));
        emit(R - q, sub_i(c2, s2));
    }
}

// Streaming form for a first stage that reads its inputs from global memory: `load(j)` delivers input j when it is
// needed, and the butterfly keeps the 2 x (R-1)/2 packed output accumulators in registers instead of the a/b half-sums
// ("j outer, q inner").  The inputs need not be resident together, so the loads are issued a batch of BATCH pairs
// ahead of the FFMA2s that consume them and the L2 latency of one batch hides behind the arithmetic of the previous
// one inside the same thread.  Every accumulator receives its terms in the same order (j = 1, 2, ...) as in
// dft_odd_prime_emit: the results are bit-identical.
// The accumulators are a value type so that a caller can keep a finished butterfly in registers and emit it later
// (acq_inverse_lw_kernel's leftover warp holds one per lane across several groups).
template <int R> struct OddPrimeAcc {
    pk64 x0;
    pk64 c2[(R - 1) / 2 + 1], s2[(R - 1) / 2 + 1];   // [0] unused
};
template <int R, bool INV, int BATCH, class Load>
__device__ __forceinline__ void dft_odd_prime_stream_acc(Load load, OddPrimeAcc<R>& h)
{
    constexpr int H = (R - 1) / 2;
    const pk64 v0 = pk(load(0));
    h.x0 = v0;
#pragma unroll
    for (int q = 1; q <= H; q++) {
        h.c2[q] = v0;
        h.s2[q] = pk(0.f, 0.f);
    }
    float2 p[BATCH], m[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
        if (1 + u <= H) {
            p[u] = load(1 + u);
            m[u] = load(R - 1 - u);
        }
    }
#pragma unroll
    for (int j0 = 1; j0 <= H; j0 += BATCH) {
        pk64 a[BATCH], b[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            if (j0 + u <= H) {
                a[u] = add2(pk(p[u]), pk(m[u]));
                b[u] = sub2(pk(p[u]), pk(m[u]));
            }
        }
        // next batch's loads are in flight while this batch is consumed
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            if (j0 + BATCH + u <= H) {
                p[u] = load(j0 + BATCH + u);
                m[u] = load(R - (j0 + BATCH + u));
            }
        }
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            if (j0 + u <= H) {
                const int j = j0 + u;
                h.x0 = add2(h.x0, a[u]);
#pragma unroll
                for (int q = 1; q <= H; q++) {
                    const int k = (j * q) % R;
                    const float c = RT<R>::c(k);
                    const float s = INV ? RT<R>::s(k) : -RT<R>::s(k);
                    h.c2[q] = fma2s(a[u], c, h.c2[q]);
                    h.s2[q] = fma2s(b[u], s, h.s2[q]);
                }
            }
        }
    }
}
template <int R, class Emit> __device__ __forceinline__ void dft_odd_prime_stream_emit(const OddPrimeAcc<R>& h, Emit emit)
{
    constexpr int H = (R - 1) / 2;
    emit(0, upk(h.x0));
#pragma unroll
    for (int q = 1; q <= H; q++) {
        const float2 cc = upk(h.c2[q]), ss = upk(h.s2[q]);
        emit(q, make_float2(cc.x - ss.y, cc.y + ss.x));
        emit(R - q, make_float2(cc.x + ss.y, cc.y - ss.x));
    }
}
template <int R, bool INV, int BATCH, class Load, class Emit>
__device__ __forceinline__ void dft_odd_prime_stream(Load load, Emit emit)
{
    OddPrimeAcc<R> h;
    dft_odd_prime_stream_acc<R, INV, BATCH>(load, h);
    dft_odd_prime_stream_emit<R>(h, emit);
}

}  // namespace gb
#include "dft31_nested.cuh"
namespace gb {

// dft_emit<R,INV>(v, emit): DFT of v with outputs delivered through emit(q, X_q).
// NESTED (Plan::NESTED31) selects the nested 31-point butterfly of dft31_nested.cuh: 27 % fewer FP32 pipe slots, ~100 live
// registers -- a gain where the kernel has them (N = 4092: 1.193 -> 1.070 ms), a loss at the 96-register cap of the
// one-CTA-per-SM plan of 16368 (0.613 -> 0.680 ms), hence a per-plan switch.  Every kernel of a plan uses the same form,
// so the chains of one plan stay bit-identical among themselves.
template <int R, bool INV, bool NESTED = false, class Emit> __device__ __forceinline__ void dft_emit(float2 (&v)[R], Emit emit)
{
    if constexpr (R == 31 && NESTED) {
        dft31_nested_emit<INV>(v, emit);
    } else if constexpr (R == 3 || R == 5 || R == 7 || R == 11 || R == 13 || R == 17 || R == 19 || R == 23 || R == 29 || R == 31) {
        dft_odd_prime_emit<R, INV>(v, emit);
    } else {
        Dft<R, INV>::run(v);
#pragma unroll
        for (int q = 0; q < R; q++) emit(q, v[q]);
    }
}

// In-place form with the same arithmetic as dft_emit<R, INV, NESTED> (kernels that must agree bit for bit with a chain
// that emits, e.g. the fused kernel's last forward stage against acq_forward_kernel).
template <int R, bool INV, bool NESTED> __device__ __forceinline__ void dft_run_like_emit(float2 (&v)[R])
{
    float2 o[R];
    dft_emit<R, INV, NESTED>(v, [&](int q, float2 y) { o[q] = y; });
#pragma unroll
    for (int q = 0; q < R; q++) v[q] = o[q];
}

template <bool INV> struct Dft<3, INV> : DftOddPrime<3, INV> {};
template <bool INV> struct Dft<5, INV> : DftOddPrime<5, INV> {};
template <bool INV> struct Dft<7, INV> : DftOddPrime<7, INV> {};
template <bool INV> struct Dft<11, INV> : DftOddPrime<11, INV> {};
template <bool INV> struct Dft<13, INV> : DftOddPrime<13, INV> {};
template <bool INV> struct Dft<17, INV> : DftOddPrime<17, INV> {};
template <bool INV> struct Dft<19, INV> : DftOddPrime<19, INV> {};
template <bool INV> struct Dft<23, INV> : DftOddPrime<23, INV> {};
template <bool INV> struct Dft<29, INV> : DftOddPrime<29, INV> {};
template <bool INV> struct Dft<31, INV> : DftOddPrime<31, INV> {};

// Composite R = A*B in registers: j = j1*B + j2, q = q1 + A*q2
template <int A, int B, bool INV> struct DftComposite {
    static constexpr int R = A * B;
    static __device__ __forceinline__ void run(float2 (&v)[R])
    {
        float2 t[A][B];
#pragma unroll
        for (int j2 = 0; j2 < B; j2++) {
            float2 c[A];
#pragma unroll
            for (int j1 = 0; j1 < A; j1++) c[j1] = v[j1 * B + j2];
            Dft<A, INV>::run(c);
#pragma unroll
            for (int q1 = 0; q1 < A; q1++) t[q1][j2] = (j2 * q1 == 0) ? c[q1] : mulw<R, INV>(c[q1], (j2 * q1) % R);
        }
#pragma unroll
        for (int q1 = 0; q1 < A; q1++) {
            float2 c[B];
#pragma unroll
            for (int j2 = 0; j2 < B; j2++) c[j2] = t[q1][j2];
            Dft<B, INV>::run(c);
#pragma unroll
            for (int q2 = 0; q2 < B; q2++) v[q1 + A * q2] = c[q2];
        }
    }
};
// Composite R = A*B with gcd(A, B) = 1 in registers, Good-Thomas form: input n = (B n1 + A n2) mod R, output k with
// k = k1 (mod A), k = k2 (mod B).  All index maps are compile-time, so they are pure register renaming and the butterfly
// has NO internal twiddle multiplies (DftComposite needs (A-1)(B-1) of them).
constexpr int pfa_inv_mod(int a, int m)
{
    int x = 1;
    while ((a % m) * x % m != 1) x++;
    return x;
}
template <int A, int B, bool INV> struct DftPfaComposite {
    static constexpr int R = A * B;
    static constexpr int E1 = B * pfa_inv_mod(B, A);   // = 1 (mod A), 0 (mod B)
    static constexpr int E2 = A * pfa_inv_mod(A, B);   // = 0 (mod A), 1 (mod B)
    static __device__ __forceinline__ void run(float2 (&v)[R])
    {
        float2 t[A][B];
#pragma unroll
        for (int n2 = 0; n2 < B; n2++) {
            float2 c[A];
#pragma unroll
            for (int n1 = 0; n1 < A; n1++) c[n1] = v[(B * n1 + A * n2) % R];
            Dft<A, INV>::run(c);
#pragma unroll
            for (int k1 = 0; k1 < A; k1++) t[k1][n2] = c[k1];
        }
#pragma unroll
        for (int k1 = 0; k1 < A; k1++) {
            float2 c[B];
#pragma unroll
            for (int n2 = 0; n2 < B; n2++) c[n2] = t[k1][n2];
            Dft<B, INV>::run(c);
#pragma unroll
            for (int k2 = 0; k2 < B; k2++) v[(k1 * E1 + k2 * E2) % R] = c[k2];
        }
    }
};
#ifdef GB_CT_RADIX12
template <bool INV> struct Dft<12, INV> : DftComposite<4, 3, INV> {};
#else
template <bool INV> struct Dft<12, INV> : DftPfaComposite<4, 3, INV> {};
#endif
template <bool INV> struct Dft<33, INV> : DftPfaComposite<3, 11, INV> {};
template <bool INV> struct Dft<25, INV> : DftComposite<5, 5, INV> {};
template <bool INV> struct Dft<32, INV> : DftComposite<4, 8, INV> {};

// ------------------------------------------------------------------ plan geometry
// A plan is a compile-time list of up to 6 radices (1 = unused), the CTA size and the padding.
template <int N_, int T_, int MINB_, int PAD_, int R0, int R1 = 1, int R2 = 1, int R3 = 1, int R4 = 1, int R5 = 1> struct Plan {
    static constexpr int N = N_;
    static constexpr int T = T_;
    static constexpr int MINB = MINB_;  // CTAs per SM the register allocation must allow
    static constexpr int PAD = PAD_;  // 0: dense; s>0: one complex of padding every 2^s
    static constexpr int NSTAGE = (R0 > 1) + (R1 > 1) + (R2 > 1) + (R3 > 1) + (R4 > 1) + (R5 > 1);
    static constexpr int radix(int s) { return s == 0 ? R0 : s == 1 ? R1 : s == 2 ? R2 : s == 3 ? R3 : s == 4 ? R4 : R5; }
    // block length seen by stage s
    static constexpr int len(int s) { return s == 0 ? N : len(s - 1) / radix(s - 1); }
    static constexpr int sub(int s) { return len(s) / radix(s); }
    static constexpr int LINE = PAD ? N + (N >> PAD) + 1 : N;  // complex elements of shared memory
    // Spectra in global memory (forward spectra, code spectra) are stored scrambled + transposed, [q][b] over the last
    // radix: row q holds output q of every last-stage butterfly b.  Rows are padded to a multiple of 16 complex (128 B)
    // so that a warp's 256-byte row segment never straddles a third cache line (4092: rows of 132 -> 144; 8184: 264 ->
    // 272; every other plan is already aligned and SPEC_LEN == N).
    static constexpr int SPEC_ROWS = radix(NSTAGE - 1);
    static constexpr int SPEC_STRIDE = (N / SPEC_ROWS + 15) / 16 * 16;
    static constexpr int SPEC_LEN = SPEC_ROWS * SPEC_STRIDE;
    // inverse kernel: double-buffer the line if MINB CTAs x 2 lines still fit one SM's shared memory
    static constexpr bool DB = (size_t)MINB_ * 2 * LINE * 8 + (size_t)MINB_ * 1024 <= 227 * 1024;
    static_assert(R0 * R1 * R2 * R3 * R4 * R5 == N, "radices must multiply to N");
    static_assert(NSTAGE >= 2, "need at least two stages");
    __device__ static __forceinline__ int phys(int i) { return PAD ? i + (i >> PAD) : i; }
    // Cooley-Tukey (inter-stage twiddles) unless overridden by PfaPlan
    static constexpr bool PFA = false;
    // inverse kernel, radix-31 first stage: 0 = all inputs resident (dft_odd_prime_emit), n > 0 = streamed in batches of
    // n input pairs (dft_odd_prime_stream)
    static constexpr int STREAM_A = 0;
    // radix-31 butterflies of this plan in the nested form (dft_emit's NESTED): where the kernel has the registers for it
    // (N = 4092: 1.193 -> 1.064 ms, N = 8184: 0.446 -> 0.404 ms; the 96-register cap of 16368's 17 warps spills it: 0.613 -> 0.680)
#ifdef GB_NESTED31_ALL   // A/B builds (tools/build_alt.sh): every plan with a 31-point stage
    static constexpr bool NESTED31 = true;
#else
    static constexpr bool NESTED31 = N_ == 4092 || N_ == 8184;
#endif
};

// Good-Thomas prime-factor plan: the stage radices are pairwise coprime, so with the index maps
//   time      n(l) = sum_s n_s (N / R_s)                      mod N   (Ruritanian)
//   frequency k(l) = sum_s k_s (N / R_s) ((N / R_s)^-1 mod R_s) mod N  (CRT)
// over the digits l = sum_s d_s SUB_s of a line position, W_N^(n k) = prod_s W_(R_s)^(n_s k_s): the transform is a
// plain multi-dimensional DFT and NO inter-stage twiddle exists (no twiddle loads, no complex multiplies between
// stages).  The same stages, butterflies and shared-memory access pattern as the Cooley-Tukey plan; the maps appear
// only where samples enter (the host pre-permutes IQ blocks, wipe-off tables and codes into line order, `npos`) and
// where a code-phase index leaves (reduce_row_to_cell reads n(l)).  The circular correlation is computed in the
// digit domain, which n(l) maps isomorphically onto Z_N, so the result is the same correlation, sample for sample.
constexpr int plan_gcd(int a, int b) { return b == 0 ? a : plan_gcd(b, a % b); }
template <int N_, int T_, int MINB_, int PAD_, int R0, int R1 = 1, int R2 = 1, int R3 = 1, int R4 = 1, int R5 = 1>
struct PfaPlan : Plan<N_, T_, MINB_, PAD_, R0, R1, R2, R3, R4, R5> {
    static constexpr bool PFA = true;
    static constexpr int STREAM_A = 0;
    static_assert(plan_gcd(R0, R1) == 1 && plan_gcd(R0, R2) == 1 && plan_gcd(R0, R3) == 1 && plan_gcd(R0, R4) == 1 &&
                      plan_gcd(R0, R5) == 1 && plan_gcd(R1, R2) == 1 && plan_gcd(R1, R3) == 1 && plan_gcd(R1, R4) == 1 &&
                      plan_gcd(R1, R5) == 1 && plan_gcd(R2, R3) == 1 && plan_gcd(R2, R4) == 1 && plan_gcd(R2, R5) == 1 &&
                      plan_gcd(R3, R4) == 1 && plan_gcd(R3, R5) == 1 && plan_gcd(R4, R5) == 1,
                  "prime-factor plans need pairwise coprime radices");
};

// PfaPlan whose inverse first stage streams its inputs (batches of SB pairs)
template <int SB, int N_, int T_, int MINB_, int PAD_, int R0, int R1 = 1, int R2 = 1, int R3 = 1, int R4 = 1, int R5 = 1>
struct PfaStreamPlan : PfaPlan<N_, T_, MINB_, PAD_, R0, R1, R2, R3, R4, R5> {
    static constexpr int STREAM_A = SB;
};

// ------------------------------------------------------------------ stage workers
// Iterates the butterflies of stage S owned by this thread.  F(b_iter, base, i) is called with the
// compile-time iteration number, the logical index of element j=0 and the position i within the
// sub-block; elements are at base + j*SUB.
template <class P, int S> struct StageGeo {
    static constexpr int R = P::radix(S);
    static constexpr int L = P::len(S);
    static constexpr int SUB = P::sub(S);
    static constexpr int NB = P::N / R;                    // butterflies in the stage
    static constexpr int ITERS = (NB + P::T - 1) / P::T;   // per-thread iterations
    // Stage twiddles W_L^(i*q) are stored per stage as [q-1][i] (coalesced across the threads of a
    // warp, which walk i); TWOFF is where this stage's block starts in the plan's twiddle buffer.
    static constexpr int twoff(int s) { return s == 0 ? 0 : twoff(s - 1) + (P::radix(s - 1) - 1) * P::sub(s - 1); }
    static constexpr int TWOFF = twoff(S);
};
// total number of twiddles of a plan
template <class P> constexpr int plan_twiddle_count() { return StageGeo<P, P::NSTAGE - 1>::TWOFF + (P::radix(P::NSTAGE - 1) - 1) * P::sub(P::NSTAGE - 1); }

// DIF stage S (INV=false: forward sign), smem -> smem
template <class P, int S, bool INV> __device__ __forceinline__ void dif_stage(float2* __restrict__ s, const float2* __restrict__ tw)
{
    using G = StageGeo<P, S>;
#pragma unroll 1
    for (int it = 0; it < G::ITERS; it++) {
        const int b = threadIdx.x + it * P::T;
        if (G::NB % P::T == 0 || b < G::NB) {
            const int blk = b / G::SUB, i = b - blk * G::SUB;
            const int base = blk * G::L + i;
            float2 v[G::R];
#pragma unroll
            for (int j = 0; j < G::R; j++) v[j] = s[P::phys(base + j * G::SUB)];
            Dft<G::R, INV>::run(v);
            s[P::phys(base)] = v[0];
#pragma unroll
            for (int q = 1; q < G::R; q++) {
                if (G::SUB == 1 || P::PFA) {
                    s[P::phys(base + q * G::SUB)] = v[q];
                } else {
                    const float2 w = __ldg(&tw[G::TWOFF + (q - 1) * G::SUB + i]);
                    s[P::phys(base + q * G::SUB)] = INV ? cmul_conj(v[q], w) : cmul(v[q], w);
                }
            }
        }
    }
}

// DIT stage S (INV=true: inverse sign), smem -> smem; exact inverse (x radix) of dif_stage<P,S,false>
template <class P, int S, bool INV> __device__ __forceinline__ void dit_stage(float2* __restrict__ s, const float2* __restrict__ tw)
{
    using G = StageGeo<P, S>;
#pragma unroll 1
    for (int it = 0; it < G::ITERS; it++) {
        const int b = threadIdx.x + it * P::T;
        if (G::NB % P::T == 0 || b < G::NB) {
            const int blk = b / G::SUB, i = b - blk * G::SUB;
            const int base = blk * G::L + i;
            float2 v[G::R];
            v[0] = s[P::phys(base)];
#pragma unroll
            for (int q = 1; q < G::R; q++) {
                const float2 u = s[P::phys(base + q * G::SUB)];
                if (G::SUB == 1 || P::PFA) {
                    v[q] = u;
                } else {
                    const float2 w = __ldg(&tw[G::TWOFF + (q - 1) * G::SUB + i]);
                    v[q] = INV ? cmul_conj(u, w) : cmul(u, w);
                }
            }
            Dft<G::R, INV>::run(v);
#pragma unroll
            for (int j = 0; j < G::R; j++) s[P::phys(base + j * G::SUB)] = v[j];
        }
    }
}

// DIT stage S of a prime-factor plan (no twiddles) with ONE sub-block per warp pass: the SUB <= 32 butterflies of a block
// read/write SUB consecutive complex per element index, so no request straddles a block boundary.  dit_stage's
// thread-linear assignment does (SUB = 31 against 32 lanes) and pays a third wavefront on every shared-memory request of
// the stage -- 19 % of the kernel's shared-memory wavefronts, on the busiest unit (L1 data pipe 78 %).  The number of
// warp passes is the same (12 blocks of 31 = 372 butterflies = 11.6 warps).
template <class P, int S, bool INV, int NWARPS> __device__ __forceinline__ void dit_stage_rows(float2* __restrict__ s)
{
    using G = StageGeo<P, S>;
    static_assert(P::PFA && G::SUB <= 32, "one sub-block per warp pass");
    constexpr int NBLK = P::N / G::L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
    for (int blk = warp; blk < NBLK; blk += NWARPS) {
        if (lane < G::SUB) {
            const int base = blk * G::L + lane;
            float2 v[G::R];
#pragma unroll
            for (int q = 0; q < G::R; q++) v[q] = s[P::phys(base + q * G::SUB)];
            Dft<G::R, INV>::run(v);
#pragma unroll
            for (int j = 0; j < G::R; j++) s[P::phys(base + j * G::SUB)] = v[j];
        }
    }
}

// DIF stages [S, LAST) with a barrier after each
template <class P, int S, int LAST, bool INV> struct DifRange {
    static __device__ __forceinline__ void run(float2* s, const float2* tw)
    {
        if constexpr (S < LAST) {
            dif_stage<P, S, INV>(s, tw);
            __syncthreads();
            DifRange<P, S + 1, LAST, INV>::run(s, tw);
        }
    }
};
// DIT stages S, S-1, ..., LAST+1, barrier after each
template <class P, int S, int LAST, bool INV> struct DitRange {
    static __device__ __forceinline__ void run(float2* s, const float2* tw)
    {
        if constexpr (S > LAST) {
            // (dit_stage_rows was measured in the one-CTA-per-SM kernels of 8184 / 16368 too: 0.609 -> 0.621 ms on config 1,
            // so only the leftover-warp kernel uses it)
            dit_stage<P, S, INV>(s, tw);
            __syncthreads();
            DitRange<P, S - 1, LAST, INV>::run(s, tw);
        }
    }
};

}  // namespace gb
