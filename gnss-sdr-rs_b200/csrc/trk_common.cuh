// trk_common.cuh -- device helpers shared by the tracking kernels (trk_kernels.cu, trk_ws.cu).  Both are compiled
// with -fmad=false: the epoch-end scalar arithmetic must round exactly like the reference's f32 code
// (do_tracking.rs:231-302); FMAs are written explicitly where they are wanted.
#pragma once
#include "trk_kernels.cuh"

namespace gb {

static __device__ __constant__ float kTwoPi = 6.28318530717958647692f;  // 2.0 * std::f32::consts::PI

// Rust `as usize` for f32: saturating, NaN -> 0 (Q7)
__device__ __forceinline__ unsigned long long f32_as_usize(float v)
{
    if (!(v > 0.f)) return 0ull;
    if (v >= 18446744073709551616.f) return ~0ull;
    return (unsigned long long)v;
}

// get_ca_chip (do_tracking.rs:274-277): floor, saturating cast, % 1023
__device__ __forceinline__ float ca_chip(const float* __restrict__ row, float phase)
{
    const float f = floorf(phase);
    unsigned idx = f > 0.f ? (f < 4.0e9f ? (unsigned)f : (unsigned)(f32_as_usize(f) % 1023ull)) : 0u;
    idx = idx % 1023u;
    return row[idx];
}

// x % 1023.0 (fmodf is exact; fast path for the only range the loops ever produce)
__device__ __forceinline__ float mod1023(float t)
{
    if (t >= 0.f && t < 2046.f) return t >= 1023.f ? t - 1023.f : t;
    return fmodf(t, 1023.f);
}

template <int MODE> __device__ __forceinline__ void carrier(float phase, float& c, float& s)
{
    if (MODE == GB_TRK_ORDERED) {
        // glibc's sinf/cosf are (nearly always) correctly rounded; so is the f64 result rounded to f32
        double sd, cd;
        sincos((double)phase, &sd, &cd);
        c = (float)cd;
        s = (float)sd;
    } else {
        sincosf(phase, &s, &c);
    }
}

// exact floor of x in (-1, 2^22) as an int without the conversion unit: a round-down add of 2^23 leaves floor(x) in
// the mantissa (FADD.RM on the FP32 pipe + one integer subtract); negative x gives a negative result (callers clamp).
__device__ __forceinline__ int floor_small(float x) { return __float_as_int(__fadd_rd(x, 8388608.0f)) - 0x4B000000; }
// rintf for |x| < 2^22 (round-to-nearest-even through the 1.5 * 2^23 magic constant), again FP32-pipe only
__device__ __forceinline__ float rint_small(float x) { return __fadd_rn(__fadd_rn(x, 12582912.0f), -12582912.0f); }

// fmodf(x, y) for y > 0 and |x| < 2^20 y, bit-exact (fmod's result is always representable, so ONE fused
// multiply-add from the original operand is exact once the integer quotient is right; a quotient that the rounded
// product x * (1/y) puts off by one is detected by the sign / size of the remainder and the FMA redone from |x|).
// Host restatement checked against glibc fmodf: tests/cpp/test_fmod_small.c.
__device__ __forceinline__ float fmod_small(float x, float y, float inv_y)
{
    const float ax = fabsf(x);
    const float kf = __fadd_rn(__fadd_rd(__fmul_rn(ax, inv_y), 8388608.0f), -8388608.0f);   // floor(ax / y), maybe +-1
    float r = fmaf(-kf, y, ax);
    if (r >= y) r = fmaf(-(kf + 1.0f), y, ax);
    else if (r < 0.f) r = fmaf(-(kf - 1.0f), y, ax);
    return copysignf(r, x);
}

// packed FP32 helpers (Blackwell FFMA2): (acc.x, acc.y) += (a.x, a.y) * s in one instruction
typedef unsigned long long pk64;
__device__ __forceinline__ pk64 pk2(float x, float y)
{
    pk64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void upk2(pk64 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ pk64 fma2s(pk64 a, float s, pk64 c)
{
    pk64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(pk2(s, s)), "l"(c));
    return r;
}
__device__ __forceinline__ pk64 add2(pk64 a, pk64 b)
{
    pk64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk64 add2_rm(pk64 a, pk64 b)   // round towards -inf (floor through the 2^23 magic constant)
{
    pk64 r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk64 fma2_rm(pk64 a, pk64 b, pk64 c)   // a * b + c rounded towards -inf (one rounding)
{
    pk64 r;
    asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
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
__device__ __forceinline__ float4 lds_f32x4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

}  // namespace gb
