// Does FFMA2 free issue slots?  8 FFMA2 (or 16 FFMA) + M independent integer LOP3 per iteration.
// ALU (LOP3) is a 16-lane pipe (2 cycles per warp instruction), FP32 32 lanes (FFMA 1 cycle, FFMA2 2 cycles of pipe).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int PACKED, int NI> __global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters)
{
    int acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    float s = 0.f;
    if (PACKED) {
        u64 x[8];
        const u64 aa = pack(a, a), bb = pack(b, b);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = pack(threadIdx.x + i, threadIdx.x - i);
#pragma unroll 1
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int u = 0; u < 16; u++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    x[i] = fma2(x[i], aa, bb);
                    if (i < NI) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[i]) : "r"(u + i), "r"(it));
                }
            }
#pragma unroll
        for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    } else {
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int u = 0; u < 16; u++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    x[2 * i] = fmaf(x[2 * i], a, b);
                    x[2 * i + 1] = fmaf(x[2 * i + 1], a, b);
                    if (i < NI) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[i]) : "r"(u + i), "r"(it));
                }
            }
#pragma unroll
        for (int i = 0; i < 16; i++) s += x[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    if (s == 1234.5f) out[0] = s;
}
template <int PACKED, int NI> void run(int sms)
{
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 8, iters = 512;
    float best = 1e9f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        k<PACKED, NI><<<blocks, 256>>>(d, 1.0000001f, 1e-9f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    // cycles per SMSP per (16 FMA-lane-ops + NI LOP3) group: 16 warps per SMSP
    const double groups = 16.0 * iters * 16.0;  // per SMSP: u-iterations x warps
    printf("%s + %d LOP3 : %.3f ms  -> %.1f cycles per group of 16 FMA (+%d LOP3) per SMSP\n", PACKED ? "8 FFMA2" : "16 FFMA",
           NI, best, best * 1e-3 * 1.965e9 / groups, NI);
    cudaFree(d);
}
int main()
{
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0, 0>(sms); run<1, 0>(sms);
    run<0, 4>(sms); run<1, 4>(sms);
    run<0, 8>(sms); run<1, 8>(sms);
    return 0;
}
