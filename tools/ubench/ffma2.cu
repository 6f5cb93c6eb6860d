// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) throughput on sm_100a, alone and mixed with integer/LDS work.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters)
{
    __shared__ float sm[256];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    if (MODE == 0) {  // scalar FFMA, 16 chains
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int u = 0; u < 16; u++)
#pragma unroll
                for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += x[i];
        if (s == 1234.5f) out[0] = s;
    } else if (MODE == 1 || MODE == 2 || MODE == 3) {  // FFMA2, 8 packed chains (= 16 scalar chains)
        u64 x[8];
        const u64 aa = pack(a, a), bb = pack(b, b);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = pack(threadIdx.x + i, threadIdx.x - i);
        int acc = threadIdx.x;
        float l = 0.f;
#pragma unroll 1
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int u = 0; u < 16; u++) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = fma2(x[i], aa, bb);
                if (MODE == 2) {  // + 8 integer ops per 8 FFMA2
#pragma unroll
                    for (int i = 0; i < 16; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc) : "r"(u + i), "r"(it));
                }
                if (MODE == 3) {  // + 4 LDS per 8 FFMA2
#pragma unroll
                    for (int i = 0; i < 4; i++) l += sm[(threadIdx.x + i * 32 + u) & 255];
                }
            }
        float s = l + acc;
#pragma unroll
        for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
        if (s == 1234.5f) out[0] = s;
    } else if (MODE == 4) {  // scalar FFMA 16 chains + 8 integer ops per 16 FFMA
        float x[16];
        int acc = threadIdx.x;
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int u = 0; u < 16; u++) {
#pragma unroll
                for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
#pragma unroll
                for (int i = 0; i < 16; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc) : "r"(u + i), "r"(it));
            }
        float s = acc;
#pragma unroll
        for (int i = 0; i < 16; i++) s += x[i];
        if (s == 1234.5f) out[0] = s;
    }
}

template <int MODE> void run(const char* name, int sms)
{
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 8, iters = 512;
    float best = 1e9f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(d, 1.0000001f, 1e-9f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    const double fmas = 16.0 * 16.0 * iters * (double)blocks * 256;
    printf("%-34s %.3f ms  %.2f TFLOP/s (FMA=2 flop)  %.1f FMA/clk/SM @1.965GHz\n", name, best, 2 * fmas / best / 1e9,
           fmas / (best * 1e-3) / 1.965e9 / sms);
    cudaFree(d);
}

int main()
{
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("FFMA x16 chains", sms);
    run<1>("FFMA2 x8 packed chains", sms);
    run<4>("FFMA x16 + 16 LOP3", sms);
    run<2>("FFMA2 x8 + 16 LOP3", sms);
    run<3>("FFMA2 x8 + 4 LDS", sms);
    return 0;
}
