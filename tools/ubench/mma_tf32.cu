// Rate of the legacy warp-level tensor path (mma.sync.m16n8k8 tf32, SASS HMMA.1688.F32.TF32) on sm_100a, alone and
// next to packed FP32 work: decides whether the radix-31 stage of acq_inverse_lw_kernel can move to the tensor pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32 mma_tf32.cu && ./mma_tf32
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ void mma(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// NCH independent accumulator chains per warp, NF FFMA2 per mma
template <int NCH, int NF> __global__ void __launch_bounds__(1024) k(float* out, int iters, float fa, float fb)
{
    float d[NCH][4];
    unsigned a[4], b0 = threadIdx.x * 77u, b1 = threadIdx.x * 131u;
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = (threadIdx.x + i) * 2654435761u & 0x3fffe000u;
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
        for (int i = 0; i < 4; i++) d[c][i] = 0.f;
    u64 x[8];
    const u64 aa = pack(fa, fa), bb = pack(fb, fb);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = pack(threadIdx.x + i, threadIdx.x - i);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                mma(d[c], a, b0, b1);
#pragma unroll
                for (int f = 0; f < NF; f++) x[(c * NF + f) & 7] = fma2(x[(c * NF + f) & 7], aa, bb);
            }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    if (s == 1234.5f) out[0] = s;
}
template <int NCH, int NF> void run(int sms, int threads, int bps, double mhz)
{
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2048;
    k<NCH, NF><<<sms * bps, threads>>>(d, 16, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NCH, NF><<<sms * bps, threads>>>(d, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)bps * threads / 32;
    const double n_mma = warps * iters * 4.0 * NCH;          // per SM
    const double clk = ms * 1e-3 * mhz * 1e6;
    printf("chains %d ffma2/mma %d warps/SM %4.0f : %.3f ms  %.2f clk per mma per SM (%.0f MAC/clk/SM, %.1f TFLOP/s tf32)  fp32 %.1f TFLOP/s\n",
           NCH, NF, warps, ms, clk / n_mma, 1024.0 * n_mma / clk, 2048.0 * n_mma * sms / (ms * 1e-3) * 1e-12,
           4.0 * NF * n_mma * 32 * sms / (ms * 1e-3) * 1e-12);
    cudaFree(d);
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz nominal\n", p.name, p.multiProcessorCount, mhz);
    const int sms = p.multiProcessorCount;
    run<1, 0>(sms, 128, 1, mhz);
    run<4, 0>(sms, 128, 1, mhz);
    run<8, 0>(sms, 128, 1, mhz);
    run<8, 0>(sms, 256, 1, mhz);
    run<8, 0>(sms, 512, 1, mhz);
    run<4, 0>(sms, 512, 2, mhz);
    run<4, 1>(sms, 512, 1, mhz);
    run<4, 2>(sms, 512, 1, mhz);
    run<4, 4>(sms, 512, 1, mhz);
    run<4, 8>(sms, 512, 1, mhz);
    run<4, 4>(sms, 160, 3, mhz);
    run<4, 8>(sms, 160, 3, mhz);
    return 0;
}
