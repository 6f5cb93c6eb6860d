// fine_doppler.cu -- sub-bin carrier estimate of an acquired satellite, sm_100a.
//
// Replaces finer_doppler (acquisition_bk.rs:215-302): the code-stripped (long_ms-1) ms of signal, zero-padded to
// M = 8 * next_power_of_two(L) samples, is transformed and the first index of the largest |X[k]| is returned.
//
// The reference runs one M-point Radix4 FFT over a buffer that is > 87 % zeros (M = 2^21 at 16.3676 Msps).  Here the
// zero padding is never materialised.  Because the signal occupies only n < L <= P2 = M/8,
//     X[8 q + r] = sum_{n < L} (x[n] W_M^{n r}) W_P2^{n q},        r = 0..7,
// i.e. eight independent P2-point transforms of the same L samples, each pre-rotated by a fraction r/8 of a bin.
// Each P2-point transform is a two-pass (four-step) FFT, P2 = A x B, n = a B + b, q = ka + A kb:
//   pass 1 (fine_cols_kernel): strip the code, remove the mean, rotate by W_M^{n r}; A-point column FFTs in shared
//           memory (TB columns per CTA, coalesced over b); multiply by W_P2^{b ka}; store Y[ka][b];
//   pass 2 (fine_rows_kernel): B-point row FFTs in shared memory, |.| = hypotf, and the arg-max folded into one
//           64-bit atomicMax per CTA (key = magnitude bits : ~index, so ties go to the lowest index like
//           Iterator::find at :277-280).  Nothing but Y (L2-resident) and 8 bytes per request ever reach HBM.
// The shared-memory FFT is an in-place radix-2 DIF (bit-reversed output, undone in the index arithmetic); the
// transforms here are ~2 % of an acquisition search, so clarity wins over radix-4/8 butterflies.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fine_doppler.cuh"

namespace gb {

namespace {

constexpr int kThreads = 256;
constexpr int kTileElems = 4096;  // complex elements of shared memory per CTA (32 KiB)

__device__ __forceinline__ float2 cmulf(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// exp(-2 pi i num / den), den a power of two, num < den <= 2^24: num/den is exact in f32
__device__ __forceinline__ float2 unit_root(unsigned num, unsigned den)
{
    float s, c;
    sincospif(-2.0f * ((float)num / (float)den), &s, &c);
    return make_float2(c, s);
}

// In-place radix-2 DIF over `nfft` transforms of length n = 1 << logn held in shared memory; element e of transform f
// is at s[e * se + f * sf].  tw[j] = exp(-2 pi i j / n), j < n/2.  Output position p holds frequency bitrev(p).
__device__ __forceinline__ void fft_dif_smem(float2* s, const float2* tw, int logn, int nfft, int se, int sf)
{
    const int n = 1 << logn, half = n >> 1;
    for (int st = 0; st < logn; st++) {
        const int h = half >> st;  // butterfly span
        for (int t = threadIdx.x; t < half * nfft; t += kThreads) {
            // consecutive threads take consecutive transforms when sf == 1 (columns), consecutive butterflies otherwise
            int f, bt;
            if (sf == 1) {
                f = t % nfft;
                bt = t / nfft;
            } else {
                f = t / half;
                bt = t - f * half;
            }
            const int blk = bt / h, j = bt - blk * h;
            const int i0 = blk * 2 * h + j;
            float2* p0 = s + i0 * se + f * sf;
            float2* p1 = p0 + h * se;
            const float2 u = *p0, v = *p1;
            *p0 = make_float2(u.x + v.x, u.y + v.y);
            *p1 = cmulf(make_float2(u.x - v.x, u.y - v.y), tw[j << st]);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

}  // namespace

// mean of the n_long samples (acquisition_bk.rs:234): f64 partial sums, one CTA
__global__ void __launch_bounds__(1024) fine_mean_kernel(const float2* __restrict__ x, unsigned long long start,
                                                        unsigned long long mask, unsigned long long n_long,
                                                        float2* __restrict__ mean_out)
{
    __shared__ double sre[32], sim[32];
    double a = 0.0, b = 0.0;
    for (unsigned long long i = threadIdx.x; i < n_long; i += blockDim.x) {
        const float2 v = x[(start + i) & mask];
        a += (double)v.x;
        b += (double)v.y;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sre[threadIdx.x >> 5] = a;
        sim[threadIdx.x >> 5] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            ta += sre[w];
            tb += sim[w];
        }
        // the reference divides the f32 sum by len as f32
        mean_out[0] = make_float2(__fdiv_rn((float)ta, (float)n_long), __fdiv_rn((float)tb, (float)n_long));
    }
}

// pass 1: grid (B / TB, 8, n_req)
__global__ void __launch_bounds__(kThreads) fine_cols_kernel(const FineArgs a)
{
    extern __shared__ float2 sm[];
    const int A = 1 << a.log_a, B = 1 << a.log_b, TB = kTileElems >> a.log_a;
    float2* tile = sm;              // [A][TB]
    float2* tw = sm + kTileElems;   // [A/2]
    const int r = blockIdx.y, q = blockIdx.z, b0 = blockIdx.x * TB;
    const unsigned P2 = 1u << (a.log_a + a.log_b), M = P2 << 3;
    const int8_t* __restrict__ code = a.codes + (size_t)q * 1023;
    const unsigned long long cp = a.code_phase[q];
    const float2 mean = a.mean[0];
    for (int j = threadIdx.x; j < A / 2; j += kThreads) tw[j] = unit_root((unsigned)j, (unsigned)A);
    for (int t = threadIdx.x; t < A * TB; t += kThreads) {
        const int c = t % TB, aa = t / TB;
        const unsigned n = (unsigned)aa * (unsigned)B + (unsigned)(b0 + c);
        float2 v = make_float2(0.f, 0.f);
        if (n < a.use) {
            // code index in the reference's f32 arithmetic (:241-247): floor((x as f32 * 1.023e6) / fs) % 1023
            const float fi = floorf(__fdiv_rn(__fmul_rn((float)n, 1.023e6f), a.fs));
            const unsigned ci = (unsigned)fi % 1023u;
            const float ch = (float)code[ci];
            const float2 xs = a.x[(a.start + cp + n) & a.mask];
            const float2 y = make_float2(__fsub_rn(xs.x, mean.x) * ch, __fsub_rn(xs.y, mean.y) * ch);
            v = cmulf(y, unit_root((n * (unsigned)r) & (M - 1u), M));
        }
        tile[aa * TB + c] = v;
    }
    __syncthreads();
    fft_dif_smem(tile, tw, a.log_a, TB, TB, 1);
    float2* __restrict__ Y = a.Y + ((size_t)q * 8 + r) * P2;
    for (int t = threadIdx.x; t < A * TB; t += kThreads) {
        const int c = t % TB, p = t / TB;
        const unsigned ka = bitrev((unsigned)p, a.log_a);
        const unsigned b = (unsigned)(b0 + c);
        Y[(size_t)ka * B + b] = cmulf(tile[p * TB + c], unit_root((ka * b) & (P2 - 1u), P2));
    }
}

// pass 2: grid (A / TR, 8, n_req)
__global__ void __launch_bounds__(kThreads) fine_rows_kernel(const FineArgs a)
{
    extern __shared__ float2 sm[];
    const int A = 1 << a.log_a, B = 1 << a.log_b, TR = kTileElems >> a.log_b;
    float2* tile = sm;             // [TR][B]
    float2* tw = sm + kTileElems;  // [B/2]
    __shared__ unsigned long long best[kThreads / 32];
    const int r = blockIdx.y, q = blockIdx.z, ka0 = blockIdx.x * TR;
    const unsigned P2 = 1u << (a.log_a + a.log_b);
    const float2* __restrict__ Y = a.Y + ((size_t)q * 8 + r) * P2 + (size_t)ka0 * B;
    for (int j = threadIdx.x; j < B / 2; j += kThreads) tw[j] = unit_root((unsigned)j, (unsigned)B);
    for (int t = threadIdx.x; t < TR * B; t += kThreads) tile[t] = Y[t];
    __syncthreads();
    fft_dif_smem(tile, tw, a.log_b, TR, 1, B);
    unsigned long long key = 0ull;
    for (int t = threadIdx.x; t < TR * B; t += kThreads) {
        const int row = t >> a.log_b, p = t & (B - 1);
        const unsigned kb = bitrev((unsigned)p, a.log_b);
        const unsigned k = (((unsigned)(ka0 + row) + (unsigned)A * kb) << 3) + (unsigned)r;
        const float2 v = tile[t];
        float m = hypotf(v.x, v.y);  // Complex::abs (:274)
        if (!(m == m)) m = 0.f;      // f32::max never returns a NaN operand (:276)
        if (a.mag_out) a.mag_out[(size_t)q * (P2 << 3) + k] = m;
        const unsigned long long kk = ((unsigned long long)__float_as_uint(m) << 32) | (unsigned long long)(0xffffffffu - k);
        key = kk > key ? kk : key;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; w++) key = best[w] > key ? best[w] : key;
        atomicMax(&a.best[q], key);
    }
}

size_t fine_smem_bytes(int log_a, int log_b)
{
    const int big = log_a > log_b ? log_a : log_b;
    return sizeof(float2) * ((size_t)kTileElems + ((size_t)1 << (big - 1)));
}

cudaError_t fine_launch(const FineArgs& a, int n_req, const float2* x, unsigned long long start, unsigned long long mask,
                        unsigned long long n_long, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(a.best, 0, sizeof(unsigned long long) * n_req, st);
    if (e != cudaSuccess) return e;
    fine_mean_kernel<<<1, 1024, 0, st>>>(x, start, mask, n_long, a.mean);
    const size_t smem = fine_smem_bytes(a.log_a, a.log_b);
    const int TB = kTileElems >> a.log_a, TR = kTileElems >> a.log_b;
    dim3 g1((1u << a.log_b) / TB, 8, n_req), g2((1u << a.log_a) / TR, 8, n_req);
    fine_cols_kernel<<<g1, kThreads, smem, st>>>(a);
    fine_rows_kernel<<<g2, kThreads, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace gb
