// acq_generic.cu -- any-length plan: the fallback behind gb_acq_configure / gb_fft_* for code periods that have no
// compiled shared-memory plan (acq_kernels.cu).  The reference plans ANY fft_size through rustfft's planner
// (do_acquisition.rs:131-142, fft.rs:12-15); here every length without a tuned plan runs as Bluestein's chirp-z over
// a power-of-two Stockham FFT kept in global memory (L2-resident for the sizes involved):
//
//   X[k] = w[k] * sum_n (x[n] w[n]) conj(w[k - n]),   w[n] = exp(-j pi n^2 / N)
//
// i.e. one length-M circular convolution, M = 2^ceil(log2(2N - 1)), = FFT_M, pointwise product with the precomputed
// spectrum of the chirp, inverse FFT_M.  Radix-4 Stockham passes (autosort: no bit reversal), a radix-2 pass when log2 M
// is odd, twiddles from sincospi on exact dyadic arguments.  Not a tuned path -- correctness for every length is its job;
// the sample rates of the BASELINE configurations all have shared-memory plans.
//
// Acquisition chain for one coherent group g and one slab of Doppler bins (same structure as the tuned shared chain:
// the forward path once per bin, the inverse path per (PRN, bin)):
//   pre   a[d][n]    = (sum_c rot[d][c] * wipe(x[g n_coh + c][n], table[d][n])) * w[n]        n < N, zero padded to M
//   conv  c[d]       = IFFT_M(FFT_M(a[d]) * Bf)
//   mid   a2[p][d][k] = (c[d][k] / M) * conj(C_p[k])                                          k < N, zero padded
//         (forward post-chirp w[k] times inverse pre-chirp conj(w[k]) = 1: neither is applied)
//   conv  c2[p][d]   = IFFT_M(FFT_M(a2[p][d]) * Bi)
//   acc   P[p][d][n] += |c2[p][d][n] / M|^2                                                   (|post-chirp| = 1)
// and reduce_rows_kernel turns every accumulated row into its cell.  Scalings by 1 / M are exact (powers of two).
#include "acq_generic.cuh"

#include <math.h>

#include <vector>

namespace gb {

namespace {

template <typename T> struct V2;
template <> struct V2<float> { typedef float2 type; };
template <> struct V2<double> { typedef double2 type; };

template <typename T> __device__ __forceinline__ typename V2<T>::type mk(T x, T y)
{
    typename V2<T>::type r;
    r.x = x; r.y = y;
    return r;
}
template <typename T2> __device__ __forceinline__ T2 cmulg(T2 a, T2 b)
{
    T2 r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
__device__ __forceinline__ void sincospi_t(float x, float* s, float* c) { sincospif(x, s, c); }
__device__ __forceinline__ void sincospi_t(double x, double* s, double* c) { sincospi(x, s, c); }

// One Stockham radix-R pass of a batch of length-M transforms (R = 2 or 4), x -> y.  Thread j owns butterfly j of its
// transform: inputs x[j + r M / R], twiddles exp(sign 2 pi i r k / (Ns R)) with k = j mod Ns, outputs y[(j - k) R + k + r Ns].
template <typename T, int R> __global__ void __launch_bounds__(256)
stockham_pass(const typename V2<T>::type* __restrict__ x, typename V2<T>::type* __restrict__ y, int M, int Ns, int inverse)
{
    typedef typename V2<T>::type T2;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = M / R;
    if (j >= q) return;
    const size_t off = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * (size_t)M;
    const int k = j & (Ns - 1);
    T2 v[R];
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = x[off + j + (size_t)r * q];
    if (Ns > 1) {
        // angle / pi = sign * 2 r k / (Ns R): a dyadic rational, exact in T
        const T base = (T)(2 * k) / (T)(Ns * R);
#pragma unroll
        for (int r = 1; r < R; r++) {
            T s, c;
            sincospi_t(base * (T)r, &s, &c);
            const T2 w = mk<T>(c, inverse ? s : -s);
            v[r] = cmulg(v[r], w);
        }
    }
    T2 o[R];
    if (R == 2) {
        o[0] = mk<T>(v[0].x + v[1].x, v[0].y + v[1].y);
        o[1] = mk<T>(v[0].x - v[1].x, v[0].y - v[1].y);
    } else {
        const T2 a = mk<T>(v[0].x + v[2].x, v[0].y + v[2].y), b = mk<T>(v[0].x - v[2].x, v[0].y - v[2].y);
        const T2 c = mk<T>(v[1].x + v[3].x, v[1].y + v[3].y);
        const T2 dd = mk<T>(v[1].x - v[3].x, v[1].y - v[3].y);
        // forward: d = -i (v1 - v3); inverse: d = +i (v1 - v3)
        const T2 d = inverse ? mk<T>(-dd.y, dd.x) : mk<T>(dd.y, -dd.x);
        o[0] = mk<T>(a.x + c.x, a.y + c.y);
        o[1] = mk<T>(b.x + d.x, b.y + d.y);
        o[2] = mk<T>(a.x - c.x, a.y - c.y);
        o[3] = mk<T>(b.x - d.x, b.y - d.y);
    }
    const size_t j0 = (size_t)(j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) y[off + j0 + (size_t)r * Ns] = o[r];
}

// batch of unnormalised length-M FFTs, ping-pong between a and b; returns the buffer that holds the result
template <typename T>
typename V2<T>::type* fft_pow2(typename V2<T>::type* a, typename V2<T>::type* b, int M, size_t batch, int inverse, cudaStream_t st,
                               cudaError_t* err)
{
    typedef typename V2<T>::type T2;
    T2 *src = a, *dst = b;
    const unsigned by = (unsigned)(batch < 32768 ? batch : 32768);
    const unsigned bz = (unsigned)((batch + by - 1) / by);
    if ((size_t)by * bz != batch) { *err = cudaErrorInvalidValue; return a; }   // callers pass batches that factor
    int Ns = 1;
    while (Ns < M) {
        const int R = (M / Ns) % 4 == 0 ? 4 : 2;
        dim3 grid((unsigned)((M / R + 255) / 256), by, bz);
        if (R == 4) stockham_pass<T, 4><<<grid, 256, 0, st>>>(src, dst, M, Ns, inverse);
        else stockham_pass<T, 2><<<grid, 256, 0, st>>>(src, dst, M, Ns, inverse);
        Ns *= R;
        T2* t = src; src = dst; dst = t;
    }
    *err = cudaGetLastError();
    return src;
}

// a[b][m] = x[b][m] * chirp[m] (conjugated for the inverse transform), zero padded to M
template <typename T> __global__ void bluestein_pre(const typename V2<T>::type* __restrict__ x, const typename V2<T>::type* __restrict__ w,
                                                    typename V2<T>::type* __restrict__ a, int N, int M, int inverse)
{
    typedef typename V2<T>::type T2;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const size_t b = blockIdx.y;
    T2 v = mk<T>((T)0, (T)0);
    if (m < N) {
        T2 c = w[m];
        if (inverse) c.y = -c.y;
        v = cmulg(x[b * N + m], c);
    }
    a[b * M + m] = v;
}
template <typename T2> __global__ void pointwise_mul(T2* __restrict__ a, const T2* __restrict__ k, int M)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const size_t o = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * (size_t)M + m;
    a[o] = cmulg(a[o], k[m]);
}
// X[b][k] = chirp[k] * c[b][k] / M
template <typename T> __global__ void bluestein_post(const typename V2<T>::type* __restrict__ c, const typename V2<T>::type* __restrict__ w,
                                                     typename V2<T>::type* __restrict__ out, int N, int M, int inverse, T inv_m)
{
    typedef typename V2<T>::type T2;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const size_t b = blockIdx.y;
    T2 ch = w[k];
    if (inverse) ch.y = -ch.y;
    T2 v = c[b * M + k];
    v.x *= inv_m; v.y *= inv_m;
    out[b * N + k] = cmulg(v, ch);
}

// ---------------------------------------------------------------- acquisition pointwise kernels (f32)
__device__ __forceinline__ float2 wipe_ref(float2 x, float2 w)   // multiply_simd_block (doppler_shift.rs:43-58): no FMA
{
    return make_float2(__fadd_rn(__fmul_rn(x.x, w.x), -__fmul_rn(x.y, w.y)), __fadd_rn(__fmul_rn(x.x, w.y), __fmul_rn(x.y, w.x)));
}
__global__ void gen_fwd_pre(const float2* __restrict__ iq, unsigned long long iq_start, unsigned long long iq_mask,
                            const float2* __restrict__ tables, const float2* __restrict__ rot, const float2* __restrict__ chirp,
                            float2* __restrict__ a, int N, int M, int n_coh, int g, int d_lo)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= M) return;
    const int dl = blockIdx.y, d = d_lo + dl;
    float2 v = make_float2(0.f, 0.f);
    if (n < N) {
        const float2 t = __ldg(&tables[(size_t)d * N + n]);
        if (n_coh == 1) {
            v = wipe_ref(__ldg(&iq[(iq_start + (unsigned long long)g * N + n) & iq_mask]), t);
        } else {
            for (int c = 0; c < n_coh; c++) {
                const float2 r = __ldg(&rot[(size_t)d * n_coh + c]);
                const float2 u = wipe_ref(__ldg(&iq[(iq_start + (unsigned long long)(g * n_coh + c) * N + n) & iq_mask]), t);
                v.x += u.x * r.x - u.y * r.y;
                v.y += u.x * r.y + u.y * r.x;
            }
        }
        v = cmulg(v, __ldg(&chirp[n]));
    }
    a[(size_t)dl * M + n] = v;
}
// a2[(pi * n_d + dl)][k] = (c[dl][k] / M) * conj(C[rows[pi]][k]), zero padded
__global__ void gen_mid(const float2* __restrict__ c, const float2* __restrict__ code_fft, const int* __restrict__ rows,
                        float2* __restrict__ a2, int N, int M, int n_d, float inv_m)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const int dl = blockIdx.y, pi = blockIdx.z;
    float2 v = make_float2(0.f, 0.f);
    if (k < N) {
        const float2 s = c[(size_t)dl * M + k];
        const float2 cc = __ldg(&code_fft[(size_t)rows[pi] * N + k]);
        const float sx = s.x * inv_m, sy = s.y * inv_m;
        v = make_float2(sx * cc.x + sy * cc.y, sy * cc.x - sx * cc.y);   // s * conj(cc)
    }
    a2[((size_t)pi * n_d + dl) * M + k] = v;
}
// acc[(pi * D + d_lo + dl)][n] (+)= |c2[(pi * n_d + dl)][n] / M|^2
__global__ void gen_acc(const float2* __restrict__ c2, float* __restrict__ acc, int N, int M, int n_d, int D, int d_lo,
                        float inv_m, int first)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int dl = blockIdx.y, pi = blockIdx.z;
    const float2 v = c2[((size_t)pi * n_d + dl) * M + n];
    const float re = v.x * inv_m, im = v.y * inv_m;
    const float p = re * re + im * im;
    float* dst = acc + ((size_t)pi * D + d_lo + dl) * N + n;
    *dst = first ? p : *dst + p;
}
__global__ void gen_code_to_complex(const int8_t* __restrict__ codes, float2* __restrict__ out, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) out[i] = make_float2((float)codes[i], 0.f);
}
__global__ void gen_real_to_complex(const float* __restrict__ in, float2* __restrict__ out, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) out[i] = make_float2(in[i], 0.f);
}

template <typename T> struct PlanT {
    typedef typename V2<T>::type T2;
    int N = 0, M = 0;
    bool pow2 = false;
    T2 *chirp = nullptr, *Bf = nullptr, *Bi = nullptr;
};

template <typename T> cudaError_t plan_build(PlanT<T>& p, int N, cudaStream_t st)
{
    typedef typename V2<T>::type T2;
    p.N = N;
    p.pow2 = (N & (N - 1)) == 0;   // the facade transforms a power of two directly; the search always convolves
    int M = 1;
    while (M < 2 * N - 1) M <<= 1;
    p.M = M;
    std::vector<T2> w(N), b(M), bc(M);
    for (int m = 0; m < M; m++) { b[m].x = b[m].y = 0; bc[m].x = bc[m].y = 0; }
    for (int n = 0; n < N; n++) {
        const long long q = ((long long)n * n) % (2LL * N);      // n^2 mod 2N: the chirp has period 2N in n^2
        const double ang = -M_PI * (double)q / (double)N;
        w[n].x = (T)cos(ang); w[n].y = (T)sin(ang);
        const T cr = (T)cos(ang), ci = (T)(-sin(ang));          // conj(w[n])
        b[n].x = cr; b[n].y = ci;
        bc[n].x = cr; bc[n].y = -ci;
        if (n > 0) { b[M - n] = b[n]; bc[M - n] = bc[n]; }
    }
    cudaError_t e;
    T2* tmp = nullptr;
    if ((e = cudaMalloc((void**)&p.chirp, sizeof(T2) * N)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&p.Bf, sizeof(T2) * M)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&p.Bi, sizeof(T2) * M)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&tmp, sizeof(T2) * M)) != cudaSuccess) return e;
    cudaMemcpyAsync(p.chirp, w.data(), sizeof(T2) * N, cudaMemcpyHostToDevice, st);
    for (int which = 0; which < 2; which++) {
        T2* dst = which ? p.Bi : p.Bf;
        cudaMemcpyAsync(dst, which ? bc.data() : b.data(), sizeof(T2) * M, cudaMemcpyHostToDevice, st);
        T2* r = fft_pow2<T>(dst, tmp, M, 1, 0, st, &e);
        if (e != cudaSuccess) break;
        if (r != dst) cudaMemcpyAsync(dst, r, sizeof(T2) * M, cudaMemcpyDeviceToDevice, st);
    }
    cudaStreamSynchronize(st);   // the host vectors go out of scope
    cudaFree(tmp);
    return e;
}
template <typename T> void plan_free(PlanT<T>& p)
{
    if (p.chirp) cudaFree(p.chirp);
    if (p.Bf) cudaFree(p.Bf);
    if (p.Bi) cudaFree(p.Bi);
    p.chirp = p.Bf = p.Bi = nullptr;
}

// batch of natural-order, unnormalised length-N DFTs: in [batch][N] -> out [batch][N]; s0, s1: batch x M scratch each
template <typename T>
cudaError_t dft_any(const PlanT<T>& p, int inverse, const typename V2<T>::type* in, typename V2<T>::type* out, int batch,
                    typename V2<T>::type* s0, typename V2<T>::type* s1, cudaStream_t st)
{
    typedef typename V2<T>::type T2;
    cudaError_t e = cudaSuccess;
    const int N = p.N, M = p.M;
    if (p.pow2) {
        if ((e = cudaMemcpyAsync(s0, in, sizeof(T2) * (size_t)batch * N, cudaMemcpyDefault, st)) != cudaSuccess) return e;
        T2* r = fft_pow2<T>(s0, s1, N, batch, inverse, st, &e);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(out, r, sizeof(T2) * (size_t)batch * N, cudaMemcpyDefault, st);
    }
    dim3 gm((M + 255) / 256, batch), gn((N + 255) / 256, batch);
    bluestein_pre<T><<<gm, 256, 0, st>>>(in, p.chirp, s0, N, M, inverse);
    T2* A = fft_pow2<T>(s0, s1, M, batch, 0, st, &e);
    if (e != cudaSuccess) return e;
    T2* other = A == s0 ? s1 : s0;
    pointwise_mul<T2><<<dim3((M + 255) / 256, batch, 1), 256, 0, st>>>(A, inverse ? p.Bi : p.Bf, M);
    T2* c = fft_pow2<T>(A, other, M, batch, 1, st, &e);
    if (e != cudaSuccess) return e;
    bluestein_post<T><<<gn, 256, 0, st>>>(c, p.chirp, out, N, M, inverse, (T)1 / (T)M);
    return cudaGetLastError();
}

}  // namespace

struct GenericPlan {
    PlanT<float> f;
    PlanT<double> d;
    bool has_d = false;
};

cudaError_t generic_plan_create(int n, GenericPlan** out, cudaStream_t st)
{
    GenericPlan* p = new GenericPlan();
    const cudaError_t e = plan_build<float>(p->f, n, st);
    if (e != cudaSuccess) { plan_free(p->f); delete p; return e; }
    *out = p;
    return cudaSuccess;
}
void generic_plan_destroy(GenericPlan* p)
{
    if (!p) return;
    plan_free(p->f);
    plan_free(p->d);
    delete p;
}
int generic_plan_m(const GenericPlan* p) { return p->f.M; }   // >= N also for a power of two

cudaError_t generic_dft_f32(GenericPlan* p, int inverse, const float2* in, float2* out, int batch, float2* s0, float2* s1, cudaStream_t st)
{
    return dft_any<float>(p->f, inverse, in, out, batch, s0, s1, st);
}
cudaError_t generic_dft_f64(GenericPlan* p, int inverse, const double2* in, double2* out, int batch, double2* s0, double2* s1,
                            cudaStream_t st)
{
    if (!p->has_d) {
        const cudaError_t e = plan_build<double>(p->d, p->f.N, st);
        if (e != cudaSuccess) return e;
        p->has_d = true;
    }
    return dft_any<double>(p->d, inverse, in, out, batch, s0, s1, st);
}

// code spectra C_p = DFT_N(code_p), natural order (AcquisitionWorker::new, do_acquisition.rs:143-150)
cudaError_t generic_code_fft(GenericPlan* p, const int8_t* codes_dev, int n_prn, float2* code_fft, float2* s0, float2* s1, cudaStream_t st)
{
    const size_t total = (size_t)n_prn * p->f.N;
    gen_code_to_complex<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(codes_dev, code_fft, total);
    return dft_any<float>(p->f, 0, code_fft, code_fft, n_prn, s0, s1, st);
}

cudaError_t generic_real_to_complex(const float* in, float2* out, size_t total, cudaStream_t st)
{
    gen_real_to_complex<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, total);
    return cudaGetLastError();
}

namespace {
template <typename T, typename T2> __global__ void gen_r2c(const T* __restrict__ in, T2* __restrict__ out, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) { out[i].x = in[i]; out[i].y = (T)0; }
}
// out[b][k] = |x[b][k]|^2 (norm_sqr, fft.rs:27, :53) or x[b][k] for k < n_out
template <typename T, typename T2> __global__ void gen_take(const T2* __restrict__ x, void* __restrict__ out, int n, int n_out, int power)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_out) return;
    const size_t b = blockIdx.y;
    const T2 v = x[b * n + k];
    if (power) static_cast<T*>(out)[b * n_out + k] = v.x * v.x + v.y * v.y;
    else static_cast<T2*>(out)[b * n_out + k] = v;
}
}  // namespace

cudaError_t generic_r2c_f64(const double* in, double2* out, size_t total, cudaStream_t st)
{
    gen_r2c<double, double2><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, total);
    return cudaGetLastError();
}
cudaError_t generic_take_f32(const float2* x, void* out, int n, int n_out, int batch, int power, cudaStream_t st)
{
    gen_take<float, float2><<<dim3((n_out + 255) / 256, batch), 256, 0, st>>>(x, out, n, n_out, power);
    return cudaGetLastError();
}
cudaError_t generic_take_f64(const double2* x, void* out, int n, int n_out, int batch, int power, cudaStream_t st)
{
    gen_take<double, double2><<<dim3((n_out + 255) / 256, batch), 256, 0, st>>>(x, out, n, n_out, power);
    return cudaGetLastError();
}

// The search over bins [d_lo, d_lo + n_d) for the a.n_active rows: accumulates |.|^2 rows into acc (n_active x D x N).
// s0, s1: scratch of n_active * n_d * M complex each.
cudaError_t generic_search_slab(GenericPlan* p, const AcqArgs& a, int d_lo, int n_d, float* acc, float2* s0, float2* s1, cudaStream_t st)
{
    const int N = p->f.N, M = p->f.M;
    const PlanT<float>& pl = p->f;
    const int n_groups = a.K / a.n_coh;
    const float inv_m = 1.0f / (float)M;
    cudaError_t e = cudaSuccess;
    for (int g = 0; g < n_groups; g++) {
        gen_fwd_pre<<<dim3((M + 255) / 256, n_d), 256, 0, st>>>(a.iq, a.iq_start, a.iq_mask, a.tables, a.rot, pl.chirp, s0, N, M, a.n_coh,
                                                                g, d_lo);
        float2* A = fft_pow2<float>(s0, s1, M, n_d, 0, st, &e);
        if (e != cudaSuccess) return e;
        float2* other = A == s0 ? s1 : s0;
        pointwise_mul<float2><<<dim3((M + 255) / 256, n_d, 1), 256, 0, st>>>(A, pl.Bf, M);
        float2* c = fft_pow2<float>(A, other, M, n_d, 1, st, &e);
        if (e != cudaSuccess) return e;
        // c holds n_d rows; the inverse path needs n_active * n_d rows in the OTHER buffer: move c to the tail of its own
        // buffer first when it sits at the head of the buffer gen_mid is about to fill
        float2* dst = c == s0 ? s1 : s0;
        gen_mid<<<dim3((M + 255) / 256, n_d, a.n_active), 256, 0, st>>>(c, a.code_fft, a.rows, dst, N, M, n_d, inv_m);
        const size_t batch = (size_t)a.n_active * n_d;
        float2* A2 = fft_pow2<float>(dst, c == s0 ? s0 : s1, M, batch, 0, st, &e);
        if (e != cudaSuccess) return e;
        float2* other2 = A2 == s0 ? s1 : s0;
        {
            const unsigned by = (unsigned)(batch < 32768 ? batch : 32768), bz = (unsigned)((batch + by - 1) / by);
            pointwise_mul<float2><<<dim3((M + 255) / 256, by, bz), 256, 0, st>>>(A2, pl.Bi, M);
        }
        float2* c2 = fft_pow2<float>(A2, other2, M, batch, 1, st, &e);
        if (e != cudaSuccess) return e;
        gen_acc<<<dim3((N + 255) / 256, n_d, a.n_active), 256, 0, st>>>(c2, acc, N, M, n_d, a.D, d_lo, inv_m, g == 0);
    }
    return cudaGetLastError();
}

}  // namespace gb
