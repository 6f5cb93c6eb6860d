// frontend.cu -- table-driven digital front-end, sm_100a.  See frontend.cuh.
//
// What is sequential in rf/frontend.rs and what is not:
//   * the NCO phase accumulator (frontend.rs:48-52) does not depend on the samples -> its orbit is computed once per
//     (f_if, fs) on the host, in the reference's f32 arithmetic, and the kernel looks LUT indices up by sample number;
//   * the 8 + 8 DC-bias lanes (dc_remove.rs:23-29) are rounded linear recurrences over every 8th sample: they stay
//     sequential (FMUL -> FADD per step, n/8 steps) and set the run time: one half-warp walks them through shared
//     memory while the other warps of the CTA load the next tile and mix + store the previous one.
#include "frontend.cuh"

#include <math.h>

namespace gb {

static inline float fe_next_phase(float acc, float step)
{
    const float s = acc + step;
    if (s >= 0.f && s < 4096.f) return s >= 2048.f ? s - 2048.f : s;   // exact, == fmodf(s, 2048)
    return fmodf(s, 2048.0f);
}

uint16_t fe_lut_index(float phase)
{
    // Rust `as usize` saturates: negative / NaN -> 0, > usize::MAX -> usize::MAX (% 2048 = 2047)
    if (!(phase > 0.f)) return 0;
    if (phase >= 18446744073709551616.f) return 2047;
    return (uint16_t)((unsigned long long)phase & 2047ull);
}

uint64_t fe_orbit_pos(uint64_t count, uint64_t mu, uint64_t period)
{
    return count < mu + period ? count : mu + (count - mu) % period;
}

bool fe_build_phase_orbit(float step, uint64_t cap, uint64_t min_period, std::vector<float>& phase, uint64_t* mu_out,
                          uint64_t* period_out)
{
    // Brent: cycle length lambda of x -> fe_next_phase(x) from x0 = 0
    uint64_t power = 1, lam = 1, steps = 0;
    float t = 0.f, h = fe_next_phase(0.f, step);
    // NaN never compares equal: a step that produces NaN has no usable orbit
    while (t != h) {
        if (++steps > 2 * cap || h != h) return false;
        if (power == lam) {
            t = h;
            power *= 2;
            lam = 0;
        }
        h = fe_next_phase(h, step);
        lam++;
    }
    if (lam > cap) return false;
    uint64_t mu = 0;
    t = 0.f;
    h = 0.f;
    for (uint64_t i = 0; i < lam; i++) h = fe_next_phase(h, step);
    while (t != h) {
        t = fe_next_phase(t, step);
        h = fe_next_phase(h, step);
        if (++mu > cap) return false;
    }
    const uint64_t reps = lam >= min_period ? 1 : (min_period + lam - 1) / lam;
    if (mu + lam * reps > cap + min_period) return false;
    phase.resize(mu + lam * reps);
    float acc = 0.f;
    for (uint64_t i = 0; i < mu + lam; i++) {
        phase[i] = acc;
        acc = fe_next_phase(acc, step);
    }
    for (uint64_t r = 1; r < reps; r++)
        for (uint64_t i = 0; i < lam; i++) phase[mu + r * lam + i] = phase[mu + i];
    *mu_out = mu;
    *period_out = lam * reps;
    return true;
}

#define FE_TILE 2048
#define FE_MOVERS 256           // threads that load / mix / store
#define FE_THREADS (32 + FE_MOVERS)

// Dynamic shared memory: three tiles of FE_TILE complex samples (load t+1 | DC-remove t | mix t-1).
__global__ void __launch_bounds__(FE_THREADS) frontend_table_kernel(const float2* __restrict__ src, float2* __restrict__ ring,
                                                                   unsigned long long head, unsigned long long mask,
                                                                   unsigned long long n, const float* __restrict__ lut,
                                                                   float* __restrict__ bias, const uint16_t* __restrict__ idx_tab,
                                                                   unsigned long long pos0, unsigned long long mu,
                                                                   unsigned long long period, float alpha, float con)
{
    extern __shared__ float2 fe_tiles[];
    const int n_tiles = (int)((n + FE_TILE - 1) / FE_TILE);
    const int warp = threadIdx.x >> 5;
    const int m = threadIdx.x - 32;   // mover index (warps 1..8)
    const unsigned long long end = mu + period;

    auto tile_len = [&](int t) { return (int)((n - (unsigned long long)t * FE_TILE) < FE_TILE ? (n - (unsigned long long)t * FE_TILE) : FE_TILE); };
    constexpr int PER = FE_TILE / FE_MOVERS;   // samples per mover and tile
    auto load = [&](int t) {
        float2* buf = fe_tiles + (t % 3) * FE_TILE;
        const int tn = tile_len(t);
        const float2* s = src + (unsigned long long)t * FE_TILE;
        float2 x[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int i = m + u * FE_MOVERS;
            if (i < tn) x[u] = __ldg(&s[i]);
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int i = m + u * FE_MOVERS;
            if (i < tn) buf[i] = x[u];
        }
    };
    auto mix = [&](int t) {
        // nco_lut.rs:8-15 verbatim: i' = I*re + Q*im, q' = I*im - Q*re with im = -sin; separate roundings.
        // The orbit-index and LUT look-ups of a thread's PER samples are issued together (two dependent global / L1
        // latencies per tile instead of two per sample).
        const float2* buf = fe_tiles + (t % 3) * FE_TILE;
        const int tn = tile_len(t);
        const unsigned long long first = (unsigned long long)t * FE_TILE;
        unsigned long long pos = pos0 + first + m;                  // < end + n
        if (pos >= end) pos = mu + (pos - mu) % period;
        unsigned k[PER];
        float lc[PER], ls[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            k[u] = (m + u * FE_MOVERS < tn) ? __ldg(&idx_tab[pos]) : 0u;
            pos += FE_MOVERS;                                       // period >= FE_MOVERS: one subtraction wraps
            if (pos >= end) pos -= period;
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            lc[u] = __ldg(&lut[k[u]]);
            ls[u] = __ldg(&lut[2048 + k[u]]);
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int i = m + u * FE_MOVERS;
            if (i < tn) {
                const float2 x = buf[i];
                float2 y;
                y.x = __fadd_rn(__fmul_rn(x.x, lc[u]), __fmul_rn(x.y, ls[u]));
                y.y = __fsub_rn(__fmul_rn(x.x, ls[u]), __fmul_rn(x.y, lc[u]));
                ring[(head + first + i) & mask] = y;
            }
        }
    };

    float b = 0.f;
    if (threadIdx.x < 16) b = bias[threadIdx.x];
    if (warp > 0) load(0);
    __syncthreads();
    for (int t = 0; t < n_tiles; t++) {
        if (warp == 0) {
            if (threadIdx.x < 16) {
                // dc_remove.rs:23-29: lane j of component c sees samples 8k + j; bias = bias*con + x*alpha (two
                // products, one sum, each rounded), out = x - bias.  Eight values are loaded ahead of the chain.
                const int lane = threadIdx.x & 7, comp = threadIdx.x >> 3;
                float* tl = reinterpret_cast<float*>(fe_tiles + (t % 3) * FE_TILE) + comp;
                const int tn = tile_len(t);
                int c = lane;
                for (; c + 56 < tn; c += 64) {
                    float x[8], xa[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        x[u] = tl[2 * (c + 8 * u)];
                        xa[u] = __fmul_rn(x[u], alpha);
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        b = __fadd_rn(__fmul_rn(b, con), xa[u]);
                        x[u] = __fsub_rn(x[u], b);
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) tl[2 * (c + 8 * u)] = x[u];
                }
                for (; c < tn; c += 8) {
                    const float x = tl[2 * c];
                    b = __fadd_rn(__fmul_rn(b, con), __fmul_rn(x, alpha));
                    tl[2 * c] = __fsub_rn(x, b);
                }
            }
        } else {
            if (t > 0) mix(t - 1);
            if (t + 1 < n_tiles) load(t + 1);
        }
        __syncthreads();
    }
    if (warp > 0) mix(n_tiles - 1);
    if (threadIdx.x < 16) bias[threadIdx.x] = b;
}

cudaError_t fe_launch_table(const float2* src, float2* ring, unsigned long long head, unsigned long long mask,
                            unsigned long long n, const float* lut, float* bias, const uint16_t* idx_tab,
                            unsigned long long pos0, unsigned long long mu, unsigned long long period, float alpha,
                            float con, cudaStream_t st)
{
    if (period < FE_MOVERS) return cudaErrorInvalidValue;
    const size_t smem = 3 * FE_TILE * sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(frontend_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    frontend_table_kernel<<<1, FE_THREADS, smem, st>>>(src, ring, head, mask, n, lut, bias, idx_tab, pos0, mu, period, alpha, con);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ tolerance mode: the DC recurrences as a segmented scan
// (gb_frontend_set_mode(h, GB_FE_PARALLEL); the default stays the bit-exact single-CTA kernel above.)
// bias_k = con * bias_{k-1} + alpha * x_k is affine in bias_{k-1}: a run of L steps maps b -> con^L * b + P with P the
// response from a zero state.  The call is cut into segments of FE_SEG steps (8 * FE_SEG samples):
//   fe_par_partial : thread (segment, lane) walks its FE_SEG samples from zero           -> P[segment][lane]  (re, im)
//   fe_par_scan    : one CTA composes the segments in order                              -> bias at each segment's entry
//   fe_par_apply   : thread (segment, lane) walks the segment again from its entry bias, subtracts, mixes, stores.
// Inside a segment every step is rounded exactly like dc_remove.rs:23-29; the composition across segments rounds
// differently from the sequential chain, so outputs agree with the reference to ~1e-6 of the bias (the recurrence
// damps injected rounding error with a time constant of 1000 steps), not bit for bit.  The NCO stays exact (orbit table).
#define FE_SEG 32

__global__ void __launch_bounds__(256) fe_par_partial(const float2* __restrict__ src, unsigned long long n,
                                                      float2* __restrict__ part, float alpha, float con)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long first = (t >> 3) * (8ull * FE_SEG);
    if (first >= n) return;
    const unsigned long long left = (n - first) / 8;
    const int steps = left < FE_SEG ? (int)left : FE_SEG;
    const float2* __restrict__ p = src + first + (t & 7);
    float br = 0.f, bi = 0.f;
    int k = 0;
    for (; k + 8 <= steps; k += 8) {
        float2 x[8];
#pragma unroll
        for (int u = 0; u < 8; u++) x[u] = __ldg(p + 8 * (k + u));
#pragma unroll
        for (int u = 0; u < 8; u++) {
            br = __fadd_rn(__fmul_rn(br, con), __fmul_rn(x[u].x, alpha));
            bi = __fadd_rn(__fmul_rn(bi, con), __fmul_rn(x[u].y, alpha));
        }
    }
    for (; k < steps; k++) {
        const float2 x = __ldg(p + 8 * k);
        br = __fadd_rn(__fmul_rn(br, con), __fmul_rn(x.x, alpha));
        bi = __fadd_rn(__fmul_rn(bi, con), __fmul_rn(x.y, alpha));
    }
    part[t] = make_float2(br, bi);
}

// 256 threads = 32 chunks of consecutive segments x 8 lanes.  a_full = con^FE_SEG, a_last = con^(steps of the last segment).
__global__ void __launch_bounds__(256) fe_par_scan(const float2* __restrict__ part, float2* __restrict__ bin, unsigned n_seg,
                                                   float* __restrict__ bias, float a_full, float a_last)
{
    __shared__ float s_a[32][8];
    __shared__ float2 s_p[32][8];
    const unsigned j = threadIdx.x & 7, c = threadIdx.x >> 3;
    const unsigned per = (n_seg + 31) / 32;
    const unsigned s0 = c * per < n_seg ? c * per : n_seg, s1 = s0 + per < n_seg ? s0 + per : n_seg;
    float A = 1.f;
    float2 P = make_float2(0.f, 0.f);
    for (unsigned s = s0; s < s1; s++) {
        const float a = s + 1 == n_seg ? a_last : a_full;
        const float2 q = part[(size_t)s * 8 + j];
        P.x = fmaf(P.x, a, q.x);
        P.y = fmaf(P.y, a, q.y);
        A *= a;
    }
    s_a[c][j] = A;
    s_p[c][j] = P;
    float2 b = make_float2(bias[j], bias[8 + j]);
    __syncthreads();
    for (unsigned cc = 0; cc < c; cc++) {
        b.x = fmaf(b.x, s_a[cc][j], s_p[cc][j].x);
        b.y = fmaf(b.y, s_a[cc][j], s_p[cc][j].y);
    }
    for (unsigned s = s0; s < s1; s++) {
        bin[(size_t)s * 8 + j] = b;
        const float a = s + 1 == n_seg ? a_last : a_full;
        const float2 q = part[(size_t)s * 8 + j];
        b.x = fmaf(b.x, a, q.x);
        b.y = fmaf(b.y, a, q.y);
    }
    if (c == 31) {   // chunks past the end are empty: the last chunk always carries the final state
        bias[j] = b.x;
        bias[8 + j] = b.y;
    }
}

__global__ void __launch_bounds__(256) fe_par_apply(const float2* __restrict__ src, float2* __restrict__ ring,
                                                    unsigned long long head, unsigned long long mask, unsigned long long n,
                                                    const float* __restrict__ lut, const float2* __restrict__ bin,
                                                    const uint16_t* __restrict__ idx_tab, unsigned long long pos0,
                                                    unsigned long long mu, unsigned long long period, float alpha, float con)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long first = (t >> 3) * (8ull * FE_SEG) + (t & 7);
    if (first >= n) return;
    const unsigned long long left = (n - first + 7) / 8;
    const int steps = left < FE_SEG ? (int)left : FE_SEG;
    const unsigned long long end = mu + period;
    const float2* __restrict__ p = src + first;
    float2 b = bin[t];
    unsigned long long pos = pos0 + first;
    if (pos >= end) pos = mu + (pos - mu) % period;
    unsigned long long o = head + first;
    auto body = [&](const float2 x, const unsigned k, const unsigned long long at) {
        // dc_remove.rs:23-29 then nco_lut.rs:8-15, each product and sum rounded separately
        b.x = __fadd_rn(__fmul_rn(b.x, con), __fmul_rn(x.x, alpha));
        b.y = __fadd_rn(__fmul_rn(b.y, con), __fmul_rn(x.y, alpha));
        const float re = __fsub_rn(x.x, b.x), im = __fsub_rn(x.y, b.y);
        const float lc = __ldg(&lut[k]), ls = __ldg(&lut[2048 + k]);
        float2 y;
        y.x = __fadd_rn(__fmul_rn(re, lc), __fmul_rn(im, ls));
        y.y = __fsub_rn(__fmul_rn(re, ls), __fmul_rn(im, lc));
        ring[at & mask] = y;
    };
    int k = 0;
    for (; k + 8 <= steps; k += 8) {
        float2 x[8];
        unsigned ix[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x[u] = __ldg(p + 8 * (k + u));
            ix[u] = __ldg(&idx_tab[pos]);
            pos += 8;                       // period >= 4096: one subtraction wraps
            if (pos >= end) pos -= period;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) body(x[u], ix[u], o + 8 * (k + u));
    }
    for (; k < steps; k++) {
        const unsigned ix = __ldg(&idx_tab[pos]);
        pos += 8;
        if (pos >= end) pos -= period;
        body(__ldg(p + 8 * k), ix, o + 8 * k);
    }
}

size_t fe_parallel_scratch(unsigned long long n)
{
    const unsigned long long n_seg = (n / 8 + FE_SEG - 1) / FE_SEG;
    return (size_t)n_seg * 8 * 2;   // float2 elements: partial responses, then entry biases
}

cudaError_t fe_launch_parallel(const float2* src, float2* ring, unsigned long long head, unsigned long long mask,
                               unsigned long long n, const float* lut, float* bias, const uint16_t* idx_tab,
                               unsigned long long pos0, unsigned long long mu, unsigned long long period, float alpha,
                               float con, float2* scratch, cudaStream_t st)
{
    if (period < 8 || n % 8 != 0 || n == 0) return cudaErrorInvalidValue;
    const unsigned long long steps = n / 8, n_seg = (steps + FE_SEG - 1) / FE_SEG;
    if (n_seg > 0xffffffffull) return cudaErrorInvalidValue;
    const unsigned long long last = steps - (n_seg - 1) * FE_SEG;
    float2* part = scratch;
    float2* bin = scratch + n_seg * 8;
    const unsigned grid = (unsigned)((n_seg * 8 + 255) / 256);
    fe_par_partial<<<grid, 256, 0, st>>>(src, n, part, alpha, con);
    fe_par_scan<<<1, 256, 0, st>>>(part, bin, (unsigned)n_seg, bias, (float)pow((double)con, (double)FE_SEG),
                                   (float)pow((double)con, (double)last));
    fe_par_apply<<<grid, 256, 0, st>>>(src, ring, head, mask, n, lut, bin, idx_tab, pos0, mu, period, alpha, con);
    return cudaGetLastError();
}


}  // namespace gb
