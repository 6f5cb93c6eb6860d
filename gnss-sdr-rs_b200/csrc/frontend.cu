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

}  // namespace gb
