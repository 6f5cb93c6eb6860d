// acq_cluster.cu -- acquisition for code periods that do not fit one CTA's shared memory (sm_100a).
//
// N = 80000 (Galileo-E1-like 4 ms code at 20 Msps, BASELINE config 4) is 640 KB of complex f32.  The line
// is spread over a thread-block CLUSTER of RO = 4 CTAs (4 SMs, 160 KB each): the first forward stage is a
// radix-4 decimation-in-frequency step whose four output sub-blocks are independent length-20000 problems,
// one per CTA (existing Plan<20000>), and the last inverse stage is the matching radix-4 DIT step that reads
// the four sub-blocks through DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank) -- the 640 KB line never
// leaves the four SMs.  The accumulated |.|^2 row (N floats) lives in L2/HBM and is reduced to a cell by
// reduce_rows_kernel.  Same arithmetic and quirks as acq_fused_kernel; n_coh = 1 only.
#include <cooperative_groups.h>

#include "acq_common.cuh"
#include "acq_generic.cuh"

namespace cg = cooperative_groups;

namespace gb {

constexpr int kRO = 4;
using PI80k = P20000;
constexpr int kN80k = PI80k::N * kRO;

// multiply by W_4^k (forward: (-i)^k, inverse: (+i)^k)
template <bool INV> __device__ __forceinline__ float2 rot4(float2 a, int k)
{
    k &= 3;
    if (k == 0) return a;
    if (k == 2) return make_float2(-a.x, -a.y);
    const bool plus_i = INV ? (k == 1) : (k == 3);
    return plus_i ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// outer forward stage for sub-block r: V_r[i] = (sum_j x[i + j*NI] W_4^(j r)) * W_N^(i r)
// UNR butterflies per thread are in flight together (their kRO loads each are issued before the first use): with one CTA
// per SM nothing else hides the L2 latency of this streaming loop -- it was 35 % of the kernel's time at one butterfly
// per trip (round 2 profile: long-scoreboard stalls 52 % of the warp samples).
template <class PI, int UNR = 4, class Load>
__device__ __forceinline__ void outer_forward(int r, const float2* __restrict__ otw, float2* __restrict__ line, Load load)
{
    constexpr int NI = PI::N;
    // every trip is full: an index past the end (the last butterflies of the last trip) is clamped for the loads and its
    // store skipped, so no array element is conditionally defined (which would send the arrays to local memory)
    for (int i0 = threadIdx.x; i0 < NI; i0 += UNR * PI::T) {
        float2 x[UNR][kRO], t[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int i = min(i0 + u * PI::T, NI - 1);
#pragma unroll
            for (int j = 0; j < kRO; j++) x[u][j] = load(i + j * NI);
            if (r > 0) t[u] = __ldg(&otw[(r - 1) * NI + i]);
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int i = i0 + u * PI::T;
            float2 v = x[u][0];
#pragma unroll
            for (int j = 1; j < kRO; j++) v = cadd(v, rot4<false>(x[u][j], j * r));
            if (r > 0) v = cmul(v, t[u]);
            if (i < NI) line[PI::phys(i)] = v;
        }
    }
}

// inner length-NI chain on this CTA's sub-block: forward FFT, x conj(code), inverse FFT (all in shared memory)
template <class PI>
__device__ __forceinline__ void inner_chain(float2* __restrict__ line, const float2* __restrict__ tw,
                                            const float2* __restrict__ code)
{
    constexpr int LASTS = PI::NSTAGE - 1;
    using GM = StageGeo<PI, LASTS>;
    DifRange<PI, 0, LASTS, false>::run(line, tw);
#pragma unroll 1
    for (int it = 0; it < GM::ITERS; it++) {
        const int b = threadIdx.x + it * PI::T;
        if (GM::NB % PI::T == 0 || b < GM::NB) {
            const int base = b * GM::R;
            float2 v[GM::R];
#pragma unroll
            for (int j = 0; j < GM::R; j++) v[j] = line[PI::phys(base + j)];
            Dft<GM::R, false>::run(v);
#pragma unroll
            for (int q = 0; q < GM::R; q++) v[q] = cmul_conj(v[q], __ldg(&code[q * GM::NB + b]));
            dft_emit<GM::R, true>(v, [&](int j, float2 y) { line[PI::phys(base + j)] = y; });
        }
    }
    __syncthreads();
    DitRange<PI, LASTS - 1, -1, true>::run(line, tw);
}

// The accumulated power of the NI outputs a CTA owns lives in TENSOR MEMORY while the periods run (thread t keeps outputs
// t + m T, m = 0 .. ACC_COLS - 1, in ACC_COLS columns of its TMEM lane; warps that share a lane quarter stack their column
// ranges: 4 x 40 = 160 -> 256 columns) and goes to the global row once, after the last period: the per-period
// read-modify-write of the row through L2 (and its latency in the middle of the last stage) is gone.
template <class PI> __global__ void __cluster_dims__(kRO, 1, 1) __launch_bounds__(PI::T, 1) acq_cluster_kernel(const AcqArgs a)
{
    extern __shared__ float2 line[];
    __shared__ uint32_t tmem_base_smem;
    constexpr int NI = PI::N;
    constexpr int N = NI * kRO;
    constexpr int UNR = 8;                                                   // outputs per thread and trip of the last stage
    constexpr int TRIPS = (NI + UNR * PI::T - 1) / (UNR * PI::T);
    constexpr int ACC_COLS = TRIPS * UNR;
    constexpr uint32_t TMEM_COLS = ((PI::T + 127) / 128) * ACC_COLS <= 32 ? 32 : ((PI::T + 127) / 128) * ACC_COLS <= 64 ? 64 :
                                   ((PI::T + 127) / 128) * ACC_COLS <= 128 ? 128 : ((PI::T + 127) / 128) * ACC_COLS <= 256 ? 256 : 512;
    static_assert(((PI::T + 127) / 128) * ACC_COLS <= 512, "accumulators must fit the SM's tensor memory");
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int cell = (int)(blockIdx.x / kRO);
    const int d = cell % a.D;
    const int row = a.rows[cell / a.D];
    const float2* __restrict__ w = a.tables + (size_t)d * N;
    const float2* __restrict__ code = a.code_fft + (size_t)row * N + (size_t)r * NI;
    const float2* __restrict__ otw = a.otw;
    float* __restrict__ acc = a.acc_rows + (size_t)cell * N + (size_t)r * NI;
    const float2* peer[kRO];
#pragma unroll
    for (int q = 0; q < kRO; q++) peer[q] = cluster.map_shared_rank(line, q);
    const uint32_t tmem_base = tmem_alloc_cta<TMEM_COLS>(&tmem_base_smem);
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t taddr = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * ACC_COLS;

    for (int k = 0; k < a.K; k++) {
        const unsigned long long blk0 = (unsigned long long)k * N;
        outer_forward<PI>(r, otw, line, [&](int n) { return wipe(ld_iq(a, blk0 + n), __ldg(&w[n])); });
        __syncthreads();
        inner_chain<PI>(line, a.tw, code);
        cluster.sync();  // every sub-block now holds u_q[i] in natural order
        // outer inverse stage (DIT radix 4) for the outputs this CTA owns: n = i + r*NI; UNR outputs per trip, all of
        // their distributed-shared-memory and twiddle loads in flight together
        if (k > 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll 1
        for (int m = 0; m < TRIPS; m++) {
            const int i0 = threadIdx.x + m * UNR * PI::T;
            float p[UNR];
            if (k > 0) tmem_ld8_nm(p, taddr + m * UNR);
            // two half-trips of HALF outputs: HALF x (kRO distributed-shared-memory + kRO - 1 twiddle) loads in flight
            constexpr int HALF = UNR / 2;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float2 u[HALF][kRO], t[HALF][kRO];
                // an index past the end (only the last output of the last trip, for threads >= NI % T) is clamped: its
                // value is computed and kept in tensor memory but never written to the row
#pragma unroll
                for (int e = 0; e < HALF; e++) {
                    const int i = min(i0 + (h * HALF + e) * PI::T, NI - 1);
#pragma unroll
                    for (int q = 0; q < kRO; q++) u[e][q] = peer[q][PI::phys(i)];
#pragma unroll
                    for (int q = 1; q < kRO; q++) t[e][q] = __ldg(&otw[(q - 1) * NI + i]);
                }
                if (h == 0 && k > 0) tmem_wait_ld8_nm(p);
#pragma unroll
                for (int e = 0; e < HALF; e++) {
                    const int i = i0 + (h * HALF + e) * PI::T;
                    float2 y = u[e][0];
#pragma unroll
                    for (int q = 1; q < kRO; q++) y = cadd(y, rot4<true>(cmul_conj(u[e][q], t[e][q]), q * r));
                    const float pw = y.x * y.x + y.y * y.y;
                    const float pe = (k == 0) ? pw : p[h * HALF + e] + pw;
                    p[h * HALF + e] = pe;
                    if (k == a.K - 1 && i < NI) acc[i] = pe;
                }
            }
            if (k < a.K - 1) tmem_st<8>(taddr + m * UNR, p);
        }
        cluster.sync();  // peers are done reading this CTA's line before the next block overwrites it
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc_warp<TMEM_COLS>(tmem_base);
}

// code spectra for the cluster plan: same forward path on the +-1 code samples, one CTA per (PRN, sub-block)
template <class PI> __global__ void __launch_bounds__(PI::T, 1) code_fft_cluster_kernel(const int8_t* __restrict__ codes,
                                                                                      float2* __restrict__ code_fft,
                                                                                      const float2* __restrict__ tw,
                                                                                      const float2* __restrict__ otw)
{
    extern __shared__ float2 line[];
    constexpr int NI = PI::N;
    constexpr int N = NI * kRO;
    constexpr int LASTS = PI::NSTAGE - 1;
    using GM = StageGeo<PI, LASTS>;
    const int r = (int)(blockIdx.x % kRO);
    const int prn = (int)(blockIdx.x / kRO);
    const int8_t* c = codes + (size_t)prn * N;
    float2* out = code_fft + (size_t)prn * N + (size_t)r * NI;
    outer_forward<PI>(r, otw, line, [&](int n) { return make_float2((float)c[n], 0.f); });
    __syncthreads();
    DifRange<PI, 0, LASTS, false>::run(line, tw);
#pragma unroll 1
    for (int it = 0; it < GM::ITERS; it++) {
        const int b = threadIdx.x + it * PI::T;
        if (GM::NB % PI::T == 0 || b < GM::NB) {
            float2 v[GM::R];
#pragma unroll
            for (int j = 0; j < GM::R; j++) v[j] = line[PI::phys(b * GM::R + j)];
            dft_emit<GM::R, false>(v, [&](int q, float2 y) { out[q * GM::NB + b] = y; });
        }
    }
}

// accumulated power row (global) -> cell; also used for diagnostics rows
__global__ void __launch_bounds__(512) reduce_rows_kernel(const float* __restrict__ acc_rows, int n, int D, const int* rows,
                                                         int spc, gb_acq_cell* cells)
{
    __shared__ float red_v[33];
    __shared__ unsigned red_i[33];
    __shared__ float red_s[33];
    const int cell = blockIdx.x;
    const float* __restrict__ acc = acc_rows + (size_t)cell * n;
    const int nsum = (n / 8) * 8;
    PeakIdx pk;
    pk.v = 0.f;
    pk.idx = 0u;
    float sum = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = acc[i];
        PeakIdx c;
        c.v = v;
        c.idx = (unsigned)i;
        if (v > 0.f) pk = peak_merge(pk, c);
        if (i < nsum) sum += v;
    }
    pk = warp_peak(pk);
    sum = warp_sum(sum);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) { red_v[warp] = pk.v; red_i[warp] = pk.idx; red_s[warp] = sum; }
    __syncthreads();
    if (warp == 0) {
        PeakIdx q;
        q.v = lane < nw ? red_v[lane] : 0.f;
        q.idx = lane < nw ? red_i[lane] : 0u;
        float s = lane < nw ? red_s[lane] : 0.f;
        q = warp_peak(q);
        s = warp_sum(s);
        if (lane == 0) { red_v[32] = q.v; red_i[32] = q.idx; red_s[32] = s; }
    }
    __syncthreads();
    const float peak = red_v[32];
    const unsigned arg = red_i[32];
    const float total = red_s[32];
    float p2 = 0.f;
    if (spc > 0) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (two_peak_searched(i, (int)arg, spc, n)) p2 = fmaxf(p2, acc[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p2 = fmaxf(p2, __shfl_xor_sync(0xffffffffu, p2, o));
        __syncthreads();
        if (lane == 0) red_v[warp] = p2;
        __syncthreads();
        if (warp == 0) {
            float v = lane < nw ? red_v[lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            p2 = v;
        }
    }
    if (threadIdx.x == 0) {
        gb_acq_cell c;
        c.peak = peak; c.argmax = arg; c.sum8 = total; c.peak2 = p2;
        cells[(size_t)rows[cell / D] * D + (cell % D)] = c;
    }
}

// ------------------------------------------------------------------ host side
cudaError_t acq_launch_reduce_rows(const float* acc_rows, int n, int D, int n_active, const int* rows, int spc, gb_acq_cell* cells,
                                   cudaStream_t st)
{
    reduce_rows_kernel<<<n_active * D, 512, 0, st>>>(acc_rows, n, D, rows, spc, cells);
    return cudaGetLastError();
}

int acq_cluster_supported(int n) { return n == kN80k; }
int acq_cluster_inner(int n) { return n == kN80k ? PI80k::N : 0; }
int acq_cluster_outer(int n) { return n == kN80k ? kRO : 0; }

static size_t cluster_smem() { return sizeof(float2) * (size_t)PI80k::LINE; }

cudaError_t acq_cluster_launch_search(const AcqArgs& a, cudaStream_t st)
{
    const size_t smem = cluster_smem();
    cudaError_t e = cudaFuncSetAttribute(acq_cluster_kernel<PI80k>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    acq_cluster_kernel<PI80k><<<a.n_active * a.D * kRO, PI80k::T, smem, st>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    reduce_rows_kernel<<<a.n_active * a.D, 512, 0, st>>>(a.acc_rows, kN80k, a.D, a.rows, a.spc, a.cells);
    return cudaGetLastError();
}

cudaError_t acq_cluster_launch_code_fft(const int8_t* codes, int n_prn, float2* code_fft, const float2* tw, const float2* otw,
                                        cudaStream_t st)
{
    const size_t smem = cluster_smem();
    cudaError_t e = cudaFuncSetAttribute(code_fft_cluster_kernel<PI80k>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    code_fft_cluster_kernel<PI80k><<<n_prn * kRO, PI80k::T, smem, st>>>(codes, code_fft, tw, otw);
    return cudaGetLastError();
}

}  // namespace gb
