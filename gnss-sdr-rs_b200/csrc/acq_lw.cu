// acq_lw.cu -- leftover-warp form of the shared-chain inverse kernel for N = 4092 (its own translation unit: the
// plan zoo of acq_kernels.cu takes two minutes to compile).
#include "acq_common.cuh"


namespace gb {

// spectra: L2 only (ld.global.cg), L1 is left to the code spectrum; CG = false is the A/B form
#define LDSPEC(p) (CG ? __ldcg(p) : __ldg(p))

// ------------------------------------------------------------------ leftover-warp form of the inverse kernel
// 4092 / 31 = 132 first-stage butterflies per code period = 4 warps + 4 threads: in acq_inverse_kernel the fifth warp
// runs the whole radix-31 instruction stream for 4 of its 32 lanes, a fifth of that stage's FMA-pipe time.  Here the CTA
// is PW::T working threads (whole warps; PW = the same radices on T = 128, which also owns stages B and C) plus ONE
// leftover warp that batches the REM = NB - PW::T ragged butterflies of 32 / REM consecutive groups into one full-warp
// pass: lane l computes butterfly PW::T + l % REM of group g0 + l / REM, keeps the finished accumulators in registers
// (OddPrimeAcc) and hands them to the line when their group comes round -- it runs one batch ahead of the working
// warps, so its arithmetic overlaps their stages B and C.  20 groups: 83 radix-31 warp passes instead of 100.
// Named barriers (the warp-specialised pattern): A_DONE = line of group g complete (working warps wait, the leftover
// warp only arrives), MID = between stages B and C (working warps only), END = line free again (everybody).
// Every output accumulates its terms in the same order as dft_odd_prime_emit, so cells are bit-identical to
// acq_inverse_kernel's (tests/test_gpu_acquisition.py::test_leftover_warp_kernel_is_bit_identical).
enum { BAR_A_DONE = 1, BAR_MID = 2, BAR_END = 3 };

// DB: the line is double-buffered (group g lives in line + (g & 1) * LINE).  Passing A_DONE(g - 1) then proves that every
// working warp has left stage C of group g - 2, the last reader of buffer g & 1, so END disappears from the loop and the
// leftover warp waits on A_DONE instead of arriving at it (no fence needed: bar.sync orders its stores).
template <class PW, bool CG, bool DB>
__device__ __noinline__ void lw_leftover_warp(const float2* __restrict__ spec, const float2* __restrict__ code,
                                              float2* __restrict__ line, int n_groups)
{
    constexpr int LASTS = PW::NSTAGE - 1;
    using GM = StageGeo<PW, LASTS>;
    constexpr int N = PW::N;
    constexpr int TW = PW::T, TALL = PW::T + 32;
    constexpr int REM = GM::NB - TW;
    constexpr int BATCH = 32 / REM;
    const int lane = threadIdx.x - TW;
    const int b = TW + lane % REM, slot = lane / REM;
    OddPrimeAcc<GM::R> h;
    auto compute = [&](int g0) {
        const int g = g0 + slot;
        if (g < n_groups) {
            const float2* __restrict__ sg = spec + (size_t)g * PW::SPEC_LEN;
            dft_odd_prime_stream_acc<GM::R, true, 3>(
                [&](int q) { return cmul_conj(LDSPEC(&sg[q * PW::SPEC_STRIDE + b]), __ldg(&code[q * PW::SPEC_STRIDE + b])); }, h);
        }
    };
    compute(0);
    for (int g = 0; g < n_groups; g++) {
        if (DB) {
            float2* __restrict__ lg = line + (g & 1) * PW::LINE;
            if (slot == g % BATCH)
                dft_odd_prime_stream_emit<GM::R>(h, [&](int j, float2 y) { lg[PW::phys(b * GM::R + j)] = y; });
            if (g % BATCH == BATCH - 1 && g + 1 < n_groups) compute(g + 1);
            __syncwarp();
            named_bar_sync(BAR_A_DONE, TALL);
        } else {
            if (g > 0) named_bar_sync(BAR_END, TALL);   // group g-1 has left the line
            if (slot == g % BATCH)
                dft_odd_prime_stream_emit<GM::R>(h, [&](int j, float2 y) { line[PW::phys(b * GM::R + j)] = y; });
            __threadfence_block();   // bar.arrive alone does not order the shared-memory stores before the arrival
            __syncwarp();
            named_bar_arrive(BAR_A_DONE, TALL);
            if (g % BATCH == BATCH - 1 && g + 1 < n_groups) compute(g + 1);
        }
    }
    named_bar_sync(BAR_END, TALL);
}

template <class PW, bool CG, bool DB = false>
__global__ void __launch_bounds__(PW::T + 32, PW::MINB) acq_inverse_lw_kernel(const AcqArgs a)
{
    extern __shared__ float2 smem_line[];
    constexpr int LASTS = PW::NSTAGE - 1;
    using G0 = StageGeo<PW, 0>;
    using GM = StageGeo<PW, LASTS>;
    constexpr int N = PW::N;
    constexpr int TW = PW::T, TALL = PW::T + 32;
    constexpr int REM = GM::NB - TW;     // ragged butterflies per group
    static_assert(PW::PFA && LASTS == 2, "three-stage prime-factor plans only");
    static_assert(REM > 0 && REM <= 16 && 32 % REM == 0, "the ragged part must tile a warp");
    const int n_groups = a.K / a.n_coh;
    // Doppler-major block order: the n_active CTAs that share one bin's spectra run together (L2 reuse).  PRN-major
    // Doppler tiles (co-resident CTAs sharing a code in L1) were measured and change nothing.
    const int dl = (int)(blockIdx.x / (unsigned)a.n_active);
    const int row = a.rows[blockIdx.x % (unsigned)a.n_active];
    const int2 sm = a.inv_map ? __ldg(&a.inv_map[a.d_lo + dl]) : make_int2(dl, 0);   // {spectrum slot, shifted code set}
    // 32-bit element offsets, pointers re-formed per group from the uniform bases (fewer long-lived registers)
    const unsigned code_off = ((unsigned)sm.y * (unsigned)a.n_prn + (unsigned)row) * (unsigned)PW::SPEC_LEN;
    const unsigned spec_off = (unsigned)sm.x * (unsigned)n_groups * (unsigned)PW::SPEC_LEN;
    float2* __restrict__ line = smem_line;

    if (threadIdx.x >= TW) {
        lw_leftover_warp<PW, CG, DB>(a.spec + spec_off, a.code_fft + code_off, line, n_groups);
        return;
    }
    {
        // ---- working warps
        float acc[G0::ITERS][G0::R];
#pragma unroll
        for (int it = 0; it < G0::ITERS; it++)
#pragma unroll
            for (int j = 0; j < G0::R; j++) acc[it][j] = 0.f;
        const int b = threadIdx.x;
        for (int g = 0; g < n_groups; g++) {
            const float2* __restrict__ sg = a.spec + (spec_off + (unsigned)g * (unsigned)PW::SPEC_LEN);
            const float2* __restrict__ code = a.code_fft + code_off;
            {
                float2 v[GM::R];
#pragma unroll
                for (int q = 0; q < GM::R; q++) v[q] = cmul_conj(LDSPEC(&sg[q * PW::SPEC_STRIDE + b]), __ldg(&code[q * PW::SPEC_STRIDE + b]));
                // END of the previous group sits HERE, between this group's global loads and its first store to the
                // line: a warp that leaves stage C early spends its L2 latency before the barrier instead of after it
                if (!DB && g > 0) named_bar_sync(BAR_END, TALL);
                float2* __restrict__ lg = DB ? line + (g & 1) * PW::LINE : line;
                dft_emit<GM::R, true>(v, [&](int j, float2 y) { lg[PW::phys(b * GM::R + j)] = y; });
            }
            float2* __restrict__ lg = DB ? line + (g & 1) * PW::LINE : line;
            named_bar_sync(BAR_A_DONE, TALL);
            dit_stage_rows<PW, 1, true, TW / 32>(lg);
            named_bar_sync(BAR_MID, TW);
            final_stage_accumulate<PW>(lg, a.tw, acc);
        }
        named_bar_sync(BAR_END, TALL);   // the reduction reuses the line as scratch
        reduce_row_to_cell<PW, BAR_MID>(acc, smem_line, a.spc, &a.cells[(size_t)row * a.D + a.d_lo + dl], a.npos);
    }
}

template <class PW, bool CG, bool DB = false> static cudaError_t launch_lw(const AcqArgs& a, int n_d, cudaStream_t st)
{
    const size_t smem = sizeof(float2) * (size_t)PW::LINE * (DB ? 2 : 1);
    cudaError_t e = cudaFuncSetAttribute(acq_inverse_lw_kernel<PW, CG, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    acq_inverse_lw_kernel<PW, CG, DB><<<n_d * a.n_active, PW::T + 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t acq_launch_inverse_lw4092(const AcqArgs& a, int n_d, cudaStream_t st)
{
    using PW = P4092W;
    static_assert(PW::T + 32 == P4092::T && PW::LINE == P4092::LINE, "same line as the default plan, one extra warp");
    // Three CTAs per SM with 128 registers: the 36 power accumulators stay in registers next to the 31 stage-A inputs
    // (four CTAs per SM at 96 registers spill them: 72 local-memory accesses per thread and group through the L1 data
    // pipe the kernel is bound by).  Config 2: 1.381 -> 1.304 ms.  gb_tuning_set("acq_lw_minb", 4 | 2) for A/B.
    const int minb = tuning("acq_lw_minb", 3);
    if (tuning("acq_lw_db", 0)) return launch_lw<P4092W3, true, true>(a, n_d, st);   // A/B: double-buffered line
    if (minb == 4) return launch_lw<PW, true>(a, n_d, st);
    if (minb == 2) return launch_lw<P4092W2, true>(a, n_d, st);
    if (tuning("acq_spec_ldg", 0)) return launch_lw<P4092W3, false>(a, n_d, st);   // A/B switch: spectra through L1
    return launch_lw<P4092W3, true>(a, n_d, st);
}

}  // namespace gb
