// acq_lw.cu -- leftover-warp form of the shared-chain inverse kernel for N = 4092 (its own translation unit: the
// plan zoo of acq_kernels.cu takes two minutes to compile).
#include "acq_common.cuh"

#include <cmath>
#include <cstring>
#include <mutex>


namespace gb {

// spectra: L2 only (ld.global.cg), L1 is left to the code spectrum; CG = false is the A/B form
#define LDSPEC(p) (CG ? __ldcg(p) : __ldg(p))

// ------------------------------------------------------------------ leftover-warp form of the inverse kernel
// 4092 / 31 = 132 first-stage butterflies per code period = 4 warps + 4 threads: in acq_inverse_kernel the fifth warp
// runs the whole radix-31 instruction stream for 4 of its 32 lanes, a fifth of that stage's FMA-pipe time.  Here the CTA
// is PW::T working threads (whole warps; PW = the same radices on T = 128, which also owns stages B and C) plus ONE
// leftover warp that batches the REM = NB - PW::T ragged butterflies of 32 / REM consecutive groups into one full-warp
// pass: lane l computes butterfly PW::T + l % REM of group g0 + l / REM, parks the finished outputs in a shared-memory
// stash and the warp hands them to the line when their group comes round -- it runs one batch ahead of the working
// warps, so its arithmetic overlaps their stages B and C.  20 groups: 83 radix-31 warp passes instead of 100.
// Named barriers (the warp-specialised pattern): A_DONE = line of group g complete (working warps wait, the leftover
// warp only arrives), MID = between stages B and C (working warps only), END = line free again (everybody).
// The leftover warp runs the same butterfly code as the working warps (dft_emit), so cells are bit-identical to
// acq_inverse_kernel's (tests/test_gpu_acquisition.py::test_leftover_warp_kernel_is_bit_identical).
enum { BAR_A_DONE = 1, BAR_MID = 2, BAR_END = 3 };

// DB: the line is double-buffered (group g lives in line + (g & 1) * LINE).  Passing A_DONE(g - 1) then proves that every
// working warp has left stage C of group g - 2, the last reader of buffer g & 1, so END disappears from the loop and the
// leftover warp waits on A_DONE instead of arriving at it (no fence needed: bar.sync orders its stores).
// The leftover warp computes its batch with the SAME butterfly as the working warps (dft_emit: resident inputs, the
// plan's radix-31 form), so every butterfly of the line rounds alike whatever warp made it; the finished outputs wait in a
// shared-memory stash (32 lanes x 31 complex behind the line) and the REM butterflies of a group -- one contiguous run of
// REM x 31 complex in the stash and in the line -- are copied by the whole warp when their group comes round.
constexpr int LW_STASH = 32 * 31;   // complex elements
template <class PW, bool CG, bool DB>
__device__ __noinline__ void lw_leftover_warp(const float2* __restrict__ spec, const float2* __restrict__ code,
                                              float2* __restrict__ line, int n_groups)
{
    constexpr int LASTS = PW::NSTAGE - 1;
    using GM = StageGeo<PW, LASTS>;
    constexpr int TW = PW::T, TALL = PW::T + 32;
    constexpr int REM = GM::NB - TW;
    constexpr int BATCH = 32 / REM;
    static_assert(GM::R == 31 && PW::PAD == 0, "31-point first stage, unpadded line");
    const int lane = threadIdx.x - TW;
    const int b = TW + lane % REM, slot = lane / REM;
    float2* __restrict__ stash = line + PW::LINE * (DB ? 2 : 1);
    auto compute = [&](int g0) {
        const int g = g0 + slot;
        if (g < n_groups) {
            const float2* __restrict__ sg = spec + (size_t)g * PW::SPEC_LEN;
            float2 v[GM::R];
#pragma unroll
            for (int q = 0; q < GM::R; q++) v[q] = cmul_conj(LDSPEC(&sg[q * PW::SPEC_STRIDE + b]), __ldg(&code[q * PW::SPEC_STRIDE + b]));
            dft_emit<GM::R, true, PW::NESTED31>(v, [&](int j, float2 y) { stash[lane * GM::R + j] = y; });
        }
        __syncwarp();
    };
    // group g: the REM x 31 outputs of slot g % BATCH go to line positions TW * 31 ...
    auto hand_over = [&](int g, float2* __restrict__ lg) {
        const float2* __restrict__ src = stash + (g % BATCH) * REM * GM::R;
        float2* __restrict__ dst = lg + TW * GM::R;
        for (int e = lane; e < REM * GM::R; e += 32) dst[e] = src[e];
    };
    compute(0);
    for (int g = 0; g < n_groups; g++) {
        if (DB) {
            hand_over(g, line + (g & 1) * PW::LINE);
            __syncwarp();
            if (g % BATCH == BATCH - 1 && g + 1 < n_groups) compute(g + 1);
            named_bar_sync(BAR_A_DONE, TALL);
        } else {
            if (g > 0) named_bar_sync(BAR_END, TALL);   // group g-1 has left the line
            hand_over(g, line);
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
                dft_emit<GM::R, true, PW::NESTED31>(v, [&](int j, float2 y) { lg[PW::phys(b * GM::R + j)] = y; });
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

// ------------------------------------------------------------------ power accumulators in tensor memory
// gb_tuning_set("acq_lw_tmem", 1).  The 36 power accumulators of a working thread are touched once per group (stage C) and
// otherwise only occupy registers: with them the kernel needs 128 registers (3 CTAs per SM), without them it fits 96
// (4 CTAs per SM, 16 working warps instead of 12).  Blackwell's tensor memory is a 256 KB per-SM accumulator store with
// its own datapath (tcgen05.ld / tcgen05.st, SASS LDTM / STTM; no L1 data-pipe traffic, unlike a spill): each CTA
// allocates 64 columns, thread (warp w, lane l) owns TMEM lane 32 w + l, columns 0 .. 35.  Stage C loads twelve
// accumulators per radix-12 butterfly while the butterfly's inputs come from the line, adds |.|^2 in the same order as
// final_stage_accumulate (results are bit-identical) and stores them back.
// CT = true (A/B, gb_tuning_set("acq_lw_tmem", 2)): the thread's 31 code-spectrum values -- the same for every group of the
// search -- are kept in tensor memory as well (columns 64 .. 125 of a 128-column allocation: four CTAs fill the SM's 512
// columns) and read back in chunks of four values per group instead of 31 loads through L1 / L2.
template <class PW, bool CG, bool CT = false>
__global__ void __launch_bounds__(PW::T + 32, PW::MINB) acq_inverse_lwt_kernel(const AcqArgs a)
{
    extern __shared__ float2 smem_line[];
    __shared__ uint32_t tmem_base_smem;
    constexpr int LASTS = PW::NSTAGE - 1;
    using G0 = StageGeo<PW, 0>;
    using GM = StageGeo<PW, LASTS>;
    constexpr int TW = PW::T, TALL = PW::T + 32;
    constexpr uint32_t TMEM_COLS = CT ? 128 : 64, CODE_COL = 64;
    static_assert(PW::PFA && LASTS == 2 && TW == 128 && G0::ITERS * G0::R <= 64, "N = 4092 plan, 36 accumulators per thread");
    static_assert(!CT || (GM::R == 31 && PW::MINB * 128 <= 512), "31 code values per thread in 62 columns");
    const int n_groups = a.K / a.n_coh;
    const int dl = (int)(blockIdx.x / (unsigned)a.n_active);
    const int row = a.rows[blockIdx.x % (unsigned)a.n_active];
    const int2 sm = a.inv_map ? __ldg(&a.inv_map[a.d_lo + dl]) : make_int2(dl, 0);
    const unsigned code_off = ((unsigned)sm.y * (unsigned)a.n_prn + (unsigned)row) * (unsigned)PW::SPEC_LEN;
    const unsigned spec_off = (unsigned)sm.x * (unsigned)n_groups * (unsigned)PW::SPEC_LEN;
    float2* __restrict__ line = smem_line;
    const int warp = threadIdx.x >> 5;

    const uint32_t tmem_base = tmem_alloc_cta<TMEM_COLS>(&tmem_base_smem);

    if (threadIdx.x >= TW) {
        lw_leftover_warp<PW, CG, false>(a.spec + spec_off, a.code_fft + code_off, line, n_groups);
        return;
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's lane quarter, column 0
    tmem_zero_accumulators<PW>(taddr);   // tensor memory comes uninitialised
    const int b = threadIdx.x;
    if constexpr (CT) {
        const float2* __restrict__ code = a.code_fft + code_off;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            float cv[8];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int q = 4 * c + u;
                const float2 w = q < GM::R ? __ldg(&code[q * PW::SPEC_STRIDE + b]) : make_float2(0.f, 0.f);
                cv[2 * u] = w.x;
                cv[2 * u + 1] = w.y;
            }
            tmem_st<8>(taddr + CODE_COL + 8 * c, cv);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    for (int g = 0; g < n_groups; g++) {
        const float2* __restrict__ sg = a.spec + (spec_off + (unsigned)g * (unsigned)PW::SPEC_LEN);
        const float2* __restrict__ code = a.code_fft + code_off;
        {
            float2 v[GM::R];
            if constexpr (CT) {
                float cv[2][8];
                tmem_ld8_nm(cv[0], taddr + CODE_COL);
#pragma unroll
                for (int q = 0; q < GM::R; q++) v[q] = LDSPEC(&sg[q * PW::SPEC_STRIDE + b]);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    tmem_wait_ld8_nm(cv[c & 1]);
                    if (c < 7) tmem_ld8_nm(cv[(c + 1) & 1], taddr + CODE_COL + 8 * (c + 1));
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int q = 4 * c + u;
                        if (q < GM::R) v[q] = cmul_conj(v[q], make_float2(cv[c & 1][2 * u], cv[c & 1][2 * u + 1]));
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < GM::R; q++) v[q] = cmul_conj(LDSPEC(&sg[q * PW::SPEC_STRIDE + b]), __ldg(&code[q * PW::SPEC_STRIDE + b]));
            }
            if (g > 0) named_bar_sync(BAR_END, TALL);
            dft_emit<GM::R, true, PW::NESTED31>(v, [&](int j, float2 y) { line[PW::phys(b * GM::R + j)] = y; });
        }
        named_bar_sync(BAR_A_DONE, TALL);
        dit_stage_rows<PW, 1, true, TW / 32>(line);
        named_bar_sync(BAR_MID, TW);
        final_stage_accumulate_tmem<PW>(line, a.tw, taddr);
    }
    named_bar_sync(BAR_END, TALL);   // the reduction reuses the line as scratch
    float acc[G0::ITERS][G0::R];
    tmem_load_accumulators<PW>(taddr, acc);
    reduce_row_to_cell<PW, BAR_MID>(acc, smem_line, a.spc, &a.cells[(size_t)row * a.D + a.d_lo + dl], a.npos);
    if (warp == 0) tmem_dealloc_warp<TMEM_COLS>(tmem_base);   // every working warp has passed reduce_row_to_cell's barriers after its last TMEM read
}

template <class PW, bool CG, bool CT = false> static cudaError_t launch_lwt(const AcqArgs& a, int n_d, cudaStream_t st)
{
    const size_t smem = sizeof(float2) * ((size_t)PW::LINE + LW_STASH);   // line + the leftover warp's stash
    cudaError_t e = cudaFuncSetAttribute(acq_inverse_lwt_kernel<PW, CG, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    acq_inverse_lwt_kernel<PW, CG, CT><<<n_d * a.n_active, PW::T + 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <class PW, bool CG, bool DB = false> static cudaError_t launch_lw(const AcqArgs& a, int n_d, cudaStream_t st)
{
    const size_t smem = sizeof(float2) * ((size_t)PW::LINE * (DB ? 2 : 1) + LW_STASH);   // line(s) + the leftover warp's stash
    cudaError_t e = cudaFuncSetAttribute(acq_inverse_lw_kernel<PW, CG, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    acq_inverse_lw_kernel<PW, CG, DB><<<n_d * a.n_active, PW::T + 32, smem, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ tensor-pipe form of the first inverse stage (A/B)
// gb_tuning_set("acq_tc", 1).  NOT the default and outside the north star ("tensor cores are not used"): an explicit A/B
// that moves the radix-31 stage -- 55 % of the inverse path's FP32 pipe slots -- to the warp-level tensor path
// (mma.sync.m16n8k8 tf32, SASS HMMA.1688.F32.TF32; measured 476 MAC/clk/SM on B200, tools/ubench/mma_tf32.cu).
// The conjugate-symmetric half form is two real 16 x 16 matrices,
//   C'[q][k] = (k == 0 ? 1 : cos(2 pi k q / 31))   applied to (x_0, a_1 .. a_15),   a_k = x_k + x_(31-k)
//   S [q][k] = (k == 0 ? 0 : sin(2 pi k q / 31))   applied to (  - , b_1 .. b_15),   b_k = x_k - x_(31-k)
// so y_q, y_(31-q) = C'a +/- i S b is four real GEMMs (re / im of a, re / im of b) of shape 16 x 16 x 8 per tile of eight
// butterflies, each as 3 x TF32 (hi * lo + lo * hi + hi * hi, hi = value rounded to 10 mantissa bits, lo = the exact
// rest; the dropped lo * lo term is 2^-22): 24 HMMA per tile instead of 8 x 1050 FP32 pipe slots.  Error per output
// <= 1e-6 of the butterfly's largest input (tests/test_gpu_acquisition.py::test_tensor_stage_*).
// Data flow: a tile's B fragments come straight from global memory -- spectra and code spectra are re-laid in
// "fragment order" (tc_relayout_kernel: tile, k-slot, lane, {x_k, x_(31-k)}; one 16-byte load per pair, 512 contiguous
// bytes per warp instruction) --, D fragments are combined in registers and stored to the line as before.  The rows of
// the two matrices are permuted (output q of row r: tc_row_q) so that the eight 64-bit stores of a tile are bank-conflict
// free per half-warp.  The ragged butterflies 128 .. 131 stay with the FP32 leftover warp and the [q][b] layout.
constexpr int TC_TILES = 16;                 // tiles of eight butterflies handled by the working warps
constexpr int TC_LEN = TC_TILES * 256;       // complex elements of one spectrum in fragment order

__host__ __device__ constexpr int tc_row_q(int r) { return 2 * (r >> 2) + (r & 1) + 8 * ((r >> 1) & 1); }

// A fragments of {C' hi, C' lo, S hi, S lo} x two k-steps, one uint4 per lane
__device__ uint4 g_tc_frag[4][2][32];

__global__ void tc_relayout_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int src_len, int stride, int n_sets)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n_sets * TC_LEN) return;
    const size_t set = e / TC_LEN;
    const int r = (int)(e % TC_LEN);
    const int tile = r >> 8, s = (r >> 6) & 3, lane = (r >> 1) & 31, pm = r & 1;
    const int g = lane >> 2, t = lane & 3, kappa = t + 4 * s;
    const int q = pm ? 31 - kappa : kappa;
    dst[e] = (pm && kappa == 0) ? make_float2(0.f, 0.f) : __ldg(&src[set * src_len + q * stride + tile * 8 + g]);
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint4 a, unsigned b0, unsigned b1)
{
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// (re, im) = hi + lo per component: hi = the value rounded to TF32 (half-up on the magnitude, two integer operations),
// lo = the exact rest (one packed subtraction for both components; the tensor path ignores its low 13 bits)
__device__ __forceinline__ void tf32_split2(pk64 v, unsigned& hre, unsigned& him, unsigned& lre, unsigned& lim)
{
    const float2 f = upk(v);
    hre = (__float_as_uint(f.x) + 0x1000u) & 0xffffe000u;
    him = (__float_as_uint(f.y) + 0x1000u) & 0xffffe000u;
    const float2 l = upk(sub2(v, pk(__uint_as_float(hre), __uint_as_float(him))));
    lre = __float_as_uint(l.x);
    lim = __float_as_uint(l.y);
}
// dre, dim (16 x 8 each) = M (16 x 16; hi / lo fragments of the two k-steps) * re / im of x (16 x 8 complex; this lane holds
// rows t, t + 4, t + 8, t + 12 of one column), as 3 x TF32 with the small terms first; the two chains run in lock-step
__device__ __forceinline__ void tc_gemm2(float (&dre)[4], float (&dim)[4], const uint4 (&mh)[2], const uint4 (&ml)[2],
                                         const pk64 (&x)[4])
{
    unsigned hre[4], him[4], lre[4], lim[4];
#pragma unroll
    for (int s = 0; s < 4; s++) tf32_split2(x[s], hre[s], him[s], lre[s], lim[s]);
    dre[0] = dre[1] = dre[2] = dre[3] = 0.f;
    dim[0] = dim[1] = dim[2] = dim[3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
        mma_tf32(dre, mh[ks], lre[2 * ks], lre[2 * ks + 1]);
        mma_tf32(dim, mh[ks], lim[2 * ks], lim[2 * ks + 1]);
        mma_tf32(dre, ml[ks], hre[2 * ks], hre[2 * ks + 1]);
        mma_tf32(dim, ml[ks], him[2 * ks], him[2 * ks + 1]);
    }
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
        mma_tf32(dre, mh[ks], hre[2 * ks], hre[2 * ks + 1]);
        mma_tf32(dim, mh[ks], him[2 * ks], him[2 * ks + 1]);
    }
}

template <class PW>
__global__ void __launch_bounds__(PW::T + 32, PW::MINB) acq_inverse_tc_kernel(const AcqArgs a, const float2* __restrict__ spec_tc,
                                                                              const float2* __restrict__ code_tc)
{
    extern __shared__ float2 smem_line[];
    constexpr int LASTS = PW::NSTAGE - 1;
    using G0 = StageGeo<PW, 0>;
    using GM = StageGeo<PW, LASTS>;
    constexpr int TW = PW::T, TALL = PW::T + 32;
    static_assert(PW::PFA && LASTS == 2 && GM::R == 31 && TW == 128 && PW::N == 4092, "the N = 4092 plan only");
    const int n_groups = a.K / a.n_coh;
    const int dl = (int)(blockIdx.x / (unsigned)a.n_active);
    const int row = a.rows[blockIdx.x % (unsigned)a.n_active];
    const int2 sm = a.inv_map ? __ldg(&a.inv_map[a.d_lo + dl]) : make_int2(dl, 0);
    const unsigned code_set = (unsigned)sm.y * (unsigned)a.n_prn + (unsigned)row;
    float2* __restrict__ line = smem_line;

    if (threadIdx.x >= TW) {
        lw_leftover_warp<PW, true, false>(a.spec + (unsigned)sm.x * (unsigned)n_groups * (unsigned)PW::SPEC_LEN,
                                          a.code_fft + code_set * (unsigned)PW::SPEC_LEN, line, n_groups);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gq = lane >> 2, t4 = lane & 3;
    uint4 frag[4][2];
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) frag[m][ks] = g_tc_frag[m][ks][lane];
    float acc[G0::ITERS][G0::R];
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++)
#pragma unroll
        for (int j = 0; j < G0::R; j++) acc[it][j] = 0.f;
    // float4 index of (tile T, k-slot s, lane): T * 128 + s * 32 + lane; this warp owns tiles 4 * warp .. 4 * warp + 3
    const float4* __restrict__ code4 = reinterpret_cast<const float4*>(code_tc + (size_t)code_set * TC_LEN) + warp * 512 + lane;
    const int q_lo = tc_row_q(gq), q_hi = q_lo + 4;   // outputs of D rows gq and gq + 8
    const float4* __restrict__ sp4 = reinterpret_cast<const float4*>(spec_tc + (size_t)sm.x * n_groups * TC_LEN) + warp * 512 + lane;
    // The spectra of a tile are loaded ONE TILE AHEAD (tile 0 of the next group during stages B and C of this one), and
    // the loads are issued after the tile's own inputs have been consumed: a scoreboard wait covers every load in flight,
    // so loads issued before that first use would be waited for as well.
    float4 xs[4];
#pragma unroll
    for (int s = 0; s < 4; s++) xs[s] = __ldcg(&sp4[s * 32]);
    for (int g = 0; g < n_groups; g++) {
#pragma unroll
        for (int tau = 0; tau < 4; tau++) {
            pk64 av[4], bv[4];
            {
                float4 c[4];
#pragma unroll
                for (int s = 0; s < 4; s++) c[s] = __ldg(&code4[tau * 128 + s * 32]);
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const pk64 vp = pk(cmul_conj(make_float2(xs[s].x, xs[s].y), make_float2(c[s].x, c[s].y)));
                    const pk64 vm = pk(cmul_conj(make_float2(xs[s].z, xs[s].w), make_float2(c[s].z, c[s].w)));
                    av[s] = add2(vp, vm);
                    bv[s] = sub2(vp, vm);
                }
            }
            if (tau < 3 || g + 1 < n_groups) {
                const float4* __restrict__ nx = sp4 + (tau < 3 ? tau + 1 : TC_TILES) * 128;   // next tile | the same warp's tile 0 of the next group
#pragma unroll
                for (int s = 0; s < 4; s++) xs[s] = __ldcg(&nx[s * 32]);
            }
            // END of the previous group between this group's first global loads and its first store to the line
            if (tau == 0 && g > 0) named_bar_sync(BAR_END, TALL);
            float pre[4], pim[4], qre[4], qim[4];
            tc_gemm2(pre, pim, frag[0], frag[1], av);
            tc_gemm2(qre, qim, frag[2], frag[3], bv);
            // D fragment: [0] (row gq, col 2 t4), [1] (gq, 2 t4 + 1), [2] (gq + 8, 2 t4), [3] (gq + 8, 2 t4 + 1); col = butterfly
            float2* __restrict__ lb = line + (warp * 32 + tau * 8 + 2 * t4) * 31;
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = 2 * h + c, q = h ? q_hi : q_lo;
                    const pk64 p = pk(pre[i], pim[i]), r = pk(-qim[i], qre[i]);   // y_q = P + i Q, y_(31-q) = P - i Q
                    lb[c * 31 + q] = upk(add2(p, r));
                    if (h || q_lo != 0) lb[c * 31 + 31 - q] = upk(sub2(p, r));
                }
        }
        sp4 += TC_LEN / 2;   // float4 units
        named_bar_sync(BAR_A_DONE, TALL);
        dit_stage_rows<PW, 1, true, TW / 32>(line);
        named_bar_sync(BAR_MID, TW);
        final_stage_accumulate<PW>(line, a.tw, acc);
    }
    named_bar_sync(BAR_END, TALL);
    reduce_row_to_cell<PW, BAR_MID>(acc, smem_line, a.spc, &a.cells[(size_t)row * a.D + a.d_lo + dl], a.npos);
}

static uint32_t tf32_round_host(double v, double* rest)
{
    float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    float hi;
    memcpy(&hi, &u, 4);
    if (rest) *rest = v - (double)hi;
    return u;
}
static cudaError_t tc_upload_fragments(cudaStream_t st)
{
    static std::mutex mu;
    static bool done[64] = {false};
    static uint4 host[4][2][32];
    std::lock_guard<std::mutex> lk(mu);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (done[dev]) return cudaSuccess;
    const double two_pi = 6.283185307179586476925286766559;
    for (int m = 0; m < 2; m++)   // 0: C', 1: S
        for (int ks = 0; ks < 2; ks++)
            for (int lane = 0; lane < 32; lane++) {
                const int g = lane >> 2, t = lane & 3;
                uint32_t hi[4], lo[4];
                for (int i = 0; i < 4; i++) {
                    const int r = g + 8 * (i & 1), k = 8 * ks + t + 4 * (i >> 1);   // a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4)
                    const int q = tc_row_q(r);
                    const double ang = two_pi * (double)((k * q) % 31) / 31.0;
                    const double v = m == 0 ? (k == 0 ? 1.0 : cos(ang)) : (k == 0 ? 0.0 : sin(ang));
                    double rest;
                    hi[i] = tf32_round_host(v, &rest);
                    lo[i] = tf32_round_host(rest, nullptr);
                }
                host[2 * m][ks][lane] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                host[2 * m + 1][ks][lane] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
    e = cudaMemcpyToSymbolAsync(g_tc_frag, host, sizeof(host), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);   // `host` is static, but the flag must not be set before the copy has happened
    if (e == cudaSuccess) done[dev] = true;
    return e;
}

int acq_tc_spec_len() { return TC_LEN; }

static cudaError_t acq_launch_inverse_tc4092(const AcqArgs& a, int n_d, int n_spec_sets, int n_code_sets, int code_fresh,
                                      float2* spec_tc, float2* code_tc, cudaStream_t st)
{
    using PW = P4092W3;
    cudaError_t e = tc_upload_fragments(st);
    if (e != cudaSuccess) return e;
    const size_t n_spec = (size_t)n_spec_sets * TC_LEN, n_code = (size_t)n_code_sets * TC_LEN;
    tc_relayout_kernel<<<(unsigned)((n_spec + 255) / 256), 256, 0, st>>>(a.spec, spec_tc, PW::SPEC_LEN, PW::SPEC_STRIDE, n_spec_sets);
    if (!code_fresh)
        tc_relayout_kernel<<<(unsigned)((n_code + 255) / 256), 256, 0, st>>>(a.code_fft, code_tc, PW::SPEC_LEN, PW::SPEC_STRIDE, n_code_sets);
    const size_t smem = sizeof(float2) * ((size_t)PW::LINE + LW_STASH);   // line + the leftover warp's stash
    e = cudaFuncSetAttribute(acq_inverse_tc_kernel<PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    acq_inverse_tc_kernel<PW><<<n_d * a.n_active, PW::T + 32, smem, st>>>(a, spec_tc, code_tc);
    return cudaGetLastError();
}

cudaError_t acq_launch_inverse_lw4092(const AcqArgs& a, int n_d, cudaStream_t st)
{
    using PW = P4092W;
    static_assert(PW::T + 32 == P4092::T && PW::LINE == P4092::LINE, "same line as the default plan, one extra warp");
    // Three CTAs per SM with 128 registers: the 36 power accumulators stay in registers next to the 31 stage-A inputs
    // (four CTAs per SM at 96 registers spill them: 72 local-memory accesses per thread and group through the L1 data
    // pipe the kernel is bound by).  Config 2: 1.381 -> 1.304 ms.  gb_tuning_set("acq_lw_minb", 4 | 2) for A/B.
    if (a.spec_tc && a.code_tc) {   // A/B: radix-31 stage on the tensor pipe
        const int n_groups = a.K / a.n_coh;
        return acq_launch_inverse_tc4092(a, n_d, (a.tc_n_fwd ? a.tc_n_fwd : n_d) * n_groups, a.tc_n_code_sets, a.tc_code_fresh,
                                         a.spec_tc, a.code_tc, st);
    }
    // The default (2): power accumulators AND the thread's 31 code-spectrum values in tensor memory, 96 registers, four CTAs
    // per SM (config 2: 1.304 -> 1.228 ms with the accumulators there, -> 1.193 ms with the code spectrum as well).
    // gb_tuning_set("acq_lw_tmem", 1) = accumulators only, 0 = the register forms below (three CTAs per SM at 128 registers:
    // the 36 accumulators next to the 31 stage-A inputs; four CTAs at 96 registers spill them: 1.382 ms), 5 = five CTAs
    // (spills: 1.68 ms).  Cells are identical in every form.
    const int tm = tuning("acq_lw_tmem", 2);
    const int minb = tuning("acq_lw_minb", 3);
    if (tm == 5) return launch_lwt<P4092W5, true>(a, n_d, st);
    if (tm == 2) return launch_lwt<P4092W, true, true>(a, n_d, st);
    if (tm == 3) return launch_lwt<P4092W3, true, true>(a, n_d, st);   // A/B: the same at three CTAs per SM and 128 registers
    if (tm) return launch_lwt<P4092W, true>(a, n_d, st);
    if (tuning("acq_lw_db", 0)) return launch_lw<P4092W3, true, true>(a, n_d, st);   // A/B: double-buffered line
    if (minb == 4) return launch_lw<PW, true>(a, n_d, st);
    if (minb == 2) return launch_lw<P4092W2, true>(a, n_d, st);
    if (tuning("acq_spec_ldg", 0)) return launch_lw<P4092W3, false>(a, n_d, st);   // A/B switch: spectra through L1
    return launch_lw<P4092W3, true>(a, n_d, st);
}

}  // namespace gb
