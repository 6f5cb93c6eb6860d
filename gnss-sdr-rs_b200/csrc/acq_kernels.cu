// acq_kernels.cu -- fused parallel code-phase search, sm_100a.
//
// One CTA owns one (PRN, Doppler bin) cell row and runs, per 1 ms block (or per coherent group),
//   carrier wipe-off (doppler_shift.rs:25-58)  -> forward FFT (do_acquisition.rs:182)
//   -> x conj(code spectrum) (:184-186)        -> inverse FFT (:188)
//   -> |.|^2 accumulate (:190-192)
// entirely in shared memory / registers, then reduces the accumulated row to
// {peak, first argmax, 8-lane sum, second peak} (:195-202, :229-234) -- 16 bytes to HBM per cell row.
// The only HBM/L2 reads are the IQ chunk, the wipe-off table and the code spectrum.
#include "acq_common.cuh"

#include <type_traits>

namespace gb {

#define GB_FOR_EACH_PRODUCTION_PLAN(X) X(0, P1024) X(1, P2048) X(2, P4092) X(3, P4096) X(4, P8184) X(5, P16368) X(6, P20000)
#ifdef GB_TUNING
// A/B scaffolding (build.sh -DGB_TUNING; selected with gb_tuning_set("acq_variant", 1..8)): not part of the shipped library
#define GB_FOR_EACH_PLAN(X) GB_FOR_EACH_PRODUCTION_PLAN(X) \
    X(7, P4092v1) X(8, P4092v2) X(9, P4092v3) X(10, P4092v4) X(11, P16368v1) X(12, P16368v2) \
    X(13, P4092v5) X(14, P4092v6) X(15, P4092v7) X(16, P4092v8)
static const int kPlanSizes[] = {1024, 2048, 4092, 4096, 8184, 16368, 20000, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
#else
#define GB_FOR_EACH_PLAN(X) GB_FOR_EACH_PRODUCTION_PLAN(X)
static const int kPlanSizes[] = {1024, 2048, 4092, 4096, 8184, 16368, 20000};
#endif
static const int kNumPlans = sizeof(kPlanSizes) / sizeof(int);
// the default plan of each size (indices 0..6); the rest are tuning variants
template <class P> constexpr bool kProductionPlan =
    std::is_same<P, P1024>::value || std::is_same<P, P2048>::value || std::is_same<P, P4092>::value || std::is_same<P, P4096>::value ||
    std::is_same<P, P8184>::value || std::is_same<P, P16368>::value || std::is_same<P, P20000>::value;
int acq_plan_supports_alias(int plan) { return plan >= 0 && plan <= 6; }

// Stage 0 of the forward DIF (L = N) with the carrier wipe-off (and, for n_coh > 1, the coherent
// pre-sum of n_coh rotated blocks) fused into the global-memory load.
template <class P>
__device__ __forceinline__ void stage0_wipe_forward(const AcqArgs& a, const float2* __restrict__ w,
                                                    const float2* __restrict__ rot, int n_coh, int g,
                                                    float2* __restrict__ line, const float2* __restrict__ tw)
{
    using G0 = StageGeo<P, 0>;
    constexpr int N = P::N;
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) {
        const int i = threadIdx.x + it * P::T;
        if (G0::NB % P::T == 0 || i < G0::NB) {
            float2 v[G0::R];
            if (n_coh == 1) {
                const unsigned long long blk0 = (unsigned long long)g * N;
#pragma unroll
                for (int j = 0; j < G0::R; j++)
                    v[j] = wipe(ld_iq(a, blk0 + i + j * G0::SUB), __ldg(&w[i + j * G0::SUB]));
            } else {
                float2 wv[G0::R];
#pragma unroll
                for (int j = 0; j < G0::R; j++) {
                    wv[j] = __ldg(&w[i + j * G0::SUB]);
                    v[j] = make_float2(0.f, 0.f);
                }
                for (int c = 0; c < n_coh; c++) {
                    const float2 r = __ldg(&rot[c]);
                    const unsigned long long blk0 = (unsigned long long)(g * n_coh + c) * N;
#pragma unroll
                    for (int j = 0; j < G0::R; j++) {
                        const float2 t = wipe(ld_iq(a, blk0 + i + j * G0::SUB), wv[j]);
                        v[j].x = fmaf(t.x, r.x, fmaf(-t.y, r.y, v[j].x));
                        v[j].y = fmaf(t.x, r.y, fmaf(t.y, r.x, v[j].y));
                    }
                }
            }
            Dft<G0::R, false>::run(v);
            line[P::phys(i)] = v[0];
#pragma unroll
            for (int q = 1; q < G0::R; q++)
                line[P::phys(i + q * G0::SUB)] = P::PFA ? v[q] : cmul(v[q], __ldg(&tw[(q - 1) * G0::SUB + i]));
        }
    }
}

template <class P, bool WANT_ROW> __global__ void __launch_bounds__(P::T, P::MINB) acq_fused_kernel(const AcqArgs a)
{
    extern __shared__ float2 line[];
    constexpr int LASTS = P::NSTAGE - 1;
    using G0 = StageGeo<P, 0>;
    using GM = StageGeo<P, LASTS>;
    constexpr int N = P::N;

    const int d = WANT_ROW ? a.d0 : (int)(blockIdx.x % (unsigned)a.D);
    const int row = a.rows[WANT_ROW ? 0 : (int)(blockIdx.x / (unsigned)a.D)];
    const float2* __restrict__ w = a.tables + (size_t)d * N;
    const float2* __restrict__ code = a.code_fft + (size_t)row * P::SPEC_LEN;
    const float2* __restrict__ tw = a.tw;
    const float2* __restrict__ rot = a.rot ? a.rot + (size_t)d * a.n_coh : nullptr;
    const int n_coh = a.n_coh;
    const int n_groups = a.K / n_coh;

    float acc[G0::ITERS][G0::R];
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++)
#pragma unroll
        for (int j = 0; j < G0::R; j++) acc[it][j] = 0.f;

    for (int g = 0; g < n_groups; g++) {
        stage0_wipe_forward<P>(a, w, rot, n_coh, g, line, tw);
        __syncthreads();
        DifRange<P, 1, LASTS, false>::run(line, tw);

        // ---- last forward stage + x conj(code) + first inverse stage, in registers
#pragma unroll 1
        for (int it = 0; it < GM::ITERS; it++) {
            const int b = threadIdx.x + it * P::T;
            if (GM::NB % P::T == 0 || b < GM::NB) {
                const int base = b * GM::R;
                float2 v[GM::R];
#pragma unroll
                for (int j = 0; j < GM::R; j++) v[j] = line[P::phys(base + j)];
                if constexpr (GM::R == 31 && P::NESTED31) dft_run_like_emit<GM::R, false, true>(v);   // = acq_forward_kernel's last stage
                else Dft<GM::R, false>::run(v);
#pragma unroll
                for (int q = 0; q < GM::R; q++) v[q] = cmul_conj(v[q], __ldg(&code[q * P::SPEC_STRIDE + b]));
                dft_emit<GM::R, true, P::NESTED31>(v, [&](int j, float2 y) { line[P::phys(base + j)] = y; });
            }
        }
        __syncthreads();
        DitRange<P, LASTS - 1, 0, true>::run(line, tw);

        final_stage_accumulate<P>(line, tw, acc);
        __syncthreads();  // the next group's stage 0 overwrites the line
    }

    if (WANT_ROW) {
#pragma unroll
        for (int it = 0; it < G0::ITERS; it++) {
            const int i = threadIdx.x + it * P::T;
            if (G0::NB % P::T == 0 || i < G0::NB)
#pragma unroll
                for (int j = 0; j < G0::R; j++) a.row_out[P::PFA ? a.npos[i + j * G0::SUB] : i + j * G0::SUB] = acc[it][j];
        }
        return;
    }

    reduce_row_to_cell<P>(acc, line, a.spc, &a.cells[(size_t)row * a.D + d], a.npos);
}

// ------------------------------------------------------------------ shared-forward chain
// The forward path (wipe-off, coherent pre-sum, forward FFT) does not depend on the PRN, yet the
// reference recomputes it in each of its 32 workers (do_acquisition.rs:176-182).  Kernel A computes
// it once per (Doppler bin, group) and leaves the scrambled spectrum in L2/HBM (transposed like the
// code spectra); kernel B then runs, per (PRN, Doppler bin), x conj(code) -> inverse FFT -> |.|^2
// accumulate -> cell.  Same arithmetic as the fused kernel, bit for bit.
template <class P> __global__ void __launch_bounds__(P::T, P::MINB) acq_forward_kernel(const AcqArgs a)
{
    extern __shared__ float2 line[];
    constexpr int LASTS = P::NSTAGE - 1;
    using GM = StageGeo<P, LASTS>;
    constexpr int N = P::N;
    // one launch covers the groups [g_lo, g_lo + g_cnt) of every bin of the slab (the host-pointer search launches
    // one per upload slice so that the H2D copy of slice s+1 overlaps the forward path of slice s)
    const int n_groups = a.K / a.n_coh;
    const int dl = (int)(blockIdx.x / (unsigned)a.g_cnt);
    const int d = a.fwd_bins ? __ldg(&a.fwd_bins[dl]) : a.d_lo + dl;
    const int g = a.g_lo + (int)(blockIdx.x % (unsigned)a.g_cnt);
    const float2* __restrict__ w = a.tables + (size_t)d * N;
    const float2* __restrict__ rot = a.rot ? a.rot + (size_t)d * a.n_coh : nullptr;
    float2* __restrict__ out = a.spec + ((size_t)dl * n_groups + g) * P::SPEC_LEN;
    stage0_wipe_forward<P>(a, w, rot, a.n_coh, g, line, a.tw);
    __syncthreads();
    DifRange<P, 1, LASTS, false>::run(line, a.tw);
#pragma unroll 1
    for (int it = 0; it < GM::ITERS; it++) {
        const int b = threadIdx.x + it * P::T;
        if (GM::NB % P::T == 0 || b < GM::NB) {
            float2 v[GM::R];
#pragma unroll
            for (int j = 0; j < GM::R; j++) v[j] = line[P::phys(b * GM::R + j)];
            dft_emit<GM::R, false, P::NESTED31>(v, [&](int q, float2 y) { out[q * P::SPEC_STRIDE + b] = y; });
        }
    }
}

// DB = true: the line is double-buffered (2 x N complex of shared memory).  The barrier at the end of a group
// disappears: warps that finish the accumulate stage of group g early start loading and transforming group g+1 into
// the other buffer, so the L2 latency at the head of stage A overlaps the slower warps' tail (the barrier after stage
// A of g+1 is what guarantees everybody has left buffer g before stage A of g+2 overwrites it).
// ALIAS = true: Doppler aliasing (AcqArgs::inv_map): the spectrum slot and the shifted code-spectrum set of the bin are
// looked up.  A separate instantiation because these kernels sit on the register cliff: the three extra lines tripled
// the spills of the ALIAS = false form (config 1: 0.62 -> 0.74 ms), which therefore stays exactly as it was.
// TMODE bit 0: the power accumulators live in tensor memory instead of registers; bit 1: so do the thread's code-spectrum
// values of the first inverse stage, which are the same for every group (acq_common.cuh; gb_tuning_set("acq_tmem", 0..3):
// identical cells in every mode).
template <class P, bool DB, bool ALIAS, int TMODE = 0> __global__ void __launch_bounds__(P::T, P::MINB) acq_inverse_kernel(const AcqArgs a)
{
    constexpr bool TM = (TMODE & 1) != 0, CT = (TMODE & 2) != 0;
    extern __shared__ float2 smem_line[];
    __shared__ uint32_t tmem_base_smem;
    constexpr int LASTS = P::NSTAGE - 1;
    using G0 = StageGeo<P, 0>;
    using GM = StageGeo<P, LASTS>;
    constexpr int N = P::N;
    const int n_groups = a.K / a.n_coh;
    // Doppler-major block order: the n_active CTAs that share one bin's spectra run together (L2 reuse)
    const int dl = (int)(blockIdx.x / (unsigned)a.n_active);
    const int row = a.rows[blockIdx.x % (unsigned)a.n_active];
    // element offsets from a.code_fft / a.spec kept as 32-bit values (both buffers are far below 2^31 elements): the
    // 64-bit pointers are re-formed per group from the uniform base, so ALIAS costs no extra long-lived register
    unsigned code_off, spec_off;
    if constexpr (ALIAS) {
        const int2 sm = __ldg(&a.inv_map[a.d_lo + dl]);   // {spectrum slot, shifted code set}
        code_off = ((unsigned)sm.y * (unsigned)a.n_prn + (unsigned)row) * (unsigned)P::SPEC_LEN;
        spec_off = (unsigned)sm.x * (unsigned)n_groups * (unsigned)P::SPEC_LEN;
    } else {
        code_off = (unsigned)row * (unsigned)P::SPEC_LEN;
        spec_off = (unsigned)dl * (unsigned)n_groups * (unsigned)P::SPEC_LEN;
    }
    const float2* __restrict__ tw = a.tw;

    float acc[G0::ITERS][G0::R];
    uint32_t tmem_base = 0, taddr = 0, tcode = 0;
    if constexpr (TMODE != 0) {
        tmem_base = tmem_alloc_cta<tmem_cols_cta<P, TMODE>()>(&tmem_base_smem);
        taddr = tmem_thread_addr<P, TMODE>(tmem_base);
        tcode = taddr + (TM ? tmem_acc_cols<P>() : 0);
    }
    if constexpr (CT) {
        // park this thread's code-spectrum values: iteration it, chunk c holds values 4 c .. 4 c + 3 (zeros past the end)
        const float2* __restrict__ code = a.code_fft + code_off;
        const int warp0 = threadIdx.x & ~31;
#pragma unroll 1
        for (int it = 0; it < GM::ITERS; it++) {
            if (warp0 + it * P::T >= GM::NB) continue;
            const int b = threadIdx.x + it * P::T;
            const bool active = GM::NB % P::T == 0 || b < GM::NB;
#pragma unroll
            for (int c = 0; c < (int)tmem_code_chunks<P>(); c++) {
                float cv[8];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int q = 4 * c + u;
                    const float2 w = (q < GM::R && active) ? __ldg(&code[q * P::SPEC_STRIDE + b]) : make_float2(0.f, 0.f);
                    cv[2 * u] = w.x;
                    cv[2 * u + 1] = w.y;
                }
                tmem_st<8>(tcode + (it * tmem_code_chunks<P>() + c) * 8, cv);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if constexpr (TM) {
        tmem_zero_accumulators<P>(taddr);
    } else {
#pragma unroll
        for (int it = 0; it < G0::ITERS; it++)
#pragma unroll
            for (int j = 0; j < G0::R; j++) acc[it][j] = 0.f;
    }

    for (int g = 0; g < n_groups; g++) {
        float2* __restrict__ line = DB ? smem_line + (g & 1) * P::LINE : smem_line;
        const float2* __restrict__ sg = a.spec + (spec_off + (unsigned)g * (unsigned)P::SPEC_LEN);
        const float2* __restrict__ code = a.code_fft + code_off;
#pragma unroll 1
        for (int it = 0; it < GM::ITERS; it++) {
            const int b = threadIdx.x + it * P::T;
            if (GM::NB % P::T == 0 || b < GM::NB) {
                if constexpr (GM::R == 31 && P::STREAM_A > 0) {
                    // radix 31 from global memory: streamed inputs, accumulators in registers (see dft_odd_prime_stream)
                    dft_odd_prime_stream<GM::R, true, P::STREAM_A>(
                        [&](int q) { return cmul_conj(__ldg(&sg[q * P::SPEC_STRIDE + b]), __ldg(&code[q * P::SPEC_STRIDE + b])); },
                        [&](int j, float2 y) { line[P::phys(b * GM::R + j)] = y; });
                } else if constexpr (!CT) {
                    float2 v[GM::R];
#pragma unroll
                    for (int q = 0; q < GM::R; q++) v[q] = cmul_conj(__ldg(&sg[q * P::SPEC_STRIDE + b]), __ldg(&code[q * P::SPEC_STRIDE + b]));
                    dft_emit<GM::R, true, P::NESTED31>(v, [&](int j, float2 y) { line[P::phys(b * GM::R + j)] = y; });
                }
            }
            if constexpr (CT) {
                // code values from tensor memory in double-buffered chunks; the tcgen05 operations are warp-wide, so they
                // sit outside the per-thread guard (a warp entirely past the last butterfly skips the iteration)
                if ((int)(threadIdx.x & ~31) + it * P::T < GM::NB) {
                    const bool active = GM::NB % P::T == 0 || b < GM::NB;
                    constexpr int NCH = (int)tmem_code_chunks<P>();
                    const uint32_t tc = tcode + it * NCH * 8;
                    float cv[2][8];
                    tmem_ld8_nm(cv[0], tc);
                    float2 v[GM::R];
                    if (active) {
#pragma unroll
                        for (int q = 0; q < GM::R; q++) v[q] = __ldg(&sg[q * P::SPEC_STRIDE + b]);
                    }
#pragma unroll
                    for (int c = 0; c < NCH; c++) {
                        tmem_wait_ld8_nm(cv[c & 1]);
                        if (c + 1 < NCH) tmem_ld8_nm(cv[(c + 1) & 1], tc + 8 * (c + 1));
                        if (active) {
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int q = 4 * c + u;
                                if (q < GM::R) v[q] = cmul_conj(v[q], make_float2(cv[c & 1][2 * u], cv[c & 1][2 * u + 1]));
                            }
                        }
                    }
                    if (active) dft_emit<GM::R, true, P::NESTED31>(v, [&](int j, float2 y) { line[P::phys(b * GM::R + j)] = y; });
                }
            }
        }
        __syncthreads();
        DitRange<P, LASTS - 1, 0, true>::run(line, tw);
        if constexpr (TM) final_stage_accumulate_tmem<P>(line, tw, taddr);
        else final_stage_accumulate<P>(line, tw, acc);
        if (!DB) __syncthreads();
    }
    if (DB) __syncthreads();  // reduce_row_to_cell reuses the line as scratch
    if constexpr (TM) tmem_load_accumulators<P>(taddr, acc);
    reduce_row_to_cell<P>(acc, smem_line, a.spc, &a.cells[(size_t)row * a.D + a.d_lo + dl], a.npos);
    if constexpr (TMODE != 0) {
        if ((threadIdx.x >> 5) == 0) tmem_dealloc_warp<tmem_cols_cta<P, TMODE>()>(tmem_base);   // after reduce_row_to_cell's barriers
    }
}

// ------------------------------------------------------------------ code spectra (AcquisitionWorker::new, :133-138)
// One CTA per PRN: +-1 code samples -> forward DIF -> scrambled spectrum, stored transposed
// ([q][b] for the last radix) so the fused middle stage reads it coalesced.
template <class P> __global__ void __launch_bounds__(P::T) code_fft_kernel(const int8_t* __restrict__ codes,
                                                                           float2* __restrict__ code_fft,
                                                                           const float2* __restrict__ tw,
                                                                           const int* __restrict__ npos)
{
    extern __shared__ float2 line[];
    constexpr int LASTS = P::NSTAGE - 1;
    using G0 = StageGeo<P, 0>;
    using GM = StageGeo<P, LASTS>;
    const int8_t* c = codes + (size_t)blockIdx.x * P::N;
    float2* out = code_fft + (size_t)blockIdx.x * P::SPEC_LEN;
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) {
        const int i = threadIdx.x + it * P::T;
        if (G0::NB % P::T == 0 || i < G0::NB) {
            float2 v[G0::R];
#pragma unroll
            for (int j = 0; j < G0::R; j++)
                v[j] = make_float2((float)c[P::PFA ? npos[i + j * G0::SUB] : i + j * G0::SUB], 0.f);
            Dft<G0::R, false>::run(v);
            line[P::phys(i)] = v[0];
#pragma unroll
            for (int q = 1; q < G0::R; q++)
                line[P::phys(i + q * G0::SUB)] = P::PFA ? v[q] : cmul(v[q], __ldg(&tw[(q - 1) * G0::SUB + i]));
        }
    }
    __syncthreads();
    DifRange<P, 1, LASTS, false>::run(line, tw);
#pragma unroll 1
    for (int it = 0; it < GM::ITERS; it++) {
        const int b = threadIdx.x + it * P::T;
        if (GM::NB % P::T == 0 || b < GM::NB) {
            float2 v[GM::R];
#pragma unroll
            for (int j = 0; j < GM::R; j++) v[j] = line[P::phys(b * GM::R + j)];
            Dft<GM::R, false>::run(v);
#pragma unroll
            for (int q = 0; q < GM::R; q++) out[q * P::SPEC_STRIDE + b] = v[q];
        }
    }
}

// ------------------------------------------------------------------ FFT facade (fft.rs:5-56)
// Natural-order in, natural-order out: DIF in the requested direction, un-scrambled on the store.
template <class P, bool INV> __global__ void __launch_bounds__(P::T) fft_c2c_kernel(const FftArgs a)
{
    extern __shared__ float2 line[];
    constexpr int LASTS = P::NSTAGE - 1;
    using G0 = StageGeo<P, 0>;
    using GM = StageGeo<P, LASTS>;
    const size_t boff = (size_t)blockIdx.x * P::N;
    const float2* __restrict__ tw = a.tw;
#pragma unroll
    for (int it = 0; it < G0::ITERS; it++) {
        const int i = threadIdx.x + it * P::T;
        if (G0::NB % P::T == 0 || i < G0::NB) {
            float2 v[G0::R];
#pragma unroll
            for (int j = 0; j < G0::R; j++) {
                const size_t idx = boff + (P::PFA ? a.npos[i + j * G0::SUB] : i + j * G0::SUB);
                v[j] = a.real_in ? make_float2(reinterpret_cast<const float*>(a.in)[idx], 0.f)
                                 : reinterpret_cast<const float2*>(a.in)[idx];
            }
            Dft<G0::R, INV>::run(v);
            line[P::phys(i)] = v[0];
#pragma unroll
            for (int q = 1; q < G0::R; q++) {
                if (P::PFA) {
                    line[P::phys(i + q * G0::SUB)] = v[q];
                } else {
                    const float2 t = __ldg(&tw[(q - 1) * G0::SUB + i]);
                    line[P::phys(i + q * G0::SUB)] = INV ? cmul_conj(v[q], t) : cmul(v[q], t);
                }
            }
        }
    }
    __syncthreads();
    DifRange<P, 1, LASTS, INV>::run(line, tw);
    const size_t ooff = (size_t)blockIdx.x * a.n_out;
#pragma unroll 1
    for (int it = 0; it < GM::ITERS; it++) {
        const int b = threadIdx.x + it * P::T;
        if (GM::NB % P::T == 0 || b < GM::NB) {
            float2 v[GM::R];
#pragma unroll
            for (int j = 0; j < GM::R; j++) v[j] = line[P::phys(b * GM::R + j)];
            Dft<GM::R, INV>::run(v);
#pragma unroll
            for (int q = 0; q < GM::R; q++) {
                const int k = a.freq_of_pos[b * GM::R + q];
                if (k < a.n_out) {
                    if (a.power_out) reinterpret_cast<float*>(a.out)[ooff + k] = v[q].x * v[q].x + v[q].y * v[q].y;
                    else reinterpret_cast<float2*>(a.out)[ooff + k] = v[q];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ DopplerShiftTable::new (doppler_shift.rs:11-21)
// steps[d] = 2*pi*(f_if+f_d)/fs is evaluated on the host in f32 in the reference's order;
// phase = (i as f32) * step, table = (cos, -sin) with the full-range device cosf/sinf.
__global__ void doppler_table_kernel(const float* __restrict__ steps, int n, float2* __restrict__ tables)
{
    const float step = steps[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float phase = __fmul_rn((float)i, step);
        tables[(size_t)blockIdx.y * n + i] = make_float2(cosf(phase), -sinf(phase));
    }
}

// ------------------------------------------------------------------ prime-factor plans: line-order permutations
// dst[b * n + l] = src[(start + b * n + npos[l]) & mask]: IQ blocks (from the ring or an uploaded chunk) and the
// wipe-off tables are put into line order once, so every kernel's global access stays coalesced.
__global__ void permute_blocks_kernel(const float2* __restrict__ src, unsigned long long start, unsigned long long mask,
                                      const int* __restrict__ npos, int n, int b0, float2* __restrict__ dst)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n) {
        const unsigned long long b = (unsigned long long)b0 + blockIdx.y;
        dst[b * n + l] = src[(start + b * n + (unsigned long long)__ldg(&npos[l])) & mask];
    }
}
__global__ void shift_codes_kernel(const float2* __restrict__ src, const int* __restrict__ gidx, int n_prn, int n,
                                   float2* __restrict__ dst)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        const int s = blockIdx.z, p = blockIdx.y;
        dst[((size_t)s * n_prn + p) * n + t] = src[(size_t)p * n + __ldg(&gidx[(size_t)s * n + t])];
    }
}
cudaError_t acq_launch_shift_codes(const float2* src, const int* gidx, int n_shift, int n_prn, int n, float2* dst, cudaStream_t st)
{
    dim3 grid((n + 255) / 256, n_prn, n_shift);
    shift_codes_kernel<<<grid, 256, 0, st>>>(src, gidx, n_prn, n, dst);
    return cudaGetLastError();
}
cudaError_t acq_launch_permute(const float2* src, unsigned long long start, unsigned long long mask, const int* npos, int n,
                               int b0, int n_blocks, float2* dst, cudaStream_t st)
{
    dim3 grid((n + 255) / 256, n_blocks);
    permute_blocks_kernel<<<grid, 256, 0, st>>>(src, start, mask, npos, n, b0, dst);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ host-side dispatch
int acq_plan_index(int n)
{
#ifdef GB_TUNING
    const int v = tuning("acq_variant", 0);
    if (n == 4092 && v >= 1 && v <= 4) return 6 + v;
    if (n == 4092 && v >= 5 && v <= 8) return 8 + v;
    if (n == 16368 && v >= 1 && v <= 2) return 10 + v;
#endif
    for (int i = 0; i < kNumPlans; i++)
        if (kPlanSizes[i] == n) return i;
    return -1;
}
int acq_plan_sizes(int* sizes, int cap)
{
    int n = 0;
    for (int i = 0; i < kNumPlans; i++)
        if (kPlanSizes[i] > 0) {
            if (n < cap) sizes[n] = kPlanSizes[i];
            n++;
        }
    return n;
}

template <class P> static int plan_radices(int* r)
{
    for (int s = 0; s < P::NSTAGE; s++) r[s] = P::radix(s);
    return P::NSTAGE;
}
template <class P> static size_t plan_smem() { return sizeof(float2) * (size_t)(P::LINE < 160 ? 160 : P::LINE); }

int acq_plan_radices(int plan, int* radices)
{
    switch (plan) {
#define X(i, P) \
    case i: return plan_radices<P>(radices);
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return 0;
}
int acq_plan_is_pfa(int plan)
{
    switch (plan) {
#define X(i, P) \
    case i: return P::PFA ? 1 : 0;
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return 0;
}
int acq_plan_twiddles(int plan)
{
    switch (plan) {
#define X(i, P) \
    case i: return plan_twiddle_count<P>();
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return 0;
}
int acq_plan_spec_len(int plan)
{
    switch (plan) {
#define X(i, P) \
    case i: return P::SPEC_LEN;
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return 0;
}
int acq_plan_spec_stride(int plan)
{
    switch (plan) {
#define X(i, P) \
    case i: return P::SPEC_STRIDE;
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return 0;
}
int acq_plan_threads(int plan)
{
    switch (plan) {
#define X(i, P) \
    case i: return P::T;
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return 0;
}
size_t acq_plan_smem(int plan)
{
    switch (plan) {
#define X(i, P) \
    case i: return plan_smem<P>();
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return 0;
}

template <class K> static cudaError_t set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

template <class P> static cudaError_t launch_search(const AcqArgs& a, cudaStream_t st)
{
    const size_t smem = plan_smem<P>();
    cudaError_t e = set_smem(acq_fused_kernel<P, false>, smem);
    if (e != cudaSuccess) return e;
    acq_fused_kernel<P, false><<<a.n_active * a.D, P::T, smem, st>>>(a);
    return cudaGetLastError();
}
template <class P> static cudaError_t launch_forward(const AcqArgs& a, int n_d, cudaStream_t st)
{
    const size_t smem = plan_smem<P>();
    cudaError_t e = set_smem(acq_forward_kernel<P>, smem);
    if (e != cudaSuccess) return e;
    acq_forward_kernel<P><<<n_d * a.g_cnt, P::T, smem, st>>>(a);
    return cudaGetLastError();
}
// the generic inverse kernel with tensor-memory mode MODE (acq_inverse_kernel's TMODE), if MINB allocations fit the SM's 512 columns
template <class P, int MODE> constexpr bool tmem_mode_fits()
{
    return ((P::T + 127) / 128) * tmem_cols_per_thread<P, MODE>() <= 512 && tmem_cols_cta<P, MODE>() * P::MINB <= 512 &&
           ((MODE & 2) == 0 || P::STREAM_A == 0);
}
template <class P, int MODE>
static cudaError_t launch_inverse_tm(const AcqArgs& a, dim3 grid, size_t sm, bool db, cudaStream_t st, bool* launched)
{
    *launched = false;
    if constexpr (kProductionPlan<P> && tmem_mode_fits<P, MODE>()) {
        cudaError_t e;
#define GB_LAUNCH_TM(DBV, ALV)                                                                           \
    do {                                                                                                 \
        if ((e = set_smem(acq_inverse_kernel<P, DBV, ALV, MODE>, sm)) != cudaSuccess) return e;          \
        acq_inverse_kernel<P, DBV, ALV, MODE><<<grid, P::T, sm, st>>>(a);                                \
    } while (0)
        *launched = true;
        if (db && a.inv_map) GB_LAUNCH_TM(true, true);
        else if (db) GB_LAUNCH_TM(true, false);
        else if (a.inv_map) GB_LAUNCH_TM(false, true);
        else GB_LAUNCH_TM(false, false);
#undef GB_LAUNCH_TM
        return cudaGetLastError();
    }
    return cudaSuccess;
}
template <class P> static cudaError_t launch_shared(const AcqArgs& a, int n_d, cudaStream_t st)
{
    const size_t smem = plan_smem<P>();
    cudaError_t e;
    if (a.g_cnt > 0) {   // g_cnt == 0: the forward path was already launched slice by slice (acq_launch_forward)
        if ((e = launch_forward<P>(a, n_d, st)) != cudaSuccess) return e;
    }
    if constexpr (std::is_same<P, P4092>::value) {
        const bool no_lw = tuning("acq_nolw", 0) != 0;   // A/B switch (tools/time_acq.py)
        if (!no_lw && !a.plain_inverse) {
            return acq_launch_inverse_lw4092(a, n_d, st);   // acq_lw.cu
        }
    }
    // double-buffered line when two lines fit the 227 KB of one SM at the plan's CTA count
    const bool no_db = tuning("acq_nodb", 0) != 0;   // A/B switch (tools/time_acq.py)
    if constexpr (kProductionPlan<P>) {
        // Tensor memory (acq_inverse_kernel's TMODE; 32 PRNs x 41 bins x 10 blocks, cells identical in every mode):
        //   accumulators (1): the default where the register form spills them -- the power-of-two plans, 2 to 8 CTAs per SM
        //     at 128 registers: N = 4096 0.177 -> 0.133 ms, 2048 0.087 -> 0.068, 1024 0.057 -> 0.045 -- and a loss for the
        //     one-CTA-per-SM plans, which do not spill them and have no second CTA to hide the TMEM round trip behind
        //     (16368: 0.613 -> 0.637 ms, 8184: 0.448 -> 0.462, 20000: 2.14 -> 2.21);
        //   code spectrum (2): the default for N = 20000 (128 registers, 25-point first stage: 2.142 -> 2.028 ms for 20 blocks);
        //     the 31-point first stage of 16368 at its 96-register cap spills the extra chunk buffers (0.613 -> 0.84 ms),
        //     8184 and the power-of-two plans are indifferent;
        //   both (3): never better than the better of the two.
        // gb_tuning_set("acq_tmem", 0 .. 3) forces a mode for A/B (a mode that does not fit the 512 columns runs mode 0).
        constexpr int tm_default = (P::N & (P::N - 1)) == 0 ? 1 : (P::N == 20000 ? 2 : 0);
        const int tmq = tuning("acq_tmem", -1);
        const int tm = tmq < 0 ? tm_default : (tmq & 3);
        const bool db = P::DB && !no_db;
        const size_t sm = db ? 2 * smem : smem;
        const dim3 grid(n_d * a.n_active);
        bool launched = false;
        if (tm == 1) e = launch_inverse_tm<P, 1>(a, grid, sm, db, st, &launched);
        else if (tm == 2) e = launch_inverse_tm<P, 2>(a, grid, sm, db, st, &launched);
        else if (tm == 3) e = launch_inverse_tm<P, 3>(a, grid, sm, db, st, &launched);
        if (launched) return e;
    }
    if (a.inv_map) {
        // aliased form: the seven production plans only (the tuning variants never request it, acq_plan_supports_alias)
        if constexpr (kProductionPlan<P>) {
            if (P::DB && !no_db) {
                if ((e = set_smem(acq_inverse_kernel<P, true, true>, 2 * smem)) != cudaSuccess) return e;
                acq_inverse_kernel<P, true, true><<<n_d * a.n_active, P::T, 2 * smem, st>>>(a);
            } else {
                if ((e = set_smem(acq_inverse_kernel<P, false, true>, smem)) != cudaSuccess) return e;
                acq_inverse_kernel<P, false, true><<<n_d * a.n_active, P::T, smem, st>>>(a);
            }
            return cudaGetLastError();
        } else {
            return cudaErrorInvalidValue;
        }
    }
    if (P::DB && !no_db) {
        if ((e = set_smem(acq_inverse_kernel<P, true, false>, 2 * smem)) != cudaSuccess) return e;
        acq_inverse_kernel<P, true, false><<<n_d * a.n_active, P::T, 2 * smem, st>>>(a);
    } else {
        if ((e = set_smem(acq_inverse_kernel<P, false, false>, smem)) != cudaSuccess) return e;
        acq_inverse_kernel<P, false, false><<<n_d * a.n_active, P::T, smem, st>>>(a);
    }
    return cudaGetLastError();
}
template <class P> static cudaError_t launch_row(const AcqArgs& a, cudaStream_t st)
{
    const size_t smem = plan_smem<P>();
    cudaError_t e = set_smem(acq_fused_kernel<P, true>, smem);
    if (e != cudaSuccess) return e;
    acq_fused_kernel<P, true><<<1, P::T, smem, st>>>(a);
    return cudaGetLastError();
}
template <class P> static cudaError_t launch_code_fft(const int8_t* codes, int n_prn, float2* code_fft, const float2* tw,
                                                     const int* npos, cudaStream_t st)
{
    const size_t smem = plan_smem<P>();
    cudaError_t e = set_smem(code_fft_kernel<P>, smem);
    if (e != cudaSuccess) return e;
    code_fft_kernel<P><<<n_prn, P::T, smem, st>>>(codes, code_fft, tw, npos);
    return cudaGetLastError();
}
template <class P> static cudaError_t launch_fft(int inverse, const FftArgs& a, int batch, cudaStream_t st)
{
    const size_t smem = plan_smem<P>();
    cudaError_t e;
    if (inverse) {
        if ((e = set_smem(fft_c2c_kernel<P, true>, smem)) != cudaSuccess) return e;
        fft_c2c_kernel<P, true><<<batch, P::T, smem, st>>>(a);
    } else {
        if ((e = set_smem(fft_c2c_kernel<P, false>, smem)) != cudaSuccess) return e;
        fft_c2c_kernel<P, false><<<batch, P::T, smem, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t acq_launch_search(int plan, const AcqArgs& a, cudaStream_t st)
{
    switch (plan) {
#define X(i, P) \
    case i: return launch_search<P>(a, st);
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
cudaError_t acq_launch_shared(int plan, const AcqArgs& a, int n_d, cudaStream_t st)
{
    switch (plan) {
#define X(i, P) \
    case i: return launch_shared<P>(a, n_d, st);
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
cudaError_t acq_launch_forward(int plan, const AcqArgs& a, int n_d, cudaStream_t st)
{
    switch (plan) {
#define X(i, P) \
    case i: return launch_forward<P>(a, n_d, st);
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
cudaError_t acq_launch_row(int plan, const AcqArgs& a, cudaStream_t st)
{
    switch (plan) {
#define X(i, P) \
    case i: return launch_row<P>(a, st);
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
cudaError_t acq_launch_code_fft(int plan, const int8_t* codes, int n_prn, float2* code_fft, const float2* tw,
                                const int* npos, cudaStream_t st)
{
    switch (plan) {
#define X(i, P) \
    case i: return launch_code_fft<P>(codes, n_prn, code_fft, tw, npos, st);
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
cudaError_t acq_launch_fft(int plan, int inverse, const FftArgs& a, int batch, cudaStream_t st)
{
    switch (plan) {
#define X(i, P) \
    case i: return launch_fft<P>(inverse, a, batch, st);
        GB_FOR_EACH_PLAN(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
cudaError_t acq_launch_doppler_tables(const float* steps_dev, int D, int n, float2* tables, cudaStream_t st)
{
    dim3 grid((n + 255) / 256, D);
    doppler_table_kernel<<<grid, 256, 0, st>>>(steps_dev, n, tables);
    return cudaGetLastError();
}

}  // namespace gb
