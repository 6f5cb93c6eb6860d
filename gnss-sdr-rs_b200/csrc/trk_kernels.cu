// trk_kernels.cu -- batched early/prompt/late correlators + DLL/PLL, sm_100a.  Compile with -fmad=false:
// the epoch-end scalar arithmetic must round exactly like the reference's f32 code
// (do_tracking.rs:231-302); FMAs are written explicitly where they are wanted.
//
// One CTA per tracking channel.  Per epoch (TrackingChannel::do_work, do_tracking.rs:183-210):
//   carrier NCO replica generated in-register (phase_i = carrier_phase + 2*pi*f*i/fs, :232-238),
//   code replica from the C/A row in shared memory (E/P/L at chip+0.5/chip/chip-0.5, :250-263),
//   six correlator sums, then thread 0 runs the lock test, the atan-Costas PLL, the normalised
//   envelope DLL and the sample bookkeeping (:186-209, :279-302).  With n_epochs > 1 the CTA is
//   persistent: state never leaves the SM between epochs (the sequential dependence of the loops
//   is inside one CTA, channels are independent => no inter-CTA communication at all).
#include "trk_common.cuh"

namespace gb {

// WIDE (ORDERED only): the six product sequences of an epoch are stored as six float arrays instead of rotated samples +
// int8 chips, so that the ordered sums are one 16-byte load per four additions (24 instead of 11 bytes per sample).
__host__ __device__ inline int trk_wide_stride(int n_max) { return ((n_max + 31) / 32) * 32 + 4; }   // = 4 mod 32: six lanes, six bank groups

template <int MODE, int TRK_T, bool WIDE = false> __global__ void __launch_bounds__(TRK_T) trk_kernel(const TrkArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* row = reinterpret_cast<float*>(smem_raw);          // 1024 floats: this channel's C/A row
    float* red = row + 1024;                                  // 8 + 6 x (TRK_T/32) partials (128 floats reserved)
    __shared__ gb_trk_channel st;
    __shared__ int s_go;
    // FAST: {chip k-1 (chip 0 for k = 0, the saturating cast of Q7), chip k, chip k+1 (chip 0 for k = 1022), 0} per chip:
    // ONE 16-byte look-up per sample delivers the prompt chip and both candidates of the early and the late replica
    float4* row4 = reinterpret_cast<float4*>(red + 128);
    float2* rot = reinterpret_cast<float2*>(red + 128);       // ORDERED: n_max rotated samples
    int8_t* chips = reinterpret_cast<int8_t*>(rot + (MODE == GB_TRK_ORDERED ? a.n_max : 0));  // 3 x n_max
    float* prod = reinterpret_cast<float*>(red + 128);        // ORDERED + WIDE: 6 x trk_wide_stride(n_max) products
    const int wstride = trk_wide_stride(a.n_max);

    const int c = blockIdx.x;
    if (threadIdx.x == 0) st = a.ch[c];
    __syncthreads();
    {
        const int8_t* src = a.ca_table + (size_t)(st.code_row < 32 ? st.code_row : 0) * 1023;
        for (int i = threadIdx.x; i < 1023; i += TRK_T) row[i] = (float)src[i];
    }
    int ran = 0, lost = 0;
    // loop-invariant filter gains, rounded exactly as the reference evaluates them inside run_loop_filters (:286, :298)
    const float pll_g1 = 0.001f / st.pll_tau1, pll_g2 = st.pll_tau2 / st.pll_tau1;
    const float dll_g1 = 0.001f / st.dll_tau1, dll_g2 = st.dll_tau2 / st.dll_tau1;
    // "may this channel consume an epoch now?" -- TrackingChannel::update (do_tracking.rs:160-172).  Evaluated for the
    // first epoch here and for every later one at the end of the serial section (one barrier per epoch less).
    auto may_go = [&]() -> int {
        int go = (st.state == GB_TRK_TRACKING) || !a.filters;
        const unsigned long long n = st.num_samples_per_code;
        if (a.offsets == nullptr) {
            if ((long long)(a.head - (st.next_sample_index + n)) < 0) go = 0;          // the ring does not hold it yet
            if (a.capacity && a.head - st.next_sample_index > a.capacity) go = 0;      // already overwritten
        }
        if (n == 0 || n > (unsigned long long)a.n_max) go = 0;
        return go;
    };
    // Per-epoch scalars every thread needs (carrier step, chip step with its IEEE division, window start, batch bound):
    // computed ONCE per epoch by the thread that owns the fields they derive from -- the carrier half (thread 0) and the
    // code half (thread 32) of the epoch end, for the NEXT epoch -- and broadcast through shared memory.  Evaluated by
    // every thread they were ~100 of the ~900 instructions a thread issues per epoch at 16 samples per thread.
    __shared__ struct {
        float w, f_turn, cp_turn, carrier_phase;   // carrier half
        float wc, ws;                              // cos / sin of the carrier advance over TRK_T samples
        float code_step, code_phase, i_end;        // code half
        int n;
        unsigned long long start;
    } ep;
    constexpr int U_BATCH = TRK_T >= 512 ? 4 : 8;
    const float fs = st.fs;                        // never changes during a run
    const float rcp_fs = 1.0f / fs;
    auto prep_carrier = [&]() {
        ep.carrier_phase = st.carrier_phase;
        ep.w = kTwoPi * st.carrier_freq;                          // 2.0 * PI * carrier_freq
        ep.f_turn = st.carrier_freq * rcp_fs;
        ep.cp_turn = st.carrier_phase * 0.15915494309189535f;
        const float dt = ep.f_turn * (float)TRK_T;                // turns between a thread's consecutive samples
        sincosf((dt - rintf(dt)) * 6.28318548202514648f, &ep.ws, &ep.wc);
    };
    auto prep_code = [&]() {
        const int n = (int)st.num_samples_per_code;
        ep.n = n;
        ep.start = a.offsets ? a.offsets[c] : st.next_sample_index;
        ep.code_phase = st.code_phase;
        ep.code_step = st.code_rate / fs;                         // (code_rate / fs)
        // the batched loop evaluates indices up to the end of the last batch; samples past n are zeros
        ep.i_end = (float)(((n + U_BATCH * TRK_T - 1) / (U_BATCH * TRK_T)) * (U_BATCH * TRK_T));
    };
    if (threadIdx.x == 0) {
        s_go = may_go();
        prep_carrier();
    }
    if (threadIdx.x == 32) prep_code();
    __syncthreads();
    if (MODE == GB_TRK_FAST) {
        for (int i = threadIdx.x; i < 1023; i += TRK_T)
            row4[i] = make_float4(row[i == 0 ? 0 : i - 1], row[i], row[i == 1022 ? 0 : i + 1], 0.f);
        __syncthreads();
    }

    for (int e = 0; e < a.n_epochs; e++) {
        if (!s_go) break;

        const unsigned lc_old = st.lost_counter;   // read before the carrier half rewrites it
        const int n = ep.n;
        const unsigned long long start = ep.start;
        const float carrier_phase = ep.carrier_phase;
        const float w = ep.w;
        const float code_phase = ep.code_phase;
        const float code_step = ep.code_step;

        float ip = 0.f, qp = 0.f, ie = 0.f, qe = 0.f, il = 0.f, ql = 0.f;
        if (MODE == GB_TRK_FAST) {
            // Throughput form.  The f32 phase / chip arguments are evaluated exactly as the reference does
            // (same roundings; the division is a reciprocal + one FMA-residual correction, correctly rounded
            // in all but vanishingly rare halfway cases); sin/cos use a 2-term Cody-Waite reduction to
            // [-pi, pi] and the SFU (abs error < 1e-6); the six sums are per-thread partials + tree reduction.
            const float inv_2pi = 0.15915494309189535f;
            const float c1 = 6.28318548202514648f;        // fl(2 pi)
            const float c2 = -1.74845553146951715e-7f;    // 2 pi - fl(2 pi)
            // Per-sample body.  `sane` (checked once per epoch, uniform) says that every chip argument of the epoch
            // lies in [0, 3*1023): then `% 1023` is at most two exact subtractions and the E/P/L indices need no
            // range checks, so the loop is branch-free.
            // (the batched loop evaluates indices up to n - 1 + (U - 1) * TRK_T; samples past n are zeros)
            constexpr int U = U_BATCH;
            const float i_end = ep.i_end;                                  // first index past the last batch
            const bool sane = code_phase >= 0.f && code_phase < 1023.f && code_step >= 0.f &&
                              code_step * i_end < 2040.f && fabsf(carrier_phase) + fabsf(w) * (i_end * rcp_fs) < 1.0e6f;
            auto body = [&](const float2 x, const float fi) {
                const float t = w * fi;
                const float q0 = t * rcp_fs;
                const float q = fmaf(fmaf(-q0, fs, t), rcp_fs, q0);      // (w * i) / fs
                const float phase = carrier_phase + q;
                const float k = rintf(phase * inv_2pi);
                const float r = fmaf(-k, c2, fmaf(-k, c1, phase));
                const float cs = __cosf(r), sn = __sinf(r);
                const float re = fmaf(x.x, cs, x.y * sn);               // x * (cos, -sin)
                const float im = fmaf(x.y, cs, -(x.x * sn));
                float tc = code_phase + (fi * code_step);
                float pc, ec, lc;
                if (sane) {
                    tc = tc >= 1023.f ? tc - 1023.f : tc;                 // exact (Sterbenz), == fmodf
                    tc = tc >= 1023.f ? tc - 1023.f : tc;
                    const int ipx = (int)tc;                              // tc in [0, 1023): trunc == floor
                    int iex = (int)(tc + 0.5f);
                    iex = iex >= 1023 ? iex - 1023 : iex;                 // (chip + 0.5).floor() % 1023
                    const int ilx = max((int)(tc - 0.5f), 0);             // Q7: negative saturates to chip 0
                    pc = row[ipx]; ec = row[iex]; lc = row[ilx];
                } else {
                    tc = fmodf(tc, 1023.f);
                    pc = ca_chip(row, tc); ec = ca_chip(row, tc + 0.5f); lc = ca_chip(row, tc - 0.5f);
                }
                ip = fmaf(re, pc, ip); qp = fmaf(im, pc, qp);
                ie = fmaf(re, ec, ie); qe = fmaf(im, ec, qe);
                il = fmaf(re, lc, il); ql = fmaf(im, lc, ql);
            };
            // the epoch's samples are contiguous unless they straddle the ring's wrap point
            const unsigned long long s0 = start & a.mask;
            const bool contiguous = (a.mask == ~0ull) || (s0 + (unsigned long long)n <= a.mask + 1ull);
            // cycles per sample and start phase in turns: phase_i / 2 pi = cp_turn + i * f_turn (one FMA per sample).  Used
            // while the epoch spans < 16 turns: the argument error is then < 1e-6 turn, two orders below the FAST-mode
            // tolerance.
            const float f_turn = ep.f_turn;
            const float cp_turn = ep.cp_turn;
            const float wc = ep.wc, ws = ep.ws;
            const bool turns_ok = fabsf(f_turn) * i_end < 16.f && fabsf(cp_turn) < 2.f;   // baseband / low-IF carriers
            // C/A look-ups straight from the raw bits of the round-down add: index = bits - 0x4B000000, address =
            // row4 + 16 * index = 16 * bits + row4_bias (mod 2^32), one LEA per look-up
            const unsigned row4_bias = (unsigned)__cvta_generic_to_shared(row4) - 16u * 0x4B000000u;
            // one exact conditional subtraction brings the chip argument into [0, 1023) when it stays below 2 * 1023
            const bool single_wrap = code_phase + code_step * i_end < 2046.f;
            if (contiguous && sane) {
                // batches of 8 samples per thread, software-pipelined by hand: all loads, then all carrier
                // phases / SFU sin-cos, then the code look-ups and the 48 FMAs -- so the load and SFU latencies
                // of one sample hide behind the arithmetic of the other seven.  Samples past n load as 0.
                // Only the two MUFU calls per sample touch the quarter-rate XU pipe: sample indices are kept as floats
                // (exact integers), floor / rint go through round-mode adds (floor_small / rint_small).
                const float2* __restrict__ px = a.samples + s0;
                pk64 accp = pk2(0.f, 0.f), acce = accp, accl = accp;
                for (int base = threadIdx.x; base < n; base += U * TRK_T) {
                    float2 x[U];
                    float cs[U], sn[U];
                    const float fbase = (float)base;
                    if (base + (U - 1) * TRK_T < n) {   // a full batch (all but the epoch's last one): no per-sample bound check
#pragma unroll
                        for (int u = 0; u < U; u++) x[u] = __ldg(px + base + u * TRK_T);
                    } else {
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            const int i = base + u * TRK_T;
                            x[u] = i < n ? __ldg(px + i) : make_float2(0.f, 0.f);
                        }
                    }
                    if (turns_ok) {
                        // the batch's first sample through the SFU; in the throughput regime the other seven by rotating it with
                        // the constant advance over TRK_T samples (4 instructions instead of 8; the 7-step recurrence adds
                        // < 1e-6 rad)
                        const float ut = fmaf(fbase, f_turn, cp_turn);
                        const float r = (ut - rint_small(ut)) * c1;        // [-pi, pi]
                        cs[0] = __cosf(r);
                        sn[0] = __sinf(r);
                        if constexpr (TRK_T <= 128) {
                            // throughput regime (many channels per SM, issue-bound): fewer instructions win
#pragma unroll
                            for (int u = 1; u < U; u++) {
                                cs[u] = fmaf(cs[u - 1], wc, -(sn[u - 1] * ws));
                                sn[u] = fmaf(sn[u - 1], wc, cs[u - 1] * ws);
                            }
                        } else {
                            // latency regime (one channel per SM): eight independent SFU evaluations beat a 7-step chain
#pragma unroll
                            for (int u = 1; u < U; u++) {
                                const float fi = fbase + (float)(u * TRK_T);   // exact: integers below 2^24
                                const float utu = fmaf(fi, f_turn, cp_turn);
                                const float ru = (utu - rint_small(utu)) * c1;
                                cs[u] = __cosf(ru);
                                sn[u] = __sinf(ru);
                            }
                        }
                    } else {
                        // IF carriers: thousands of turns per epoch -- follow the reference's f32 roundings of the phase
                        // argument (the oracle shares them) and reduce with a two-term Cody-Waite step
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            const float fi = fbase + (float)(u * TRK_T);
                            const float t = w * fi;
                            const float q0 = t * rcp_fs;
                            const float q = fmaf(fmaf(-q0, fs, t), rcp_fs, q0);     // (w * i) / fs, correctly rounded
                            const float phase = carrier_phase + q;
                            const float k = rint_small(phase * inv_2pi);
                            const float r = fmaf(-k, c2, fmaf(-k, c1, phase));
                            cs[u] = __cosf(r);
                            sn[u] = __sinf(r);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const float fi = fbase + (float)(u * TRK_T);
                        const float re = fmaf(x[u].x, cs[u], x[u].y * sn[u]);
                        const float im = fmaf(x[u].y, cs[u], -(x[u].x * sn[u]));
                        float tc = code_phase + (fi * code_step);
                        tc = tc >= 1023.f ? tc - 1023.f : tc;
                        if (!single_wrap) tc = tc >= 1023.f ? tc - 1023.f : tc;
                        // floor(tc) sits in the mantissa of the round-down add; the early / late chips are
                        // floor(tc + 0.5) % 1023 in {k, k+1} and max(floor(tc - 0.5), 0) in {k-1, k} (Q7), decided on the
                        // reference's own f32 sums tc + 0.5 and tc - 0.5
                        const float pf = __fadd_rd(tc, 8388608.0f);
                        const float4 q = lds_f32x4(16u * (unsigned)__float_as_int(pf) + row4_bias);
                        const float fl = pf - 8388608.0f;                     // exact
                        const float pc = q.y;
                        const float ec = (tc + 0.5f) >= (fl + 1.0f) ? q.z : q.y;
                        const float lc = (tc - 0.5f) >= fl ? q.y : q.x;
                        const pk64 z = pk2(re, im);
                        accp = fma2s(z, pc, accp);
                        acce = fma2s(z, ec, acce);
                        accl = fma2s(z, lc, accl);
                    }
                }
                upk2(accp, ip, qp);
                upk2(acce, ie, qe);
                upk2(accl, il, ql);
            } else if (contiguous) {
                const float2* __restrict__ px = a.samples + s0;
#pragma unroll 4
                for (int i = threadIdx.x; i < n; i += TRK_T) body(__ldg(px + i), (float)i);
            } else {
#pragma unroll 4
                for (int i = threadIdx.x; i < n; i += TRK_T)
                    body(__ldg(&a.samples[(start + (unsigned long long)i) & a.mask]), (float)i);
            }
        } else {
            for (int i = threadIdx.x; i < n; i += TRK_T) {
                const float2 x = __ldg(&a.samples[(start + (unsigned long long)i) & a.mask]);
                const float phase = carrier_phase + (w * (float)i) / fs;  // :233
                float cs, sn;
                carrier<MODE>(phase, cs, sn);
                const float sin_p = -sn;
                const float re = x.x * cs - x.y * sin_p;   // Complex32 multiply (:237)
                const float im = x.x * sin_p + x.y * cs;
                const float chip_idx = mod1023(code_phase + ((float)i * code_step));  // :251
                if (WIDE) {
                    // the six terms of :256-262, each an exact product with +-1
                    const float pc = ca_chip(row, chip_idx), ec = ca_chip(row, chip_idx + 0.5f), lc = ca_chip(row, chip_idx - 0.5f);
                    prod[i] = re * pc; prod[wstride + i] = im * pc;
                    prod[2 * wstride + i] = re * ec; prod[3 * wstride + i] = im * ec;
                    prod[4 * wstride + i] = re * lc; prod[5 * wstride + i] = im * lc;
                } else {
                    rot[i] = make_float2(re, im);
                    chips[i] = (int8_t)ca_chip(row, chip_idx);
                    chips[a.n_max + i] = (int8_t)ca_chip(row, chip_idx + 0.5f);
                    chips[2 * a.n_max + i] = (int8_t)ca_chip(row, chip_idx - 0.5f);
                }
            }
        }
        if (MODE == GB_TRK_ORDERED) {
            __syncthreads();
            // six sums in sample order, one thread each (do_tracking.rs:256-262)
            if (WIDE && threadIdx.x < 6) {
                // one dependent chain of additions per sum; a single warp issues an instruction every ~3.5 cycles, so
                // the chain is fed with one 16-byte load per four additions and nothing else
                const float* v = prod + threadIdx.x * wstride;
                float acc = 0.f;
                int i = 0;
#pragma unroll 4
                for (; i + 4 <= n; i += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(v + i);
                    acc = acc + q.x;
                    acc = acc + q.y;
                    acc = acc + q.z;
                    acc = acc + q.w;
                }
                for (; i < n; i++) acc = acc + v[i];
                red[threadIdx.x] = acc;
            } else if (threadIdx.x < 6) {
                const int comp = threadIdx.x & 1, which = threadIdx.x >> 1;
                const float* xs = reinterpret_cast<const float*>(rot) + comp;
                const int8_t* ch = chips + which * a.n_max;
                float acc = 0.f;
#pragma unroll 8
                for (int i = 0; i < n; i++) acc = acc + xs[2 * i] * (float)ch[i];
                red[threadIdx.x] = acc;
            }
            __syncthreads();
        } else {
            // Six sums per warp with 8 shuffles instead of 30: every exchange halves the number of values a lane still
            // carries (xor 16: {ip, qp, ie} | {qe, il, ql}; xor 8: two | one of those three; xor 4: one of two), the last two
            // exchanges finish the single value left.  Lane bits 4..2 then say which sum a lane holds.
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
            float k0 = (h16 ? qe : ip) + __shfl_xor_sync(0xffffffffu, h16 ? ip : qe, 16);
            float k1 = (h16 ? il : qp) + __shfl_xor_sync(0xffffffffu, h16 ? qp : il, 16);
            float k2 = (h16 ? ql : ie) + __shfl_xor_sync(0xffffffffu, h16 ? ie : ql, 16);
            float m0 = (h8 ? k2 : k0) + __shfl_xor_sync(0xffffffffu, h8 ? k0 : k2, 8);
            float m1 = (h8 ? 0.f : k1) + __shfl_xor_sync(0xffffffffu, h8 ? k1 : 0.f, 8);
            float r = (h4 ? m1 : m0) + __shfl_xor_sync(0xffffffffu, h4 ? m0 : m1, 4);
            r += __shfl_xor_sync(0xffffffffu, r, 2);
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            // lanes 0, 4, 8 (bits 4..2 = 000, 001, 010) hold ip, qp, ie; lanes 16, 20, 24 hold qe, il, ql
            if ((lane & 3) == 0 && !(h8 && h4)) red[8 + warp * 6 + (h16 ? 3 : 0) + (h8 ? 2 : (h4 ? 1 : 0))] = r;
            __syncthreads();
        }

        // ---- epoch end (do_work, :183-210): the carrier half (lock bookkeeping, Costas PLL, prompt outputs) runs on
        // thread 0 and the code half (DLL, sample bookkeeping, next epoch's go/no-go) on the first thread of warp 1, at
        // the same time; they own disjoint fields of the channel.  Both read the six sums and the old lost_counter
        // before anything is written; in FAST mode each sums the per-warp partials itself (no extra barrier).
        if (threadIdx.x == 0 || threadIdx.x == 32) {
            float six[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                if (MODE == GB_TRK_ORDERED) {
                    six[k] = red[k];
                } else {
                    float sfin = 0.f;
#pragma unroll
                    for (int wv = 0; wv < TRK_T / 32; wv++) sfin += red[8 + wv * 6 + k];
                    six[k] = sfin;
                }
            }
            const float i_p = six[0], q_p = six[1], i_e = six[2], q_e = six[3], i_l = six[4], q_l = six[5];
            const float nf = (float)n;
            const float power = i_p * i_p + q_p * q_p;                       // :186
            const bool locked = power > 15.0f;
            const bool resets = a.filters && !locked && lc_old + 1u >= 20u;   // reset() this epoch (:199-201)
            if (threadIdx.x == 0) {
                // :240-242 (fmod_small == fmodf bit for bit in the range the loops produce)
                const float cph = st.carrier_phase + w * (nf / fs);
                st.carrier_phase = fabsf(cph) < 1.0e6f ? fmod_small(cph, kTwoPi, 0.15915494309189535f) : fmodf(cph, kTwoPi);
                st.i_prompt = i_p;
                st.q_prompt = q_p;
                if (MODE != GB_TRK_ORDERED) {
#pragma unroll
                    for (int k = 0; k < 6; k++) red[k] = six[k];             // the last epoch's sums leave at kernel end
                }
                if (a.prompt_hist)
                    reinterpret_cast<float2*>(a.prompt_hist)[(size_t)e * a.n_channels + c] = make_float2(i_p, q_p);
                if (a.filters) {
                    if (locked) {
                        st.lost_counter = 0;
                        // run_loop_filters, carrier part (:279-290)
                        // FAST: reciprocal-multiply division (2 ulp) and a multiply by 1/(2 pi) -- the serial section is a
                        // dependent chain that competes with the other channels' sample loops for issue slots, so every
                        // IEEE division (~10 dependent instructions) removed shortens the epoch
                        float pll_err;
                        if (MODE == GB_TRK_ORDERED) pll_err = (float)atan((double)(q_p / i_p)) / kTwoPi;
                        else pll_err = atanf(__fdividef(q_p, i_p)) * 0.15915494309189535f;
                        st.carrier_nco = pll_err * pll_g1 + (pll_err - st.carrier_error) * pll_g2;
                        st.carrier_error = pll_err;
                        st.carrier_freq += st.carrier_nco;
                    } else if (resets) {                                      // reset(), carrier fields (:311-326, Q9)
                        st.lost_counter = 0;
                        st.carrier_freq = 0.f; st.carrier_phase = 0.f; st.carrier_error = 0.f; st.carrier_nco = 0.f;
                        st.i_prompt = 0.f; st.q_prompt = 0.f;
                        lost = 1;
                    } else {
                        st.lost_counter += 1;
                    }
                }
                ran += 1;
                prep_carrier();   // next epoch's carrier scalars
            } else {
                // :265-267
                const float cdp = st.code_phase + code_step * nf;
                st.code_phase = fabsf(cdp) < 1.0e8f ? fmod_small(cdp, 1023.f, 9.775171065493646e-4f) : fmodf(cdp, 1023.f);
                if (a.filters) {
                    if (locked) {
                        // run_loop_filters, code part (:291-301)
                        const float pow_e = sqrtf(i_e * i_e + q_e * q_e);
                        const float pow_l = sqrtf(i_l * i_l + q_l * q_l);
                        float dll_err;
                        if (MODE == GB_TRK_ORDERED) dll_err = ((pow_e + pow_l) != 0.f) ? (pow_e - pow_l) / (pow_e + pow_l) : 0.f;
                        else dll_err = ((pow_e + pow_l) != 0.f) ? __fdividef(pow_e - pow_l, pow_e + pow_l) : 0.f;
                        st.code_nco = dll_err * dll_g1 + (dll_err - st.code_error) * dll_g2;
                        st.code_error = dll_err;
                        st.code_rate += st.code_nco;
                    }
                    if (resets) {                                             // reset(), code / bookkeeping fields
                        st.prn = 0; st.code_row = 0; st.state = GB_TRK_IDLE;
                        st.next_sample_index = 0;
                        st.code_phase = 0.f; st.code_error = 0.f; st.code_nco = 0.f; st.code_rate = 0.f;
                    } else {
                        st.next_sample_index += st.num_samples_per_code;
                        const float spc = roundf(st.fs / (st.code_rate / 1023.0f));
                        // `as usize` (saturating, NaN -> 0); the 32-bit path covers every finite sample rate in use
                        st.num_samples_per_code = (spc >= 0.f && spc < 2.0e9f) ? (unsigned long long)(unsigned)spc : f32_as_usize(spc);
                    }
                }
                st.epochs_done += 1;
                prep_code();      // next epoch's code scalars and sample window
                s_go = (e + 1 < a.n_epochs) ? may_go() : 0;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.ch[c] = st;
        if (ran > 0) {
            gb_trk_corr out;
            out.i_p = red[0]; out.q_p = red[1]; out.i_e = red[2]; out.q_e = red[3]; out.i_l = red[4]; out.q_l = red[5];
            a.corr[c] = out;
        }
        if (a.ran) a.ran[c] = (uint8_t)(ran > 255 ? 255 : ran);
        if (a.lost) a.lost[c] = (uint8_t)lost;
    }
}

size_t trk_ordered_smem_bytes(int n_max) { return (1024 + 128) * sizeof(float) + (size_t)n_max * (sizeof(float2) + 3) + 16; }
static size_t trk_ordered_wide_bytes(int n_max) { return (1024 + 128) * sizeof(float) + 6 * (size_t)trk_wide_stride(n_max) * sizeof(float) + 16; }

template <int T> static cudaError_t launch_t(const TrkArgs& a, int mode, cudaStream_t st)
{
    size_t smem = (1024 + 128) * sizeof(float) + (mode == GB_TRK_ORDERED ? 0 : 1024 * sizeof(float4));
    // the wide layout while four channels still fit an SM (n_max up to ~2300 samples: the 2.048 Msps configs)
    if (mode == GB_TRK_ORDERED && trk_ordered_wide_bytes(a.n_max) <= 56 * 1024 && tuning("trk_narrow", 0) == 0) {
        smem = trk_ordered_wide_bytes(a.n_max);
        cudaError_t e = cudaFuncSetAttribute(trk_kernel<GB_TRK_ORDERED, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        trk_kernel<GB_TRK_ORDERED, T, true><<<a.n_channels, T, smem, st>>>(a);
    } else if (mode == GB_TRK_ORDERED) {
        smem = trk_ordered_smem_bytes(a.n_max);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(trk_kernel<GB_TRK_ORDERED, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        trk_kernel<GB_TRK_ORDERED, T><<<a.n_channels, T, smem, st>>>(a);
    } else {
        trk_kernel<GB_TRK_FAST, T><<<a.n_channels, T, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// FAST ring-fed epochs with loop filters (gb_trk_run / gb_trk_epoch) run the warp-specialised kernel of trk_ws.cu;
// ORDERED mode, the open-loop gb_trk_correlate and linear buffers run trk_kernel.  CTA size of trk_kernel: with many
// channels 128 threads keep every SM full of independent channels; with few the run is bound by the per-epoch latency
// of one CTA, so each channel gets more threads.  gb_tuning_set("trk_ws", -1) selects trk_kernel everywhere (A/B),
// gb_tuning_set("trk_t", 64 | 128 | 256 | 512) its CTA size.
cudaError_t trk_launch(const TrkArgs& a0, int mode, cudaStream_t st)
{
    if (a0.n_channels <= 0) return cudaSuccess;
    TrkArgs a = a0;
    a.dbg = tuning("trk_dbg", 0);
    const int ws = tuning("trk_ws", 0);
    if (mode == GB_TRK_FAST && ws >= 0 && trk_ws_supported(a)) return trk_ws_launch(a, st, ws);
    switch (tuning("trk_t", 0)) {
    case 64: return launch_t<64>(a, mode, st);
    case 128: return launch_t<128>(a, mode, st);
    case 256: return launch_t<256>(a, mode, st);
    case 512: return launch_t<512>(a, mode, st);
    default: break;
    }
    if (mode == GB_TRK_ORDERED) {
        // the parallel part of an ORDERED epoch is the f64 sin/cos per sample: one channel per SM takes 512 threads for
        // it, more channels 256 (measured: 128 ch 10.3 us / epoch, 1024 ch 28.1 us; tools/time_trk_ordered.py)
        return a.n_channels <= 148 ? launch_t<512>(a, mode, st) : launch_t<256>(a, mode, st);
    }
    if (a.n_channels > 300) return launch_t<128>(a, mode, st);
    return launch_t<256>(a, mode, st);
}

}  // namespace gb
