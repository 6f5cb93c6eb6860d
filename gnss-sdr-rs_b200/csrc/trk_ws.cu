// trk_ws.cu -- warp-specialised form of the FAST tracking kernel (ring-fed do_work epochs with on-device loop filters).
// Compile with -fmad=false like trk_kernels.cu.
//
// One CTA per channel, as in trk_kernels.cu, but the CTA is split into NSW *sample warps* and two *control warps*:
//
//   sample warps   early_late_correlation (do_tracking.rs:231-272): carrier / code replicas in registers, six partial
//                  sums, 8-shuffle warp reduction, one float per (sum, warp) to shared memory.  They never execute the
//                  serial code.  The FIRST batch of the next epoch's window is loaded while the current epoch is being
//                  computed: its start (next_sample_index + n) is known when the epoch begins, only its length is not,
//                  and samples past the length are masked at use -- the L2 latency of the window leaves the per-epoch
//                  critical path (it was ~1/5 of it with one channel per SM).
//   carrier warp   lane 0: lock test, atan-Costas PLL (run_loop_filters, :279-290), carrier_phase advance (:240-242),
//                  prompt outputs; publishes the next epoch's carrier scalars.
//   code warp      lane 0: lock test, envelope DLL (:291-301), code_phase advance (:265-267), sample bookkeeping of
//                  do_work / update (:160-210), next epoch's go / no-go; publishes the code scalars.
//
// Everything the two control lanes can know before the sums arrive is computed while the sample warps run: the phase
// advances (they use the PRE-filter carrier_freq / code_rate of the running epoch), the window start, the reset
// bookkeeping.  Between "sums complete" and "next epoch may start" only the discriminator -> filter -> NCO -> per-sample
// step chain remains (sum 8 partials, divide, atan, two multiply-adds | two square roots, divide, two multiply-adds,
// one division by fs).  Hand-offs are named barriers, never a CTA-wide __syncthreads():
//   PART  sample warps have stored their partials / control warps may read them;
//   GO    control warps have published the epoch's scalars / sample warps may read them;
//   X     code warp -> carrier warp (n and go of the epoch just published).
// All of them are bar.sync on both sides (bar.arrive does not order the arriving thread's shared-memory stores, and a
// fence costs more than the wait it would save: the side that "only arrives" is the one that waits next anyway --
// sample warps go from PART straight to GO, and by the time a control warp reaches GO they are already parked there).
#include "trk_common.cuh"

#include <type_traits>

namespace gb {

namespace {

enum { BAR_GO = 1, BAR_PART = 2, BAR_X = 3 };

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// scalars of one epoch, written by the control lanes and read by every sample thread after GO
struct __align__(16) EpochParams {
    // carrier warp
    float f_turn, cp_turn, w, carrier_phase;
    float wc, ws;
    int carr_ok;        // the turns form of the carrier argument is valid for this epoch
    int pad0;
    // code warp
    float code_step, code_phase;
    int n;
    int flags;          // bit 0: go, bit 1: chip arguments stay in [0, 3 * 1023), bit 2: ... in [0, 2 * 1023)
    unsigned s0;        // (window start) & ring mask
    int pad1[3];
};

__device__ __forceinline__ float sqrt_fast(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

}  // namespace

// atan(q / i) for the Costas discriminator (FAST mode): ONE reciprocal (of the larger magnitude) instead of the two of
// atanf(q / i), the minimax polynomial of atanf on [0, 1] evaluated in Estrin form (depth 5 instead of 9); <= 3 ulp.
__device__ __forceinline__ float atan_ratio_fast(float q, float i)
{
    const float aq = fabsf(q), ai = fabsf(i);
    const float mx = fmaxf(aq, ai), mn = fminf(aq, ai);
    const float t = mn * rcp_fast(mx);                 // in [0, 1]; 0/0 -> NaN like atanf(0/0)
    const float s = t * t, s2 = s * s, s4 = s2 * s2;
    const float p01 = fmaf(0.19988775253295898438f, s, -0.33332940936088562012f);
    const float p23 = fmaf(0.10518480092287063599f, s, -0.14171802997589111328f);
    const float p45 = fmaf(0.039849750697612762451f, s, -0.072529748082160949707f);
    const float p67 = fmaf(0.0024500258732587099075f, s, -0.014396979473531246185f);
    const float lo = fmaf(s2, p23, p01), hi = fmaf(s2, p67, p45);
    const float P = fmaf(s4, hi, lo);
    float r = fmaf(t * s, P, t);
    if (aq > ai) r = 1.57079637050628662109f - r;
    // sign of q / i
    return __int_as_float(__float_as_int(r) ^ ((__float_as_int(q) ^ __float_as_int(i)) & 0x80000000));
}

// NSW sample warps, U samples per thread and batch, NSFU carrier evaluations per batch through the SFU (the other samples
// of each group of U / NSFU by rotating the previous one; NSFU == U: no rotation)
template <int NSW, int U, int NSFU, int MINB = (NSW >= 8 ? 2 : 4)>
__global__ void __launch_bounds__((NSW + 2) * 32, MINB) trk_ws_kernel(const TrkArgs a)
{
    constexpr int T = NSW * 32;          // sample threads
    constexpr int NT = T + 64;           // + two control warps
    constexpr bool ROT = NSFU < U;
    constexpr int CHAIN = U / NSFU;
    static_assert(U % NSFU == 0, "whole rotation chains");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* row4 = reinterpret_cast<float4*>(smem_raw);               // 2048 half chips x {early, prompt, late, 0}
    float* row = reinterpret_cast<float*>(row4 + 2048);               // 1024: plain row (general path)
    float* red = row + 1024;                                          // [6][NSW] per-warp sums
    __shared__ gb_trk_channel st;
    __shared__ EpochParams P;
    __shared__ unsigned s_bias;

    const int c = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        st = a.ch[c];
        s_bias = (unsigned)__cvta_generic_to_shared(row4) - 16u * 0x4B000000u;
    }
    __syncthreads();
    {
        const int8_t* src = a.ca_table + (size_t)(st.code_row < 32 ? st.code_row : 0) * 1023;
        for (int i = tid; i < 1023; i += NT) row[i] = (float)src[i];
        // entry h = floor(2 * chip argument): the chips get_ca_chip returns for tc + 0.5, tc and tc - 0.5 when tc lies in
        // [h / 2, (h + 1) / 2) -- floor(tc + 0.5) % 1023 (1023 -> 0), floor(tc), max(floor(tc - 0.5), 0) (Q7)
        for (int h = tid; h < 2048; h += NT) {
            const int e = ((h + 1) >> 1) % 1023, p = (h >> 1) % 1023, l = h == 0 ? 0 : ((h - 1) >> 1) % 1023;
            row4[h] = make_float4((float)src[e], (float)src[p], (float)src[l], 0.f);
        }
    }
    const float fs = st.fs;                          // never changes during a run
    const float rcp_fs = 1.0f / fs;
    const unsigned mask32 = (unsigned)a.mask;        // ring capacity <= 2^32 samples (checked by the launcher)
    // first index past the last batch of the longest epoch this launch may see: bounds of the per-epoch range checks
    const float i_end_max = (float)(((a.n_max + U * T - 1) / (U * T)) * (U * T));
    __syncthreads();

    if (warp < NSW) {
        // ------------------------------------------------------------------------------------------- sample warps
        const float2* __restrict__ smp = a.samples;
        // look-up address = 16 * (bits of the round-down add) + bias, ONE integer multiply-add per sample: the bias comes
        // back from shared memory, so the assembler cannot split it into "window base + constant" again (it then adds the
        // constant part with a second instruction per look-up)
        const unsigned row4_bias = *reinterpret_cast<volatile unsigned*>(&s_bias);
        const float c1 = 6.28318548202514648f;        // fl(2 pi)
        const float c2 = -1.74845553146951715e-7f;    // 2 pi - fl(2 pi)
        const float inv_2pi = 0.15915494309189535f;
        // a batch: U samples per thread, sample u at base + tid + u * T; one pointer + immediate offsets unless the
        // batch straddles the ring's wrap point
        auto load_batch = [&](float2(&v)[U], unsigned base) {
            const unsigned b0 = base & mask32;
            if (b0 + (unsigned)(U * T) <= mask32 + 1u && b0 + (unsigned)(U * T) >= b0) {
                const float2* __restrict__ p = smp + b0 + tid;
#pragma unroll
                for (int u = 0; u < U; u++) v[u] = __ldg(p + u * T);
            } else {
#pragma unroll
                for (int u = 0; u < U; u++) v[u] = __ldg(smp + ((base + (unsigned)(tid + u * T)) & mask32));
            }
        };
        // One epoch on the batch already in `cur`.  The first batch of the NEXT epoch's window (it starts at start + n
        // whatever the filters decide) is loaded into the same registers as soon as this epoch's arithmetic has consumed
        // them.  Returns false when the channel stops.
        float2 cur[U];
        auto epoch = [&]() -> bool {
            const float4 pc4 = *reinterpret_cast<const float4*>(&P.f_turn);
            const float4 pd4 = *reinterpret_cast<const float4*>(&P.code_step);
            const float f_turn = pc4.x, cp_turn = pc4.y, w = pc4.z, carrier_phase = pc4.w;
            const float code_step = pd4.x, code_phase = pd4.y;
            const int n = __float_as_int(pd4.z), flags = __float_as_int(pd4.w);
            const unsigned s0 = P.s0;
            float wc = 1.f, ws = 0.f;
            int carr_ok;
            if (ROT) {
                const float4 pr4 = *reinterpret_cast<const float4*>(&P.wc);
                wc = pr4.x; ws = pr4.y; carr_ok = __float_as_int(pr4.z);
            } else {
                carr_ok = P.carr_ok;
            }
            const bool fastp = (flags & 2) && carr_ok;
            const bool single_wrap = flags & 4;
            float ip = 0.f, qp = 0.f, ie = 0.f, qe = 0.f, il = 0.f, ql = 0.f;
            pk64 accp = pk2(0.f, 0.f), acce = accp, accl = accp;
            auto compute_batch = [&](int b0) {
                const int base = b0 + tid;
                if (base + (U - 1) * T >= n) {          // the epoch's ragged last batch: samples past n count as zeros
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if (base + u * T >= n) cur[u] = make_float2(0.f, 0.f);
                }
                const float fbase = (float)base;
                if (fastp) {
                    // Packed FP32 (FFMA2 / FADD2 / FMUL2): the carrier is kept as (cos, sin) pairs, the wiped sample as a
                    // (re, im) pair, and the chip arguments of two consecutive samples of the thread share every add --
                    // each lane rounds exactly like the scalar instruction.
                    pk64 cs2[U];
                    const pk64 wc2 = pk2(wc, wc), ws2 = pk2(ws, ws);
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (u % CHAIN == 0) {
                            const float fi = fbase + (float)(u * T);   // exact: integers below 2^24
                            const float ut = fmaf(fi, f_turn, cp_turn);
                            const float r = (ut - rint_small(ut)) * c1;        // [-pi, pi]
                            cs2[u] = pk2(__cosf(r), __sinf(r));
                        } else {
                            float cv, sv;
                            upk2(cs2[u - 1], cv, sv);
                            cs2[u] = fma2(pk2(-sv, cv), ws2, mul2(cs2[u - 1], wc2));   // rotate by the advance over T samples
                        }
                    }
                    auto body = [&](auto sw) {
                        const pk64 fb2 = pk2(fbase, fbase), st2 = pk2(code_step, code_step), cp2 = pk2(code_phase, code_phase);
#pragma unroll
                        for (int u = 0; u < U; u += 2) {
                            const pk64 fi2 = add2(fb2, pk2((float)(u * T), (float)((u + 1) * T)));
                            float t0, t1;
                            upk2(add2(cp2, mul2(fi2, st2)), t0, t1);              // code_phase + (i * code_step)
                            t0 = t0 >= 1023.f ? t0 - 1023.f : t0;
                            t1 = t1 >= 1023.f ? t1 - 1023.f : t1;
                            if (!decltype(sw)::value) {
                                t0 = t0 >= 1023.f ? t0 - 1023.f : t0;
                                t1 = t1 >= 1023.f ? t1 - 1023.f : t1;
                            }
                            // floor(2 tc) sits in the mantissa of ONE round-down multiply-add (2 tc is exact); the table
                            // entry of that half chip holds the three chips.  The reference decides on its f32 sums
                            // tc + 0.5 and tc - 0.5: floor(tc - 0.5) changes exactly at frac(tc) = 0.5, floor(tc + 0.5) at
                            // frac(tc) = 0.5 too EXCEPT for tc = 0.49999997 = 0.5 - 2^-25, whose sum ties up to 1.0 -- the
                            // code warp looks for that one argument and corrects the early sums (checked against the literal
                            // sums on every float within 4 ulp of every half chip and 2.5 M random arguments, DESIGN 4.3,
                            // and by the exact-selection test)
                            const pk64 pf2 = fma2_rm(pk2(t0, t1), pk2(2.0f, 2.0f), pk2(8388608.0f, 8388608.0f));
                            float pf[2];
                            upk2(pf2, pf[0], pf[1]);
#pragma unroll
                            for (int k = 0; k < 2; k++) {
                                unsigned qa;
                                asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(qa) : "r"((unsigned)__float_as_int(pf[k])), "r"(row4_bias));
                                const float4 q = lds_f32x4(qa);
                                const float ec = q.x, lc = q.z;
                                float cv, sv;
                                upk2(cs2[u + k], cv, sv);
                                const float2 x = cur[u + k];
                                const pk64 z = fma2(pk2(x.y, -x.x), pk2(sv, sv), mul2(pk2(x.x, x.y), pk2(cv, cv)));   // x * (cos, -sin)
                                accp = fma2s(z, q.y, accp);
                                acce = fma2s(z, ec, acce);
                                accl = fma2s(z, lc, accl);
                            }
                        }
                    };
                    if (single_wrap) body(std::true_type{});
                    else body(std::false_type{});
                } else {
                    // general form (IF carriers spanning thousands of turns, chip arguments outside the checked range):
                    // the reference's f32 roundings of both arguments, Cody-Waite reduction, full get_ca_chip
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const float fi = fbase + (float)(u * T);
                        const float t = w * fi;
                        const float q0 = t * rcp_fs;
                        const float q = fmaf(fmaf(-q0, fs, t), rcp_fs, q0);      // (w * i) / fs
                        const float phase = carrier_phase + q;
                        const float k = rintf(phase * inv_2pi);
                        const float r = fmaf(-k, c2, fmaf(-k, c1, phase));
                        float csv, snv;
                        if (fabsf(phase) < 1.0e6f) {
                            csv = __cosf(r);
                            snv = __sinf(r);
                        } else {
                            sincosf(phase, &snv, &csv);
                        }
                        const float re = fmaf(cur[u].x, csv, cur[u].y * snv);
                        const float im = fmaf(cur[u].y, csv, -(cur[u].x * snv));
                        const float tc = fmodf(code_phase + (fi * code_step), 1023.f);
                        const float pcv = ca_chip(row, tc), ec = ca_chip(row, tc + 0.5f), lc = ca_chip(row, tc - 0.5f);
                        ip = fmaf(re, pcv, ip); qp = fmaf(im, pcv, qp);
                        ie = fmaf(re, ec, ie); qe = fmaf(im, ec, qe);
                        il = fmaf(re, lc, il); ql = fmaf(im, lc, ql);
                    }
                }
            };
            if (n <= U * T) {
                // One batch per epoch (the latency-critical case: 2.048 Msps on 8 x 256 threads).  The next window's loads
                // are issued AFTER the arithmetic: a scoreboard wait covers every load in flight on that scoreboard, so
                // loads issued before the first use of `cur` would make that use wait for them too (and a predicated-off
                // register copy waits as well).  They have the reduction, the control section and the next epoch's
                // start-up to arrive.
                compute_batch(0);
                load_batch(cur, s0 + (unsigned)n);
            } else {
                // several batches per epoch: the next batch (or the next window's first one) is loaded while this one is
                // being computed
                float2 nxt[U];
                for (int b0 = 0;; b0 += U * T) {
                    const bool last = b0 + U * T >= n;
                    load_batch(nxt, last ? s0 + (unsigned)n : s0 + (unsigned)(b0 + U * T));
                    compute_batch(b0);
#pragma unroll
                    for (int u = 0; u < U; u++) cur[u] = nxt[u];
                    if (last) break;
                }
            }
            if (fastp) {
                upk2(accp, ip, qp);
                upk2(acce, ie, qe);
                upk2(accl, il, ql);
            }
            // Six sums per warp with 8 shuffles: every exchange halves the number of values a lane still carries
            // (xor 16: {ip, qp, ie} | {qe, il, ql}; xor 8: two | one of those three; xor 4: one of two), the last two
            // exchanges finish the single value left.  Lane bits 4..2 then say which sum a lane holds.
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
            if ((lane & 3) == 0 && !(h8 && h4)) red[((h16 ? 3 : 0) + (h8 ? 2 : (h4 ? 1 : 0))) * NSW + warp] = r;
            named_sync(BAR_PART, NT);
            named_sync(BAR_GO, NT);
            return P.flags & 1;
        };
        named_sync(BAR_GO, NT);
        if (P.flags & 1) {
            load_batch(cur, P.s0);
            while (epoch()) {}
        }
    } else if (warp == NSW) {
        // ------------------------------------------------------------------------------------------- carrier warp
        float freq = st.carrier_freq, phi = st.carrier_phase, cerr = st.carrier_error, cnco = st.carrier_nco;
        float i_prompt = st.i_prompt, q_prompt = st.q_prompt;
        unsigned lostc = st.lost_counter;
        int ran = 0, lost = 0;
        const float g1 = 0.001f / st.pll_tau1, g2 = st.pll_tau2 / st.pll_tau1;   // as evaluated inside run_loop_filters (:286)
        int n_cached = -1;
        float n_over_fs = 0.f;
        // publishes the scalars of the epoch that starts at phase `ph` with carrier `f`; ph_ok = |ph / 2 pi| < 2
        auto publish = [&](float f, float ph, float ph_turn, bool ph_ok) {
            const float f_turn = f * rcp_fs;
            const float w = kTwoPi * f;                                          // 2.0 * PI * carrier_freq
            *reinterpret_cast<float4*>(&P.f_turn) = make_float4(f_turn, ph_turn, w, ph);
            // turns form: the epoch spans < 16 turns (argument error < 1e-6 turn)
            const int ok = (fabsf(f_turn) * i_end_max < 16.f && ph_ok) ? 1 : 0;
            if (ROT) {
                const float dt = f_turn * (float)T;                              // turns between a thread's consecutive samples
                const float ang = (dt - rint_small(dt)) * 6.28318548202514648f;
                *reinterpret_cast<float4*>(&P.wc) = make_float4(__cosf(ang), __sinf(ang), __int_as_float(ok), 0.f);
            } else {
                P.carr_ok = ok;
            }
        };
        if (lane == 0) {
            const float pt = phi * 0.15915494309189535f;
            publish(freq, phi, pt, fabsf(pt) < 2.f);
        }
        __syncwarp();
        named_sync(BAR_GO, NT);
        named_sync(BAR_X, 64);
        for (int e = 0;; e++) {
            if (!(P.flags & 1)) break;                                           // published by the code warp before X
            float phi_next = 0.f, pt_next = 0.f;
            bool pt_ok = true;
            if (lane == 0) {
                // :240-242 with the carrier_freq the running epoch uses (the filters update it afterwards)
                const int n_now = P.n;
                if (n_now != n_cached) {                                         // (n as f32 / fs): one IEEE division per length
                    n_cached = n_now;
                    n_over_fs = (float)n_now / fs;
                }
                const float cph = phi + (kTwoPi * freq) * n_over_fs;
                phi_next = fabsf(cph) < 1.0e6f ? fmod_small(cph, kTwoPi, 0.15915494309189535f) : fmodf(cph, kTwoPi);
                pt_next = phi_next * 0.15915494309189535f;
                pt_ok = fabsf(pt_next) < 2.f;
            }
            named_sync(BAR_PART, NT);
            if (lane == 0) {
                float i_p, q_p;
                if (NSW % 4 == 0) {
                    float4 u = *reinterpret_cast<const float4*>(red + 0 * NSW);
                    float4 v = *reinterpret_cast<const float4*>(red + 1 * NSW);
#pragma unroll
                    for (int k = 1; k < NSW / 4; k++) {
                        const float4 u2 = *reinterpret_cast<const float4*>(red + 0 * NSW + 4 * k);
                        const float4 v2 = *reinterpret_cast<const float4*>(red + 1 * NSW + 4 * k);
                        u.x += u2.x; u.y += u2.y; u.z += u2.z; u.w += u2.w;
                        v.x += v2.x; v.y += v2.y; v.z += v2.z; v.w += v2.w;
                    }
                    i_p = (u.x + u.y) + (u.z + u.w);
                    q_p = (v.x + v.y) + (v.z + v.w);
                } else {
                    i_p = red[0]; q_p = red[NSW];
                    for (int k = 1; k < NSW; k++) { i_p += red[k]; q_p += red[NSW + k]; }
                }
                const float power = i_p * i_p + q_p * q_p;                       // :186
                const bool locked = power > 15.0f;
                const bool resets = !locked && lostc + 1u >= 20u;                // reset() this epoch (:199-201)
                phi = phi_next;
                if (locked) {
                    // run_loop_filters, carrier part (:279-290); FAST: one reciprocal, multiply by 1/(2 pi)
                    const float pll_err = atan_ratio_fast(q_p, i_p) * 0.15915494309189535f;
                    cnco = pll_err * g1 + (pll_err - cerr) * g2;
                    cerr = pll_err;
                    freq += cnco;
                }
                if (resets) { freq = 0.f; phi = 0.f; pt_next = 0.f; pt_ok = true; }
                publish(freq, phi, pt_next, pt_ok);
                // ---- off the critical path
                i_prompt = i_p;
                q_prompt = q_p;
                if (a.prompt_hist)
                    reinterpret_cast<float2*>(a.prompt_hist)[(size_t)e * a.n_channels + c] = make_float2(i_p, q_p);
                if (locked) {
                    lostc = 0;
                } else if (resets) {                                             // reset(), carrier fields (:311-326, Q9)
                    lostc = 0;
                    cerr = 0.f; cnco = 0.f; i_prompt = 0.f; q_prompt = 0.f;
                    lost = 1;
                } else {
                    lostc += 1;
                }
                ran += 1;
            }
            __syncwarp();
            named_sync(BAR_GO, NT);
            named_sync(BAR_X, 64);
        }
        if (lane == 0) {
            st.carrier_freq = freq; st.carrier_phase = phi; st.carrier_error = cerr; st.carrier_nco = cnco;
            st.i_prompt = i_prompt; st.q_prompt = q_prompt; st.lost_counter = lostc;
            if (a.ran) a.ran[c] = (uint8_t)(ran > 255 ? 255 : ran);
            if (a.lost) a.lost[c] = (uint8_t)lost;
        }
    } else {
        // ------------------------------------------------------------------------------------------- code warp
        float cphase = st.code_phase, cerr = st.code_error, cnco = st.code_nco, crate = st.code_rate;
        unsigned long long next_idx = st.next_sample_index, n64 = st.num_samples_per_code;
        unsigned epochs_done = st.epochs_done, lostc = st.lost_counter;
        int state = st.state, prn = st.prn, code_row = st.code_row;
        float step = crate / fs;                                                  // (code_rate / fs)
        float six[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int ran = 0;
        const float g1 = 0.001f / st.dll_tau1, g2 = st.dll_tau2 / st.dll_tau1;    // :298
        const float fs1023 = fs * 1023.0f;
        unsigned long long iv_n = ~0ull;   // epoch length the cached interval belongs to
        float iv_lo = 1.f, iv_hi = 0.f, iv_scale = 0.f;
        // "may this channel consume an epoch now?" -- TrackingChannel::update (do_tracking.rs:160-172)
        auto may_go = [&]() -> int {
            int go = state == GB_TRK_TRACKING;
            if ((long long)(a.head - (next_idx + n64)) < 0) go = 0;               // the ring does not hold it yet
            if (a.capacity && a.head - next_idx > a.capacity) go = 0;             // already overwritten
            if (n64 == 0 || n64 > (unsigned long long)a.n_max) go = 0;
            return go;
        };
        // the full evaluation of an epoch's code scalars (first epoch, and whenever the prediction below does not hold)
        auto publish = [&](int go) {
            const int n = (int)n64;
            const float i_end = (float)(((n + U * T - 1) / (U * T)) * (U * T));
            const bool sane = cphase >= 0.f && cphase < 1023.f && step >= 0.f && step * i_end < 2040.f;
            const bool single = cphase + step * i_end < 2046.f;
            const int flags = (go ? 1 : 0) | (sane ? 2 : 0) | (sane && single ? 4 : 0);
            *reinterpret_cast<float4*>(&P.code_step) = make_float4(step, cphase, __int_as_float(n), __int_as_float(flags));
            P.s0 = (unsigned)(next_idx & a.mask);
        };
        int go = 0;
        if (lane == 0) {
            go = a.n_epochs > 0 ? may_go() : 0;
            publish(go);
        }
        go = __shfl_sync(0xffffffffu, go, 0);
        named_sync(BAR_GO, NT);
        named_sync(BAR_X, 64);
        for (int e = 0; go; e++) {
            // ---- while the sample warps run: everything about the next epoch that does not depend on this epoch's sums.
            // Its code phase and window start are exact (they use the running epoch's code_rate); its length n' and the
            // range-check flags depend on the DLL output only through "code_rate' lies in (r_lo, r_hi)", the interval in
            // which the length stays n and the checks hold -- the critical section then only compares.
            float cphase_next = 0.f, r_lo = 1.f, r_hi = 0.f, r_single = 0.f;
            unsigned long long next_idx_next = 0;
            int go_pred = 0;
            if (lane == 0) {
                // :265-267 with the code_rate the running epoch uses
                const float nf = (float)n64;
                const float cdp = cphase + step * nf;
                cphase_next = fabsf(cdp) < 1.0e8f ? fmod_small(cdp, 1023.f, 9.775171065493646e-4f) : fmodf(cdp, 1023.f);
                next_idx_next = next_idx + n64;
                // n' = round(fs / (code_rate' / 1023)) == n  <=  fs * 1023 / code_rate' in (n - 0.45, n + 0.45); and
                // step' * i_end < 2040, cphase' + step' * i_end < 2046 as bounds on code_rate' (a little inside).  All of it
                // depends on n alone (but for one multiply): re-derived only when the epoch length changes
                if (n64 != iv_n) {
                    iv_n = n64;
                    iv_lo = fs1023 / (nf + 0.45f);
                    iv_hi = fs1023 / (nf - 0.45f);
                    const int n = (int)n64;
                    const float i_end = (float)(((n + U * T - 1) / (U * T)) * (U * T));
                    const float r_sane = (2038.f / i_end) * fs;
                    if (r_sane < iv_hi) iv_hi = r_sane;
                    if (n64 >= 100000ull) iv_hi = 0.f;
                    iv_scale = (fs / i_end) * 0.999999f;
                }
                r_lo = iv_lo;
                r_hi = iv_hi;
                r_single = (2046.f - cphase_next) * iv_scale;
                if (!(cphase_next >= 0.f && cphase_next < 1023.f)) r_hi = 0.f;   // no fast path
                const unsigned long long saved_idx = next_idx;
                next_idx = next_idx_next;
                go_pred = (e + 1 < a.n_epochs) ? may_go() : 0;                    // with n' = n, state unchanged
                next_idx = saved_idx;
            }
            // ---- the one chip argument the half-chip table gets wrong (see the sample warps): tc == 0.5 - 2^-25, where the
            // reference's early replica is already the next chip.  Only a sample before the first code wrap can produce it
            // (after a wrap tc = t - 1023 with t in [1023, 1024) is a multiple of 2^-14), and tc is monotonic there, so at
            // most one sample hits: three lanes test the samples around the crossing of 0.5 with the sample warps' own
            // arithmetic; a hit (about once in 1e7 epochs) adds x * carrier * (chip 1 - chip 0) to the early sums.
            float fix_re = 0.f, fix_im = 0.f;
            {
                const float4 pd4 = *reinterpret_cast<const float4*>(&P.code_step);
                const float step_r = pd4.x, cph_r = pd4.y;
                const int n_r = __float_as_int(pd4.z), fl_r = __float_as_int(pd4.w);
                bool hit = false;
                int i_hit = 0;
                if (lane < 3 && (fl_r & 2) && cph_r < 0.5f) {
                    const float est = step_r > 0.f ? (0.5f - cph_r) * rcp_fast(step_r) : 0.f;
                    if (est < 1.0e6f) {
                        const int i = (int)rintf(est) + lane - 1;
                        if (i >= 0 && i < n_r) {
                            const float t = __fadd_rn(cph_r, __fmul_rn((float)i, step_r));
                            hit = __float_as_int(t) == 0x3EFFFFFF;
                            i_hit = i;
                        }
                    }
                }
                if (__ballot_sync(0xffffffffu, hit)) {
                    if (hit && P.carr_ok) {
                        const float2 x = __ldg(a.samples + ((P.s0 + (unsigned)i_hit) & mask32));
                        const float ut = fmaf((float)i_hit, P.f_turn, P.cp_turn);
                        const float r = (ut - rint_small(ut)) * 6.28318548202514648f;
                        const float cv = __cosf(r), sv = __sinf(r);
                        const float d = row[1] - row[0];
                        fix_re = (x.x * cv + x.y * sv) * d;
                        fix_im = (x.y * cv - x.x * sv) * d;
                    }
                    // lanes 0..3 all end with the total (lanes 2 and 3 own the early sums below)
                    fix_re += __shfl_xor_sync(0xffffffffu, fix_re, 1);
                    fix_im += __shfl_xor_sync(0xffffffffu, fix_im, 1);
                    fix_re += __shfl_xor_sync(0xffffffffu, fix_re, 2);
                    fix_im += __shfl_xor_sync(0xffffffffu, fix_im, 2);
                }
            }
            named_sync(BAR_PART, NT);
            {
                // lane k < 6 sums the per-warp partials of sum k; squares meet their partner through one shuffle, the
                // three results (prompt power, |E|^2, |L|^2) and the six sums are then gathered on lane 0
                const int k = lane < 6 ? lane : 0;
                float v;
                if (NSW % 4 == 0) {
                    float4 u = *reinterpret_cast<const float4*>(red + k * NSW);
#pragma unroll
                    for (int j = 1; j < NSW / 4; j++) {
                        const float4 u2 = *reinterpret_cast<const float4*>(red + k * NSW + 4 * j);
                        u.x += u2.x; u.y += u2.y; u.z += u2.z; u.w += u2.w;
                    }
                    v = (u.x + u.y) + (u.z + u.w);
                } else {
                    v = red[k * NSW];
                    for (int j = 1; j < NSW; j++) v += red[k * NSW + j];
                }
                if (lane == 2) v += fix_re;
                if (lane == 3) v += fix_im;
#pragma unroll
                for (int j = 0; j < 6; j++) six[j] = __shfl_sync(0xffffffffu, v, j);
            }
            if (lane == 0) {
                const float i_p = six[0], q_p = six[1], i_e = six[2], q_e = six[3], i_l = six[4], q_l = six[5];
                const float power = i_p * i_p + q_p * q_p;                        // :186
                const bool locked = power > 15.0f;
                const bool resets = !locked && lostc + 1u >= 20u;
                if (locked) {
                    // run_loop_filters, code part (:291-301); FAST: approximate square roots and division (2 ulp)
                    const float pow_e = sqrt_fast(i_e * i_e + q_e * q_e);
                    const float pow_l = sqrt_fast(i_l * i_l + q_l * q_l);
                    const float dll_err = ((pow_e + pow_l) != 0.f) ? (pow_e - pow_l) * rcp_fast(pow_e + pow_l) : 0.f;
                    cnco = dll_err * g1 + (dll_err - cerr) * g2;
                    cerr = dll_err;
                    crate += cnco;
                    lostc = 0;
                } else if (resets) {
                    lostc = 0;
                } else {
                    lostc += 1;
                }
                // (code_rate / fs): reciprocal + exact-residual correction = the correctly rounded quotient
                const float q0 = crate * rcp_fs;
                step = fmaf(fmaf(-q0, fs, crate), rcp_fs, q0);
                if (!resets && crate > r_lo && crate < r_hi) {
                    // the prediction holds: n' = n, chip arguments in range, go as predicted
                    cphase = cphase_next;
                    next_idx = next_idx_next;
                    go = go_pred;
                    const int flags = (go ? 1 : 0) | 2 | (crate < r_single ? 4 : 0);
                    *reinterpret_cast<float4*>(&P.code_step) = make_float4(step, cphase, __int_as_float((int)n64), __int_as_float(flags));
                    P.s0 = (unsigned)(next_idx & a.mask);
                } else {
                    if (resets) {                                                 // reset(), code / bookkeeping fields
                        prn = 0; code_row = 0; state = GB_TRK_IDLE;
                        next_idx = 0;
                        cphase = 0.f; cerr = 0.f; cnco = 0.f; crate = 0.f; step = 0.f;
                    } else {
                        cphase = cphase_next;
                        next_idx = next_idx_next;
                        // round(fs / (code_rate / 1023)) as usize (:165 via generate_ca_code_samples' length)
                        const float spc = roundf(fs / (crate / 1023.0f));
                        n64 = (spc >= 0.f && spc < 2.0e9f) ? (unsigned long long)(unsigned)spc : f32_as_usize(spc);
                    }
                    go = (e + 1 < a.n_epochs) ? may_go() : 0;
                    publish(go);
                }
                epochs_done += 1;
                ran += 1;
            }
            go = __shfl_sync(0xffffffffu, go, 0);
            named_sync(BAR_GO, NT);
            named_sync(BAR_X, 64);
        }
        if (lane == 0) {
            st.code_phase = cphase; st.code_error = cerr; st.code_nco = cnco; st.code_rate = crate;
            st.next_sample_index = next_idx; st.num_samples_per_code = n64; st.epochs_done = epochs_done;
            st.state = (uint8_t)state; st.prn = (uint8_t)prn; st.code_row = (uint8_t)code_row;
            if (ran > 0) {
                gb_trk_corr out;
                out.i_p = six[0]; out.q_p = six[1]; out.i_e = six[2]; out.q_e = six[3]; out.i_l = six[4]; out.q_l = six[5];
                a.corr[c] = out;
            }
        }
    }
    __syncthreads();
    if (tid == 0) a.ch[c] = st;
}

template <int NSW, int U, int NSFU, int MINB = (NSW >= 8 ? 2 : 4)> static cudaError_t launch_ws(const TrkArgs& a, cudaStream_t st)
{
    const size_t smem = 2048 * sizeof(float4) + 1024 * sizeof(float) + 6 * NSW * sizeof(float) + 64;
    trk_ws_kernel<NSW, U, NSFU, MINB><<<a.n_channels, (NSW + 2) * 32, smem, st>>>(a);
    return cudaGetLastError();
}

bool trk_ws_supported(const TrkArgs& a)
{
    return a.filters && a.offsets == nullptr && a.mask != ~0ull && a.mask < (1ull << 32);
}

// variant: 0 = by channel count, else NSW * 100 + U * 10 + NSFU (tuning, gb_tuning_set("trk_ws", v))
cudaError_t trk_ws_launch(const TrkArgs& a, cudaStream_t st, int variant)
{
    if (a.n_channels <= 0) return cudaSuccess;
    switch (variant) {
    case 888: return launch_ws<8, 8, 8>(a, st);
    case 884: return launch_ws<8, 8, 4>(a, st);
    case 882: return launch_ws<8, 8, 2>(a, st);
    case 881: return launch_ws<8, 8, 1>(a, st);
    case 1644: return launch_ws<16, 4, 4>(a, st);
    case 1642: return launch_ws<16, 4, 2>(a, st);
    case 488: return launch_ws<4, 8, 8>(a, st);
    case 482: return launch_ws<4, 8, 2>(a, st);
    case 481: return launch_ws<4, 8, 1>(a, st);
    case 4415: return launch_ws<4, 4, 1, 5>(a, st);
    case 4416: return launch_ws<4, 4, 1, 6>(a, st);
    case 4815: return launch_ws<4, 8, 1, 5>(a, st);
    case 8413: return launch_ws<8, 4, 1, 3>(a, st);
    case 2817: return launch_ws<2, 8, 1, 7>(a, st);
    case 2816: return launch_ws<2, 8, 1, 6>(a, st);
    case 2418: return launch_ws<2, 4, 1, 8>(a, st);
    case 2417: return launch_ws<2, 4, 1, 7>(a, st);
    default: break;
    }
    // up to two channels per SM: 8 sample warps per channel (latency regime, one batch per epoch at 2.048 Msps); more:
    // 4 sample warps (throughput regime).  Carrier by rotation in both (measured: tools/time_trk.py).
    if (a.n_channels <= 296) return launch_ws<8, 8, 1>(a, st);
    return launch_ws<4, 8, 1>(a, st);
}

}  // namespace gb
