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
#include "trk_kernels.cuh"

namespace gb {

static __device__ __constant__ float kTwoPi = 6.28318530717958647692f;  // 2.0 * std::f32::consts::PI

// Rust `as usize` for f32: saturating, NaN -> 0 (Q7)
__device__ __forceinline__ unsigned long long f32_as_usize(float v)
{
    if (!(v > 0.f)) return 0ull;
    if (v >= 18446744073709551616.f) return ~0ull;
    return (unsigned long long)v;
}

// get_ca_chip (do_tracking.rs:274-277): floor, saturating cast, % 1023
__device__ __forceinline__ float ca_chip(const float* __restrict__ row, float phase)
{
    const float f = floorf(phase);
    unsigned idx = f > 0.f ? (f < 4.0e9f ? (unsigned)f : (unsigned)(f32_as_usize(f) % 1023ull)) : 0u;
    idx = idx % 1023u;
    return row[idx];
}

// x % 1023.0 (fmodf is exact; fast path for the only range the loops ever produce)
__device__ __forceinline__ float mod1023(float t)
{
    if (t >= 0.f && t < 2046.f) return t >= 1023.f ? t - 1023.f : t;
    return fmodf(t, 1023.f);
}

template <int MODE> __device__ __forceinline__ void carrier(float phase, float& c, float& s)
{
    if (MODE == GB_TRK_ORDERED) {
        // glibc's sinf/cosf are (nearly always) correctly rounded; so is the f64 result rounded to f32
        double sd, cd;
        sincos((double)phase, &sd, &cd);
        c = (float)cd;
        s = (float)sd;
    } else {
        sincosf(phase, &s, &c);
    }
}

__device__ __forceinline__ float block_sum(float v, float* red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int MODE, int TRK_T> __global__ void __launch_bounds__(TRK_T) trk_kernel(const TrkArgs a)
{
    extern __shared__ unsigned char smem_raw[];
    float* row = reinterpret_cast<float*>(smem_raw);          // 1024 floats: this channel's C/A row
    float* red = row + 1024;                                  // 8 + 6 x (TRK_T/32) partials (128 floats reserved)
    __shared__ gb_trk_channel st;
    __shared__ int s_go;
    float2* rot = reinterpret_cast<float2*>(red + 128);       // ORDERED: n_max rotated samples
    int8_t* chips = reinterpret_cast<int8_t*>(rot + (MODE == GB_TRK_ORDERED ? a.n_max : 0));  // 3 x n_max

    const int c = blockIdx.x;
    if (threadIdx.x == 0) st = a.ch[c];
    __syncthreads();
    {
        const int8_t* src = a.ca_table + (size_t)(st.code_row < 32 ? st.code_row : 0) * 1023;
        for (int i = threadIdx.x; i < 1023; i += TRK_T) row[i] = (float)src[i];
    }
    int ran = 0, lost = 0;

    for (int e = 0; e < a.n_epochs; e++) {
        if (threadIdx.x == 0) {
            int go = (st.state == GB_TRK_TRACKING) || !a.filters;
            const unsigned long long n = st.num_samples_per_code;
            if (a.offsets == nullptr) {
                // TrackingChannel::update (do_tracking.rs:168-172): wait until the ring holds the epoch
                if ((long long)(a.head - (st.next_sample_index + n)) < 0) go = 0;
                if (a.capacity && a.head - st.next_sample_index > a.capacity) go = 0;  // overwritten
            }
            if (n == 0 || n > (unsigned long long)a.n_max) go = 0;
            s_go = go;
        }
        __syncthreads();
        if (!s_go) break;

        const int n = (int)st.num_samples_per_code;
        const unsigned long long start = a.offsets ? a.offsets[c] : st.next_sample_index;
        const float carrier_phase = st.carrier_phase, fs = st.fs;
        const float w = kTwoPi * st.carrier_freq;                 // 2.0 * PI * carrier_freq
        const float code_phase = st.code_phase;
        const float code_step = st.code_rate / fs;                 // (code_rate / fs)

        float ip = 0.f, qp = 0.f, ie = 0.f, qe = 0.f, il = 0.f, ql = 0.f;
        if (MODE == GB_TRK_FAST) {
            // Throughput form.  The f32 phase / chip arguments are evaluated exactly as the reference does
            // (same roundings; the division is a reciprocal + one FMA-residual correction, correctly rounded
            // in all but vanishingly rare halfway cases); sin/cos use a 2-term Cody-Waite reduction to
            // [-pi, pi] and the SFU (abs error < 1e-6); the six sums are per-thread partials + tree reduction.
            const float rcp_fs = 1.0f / fs;
            const float inv_2pi = 0.15915494309189535f;
            const float c1 = 6.28318548202514648f;        // fl(2 pi)
            const float c2 = -1.74845553146951715e-7f;    // 2 pi - fl(2 pi)
            // Per-sample body.  `sane` (checked once per epoch, uniform) says that every chip argument of the epoch
            // lies in [0, 3*1023): then `% 1023` is at most two exact subtractions and the E/P/L indices need no
            // range checks, so the loop is branch-free.
            const bool sane = code_phase >= 0.f && code_phase < 1023.f && code_step >= 0.f &&
                              code_step * (float)(n + 8 * TRK_T) < 2040.f;
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
            if (contiguous && sane) {
                // batches of 8 samples per thread, software-pipelined by hand: all loads, then all carrier
                // phases / SFU sin-cos, then the code look-ups and the 48 FMAs -- so the load and SFU latencies
                // of one sample hide behind the arithmetic of the other seven.  Samples past n load as 0.
                const float2* __restrict__ px = a.samples + s0;
                constexpr int U = TRK_T >= 512 ? 4 : 8;
                for (int base = threadIdx.x; base < n; base += U * TRK_T) {
                    float2 x[U];
                    float cs[U], sn[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int i = base + u * TRK_T;
                        x[u] = i < n ? __ldg(px + i) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const float fi = (float)(base + u * TRK_T);
                        const float t = w * fi;
                        const float q0 = t * rcp_fs;
                        const float q = fmaf(fmaf(-q0, fs, t), rcp_fs, q0);
                        const float phase = carrier_phase + q;
                        const float k = rintf(phase * inv_2pi);
                        const float r = fmaf(-k, c2, fmaf(-k, c1, phase));
                        cs[u] = __cosf(r);
                        sn[u] = __sinf(r);
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const float fi = (float)(base + u * TRK_T);
                        const float re = fmaf(x[u].x, cs[u], x[u].y * sn[u]);
                        const float im = fmaf(x[u].y, cs[u], -(x[u].x * sn[u]));
                        float tc = code_phase + (fi * code_step);
                        tc = tc >= 1023.f ? tc - 1023.f : tc;
                        tc = tc >= 1023.f ? tc - 1023.f : tc;
                        const int ipx = (int)tc;
                        int iex = (int)(tc + 0.5f);
                        iex = iex >= 1023 ? iex - 1023 : iex;
                        const int ilx = max((int)(tc - 0.5f), 0);
                        const float pc = row[ipx], ec = row[iex], lc = row[ilx];
                        ip = fmaf(re, pc, ip); qp = fmaf(im, pc, qp);
                        ie = fmaf(re, ec, ie); qe = fmaf(im, ec, qe);
                        il = fmaf(re, lc, il); ql = fmaf(im, lc, ql);
                    }
                }
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
                rot[i] = make_float2(re, im);
                chips[i] = (int8_t)ca_chip(row, chip_idx);
                chips[a.n_max + i] = (int8_t)ca_chip(row, chip_idx + 0.5f);
                chips[2 * a.n_max + i] = (int8_t)ca_chip(row, chip_idx - 0.5f);
            }
        }
        if (MODE == GB_TRK_ORDERED) {
            __syncthreads();
            // six sums in sample order, one thread each (do_tracking.rs:256-262)
            if (threadIdx.x < 6) {
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
            ip = block_sum(ip, red); qp = block_sum(qp, red); ie = block_sum(ie, red);
            qe = block_sum(qe, red); il = block_sum(il, red); ql = block_sum(ql, red);
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (lane == 0) {
                red[8 + warp * 6 + 0] = ip; red[8 + warp * 6 + 1] = qp; red[8 + warp * 6 + 2] = ie;
                red[8 + warp * 6 + 3] = qe; red[8 + warp * 6 + 4] = il; red[8 + warp * 6 + 5] = ql;
            }
            __syncthreads();
            if (threadIdx.x < 6) {
                float s = 0.f;
#pragma unroll
                for (int wv = 0; wv < TRK_T / 32; wv++) s += red[8 + wv * 6 + threadIdx.x];
                red[threadIdx.x] = s;
            }
            __syncthreads();
        }

        if (threadIdx.x == 0) {
            const float i_p = red[0], q_p = red[1], i_e = red[2], q_e = red[3], i_l = red[4], q_l = red[5];
            const float nf = (float)n;
            // :240-242 and :265-267
            st.carrier_phase = fmodf(st.carrier_phase + w * (nf / fs), kTwoPi);
            st.code_phase = fmodf(st.code_phase + code_step * nf, 1023.f);
            st.i_prompt = i_p;
            st.q_prompt = q_p;
            gb_trk_corr out;
            out.i_p = i_p; out.q_p = q_p; out.i_e = i_e; out.q_e = q_e; out.i_l = i_l; out.q_l = q_l;
            a.corr[c] = out;
            if (a.prompt_hist) {
                a.prompt_hist[((size_t)e * a.n_channels + c) * 2 + 0] = i_p;
                a.prompt_hist[((size_t)e * a.n_channels + c) * 2 + 1] = q_p;
            }
            if (a.filters) {
                const float power = i_p * i_p + q_p * q_p;  // :186
                bool advance = true;
                if (power > 15.0f) {
                    st.lost_counter = 0;
                    // run_loop_filters (:279-302)
                    float pll_err;
                    if (MODE == GB_TRK_ORDERED) pll_err = (float)atan((double)(q_p / i_p)) / kTwoPi;
                    else pll_err = atanf(q_p / i_p) / kTwoPi;
                    st.carrier_nco = pll_err * (0.001f / st.pll_tau1) + (pll_err - st.carrier_error) * (st.pll_tau2 / st.pll_tau1);
                    st.carrier_error = pll_err;
                    st.carrier_freq += st.carrier_nco;
                    const float pow_e = sqrtf(i_e * i_e + q_e * q_e);
                    const float pow_l = sqrtf(i_l * i_l + q_l * q_l);
                    const float dll_err = ((pow_e + pow_l) != 0.f) ? (pow_e - pow_l) / (pow_e + pow_l) : 0.f;
                    st.code_nco = dll_err * (0.001f / st.dll_tau1) + (dll_err - st.code_error) * (st.dll_tau2 / st.dll_tau1);
                    st.code_error = dll_err;
                    st.code_rate += st.code_nco;
                } else {
                    st.lost_counter += 1;
                    if (st.lost_counter >= 20) {  // reset() (:311-326, Q9)
                        st.prn = 0; st.code_row = 0; st.state = GB_TRK_IDLE; st.lost_counter = 0;
                        st.next_sample_index = 0;
                        st.carrier_freq = 0.f; st.carrier_phase = 0.f; st.carrier_error = 0.f; st.carrier_nco = 0.f;
                        st.code_phase = 0.f; st.code_error = 0.f; st.code_nco = 0.f; st.code_rate = 0.f;
                        st.i_prompt = 0.f; st.q_prompt = 0.f;
                        lost = 1;
                        advance = false;
                    }
                }
                if (advance) {
                    st.next_sample_index += st.num_samples_per_code;
                    st.num_samples_per_code = f32_as_usize(roundf(st.fs / (st.code_rate / 1023.0f)));
                }
            }
            st.epochs_done += 1;
            ran += 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.ch[c] = st;
        if (a.ran) a.ran[c] = (uint8_t)(ran > 255 ? 255 : ran);
        if (a.lost) a.lost[c] = (uint8_t)lost;
    }
}

template <int T> static cudaError_t launch_t(const TrkArgs& a, int mode, cudaStream_t st)
{
    size_t smem = (1024 + 128) * sizeof(float);
    if (mode == GB_TRK_ORDERED) {
        smem += (size_t)a.n_max * (sizeof(float2) + 3) + 16;
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

// CTA size: with >= 600 channels 128 threads keep every SM full of independent channels; with fewer channels the
// run is bound by the per-epoch latency of one CTA, so each channel gets more threads.
cudaError_t trk_launch(const TrkArgs& a, int mode, cudaStream_t st)
{
    if (a.n_channels <= 0) return cudaSuccess;
    if (a.n_channels >= 600) return launch_t<128>(a, mode, st);
    if (a.n_channels >= 250) return launch_t<256>(a, mode, st);
    return launch_t<512>(a, mode, st);
}

}  // namespace gb
