// gnss_b200.cu -- the C-ABI of libgnss_b200 (include/gnss_b200.h): handle, device memory, streams,
// plan set-up and the host-side O(D) decision scan.  All sample-rate work runs in acq_kernels.cu /
// trk_kernels.cu; there is no CPU implementation of the hot path in this library.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gnss_b200.h"
#include "acq_kernels.cuh"
#include "acq_generic.cuh"
#include "trk_kernels.cuh"
#include "fine_doppler.cuh"
#include "frontend.cuh"

namespace {

const float kCodeRate = 1.023e6f;  // GPS_L1_CA_CODE_RATE_CHIPS_PER_S (gps_property_constants.rs:4)
const float kPiF = 3.14159265358979323846f;
const size_t kStageSamples = (size_t)1 << 19;  // 4 MiB per pinned staging slot

struct FftRes {
    float2* tw = nullptr;
    int* fop = nullptr;
    int* npos = nullptr;  // prime-factor plans: input index of line position l
    std::vector<int> fop_host;  // natural frequency at scrambled position l (host copy of fop)
};

}  // namespace

// ------------------------------------------------------------------ tuning switches (explicit, process-wide)
namespace {
std::mutex g_tuning_mu;
std::map<std::string, int> g_tuning;
}  // namespace

namespace gb {
int tuning(const char* key, int dflt)
{
    std::lock_guard<std::mutex> lk(g_tuning_mu);
    auto it = g_tuning.find(key);
    return it == g_tuning.end() ? dflt : it->second;
}
}  // namespace gb

extern "C" int gb_tuning_set(const char* key, int value)
{
    if (!key) return GB_EINVAL;
    std::lock_guard<std::mutex> lk(g_tuning_mu);
    g_tuning[key] = value;
    return GB_OK;
}

extern "C" int gb_tuning_get(const char* key, int dflt) { return key ? gb::tuning(key, dflt) : dflt; }

struct gb_handle {
    int device = 0;
    cudaStream_t s_acq = nullptr, s_trk = nullptr, s_copy = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_a0 = nullptr, ev_a1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    std::string last_err;
    // One lock per call family: the reference calls search_satellite from 32 rayon threads (do_acquisition.rs:302-313)
    // and runs acquisition, tracking and the sample writer on separate threads (main.rs:205-227).  Calls of one family
    // on one handle are serialised here; different families run concurrently on their own streams.
    std::recursive_mutex mu_acq, mu_trk, mu_ring;

    // sample ring
    float2* ring = nullptr;
    uint64_t ring_cap = 0, ring_head = 0;
    int8_t* i8_stage = nullptr;
    size_t i8_cap = 0;
    // pinned host staging (two slots) so pageable caller buffers are released at once and the H2D DMA is asynchronous
    float2* pin_stage[2] = {nullptr, nullptr};
    cudaEvent_t pin_free[2] = {nullptr, nullptr};
    int pin_next = 0;

    // digital front-end (rf/frontend.rs): LUT + persistent NCO / DC-bias state on the device
    float* fe_lut = nullptr;     // 2 x 2048 floats (re, im)
    float* fe_state = nullptr;   // phase_accumulator, bias_re[8], bias_im[8]
    float2* fe_stage = nullptr;
    size_t fe_cap = 0;
    float fe_step = 0.f;
    bool fe_ready = false;
    // table-driven NCO (frontend.cu): orbit of the f32 phase accumulator, LUT indices on the device
    bool fe_table = false;
    int fe_mode = 0;             // GB_FE_EXACT | GB_FE_PARALLEL
    float2* fe_scratch = nullptr;
    size_t fe_scratch_cap = 0;
    uint16_t* fe_idx = nullptr;
    std::vector<float> fe_phase;
    uint64_t fe_mu = 0, fe_period = 0, fe_count = 0;

    // acquisition
    int plan = -1, N = 0, n_prn = 0, D = 0, n_coh = 1, spc = 0, mode = GB_ACQ_SHARED;
    int spec_len = 0;   // complex elements of one stored spectrum (rows padded to 128 B, Plan::SPEC_LEN; == N for cluster plans)
    float2* spec = nullptr;
    size_t spec_cap = 0;
    // tensor-pipe A/B (gb_tuning_set("acq_tc", 1)): spectra / code spectra in fragment order; code_gen counts rebuilds of
    // code_fft / code_fft_shift, code_tc_gen / code_tc_src say what code_tc holds
    float2 *spec_tc = nullptr, *code_tc = nullptr;
    size_t spec_tc_cap = 0, code_tc_cap = 0;
    uint64_t code_gen = 1, code_tc_gen = 0;
    const float2* code_tc_src = nullptr;
    // cluster plans (code period > one CTA's shared memory)
    bool cluster = false;
    float2* otw = nullptr;
    float* acc_rows = nullptr;
    size_t acc_cap = 0;
    float fs = 0.f, threshold = 7.0f;
    float2 *tw = nullptr, *code_fft = nullptr, *tables = nullptr, *rot = nullptr, *chunk = nullptr;
    // prime-factor plans: wipe-off tables and IQ blocks in line order (PfaPlan, fft_smem.cuh)
    bool pfa = false;
    const int* npos = nullptr;
    float2 *tables_perm = nullptr, *iq_perm = nullptr;
    size_t tables_perm_cap = 0, iq_perm_cap = 0;
    int8_t* codes_dev = nullptr;
    size_t chunk_cap = 0, tables_cap = 0, rot_cap = 0;
    std::vector<float> carr;
    gb_acq_cell* cells_dev = nullptr;
    size_t cells_cap = 0;
    int *rows_dev = nullptr, *rows_pin = nullptr;   // rows_pin = rows_pin_slot[0]
    // two result slots: gb_acq_search_enqueue(slot) / gb_acq_search_wait(slot) keep one search in flight while the host
    // consumes the previous one (the synchronous calls use slot 0)
    struct Pending {
        bool active = false;
        uint64_t local_tail = 0;
        uint32_t prn_mask = 0;
        bool has_enable = false;
        std::vector<uint8_t> enable;
        int n_active = 0;
        size_t n_cells = 0;
    } pend[2];
    gb_acq_cell* cells_pin_slot[2] = {nullptr, nullptr};
    size_t cells_pin_cap[2] = {0, 0};
    int* rows_pin_slot[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_s0[2] = {nullptr, nullptr}, ev_s1[2] = {nullptr, nullptr};
    cudaEvent_t ev_chunk_free = nullptr;   // the last kernel that reads `chunk` has been enqueued behind this event
    bool chunk_free_valid = false;
    bool builtin_codes = false;            // configured with codes == NULL (GPS C/A): an identical re-configure is a no-op
    // any-length fallback plans (acq_generic.cu), one per length, shared by the acquisition and the FFT facade
    std::map<int, gb::GenericPlan*> gen_plans;
    bool generic = false;                  // the configured acquisition length has no tuned plan
    gb::GenericPlan* gen = nullptr;
    float2 *gen_s0 = nullptr, *gen_s1 = nullptr;
    size_t gen_s0_cap = 0, gen_s1_cap = 0;
    float* row_dev = nullptr;
    float last_acq_ms = 0.f;
    cudaEvent_t ev_slice[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // upload slices
    // Doppler aliasing: bins whose carriers differ by a whole number of FFT bins share one forward spectrum
    bool alias_enabled = false, alias_ok = false;
    int n_base = 0, n_shift = 0;
    int* fwd_bins_dev = nullptr;
    int2* inv_map_dev = nullptr;
    size_t fwd_bins_cap = 0, inv_map_cap = 0;
    float2* code_fft_shift = nullptr;
    size_t code_fft_shift_cap = 0;

    // fine Doppler (N3)
    float2* fine_x = nullptr;
    size_t fine_x_cap = 0;
    float2* fine_y = nullptr;
    size_t fine_y_cap = 0;
    int8_t* fine_codes = nullptr;
    size_t fine_codes_cap = 0;
    unsigned long long* fine_u64 = nullptr;  // code_phase[n] then best[n]
    size_t fine_u64_cap = 0;
    float2* fine_mean = nullptr;
    float* fine_mag = nullptr;
    size_t fine_mag_cap = 0;
    float last_fine_ms = 0.f;

    // FFT facade
    FftRes fft[32];

    // tracking
    int8_t* ca_table_dev = nullptr;
    gb_trk_channel* ch_dev = nullptr;
    int n_ch = 0, ch_cap = 0;
    gb_trk_corr* corr_dev = nullptr;
    uint8_t *ran_dev = nullptr, *lost_dev = nullptr;
    float* hist_dev = nullptr;
    size_t hist_cap = 0;
    int hist_epochs = 0, hist_channels = 0;   // shape of the prompt history the last gb_trk_run left on the device
    float2* trk_data = nullptr;
    size_t trk_data_cap = 0;
    unsigned long long* offs_dev = nullptr;
    float trk_fs_max = 0.f;
    float last_trk_ms = 0.f;
    std::vector<uint8_t> trk_bad;            // channels idled by the last upload (code_row out of the table)
    std::vector<gb_trk_channel> trk_stage;   // host copy the upload is made from
};

namespace {

int fail(gb_handle* h, cudaError_t e, const char* what)
{
    if (h) {
        char buf[512];
        snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
        h->last_err = buf;
    }
    return (e == cudaErrorMemoryAllocation) ? GB_ENOMEM : GB_ECUDA;
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return fail(h, e__, #call);   \
    } while (0)

template <class T> int ensure(gb_handle* h, T** p, size_t* cap, size_t need)
{
    if (*cap >= need && *p) return GB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc((void**)p, need * sizeof(T));
    if (e != cudaSuccess) return fail(h, e, "cudaMalloc");
    *cap = need;
    return GB_OK;
}

// Set-up copies from pageable host memory.  cudaMemcpy returns once the bytes are staged, not necessarily once the DMA has
// landed, and it runs on the legacy default stream, which the handle's non-blocking streams do not wait for: finish it.
static cudaError_t h2d_blocking(void* dst, const void* src, size_t bytes)
{
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(cudaStreamLegacy);
}

// Rust `as usize` on f32 (saturating, NaN -> 0)
size_t f32_as_usize(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return (size_t)-1;
    return (size_t)v;
}

// G2 phase-selector taps for PRN 1..32 (IS-GPS-200)
const uint8_t kG2Taps[32][2] = {{2, 6}, {3, 7}, {4, 8}, {5, 9}, {1, 9}, {2, 10}, {1, 8}, {2, 9}, {3, 10}, {2, 3}, {3, 4},
                                {5, 6}, {6, 7}, {7, 8}, {8, 9}, {9, 10}, {1, 4}, {2, 5}, {3, 6}, {4, 7}, {5, 8}, {6, 9},
                                {1, 3}, {4, 6}, {5, 7}, {6, 8}, {7, 9}, {8, 10}, {1, 6}, {2, 7}, {3, 8}, {4, 9}};

void ca_chips(int prn, int8_t* out)
{
    // two 10-bit Fibonacci LFSRs kept as bit masks: bit (k-1) = stage k
    unsigned g1 = 0x3ff, g2 = 0x3ff;
    const int ta = kG2Taps[prn - 1][0] - 1, tb = kG2Taps[prn - 1][1] - 1;
    for (int c = 0; c < 1023; c++) {
        const unsigned bit = ((g1 >> 9) ^ (g2 >> ta) ^ (g2 >> tb)) & 1u;
        out[c] = bit ? 1 : -1;
        const unsigned f1 = ((g1 >> 2) ^ (g1 >> 9)) & 1u;
        const unsigned f2 = ((g2 >> 1) ^ (g2 >> 2) ^ (g2 >> 5) ^ (g2 >> 7) ^ (g2 >> 8) ^ (g2 >> 9)) & 1u;
        g1 = ((g1 << 1) | f1) & 0x3ff;
        g2 = ((g2 << 1) | f2) & 0x3ff;
    }
}

// position of natural frequency k after the forward DIF with these radices
int scrambled_pos(int k, int n, const int* radix, int ns)
{
    int pos = 0;
    for (int s = 0; s < ns; s++) {
        const int r = radix[s], m = n / r;
        pos += (k % r) * m;
        k /= r;
        n = m;
    }
    return pos;
}

int fft_resources(gb_handle* h, int plan, int n, FftRes** out)
{
    FftRes& r = h->fft[plan];
    if (!r.tw) {
        int radix[8];
        const int ns = gb::acq_plan_radices(plan, radix);
        // per-stage twiddles W_L^(i q) = exp(-2 pi i (i q) / L), layout [stage][q-1][i], f64-evaluated
        std::vector<float2> tw;
        tw.reserve(gb::acq_plan_twiddles(plan));
        for (int s = 0, L = n; s < ns; s++) {
            const int r = radix[s], sub = L / r;
            for (int q = 1; q < r; q++)
                for (int i = 0; i < sub; i++) {
                    const double ang = -2.0 * M_PI * (double)(((long long)i * q) % L) / (double)L;
                    tw.push_back(make_float2((float)cos(ang), (float)sin(ang)));
                }
            L = sub;
        }
        if ((int)tw.size() != gb::acq_plan_twiddles(plan)) return GB_EINVAL;
        std::vector<int> fop(n);
        if (gb::acq_plan_is_pfa(plan)) {
            // Good-Thomas maps over the digits d_s = (l / SUB_s) % R_s of a line position l (SUB_s = n / (R_0..R_s)):
            //   input   n(l) = sum_s d_s (n/R_s)                       mod n   (Ruritanian)
            //   output  k(l) = sum_s d_s (n/R_s) ((n/R_s)^-1 mod R_s)  mod n   (CRT: k = d_s mod R_s)
            std::vector<int> npos(n);
            long long cin[8], cout[8];
            int sub[8];
            for (int s = 0, L = n; s < ns; s++) {
                const int rs = radix[s], m = n / rs;
                sub[s] = L / rs;
                L = sub[s];
                int inv = 1;
                while (((long long)(m % rs) * inv) % rs != 1) inv++;
                cin[s] = m;
                cout[s] = ((long long)m * inv) % n;
            }
            for (int l = 0; l < n; l++) {
                long long ni = 0, ko = 0;
                for (int s = 0; s < ns; s++) {
                    const int d = (l / sub[s]) % radix[s];
                    ni += d * cin[s];
                    ko += d * cout[s];
                }
                npos[l] = (int)(ni % n);
                fop[l] = (int)(ko % n);
            }
            CK(cudaMalloc((void**)&r.npos, sizeof(int) * n));
            CK(h2d_blocking(r.npos, npos.data(), sizeof(int) * n));
        } else {
            for (int k = 0; k < n; k++) fop[scrambled_pos(k, n, radix, ns)] = k;
        }
        CK(cudaMalloc((void**)&r.tw, sizeof(float2) * tw.size()));
        CK(cudaMalloc((void**)&r.fop, sizeof(int) * n));
        CK(h2d_blocking(r.tw, tw.data(), sizeof(float2) * tw.size()));
        CK(h2d_blocking(r.fop, fop.data(), sizeof(int) * n));
        r.fop_host = fop;
    }
    *out = &r;
    return GB_OK;
}

// FP32 FMA throughput probe: 8 independent FMA chains per thread; the loop body is unrolled 32x (256 FFMA per
// counter update / branch) so loop overhead is < 1 % and the figure is the FFMA issue rate, not the loop's.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, float a, float b, int iters)
{
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 32; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678f) out[0] = s;
}

// DigitalFrontend::process_block (rf/frontend.rs:32-62) + write into the ring.  One CTA; per 2048-sample tile:
// thread 0 runs the sequential f32 NCO phase accumulator (:48-52), 16 lanes of warp 1 run the 8+8 independent DC-bias
// recurrences (rf/dc_remove.rs:23-29: lane j sees samples 8c+j), then all threads do the LUT mix (rf/nco_lut.rs:8-15,
// verbatim: i' = I*re + Q*im, q' = I*im - Q*re with im = -sin) and store.  All roundings as in the reference.
#define FE_TILE 2048
__global__ void __launch_bounds__(256) frontend_kernel(const float2* __restrict__ src, float2* __restrict__ ring,
                                                      unsigned long long head, unsigned long long mask, unsigned long long n,
                                                      const float* __restrict__ lut, float* __restrict__ state, float step,
                                                      float alpha, float con)
{
    __shared__ float2 tile[FE_TILE];
    __shared__ __align__(4) unsigned short idx[FE_TILE];
    __shared__ float s_acc;
    __shared__ float s_bias[16];
    if (threadIdx.x == 0) s_acc = state[0];
    if (threadIdx.x < 16) s_bias[threadIdx.x] = state[1 + threadIdx.x];
    __syncthreads();
    for (unsigned long long t0 = 0; t0 < n; t0 += FE_TILE) {
        const int tn = (int)((n - t0) < FE_TILE ? (n - t0) : FE_TILE);
        for (int i = threadIdx.x; i < tn; i += blockDim.x) tile[i] = src[t0 + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            // sequential f32 phase accumulator (frontend.rs:48-52).  Dependent chain per sample: FADD -> {FADD, FSETP}
            // -> FSEL; the index conversion and the store hang off it.  Unrolled by 8, indices packed 2 per store.
            float acc = s_acc;
            const bool fast = step >= 0.f && step < 2048.f && acc >= 0.f && acc < 2048.f;
            int i = 0;
            if (fast) {
                unsigned* idx32 = reinterpret_cast<unsigned*>(idx);
                for (; i + 8 <= tn; i += 8) {
                    unsigned k[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        k[u] = (unsigned)acc;                    // acc in [0, 2048): `as usize % LUT_SIZE`
                        const float s = __fadd_rn(acc, step);    // < 4096
                        const float s2 = __fsub_rn(s, 2048.f);   // exact when s >= 2048 (== fmodf)
                        acc = s >= 2048.f ? s2 : s;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u += 2) idx32[(i + u) >> 1] = k[u] | (k[u + 1] << 16);
                }
            }
            for (; i < tn; i++) {
                idx[i] = (unsigned short)((acc > 0.f ? (unsigned)acc : 0u) & 2047u);
                const float s = __fadd_rn(acc, step);
                acc = (s >= 0.f && s < 4096.f) ? (s >= 2048.f ? s - 2048.f : s) : fmodf(s, 2048.f);
            }
            s_acc = acc;
        } else if (threadIdx.x >= 32 && threadIdx.x < 48) {
            // 8 + 8 independent DC-bias recurrences (dc_remove.rs:23-29); 8 values are loaded ahead of the chain
            const int l = threadIdx.x - 32, lane = l & 7, comp = l >> 3;
            float b = s_bias[l];
            float* t = reinterpret_cast<float*>(tile) + comp;
            int cidx = lane;
            for (; cidx + 56 < tn; cidx += 64) {
                float x[8];
#pragma unroll
                for (int u = 0; u < 8; u++) x[u] = t[2 * (cidx + 8 * u)];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    b = __fadd_rn(__fmul_rn(b, con), __fmul_rn(x[u], alpha));
                    x[u] = __fsub_rn(x[u], b);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) t[2 * (cidx + 8 * u)] = x[u];
            }
            for (; cidx < tn; cidx += 8) {
                const float x = t[2 * cidx];
                b = __fadd_rn(__fmul_rn(b, con), __fmul_rn(x, alpha));
                t[2 * cidx] = __fsub_rn(x, b);
            }
            s_bias[l] = b;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += blockDim.x) {
            const float2 x = tile[i];
            const float lc = lut[idx[i]], ls = lut[2048 + idx[i]];
            float2 y;
            y.x = __fadd_rn(__fmul_rn(x.x, lc), __fmul_rn(x.y, ls));
            y.y = __fsub_rn(__fmul_rn(x.x, ls), __fmul_rn(x.y, lc));
            ring[(head + t0 + i) & mask] = y;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) state[0] = s_acc;
    if (threadIdx.x < 16) state[1 + threadIdx.x] = s_bias[threadIdx.x];
}

// Bit synchronisation + 20 ms prompt accumulation on the batched prompt history (legacy
// decoding.rs:115-127, 164-213): one thread per channel walks its prompt-I column in epoch order.
__global__ void nav_bit_sync_kernel(const float* __restrict__ hist, int n_epochs, int n_channels, gb_nav_sync* __restrict__ st,
                                    int8_t* __restrict__ bits, int max_bits)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_channels) return;
    gb_nav_sync s;
    s.flag_bit_sync = 0; s.frame_sync_ind = 0; s.sync_epoch = -1; s.n_bits = 0;
    s.preamble_bit = -1; s.polarity = 0; s.ref_frame_sync = 0; s.ref_polarity = 0;
    unsigned buff[20];
#pragma unroll
    for (int i = 0; i < 20; i++) buff[i] = 0u;
    float old_ip = 0.f, acc = 0.f;
    int biti = 0;
    for (int cnt = 0; cnt < n_epochs; cnt++) {
        const float ip = hist[((size_t)cnt * n_channels + c) * 2];
        if (!s.flag_bit_sync && cnt > 1000 && __fmul_rn(old_ip, ip) < 0.f) {
            unsigned v_max = 0u;
            int i_max = 0;
#pragma unroll
            for (int i = 0; i < 20; i++) {
                if (i == biti) buff[i] += 1u;
                if (buff[i] >= v_max) { v_max = buff[i]; i_max = i; }   // Iterator::max_by keeps the last maximum
            }
            s.frame_sync_ind = i_max;
            if (v_max == 30u) { s.flag_bit_sync = 1; s.sync_epoch = cnt; }
        }
        if (s.flag_bit_sync) {
            acc = (biti == s.frame_sync_ind) ? ip : __fadd_rn(acc, ip);
            const int last = s.frame_sync_ind + 19 >= 20 ? s.frame_sync_ind - 1 : s.frame_sync_ind + 19;
            if (biti == last) {
                if (s.n_bits < max_bits) bits[(size_t)c * max_bits + s.n_bits] = acc > 0.f ? 1 : -1;
                s.n_bits++;
            }
        }
        old_ip = ip;
        biti = biti == 19 ? 0 : biti + 1;
    }
#pragma unroll
    for (int i = 0; i < 20; i++) s.bit_sync_buff[i] = buff[i];
    // Preamble search on the bit stream (check_preamble_syn, decoding.rs:215-226): correlation of 8 consecutive bits with
    // GPS_CA_PREAMBLE, frame sync when it is +-8 (polarity = its sign).  The legacy pushes every bit into buff_preamble
    // and tests only while its length is exactly 8 (:131-136) -- the VecDeque is never popped, so ONLY the first 8 bits
    // are ever tested: that literal outcome is ref_frame_sync / ref_polarity.  preamble_bit / polarity are the intended
    // sliding search: the first bit index at which the last 8 bits match.
    {
        const int pre[8] = {1, -1, -1, -1, 1, -1, 1, 1};
        const int nb = s.n_bits < max_bits ? s.n_bits : max_bits;
        const int8_t* b = bits + (size_t)c * max_bits;
        for (int i0 = 0; i0 + 8 <= nb; i0++) {
            int corr = 0;
#pragma unroll
            for (int x = 0; x < 8; x++) corr += (int)b[i0 + x] * pre[x];
            const bool hit = corr == 8 || corr == -8;
            if (i0 == 0 && hit) { s.ref_frame_sync = 1; s.ref_polarity = corr > 0 ? 1 : -1; }
            if (hit && s.preamble_bit < 0) { s.preamble_bit = i0; s.polarity = corr > 0 ? 1 : -1; }
            if (s.preamble_bit >= 0) break;
        }
    }
    st[c] = s;
}

__global__ void i8_to_ring_kernel(const int8_t* __restrict__ src, float2* __restrict__ ring, unsigned long long start,
                                  unsigned long long mask, unsigned long long n)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ring[(start + i) & mask] = make_float2((float)src[i], 0.f);
}

}  // namespace

// ------------------------------------------------------------------ misc
extern "C" const char* gb_strerror(int code)
{
    switch (code) {
        case GB_OK: return "ok";
        case GB_EINVAL: return "invalid argument";
        case GB_ENODEVICE: return "no CUDA device";
        case GB_ECUDA: return "CUDA error";
        case GB_EUNSUPPORTED: return "fft_size has no sm_100a plan";
        case GB_ESTATE: return "call out of order";
        case GB_ENOMEM: return "out of memory";
        case GB_ENCCL: return "NCCL is not available or a collective failed";
    case GB_ERANGE: return "samples not in the ring";
    }
    return "unknown error";
}
extern "C" const char* gb_last_cuda_error(gb_handle* h) { return h ? h->last_err.c_str() : ""; }
extern "C" int gb_version(void) { return GB_VERSION; }
extern "C" int gb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int gb_create(const gb_config* cfg, gb_handle** out)
{
    if (!out) return GB_EINVAL;
    *out = nullptr;
    const int dev = cfg ? cfg->device : 0;
    if (dev < 0 || dev >= gb_device_count()) return GB_ENODEVICE;
    gb_handle* h = new gb_handle();
    h->device = dev;
    *out = h;
    CK(cudaSetDevice(dev));
    CK(cudaStreamCreateWithFlags(&h->s_acq, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_trk, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    CK(cudaEventCreate(&h->ev_a0));
    CK(cudaEventCreate(&h->ev_a1));
    CK(cudaEventCreate(&h->ev_t0));
    CK(cudaEventCreate(&h->ev_t1));
    {
        std::vector<int8_t> tab(32 * 1023);
        for (int p = 1; p <= 32; p++) ca_chips(p, tab.data() + (p - 1) * 1023);
        CK(cudaMalloc((void**)&h->ca_table_dev, tab.size()));
        CK(h2d_blocking(h->ca_table_dev, tab.data(), tab.size()));
    }
    for (int s = 0; s < 2; s++) {
        CK(cudaMallocHost((void**)&h->pin_stage[s], kStageSamples * sizeof(float2)));
        CK(cudaEventCreateWithFlags(&h->pin_free[s], cudaEventDisableTiming));
    }
    for (int s = 0; s < 2; s++) {
        CK(cudaMallocHost((void**)&h->rows_pin_slot[s], sizeof(int) * 256));
        CK(cudaEventCreateWithFlags(&h->ev_done[s], cudaEventDisableTiming));
        CK(cudaEventCreate(&h->ev_s0[s]));
        CK(cudaEventCreate(&h->ev_s1[s]));
    }
    h->rows_pin = h->rows_pin_slot[0];
    CK(cudaEventCreateWithFlags(&h->ev_chunk_free, cudaEventDisableTiming));
    CK(cudaMalloc((void**)&h->rows_dev, sizeof(int) * 2 * 256));
    if (cfg && cfg->ring_capacity) return gb_ring_create(h, cfg->ring_capacity);
    return GB_OK;
}

extern "C" int gb_destroy(gb_handle* h)
{
    if (!h) return GB_EINVAL;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    void* dev_ptrs[] = {h->fwd_bins_dev, h->inv_map_dev, h->code_fft_shift, h->fe_idx, h->fe_lut, h->fe_state, h->fe_stage, h->fe_scratch, h->otw, h->acc_rows, h->spec, h->spec_tc, h->code_tc, h->ring, h->i8_stage, /* h->tw aliases fft[plan].tw, freed below */ h->code_fft, h->tables, h->rot, h->chunk, h->codes_dev, h->cells_dev,
                        h->rows_dev, h->row_dev, h->ca_table_dev, h->ch_dev, h->corr_dev, h->ran_dev, h->lost_dev,
                        h->hist_dev, h->trk_data, h->offs_dev, h->fine_x, h->fine_y, h->fine_codes, h->fine_u64, h->fine_mean,
                        h->fine_mag, h->tables_perm, h->iq_perm};
    for (void* p : dev_ptrs)
        if (p) cudaFree(p);
    for (auto& kv : h->gen_plans) gb::generic_plan_destroy(kv.second);
    if (h->gen_s0) cudaFree(h->gen_s0);
    if (h->gen_s1) cudaFree(h->gen_s1);
    for (auto& r : h->fft) {
        if (r.tw) cudaFree(r.tw);
        if (r.fop) cudaFree(r.fop);
        if (r.npos) cudaFree(r.npos);
    }
    for (int s = 0; s < 2; s++) {
        if (h->pin_stage[s]) cudaFreeHost(h->pin_stage[s]);
        if (h->pin_free[s]) cudaEventDestroy(h->pin_free[s]);
    }
    for (int s = 0; s < 2; s++) {
        if (h->cells_pin_slot[s]) cudaFreeHost(h->cells_pin_slot[s]);
        if (h->rows_pin_slot[s]) cudaFreeHost(h->rows_pin_slot[s]);
        cudaEvent_t es[] = {h->ev_done[s], h->ev_s0[s], h->ev_s1[s]};
        for (cudaEvent_t e : es)
            if (e) cudaEventDestroy(e);
    }
    if (h->ev_chunk_free) cudaEventDestroy(h->ev_chunk_free);
    if (h->s_acq) cudaStreamDestroy(h->s_acq);
    if (h->s_trk) cudaStreamDestroy(h->s_trk);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    for (cudaEvent_t e : h->ev_slice)
        if (e) cudaEventDestroy(e);
    cudaEvent_t evs[] = {h->ev_copy, h->ev_a0, h->ev_a1, h->ev_t0, h->ev_t1};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    cudaGetLastError();   // a failed free must not surface as the "last error" of another handle's next launch
    delete h;
    return GB_OK;
}

extern "C" int gb_synchronize(gb_handle* h)
{
    if (!h) return GB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_copy));
    CK(cudaStreamSynchronize(h->s_acq));
    CK(cudaStreamSynchronize(h->s_trk));
    return GB_OK;
}

// ------------------------------------------------------------------ C/A code (host)
extern "C" int gb_ca_code_chips(int prn, int8_t* out1023)
{
    if (prn < 1 || prn > 32 || !out1023) return GB_EINVAL;
    ca_chips(prn, out1023);
    return GB_OK;
}
// ca_code.rs:13-17
extern "C" int gb_num_samples_per_code(float code_rate, float fs) { return (int)f32_as_usize(roundf(fs / (code_rate / 1023.0f))); }
// ca_code.rs:12-27; the index arithmetic stays in f32 in the reference's order (Q3)
extern "C" int gb_generate_ca_code_samples(int prn, float code_rate, float fs, int8_t* out, int cap)
{
    if (prn < 1 || prn > 32 || !out) return GB_EINVAL;
    int8_t chips[1023];
    ca_chips(prn, chips);
    const int n = gb_num_samples_per_code(code_rate, fs);
    for (int x = 0; x < n && x < cap; x++) {
        const size_t ind = f32_as_usize(floorf((float)x * code_rate / fs));
        if (ind >= 1023) return GB_EINVAL;  // the reference panics (index out of bounds)
        out[x] = chips[ind];
    }
    return n;
}

// ------------------------------------------------------------------ sample ring
extern "C" int gb_ring_create(gb_handle* h, uint64_t cap)
{
    if (!h || cap == 0 || (cap & (cap - 1))) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_ring);  // MulticastRingBuffer::new asserts power of two
    CK(cudaSetDevice(h->device));
    if (h->ring) cudaFree(h->ring);
    h->ring = nullptr;
    CK(cudaMalloc((void**)&h->ring, cap * sizeof(float2)));
    // on the copy stream: cudaMemset runs asynchronously on the legacy default stream, which the (non-blocking) copy
    // stream does not wait for -- a late memset would wipe samples written right after the ring was created
    CK(cudaMemsetAsync(h->ring, 0, cap * sizeof(float2), h->s_copy));
    CK(cudaEventRecord(h->ev_copy, h->s_copy));
    h->ring_cap = cap;
    h->ring_head = 0;
    return GB_OK;
}
extern "C" int gb_ring_reset(gb_handle* h)
{
    if (!h || !h->ring) return GB_ESTATE;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_copy));
    h->ring_head = 0;
    return GB_OK;
}
// write_samples (multicast_ring_buffer.rs:66-101): wrap-aware copy at head & mask, then head += n.
// The caller's buffer is copied into a pinned staging slot (double-buffered) and DMA'd with cudaMemcpyAsync on
// the copy stream, so the call returns as soon as the bytes are staged and the caller may reuse its buffer.
static int ring_put(gb_handle* h, const float2* src, uint64_t n)
{
    uint64_t done = 0;
    while (done < n) {
        const uint64_t chunk = (n - done) < kStageSamples ? (n - done) : kStageSamples;
        const int s = h->pin_next;
        h->pin_next ^= 1;
        CK(cudaEventSynchronize(h->pin_free[s]));  // the DMA that last used this slot has finished
        memcpy(h->pin_stage[s], src + done, chunk * sizeof(float2));
        const uint64_t start = (h->ring_head + done) & (h->ring_cap - 1);
        const uint64_t first = (start + chunk <= h->ring_cap) ? chunk : h->ring_cap - start;
        CK(cudaMemcpyAsync(h->ring + start, h->pin_stage[s], first * sizeof(float2), cudaMemcpyHostToDevice, h->s_copy));
        if (first < chunk)
            CK(cudaMemcpyAsync(h->ring, h->pin_stage[s] + first, (chunk - first) * sizeof(float2), cudaMemcpyHostToDevice,
                               h->s_copy));
        CK(cudaEventRecord(h->pin_free[s], h->s_copy));
        done += chunk;
    }
    CK(cudaEventRecord(h->ev_copy, h->s_copy));
    h->ring_head += n;
    return GB_OK;
}

extern "C" int gb_ring_write(gb_handle* h, const gb_c32* samples, uint64_t n)
{
    if (!h || !h->ring) return GB_ESTATE;
    if (!samples || n > h->ring_cap) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_ring);
    CK(cudaSetDevice(h->device));
    return ring_put(h, reinterpret_cast<const float2*>(samples), n);
}
extern "C" int gb_ring_write_i8(gb_handle* h, const int8_t* samples, uint64_t n)
{
    if (!h || !h->ring) return GB_ESTATE;
    if (!samples || n > h->ring_cap) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_ring);
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_copy));  // the staging buffer is reused
    int rc = ensure(h, &h->i8_stage, &h->i8_cap, (size_t)n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->i8_stage, samples, n, cudaMemcpyHostToDevice, h->s_copy));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    i8_to_ring_kernel<<<blocks, 256, 0, h->s_copy>>>(h->i8_stage, h->ring, h->ring_head, h->ring_cap - 1, n);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev_copy, h->s_copy));
    h->ring_head += n;
    return GB_OK;
}
// DigitalFrontend::new (rf/frontend.rs:18-30): NCO LUT (host libm, rf/nco_lut.rs:24-42), alpha = 0.001, zero state
extern "C" int gb_frontend_configure(gb_handle* h, float f_if, float fs_in)
{
    if (!h || !(fs_in > 0.f)) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_ring);
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_copy));
    std::vector<float> lut(4096);
    for (int i = 0; i < 2048; i++) {
        const float angle = (2.0f * kPiF * (float)i) / 2048.0f;
        lut[i] = cosf(angle);
        lut[2048 + i] = -sinf(angle);
    }
    if (!h->fe_lut) CK(cudaMalloc((void**)&h->fe_lut, sizeof(float) * 4096));
    if (!h->fe_state) CK(cudaMalloc((void**)&h->fe_state, sizeof(float) * 17));
    CK(h2d_blocking(h->fe_lut, lut.data(), sizeof(float) * 4096));
    CK(cudaMemsetAsync(h->fe_state, 0, sizeof(float) * 17, h->s_copy));   // the stream the front-end kernels run on
    h->fe_step = (f_if / fs_in) * 2048.0f;  // nco_lut.rs:34
    // The phase accumulator does not depend on the samples: its orbit from 0 (tail + cycle, at most 2^24 states) is
    // computed here in the reference's f32 arithmetic and the kernel looks the LUT index up by sample number.
    // gb_tuning_set("fe_sequential", 1) (or an orbit longer than the cap) keeps the one-thread sequential accumulator.
    h->fe_table = false;
    h->fe_count = 0;
    if (h->fe_idx) { cudaFree(h->fe_idx); h->fe_idx = nullptr; }
    if (!gb::tuning("fe_sequential", 0) &&
        gb::fe_build_phase_orbit(h->fe_step, (uint64_t)1 << 24, 4096, h->fe_phase, &h->fe_mu, &h->fe_period)) {
        std::vector<uint16_t> idx(h->fe_phase.size());
        for (size_t i = 0; i < idx.size(); i++) idx[i] = gb::fe_lut_index(h->fe_phase[i]);
        CK(cudaMalloc((void**)&h->fe_idx, idx.size() * sizeof(uint16_t)));
        CK(h2d_blocking(h->fe_idx, idx.data(), idx.size() * sizeof(uint16_t)));
        h->fe_table = true;
    } else {
        h->fe_phase.clear();
    }
    h->fe_ready = true;
    return GB_OK;
}

// rf_thread's per-block work (rf/rf_thread.rs:44-48): front-end on n raw complex samples, result appended to the ring
extern "C" int gb_frontend_write(gb_handle* h, const gb_c32* raw, uint64_t n)
{
    if (!h || !h->ring || !h->fe_ready) return GB_ESTATE;
    if (!raw || n == 0 || n > h->ring_cap || (n % 8) != 0) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_ring);  // chunks_exact_mut(16 floats)
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_copy));  // the staging buffer is reused
    int rc = ensure(h, &h->fe_stage, &h->fe_cap, (size_t)n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->fe_stage, raw, n * sizeof(float2), cudaMemcpyHostToDevice, h->s_copy));
    if (h->fe_table && h->fe_mode == GB_FE_PARALLEL) {
        if ((rc = ensure(h, &h->fe_scratch, &h->fe_scratch_cap, gb::fe_parallel_scratch(n)))) return rc;
        CK(gb::fe_launch_parallel(h->fe_stage, h->ring, h->ring_head, h->ring_cap - 1, n, h->fe_lut, h->fe_state + 1, h->fe_idx,
                                  gb::fe_orbit_pos(h->fe_count, h->fe_mu, h->fe_period), h->fe_mu, h->fe_period, 0.001f,
                                  1.0f - 0.001f, h->fe_scratch, h->s_copy));
    } else if (h->fe_table) {
        CK(gb::fe_launch_table(h->fe_stage, h->ring, h->ring_head, h->ring_cap - 1, n, h->fe_lut, h->fe_state + 1, h->fe_idx,
                               gb::fe_orbit_pos(h->fe_count, h->fe_mu, h->fe_period), h->fe_mu, h->fe_period, 0.001f,
                               1.0f - 0.001f, h->s_copy));
    } else {
        frontend_kernel<<<1, 256, 0, h->s_copy>>>(h->fe_stage, h->ring, h->ring_head, h->ring_cap - 1, n, h->fe_lut, h->fe_state,
                                                   h->fe_step, 0.001f, 1.0f - 0.001f);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(h->ev_copy, h->s_copy));
    h->fe_count += n;
    h->ring_head += n;
    return GB_OK;
}

extern "C" int gb_frontend_set_mode(gb_handle* h, int mode)
{
    if (!h) return GB_ESTATE;
    if (mode != GB_FE_EXACT && mode != GB_FE_PARALLEL) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_ring);
    // the segmented scan looks the NCO up by sample number: without an orbit table there is nothing to parallelise
    if (mode == GB_FE_PARALLEL && h->fe_ready && !h->fe_table) return GB_EUNSUPPORTED;
    h->fe_mode = mode;
    return GB_OK;
}

extern "C" int gb_frontend_orbit(float f_if, float fs_in, uint64_t* mu, uint64_t* lambda, uint16_t* idx_out, uint64_t n_idx)
{
    if (!(fs_in > 0.f) || !mu || !lambda) return GB_EINVAL;
    std::vector<float> phase;
    uint64_t period = 0;
    if (!gb::fe_build_phase_orbit((f_if / fs_in) * 2048.0f, (uint64_t)1 << 24, 1, phase, mu, &period)) return GB_EUNSUPPORTED;
    *lambda = period;   // min_period 1: no replication
    if (idx_out)
        for (uint64_t i = 0; i < n_idx; i++) idx_out[i] = gb::fe_lut_index(phase[gb::fe_orbit_pos(i, *mu, period)]);
    return GB_OK;
}

// phase_accumulator, bias_re[8], bias_im[8] (diagnostics / tests)
extern "C" int gb_frontend_state(gb_handle* h, float* state17)
{
    if (!h || !state17) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_ring);
    if (!h->fe_ready) return GB_ESTATE;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_copy));
    CK(cudaMemcpy(state17, h->fe_state, sizeof(float) * 17, cudaMemcpyDeviceToHost));
    if (h->fe_table) state17[0] = h->fe_phase[gb::fe_orbit_pos(h->fe_count, h->fe_mu, h->fe_period)];
    return GB_OK;
}

extern "C" uint64_t gb_ring_head(gb_handle* h) { return h ? h->ring_head : 0; }
// copy_to_slice (multicast_ring_buffer.rs:107-129)
extern "C" int gb_ring_copy_to_slice(gb_handle* h, uint64_t start, gb_c32* dest, uint64_t n)
{
    if (!h || !h->ring) return GB_ESTATE;
    if (!dest || n > h->ring_cap) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_ring);
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_copy));
    const uint64_t ps = start & (h->ring_cap - 1);
    const uint64_t first = (ps + n <= h->ring_cap) ? n : h->ring_cap - ps;
    CK(cudaMemcpy(dest, h->ring + ps, first * sizeof(float2), cudaMemcpyDeviceToHost));
    if (first < n) CK(cudaMemcpy(dest + first, h->ring, (n - first) * sizeof(float2), cudaMemcpyDeviceToHost));
    return GB_OK;
}

// ------------------------------------------------------------------ acquisition set-up
const int kPlanGeneric = 1000;   // h->plan of a length without a tuned plan
const int kGenericMaxN = 131072;

static int generic_plan_for(gb_handle* h, int n, gb::GenericPlan** out)
{
    auto it = h->gen_plans.find(n);
    if (it != h->gen_plans.end()) { *out = it->second; return GB_OK; }
    gb::GenericPlan* p = nullptr;
    CK(gb::generic_plan_create(n, &p, h->s_acq));
    h->gen_plans[n] = p;
    *out = p;
    return GB_OK;
}

extern "C" int gb_acq_supported_sizes(int* sizes, int cap)
{
    int n = gb::acq_plan_sizes(sizes, cap);
    if (n < cap && sizes) sizes[n] = 80000;  // thread-block-cluster plan (acq_cluster.cu)
    return n + 1;
}

extern "C" int gb_acq_configure(gb_handle* h, int fft_size, float fs, int n_prn, const int8_t* codes)
{
    if (!h || n_prn < 1 || n_prn > 255 || !(fs > 0.f)) return GB_EINVAL;
    if (!codes && n_prn > 32) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    if (h->pend[0].active || h->pend[1].active) return GB_ESTATE;
    // fft_size % 4 != 0: apply_doppler_shift writes only 4 * floor(len / 4) samples (A3); the tail of result_buf keeps the
    // PREVIOUS block's unnormalised IFFT output, which is fed back as input N times larger every block -- the reference's
    // own arithmetic overflows to inf / NaN within a dozen blocks (tests/test_oracle_golden.py shows it on the oracle).
    // There is no behaviour to be a drop-in for, so these lengths are refused.
    if (fft_size < 4 || fft_size % 4 != 0) return GB_EUNSUPPORTED;
    const bool cluster = gb::acq_cluster_supported(fft_size) != 0;
    if (cluster && !codes) return GB_EINVAL;  // no built-in code has an 80000-sample period
    const int inner = cluster ? gb::acq_cluster_inner(fft_size) : fft_size;
    int plan = gb::acq_plan_index(inner);
    // no tuned shared-memory plan: the any-length plan (Bluestein over a power-of-two Stockham FFT, acq_generic.cu)
    const bool generic = plan < 0;
    if (generic && (cluster || fft_size > kGenericMaxN)) return GB_EUNSUPPORTED;
    if (generic) plan = kPlanGeneric;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_acq));
    std::vector<int8_t> host_codes;
    if (!codes) {
        // AcquisitionWorker::new (:133-135): generate_ca_code_samples(prn, 1.023e6, fs)
        const int n_code = gb_num_samples_per_code(kCodeRate, fs);
        if (n_code != fft_size) return GB_EINVAL;  // rustfft would panic on the length mismatch
        host_codes.resize((size_t)n_prn * fft_size);
        for (int p = 1; p <= n_prn; p++) {
            const int rc = gb_generate_ca_code_samples(p, kCodeRate, fs, host_codes.data() + (size_t)(p - 1) * fft_size, fft_size);
            if (rc < 0) return rc;
        }
        codes = host_codes.data();
    }
    FftRes* fr = nullptr;
    int rc = GB_OK;
    gb::GenericPlan* gen = nullptr;
    if (generic) rc = generic_plan_for(h, fft_size, &gen);
    else rc = fft_resources(h, plan, inner, &fr);
    if (rc) return rc;
    if (cluster) {
        // outer twiddles W_N^(i q), q = 1..RO-1, i < inner, layout [q-1][i], f64-evaluated
        const int ro = gb::acq_cluster_outer(fft_size);
        std::vector<float2> otw((size_t)(ro - 1) * inner);
        for (int q = 1; q < ro; q++)
            for (int i = 0; i < inner; i++) {
                const double ang = -2.0 * M_PI * (double)(((long long)i * q) % fft_size) / (double)fft_size;
                otw[(size_t)(q - 1) * inner + i] = make_float2((float)cos(ang), (float)sin(ang));
            }
        if (h->otw) cudaFree(h->otw);
        h->otw = nullptr;
        CK(cudaMalloc((void**)&h->otw, otw.size() * sizeof(float2)));
        CK(h2d_blocking(h->otw, otw.data(), otw.size() * sizeof(float2)));
    }
    if (h->code_fft) cudaFree(h->code_fft);
    if (h->codes_dev) cudaFree(h->codes_dev);
    if (h->row_dev) cudaFree(h->row_dev);
    h->code_fft = nullptr; h->codes_dev = nullptr; h->row_dev = nullptr;
    const size_t total = (size_t)n_prn * fft_size;
    const int spec_len = (cluster || generic) ? fft_size : gb::acq_plan_spec_len(plan);
    CK(cudaMalloc((void**)&h->code_fft, (size_t)n_prn * spec_len * sizeof(float2)));
    h->code_gen++;
    CK(cudaMemsetAsync(h->code_fft, 0, (size_t)n_prn * spec_len * sizeof(float2), h->s_acq));   // row padding
    CK(cudaMalloc((void**)&h->codes_dev, total));
    CK(cudaMalloc((void**)&h->row_dev, sizeof(float) * fft_size));
    CK(cudaMemcpyAsync(h->codes_dev, codes, total, cudaMemcpyHostToDevice, h->s_acq));
    if (generic) {
        const size_t need = (size_t)n_prn * gb::generic_plan_m(gen);
        rc = ensure(h, &h->gen_s0, &h->gen_s0_cap, need);
        if (!rc) rc = ensure(h, &h->gen_s1, &h->gen_s1_cap, need);
        if (rc) return rc;
        CK(gb::generic_code_fft(gen, h->codes_dev, n_prn, h->code_fft, h->gen_s0, h->gen_s1, h->s_acq));
    } else if (cluster) {
        CK(gb::acq_cluster_launch_code_fft(h->codes_dev, n_prn, h->code_fft, fr->tw, h->otw, h->s_acq));
    } else {
        CK(gb::acq_launch_code_fft(plan, h->codes_dev, n_prn, h->code_fft, fr->tw, fr->npos, h->s_acq));
    }
    CK(cudaStreamSynchronize(h->s_acq));
    h->cluster = cluster;
    h->generic = generic; h->gen = gen;
    h->plan = plan; h->N = fft_size; h->n_prn = n_prn; h->fs = fs; h->tw = generic ? nullptr : fr->tw;
    h->spec_len = spec_len;
    h->pfa = !cluster && !generic && gb::acq_plan_is_pfa(plan) != 0;
    h->npos = h->pfa ? fr->npos : nullptr;
    h->D = 0; h->n_coh = 1;
    h->carr.clear();
    h->alias_ok = false;
    h->alias_enabled = false;   // like n_coh, a per-configuration setting; the default is the reference's per-bin arithmetic
    h->builtin_codes = host_codes.size() != 0;
    h->chunk_free_valid = false;
    return GB_OK;
}

// AcquisitionWorker::new runs once per PRN in the reference (32 workers built in a rayon loop, do_acquisition.rs:268-271):
// the per-worker constructor of a drop-in calls THIS -- the first call plans, an identical later call (built-in GPS C/A
// codes, same fft_size / fs / n_prn) returns at once and keeps the Doppler tables and settings.
extern "C" int gb_acq_configure_once(gb_handle* h, int fft_size, float fs, int n_prn)
{
    if (!h) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    if (h->builtin_codes && h->plan >= 0 && h->N == fft_size && h->fs == fs && h->n_prn == n_prn) return GB_OK;
    return gb_acq_configure(h, fft_size, fs, n_prn, nullptr);
}

// prime-factor plans: keep a copy of the wipe-off tables in line order (tables_perm[d][l] = tables[d][n(l)])
static int permute_tables(gb_handle* h)
{
    if (!h->pfa || h->D == 0) return GB_OK;
    int rc = ensure(h, &h->tables_perm, &h->tables_perm_cap, (size_t)h->D * h->N);
    if (rc) return rc;
    CK(gb::acq_launch_permute(h->tables, 0, ~0ull, h->npos, h->N, 0, h->D, h->tables_perm, h->s_acq));
    CK(cudaStreamSynchronize(h->s_acq));
    return GB_OK;
}

static int upload_rotators(gb_handle* h)
{
    int prc = permute_tables(h);
    if (prc) return prc;
    if (h->n_coh <= 1 || h->D == 0) return GB_OK;
    // rot[d][c] = exp(-j 2 pi carr_d c N / fs), f64-evaluated
    std::vector<float2> rot((size_t)h->D * h->n_coh);
    for (int d = 0; d < h->D; d++)
        for (int c = 0; c < h->n_coh; c++) {
            const double cyc = (double)h->carr[d] * (double)c * (double)h->N / (double)h->fs;
            const double ang = -2.0 * M_PI * (cyc - floor(cyc));
            rot[(size_t)d * h->n_coh + c] = make_float2((float)cos(ang), (float)sin(ang));
        }
    int rc = ensure(h, &h->rot, &h->rot_cap, rot.size());
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->rot, rot.data(), rot.size() * sizeof(float2), cudaMemcpyHostToDevice, h->s_acq));
    CK(cudaStreamSynchronize(h->s_acq));
    return GB_OK;
}

// Doppler aliasing.  With carr_d = carr_b + m * fs/N (m integer, exactly) the wiped block of bin d is the wiped block of
// bin b times exp(-j 2 pi m n / N), so its spectrum is a circular shift: X_d[k] = X_b[k + m].  The forward path (wipe-
// off, coherent pre-sum, forward FFT) is then needed for the base bin only, and
//   IFFT_k{ X_d[k] conj(C[k]) }[n] = exp(-j 2 pi m n / N) * IFFT_j{ X_b[j] conj(C[j - m]) }[n]:
// the inverse kernel pairs the base spectrum with the code spectrum shifted by m (one re-indexed copy of the code
// spectra per distinct m, built here) and the unit-modulus ramp vanishes in |.|^2.  The coherent rotators
// exp(-j 2 pi carr c N / fs) of d and b differ by whole turns.  Exact up to the f32 rounding of the reference's table
// phases (cos(i * step_d) vs cos(i * step_b) shifted): ~1e-6 relative on the correlation power.
// Only tables built by gb_acq_make_doppler_tables are analysed; caller-supplied tables are used as given.
static int build_alias_map(gb_handle* h)
{
    h->alias_ok = false;
    h->n_base = h->n_shift = 0;
    if (h->cluster || h->D < 2 || h->plan < 0 || !gb::acq_plan_supports_alias(h->plan)) return GB_OK;
    const int D = h->D, N = h->N;
    const double fs = (double)h->fs;
    std::vector<int> bases, shift_of(D);
    std::vector<int2> inv(D);
    for (int d = 0; d < D; d++) {
        bool found = false;
        for (size_t bi = 0; bi < bases.size() && !found; bi++) {
            const double mN = ((double)h->carr[d] - (double)h->carr[bases[bi]]) * (double)N;   // exact in f64
            const double m = nearbyint(mN / fs);
            if (fabs(m) < (double)N && m * fs == mN) {
                inv[d].x = (int)bi;
                shift_of[d] = (int)m;
                found = true;
            }
        }
        if (!found) {
            inv[d].x = (int)bases.size();
            shift_of[d] = 0;
            bases.push_back(d);
        }
    }
    if ((int)bases.size() == D) return GB_OK;   // no two bins a whole number of FFT bins apart
    std::vector<int> shifts(shift_of);
    std::sort(shifts.begin(), shifts.end());
    shifts.erase(std::unique(shifts.begin(), shifts.end()), shifts.end());
    const int SL = h->spec_len;
    const size_t set = (size_t)h->n_prn * SL;
    if (shifts.size() * set * sizeof(float2) > ((size_t)256 << 20)) return GB_OK;
    for (int d = 0; d < D; d++)
        inv[d].y = (int)(std::lower_bound(shifts.begin(), shifts.end(), shift_of[d]) - shifts.begin());
    // code spectra are stored scrambled + transposed with padded rows: element t = q * STRIDE + b (b < NB) is line position
    // l = b * R + q (R = last radix), which holds natural frequency fop[l].  Shifted set s: dst frequency j takes src
    // frequency j - m; padding elements map to themselves.
    const std::vector<int>& fop = h->fft[h->plan].fop_host;
    if ((int)fop.size() != N) return GB_OK;
    int radix[8];
    const int ns = gb::acq_plan_radices(h->plan, radix);
    const int R = radix[ns - 1], NB = N / R, STRIDE = gb::acq_plan_spec_stride(h->plan);
    if (R * STRIDE != SL) return GB_OK;
    std::vector<int> pof(N);
    for (int l = 0; l < N; l++) pof[fop[l]] = l;
    std::vector<int> gidx(shifts.size() * (size_t)SL);
    for (size_t si = 0; si < shifts.size(); si++)
        for (int t = 0; t < SL; t++) {
            const int q = t / STRIDE, b = t % STRIDE;
            if (b >= NB) {
                gidx[si * SL + t] = t;
                continue;
            }
            const int l = b * R + q;
            const int js = (int)((((long long)fop[l] - shifts[si]) % N + N) % N);
            const int l2 = pof[js];
            gidx[si * SL + t] = (l2 % R) * STRIDE + l2 / R;
        }
    int rc;
    int* gidx_dev = nullptr;
    if ((rc = ensure(h, &h->code_fft_shift, &h->code_fft_shift_cap, shifts.size() * set))) return rc;
    h->code_gen++;
    if ((rc = ensure(h, &h->fwd_bins_dev, &h->fwd_bins_cap, bases.size()))) return rc;
    if ((rc = ensure(h, &h->inv_map_dev, &h->inv_map_cap, (size_t)D))) return rc;
    CK(cudaMalloc((void**)&gidx_dev, gidx.size() * sizeof(int)));
    cudaError_t e = cudaMemcpyAsync(gidx_dev, gidx.data(), gidx.size() * sizeof(int), cudaMemcpyHostToDevice, h->s_acq);
    if (e == cudaSuccess) e = gb::acq_launch_shift_codes(h->code_fft, gidx_dev, (int)shifts.size(), h->n_prn, SL, h->code_fft_shift, h->s_acq);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->fwd_bins_dev, bases.data(), bases.size() * sizeof(int), cudaMemcpyHostToDevice, h->s_acq);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->inv_map_dev, inv.data(), (size_t)D * sizeof(int2), cudaMemcpyHostToDevice, h->s_acq);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_acq);
    cudaFree(gidx_dev);
    if (e != cudaSuccess) return fail(h, e, "build_alias_map");
    h->n_base = (int)bases.size();
    h->n_shift = (int)shifts.size();
    h->alias_ok = true;
    return GB_OK;
}

extern "C" int gb_acq_set_doppler_aliasing(gb_handle* h, int on)
{
    if (!h) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    h->alias_enabled = on != 0;
    return GB_OK;
}
// number of forward spectra the next shared-chain search computes (== n_doppler when nothing is shared)
extern "C" int gb_acq_forward_bins(gb_handle* h)
{
    if (!h) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    return (h->alias_ok && h->alias_enabled) ? h->n_base : h->D;
}

extern "C" int gb_acq_make_doppler_tables(gb_handle* h, float f_if, const float* dopplers, int D, float* carr_out)
{
    if (!h || !dopplers || D < 1 || D > 32767) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    if (h->plan < 0) return GB_ESTATE;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_acq));
    int rc = ensure(h, &h->tables, &h->tables_cap, (size_t)D * h->N);
    if (rc) return rc;
    std::vector<float> steps(D);
    h->carr.resize(D);
    for (int d = 0; d < D; d++) {
        const float carr = f_if + dopplers[d];                 // doppler_shift.rs:13
        steps[d] = 2.0f * kPiF * carr / h->fs;                 // :14
        h->carr[d] = carr;                                     // :20
        if (carr_out) carr_out[d] = carr;
    }
    float* steps_dev = nullptr;
    CK(cudaMalloc((void**)&steps_dev, sizeof(float) * D));
    CK(cudaMemcpyAsync(steps_dev, steps.data(), sizeof(float) * D, cudaMemcpyHostToDevice, h->s_acq));
    cudaError_t e = gb::acq_launch_doppler_tables(steps_dev, D, h->N, h->tables, h->s_acq);
    cudaStreamSynchronize(h->s_acq);
    cudaFree(steps_dev);
    if (e != cudaSuccess) return fail(h, e, "doppler_table_kernel");
    h->D = D;
    if ((rc = upload_rotators(h))) return rc;
    return build_alias_map(h);
}

extern "C" int gb_acq_set_doppler_tables(gb_handle* h, const gb_c32* tables, const float* carr, int D)
{
    if (!h || !tables || !carr || D < 1 || D > 32767) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    if (h->plan < 0) return GB_ESTATE;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_acq));
    int rc = ensure(h, &h->tables, &h->tables_cap, (size_t)D * h->N);
    if (rc) return rc;
    CK(h2d_blocking(h->tables, tables, (size_t)D * h->N * sizeof(float2)));
    h->carr.assign(carr, carr + D);
    h->D = D;
    h->alias_ok = false;   // caller-supplied tables are used as given
    return upload_rotators(h);
}

extern "C" int gb_acq_get_doppler_tables(gb_handle* h, gb_c32* tables_out, float* carr_out)
{
    if (!h) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    if (h->plan < 0 || h->D == 0) return GB_ESTATE;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_acq));   // the tables may still be being built (doppler_table_kernel)
    if (tables_out) CK(cudaMemcpy(tables_out, h->tables, (size_t)h->D * h->N * sizeof(float2), cudaMemcpyDeviceToHost));
    if (carr_out) memcpy(carr_out, h->carr.data(), sizeof(float) * h->D);
    return GB_OK;
}

extern "C" int gb_acq_set_coherent(gb_handle* h, int n_coh)
{
    if (!h || n_coh < 1 || n_coh > 1024) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    if (h->plan < 0) return GB_ESTATE;
    CK(cudaSetDevice(h->device));
    h->n_coh = n_coh;
    return upload_rotators(h);
}

extern "C" int gb_acq_set_mode(gb_handle* h, int mode)
{
    if (!h || (mode != GB_ACQ_FUSED && mode != GB_ACQ_SHARED && mode != GB_ACQ_SHARED_PLAIN)) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    h->mode = mode;
    return GB_OK;
}

extern "C" int gb_acq_set_detector(gb_handle* h, float threshold, int samples_per_chip)
{
    if (!h || samples_per_chip < 0) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    h->threshold = threshold;
    h->spc = samples_per_chip;
    return GB_OK;
}

// ------------------------------------------------------------------ acquisition search
static int build_rows(gb_handle* h, uint32_t prn_mask, const uint8_t* enable, int slot)
{
    int n = 0;
    for (int p = 0; p < h->n_prn; p++) {
        const bool on = enable ? enable[p] != 0 : (p < 32 ? ((prn_mask >> p) & 1u) != 0 : true);
        if (on) h->rows_pin_slot[slot][n++] = p;
    }
    return n;
}

// prime-factor plans: put the K IQ blocks into line order (inside the timed region) and point the kernels at the
// permuted copies; the kernels' global reads then stay as coalesced as in the Cooley-Tukey plans
static cudaError_t pfa_inputs(gb_handle* h, gb::AcqArgs& a, int K)
{
    cudaError_t e = gb::acq_launch_permute(a.iq, a.iq_start, a.iq_mask, h->npos, h->N, 0, K, h->iq_perm, h->s_acq);
    a.iq = h->iq_perm; a.iq_start = 0; a.iq_mask = ~0ull;
    a.tables = h->tables_perm;
    return e;
}

// Enqueues one search on the acquisition stream and returns without waiting: memset of the cells, row list, (upload,)
// kernels, D2H of the cells into the slot's pinned buffer, then the slot's "done" event.  search_finish() waits for it.
// host_iq != nullptr: the chunk has NOT been uploaded yet.  In the shared-forward chain it is then uploaded in slices of
// whole coherent groups on the copy stream while the forward path (which needs only its own group's blocks) of the
// previous slice runs: cudaMemcpyAsync on a dedicated stream, events to the acquisition stream.
static int search_enqueue(gb_handle* h, const float2* iq_dev, uint64_t start, uint64_t mask, int K, uint32_t prn_mask,
                          const uint8_t* enable, const gb_c32* host_iq, uint64_t local_tail, int slot)
{
    if (h->plan < 0 || h->D == 0) return GB_ESTATE;
    if (K < 1 || K % h->n_coh != 0 || slot < 0 || slot > 1) return GB_EINVAL;
    if (h->pend[slot].active) return GB_ESTATE;   // the slot still holds a search nobody waited for
    const size_t n_cells = (size_t)h->n_prn * h->D;
    if (h->cells_cap < n_cells) {
        if (h->cells_dev) cudaFree(h->cells_dev);
        h->cells_dev = nullptr; h->cells_cap = 0;
        CK(cudaMalloc((void**)&h->cells_dev, n_cells * sizeof(gb_acq_cell)));
        h->cells_cap = n_cells;
    }
    if (h->cells_pin_cap[slot] < n_cells) {
        if (h->cells_pin_slot[slot]) cudaFreeHost(h->cells_pin_slot[slot]);
        h->cells_pin_slot[slot] = nullptr; h->cells_pin_cap[slot] = 0;
        CK(cudaMallocHost((void**)&h->cells_pin_slot[slot], n_cells * sizeof(gb_acq_cell)));
        h->cells_pin_cap[slot] = n_cells;
    }
    const int n_active = build_rows(h, prn_mask, enable, slot);
    CK(cudaMemsetAsync(h->cells_dev, 0, n_cells * sizeof(gb_acq_cell), h->s_acq));
    if (n_active > 0) {
        CK(cudaMemcpyAsync(h->rows_dev + 256 * slot, h->rows_pin_slot[slot], sizeof(int) * n_active, cudaMemcpyHostToDevice, h->s_acq));
        gb::AcqArgs a;
        a.iq = iq_dev; a.iq_start = start; a.iq_mask = mask;
        a.tables = h->tables; a.code_fft = h->code_fft; a.tw = h->tw;
        a.rot = h->n_coh > 1 ? h->rot : nullptr;
        a.rows = h->rows_dev + 256 * slot;
        a.D = h->D; a.K = K; a.n_coh = h->n_coh; a.n_active = n_active; a.spc = h->spc;
        a.cells = h->cells_dev; a.row_out = nullptr; a.d0 = 0; a.spec = nullptr; a.d_lo = 0;
        a.otw = h->otw; a.acc_rows = nullptr; a.npos = h->npos; a.g_lo = 0; a.g_cnt = K / h->n_coh;
        a.plain_inverse = h->mode == GB_ACQ_SHARED_PLAIN;
        a.fwd_bins = nullptr; a.inv_map = nullptr; a.n_prn = h->n_prn;
        a.spec_tc = nullptr; a.code_tc = nullptr; a.tc_n_fwd = 0; a.tc_n_code_sets = 0; a.tc_code_fresh = 0;
        if (h->pfa) {
            int rc = ensure(h, &h->iq_perm, &h->iq_perm_cap, (size_t)K * h->N);
            if (rc) return rc;
        }
        if (host_iq && (h->cluster || h->generic || h->mode == GB_ACQ_FUSED))
            CK(cudaMemcpyAsync(h->chunk, host_iq, (size_t)K * h->N * sizeof(float2), cudaMemcpyHostToDevice, h->s_acq));
        if (h->generic) {
            // any-length plan: Doppler slabs sized to the scratch (n_active * n_d transforms of M points, <= 1 GiB each)
            const int M = gb::generic_plan_m(h->gen);
            size_t max_batch = ((size_t)1 << 27) / (size_t)M;
            if (max_batch > 32768) max_batch = 32768;
            int slab = (int)(max_batch / (size_t)n_active);
            if (slab < 1) return GB_ENOMEM;
            if (slab > h->D) slab = h->D;
            int rc = ensure(h, &h->acc_rows, &h->acc_cap, (size_t)n_active * h->D * h->N);
            if (!rc) rc = ensure(h, &h->gen_s0, &h->gen_s0_cap, (size_t)n_active * slab * M);
            if (!rc) rc = ensure(h, &h->gen_s1, &h->gen_s1_cap, (size_t)n_active * slab * M);
            if (rc) return rc;
            CK(cudaEventRecord(h->ev_s0[slot], h->s_acq));
            for (int d_lo = 0; d_lo < h->D; d_lo += slab) {
                const int n_d = (h->D - d_lo) < slab ? (h->D - d_lo) : slab;
                CK(gb::generic_search_slab(h->gen, a, d_lo, n_d, h->acc_rows, h->gen_s0, h->gen_s1, h->s_acq));
            }
            CK(gb::acq_launch_reduce_rows(h->acc_rows, h->N, h->D, n_active, a.rows, h->spc, h->cells_dev, h->s_acq));
            CK(cudaEventRecord(h->ev_s1[slot], h->s_acq));
        } else if (h->cluster) {
            if (h->n_coh != 1) return GB_EUNSUPPORTED;
            int rc = ensure(h, &h->acc_rows, &h->acc_cap, (size_t)n_active * h->D * h->N);
            if (rc) return rc;
            a.acc_rows = h->acc_rows;
            CK(cudaEventRecord(h->ev_s0[slot], h->s_acq));
            CK(gb::acq_cluster_launch_search(a, h->s_acq));
            CK(cudaEventRecord(h->ev_s1[slot], h->s_acq));
        } else if (h->mode != GB_ACQ_FUSED) {
            // scratch for the forward spectra, processed in Doppler slabs of at most 1 GiB
            const size_t per_d = (size_t)(K / h->n_coh) * h->spec_len;
            size_t slab = ((size_t)1 << 27) / per_d;  // complex elements: 2^27 * 8 B = 1 GiB
            if (slab < 1) slab = 1;
            // Doppler aliasing: forward spectra for the base bins only, every bin's inverse pass in one launch
            const bool alias = h->alias_ok && h->alias_enabled && (size_t)h->n_base <= slab;
            const int n_fwd = alias ? h->n_base : h->D;
            if (alias) {
                a.fwd_bins = h->fwd_bins_dev; a.inv_map = h->inv_map_dev; a.code_fft = h->code_fft_shift;
                slab = h->D;
            }
            if (slab > (size_t)h->D) slab = h->D;
            int rc = ensure(h, &h->spec, &h->spec_cap, (alias ? (size_t)n_fwd : slab) * per_d);
            if (rc) return rc;
            a.spec = h->spec;
            const int n_groups = K / h->n_coh;
            a.g_lo = 0; a.g_cnt = n_groups;
            if (h->N == 4092 && !a.plain_inverse && gb::tuning("acq_tc", 0) && !gb::tuning("acq_nolw", 0)) {
                const size_t tl = (size_t)gb::acq_tc_spec_len();
                const int n_sets = (alias ? h->n_shift : 1) * h->n_prn;
                rc = ensure(h, &h->spec_tc, &h->spec_tc_cap, (alias ? (size_t)n_fwd : slab) * n_groups * tl);
                if (!rc) rc = ensure(h, &h->code_tc, &h->code_tc_cap, (size_t)n_sets * tl);
                if (rc) return rc;
                a.spec_tc = h->spec_tc; a.code_tc = h->code_tc;
                a.tc_n_fwd = alias ? n_fwd : 0; a.tc_n_code_sets = n_sets;
                a.tc_code_fresh = h->code_tc_gen == h->code_gen && h->code_tc_src == a.code_fft;
                h->code_tc_gen = h->code_gen; h->code_tc_src = a.code_fft;
            }
            CK(cudaEventRecord(h->ev_s0[slot], h->s_acq));
            if (host_iq && slab >= (size_t)h->D) {
                // sliced upload overlapped with the forward path
                // one call at a time: slices, so that the forward path of slice s overlaps the upload of slice s + 1.  With
                // another search in flight (enqueue / wait pair) the whole upload already hides behind that search's
                // inverse kernel, and one full-size forward launch beats four small ones
                const bool other_in_flight = h->pend[slot ^ 1].active;
                const int n_slices = other_in_flight ? 1 : (n_groups >= 8 ? 4 : (n_groups >= 2 ? 2 : 1));
                if (h->pfa) { a.iq = h->iq_perm; a.iq_start = 0; a.iq_mask = ~0ull; a.tables = h->tables_perm; }
                a.d_lo = 0;
                // the previous search's last reader of `chunk` (its permute / forward kernels -- NOT its inverse kernel)
                // must have run before this upload overwrites it; the upload then overlaps that inverse kernel
                if (h->chunk_free_valid) CK(cudaStreamWaitEvent(h->s_copy, h->ev_chunk_free, 0));
                for (int sl = 0; sl < n_slices; sl++) {
                    const int g0 = (int)((long long)n_groups * sl / n_slices), g1 = (int)((long long)n_groups * (sl + 1) / n_slices);
                    const size_t b0 = (size_t)g0 * h->n_coh, nb = (size_t)(g1 - g0) * h->n_coh;
                    if (!h->ev_slice[sl]) CK(cudaEventCreateWithFlags(&h->ev_slice[sl], cudaEventDisableTiming));
                    CK(cudaMemcpyAsync(h->chunk + b0 * h->N, host_iq + b0 * h->N, nb * h->N * sizeof(float2),
                                       cudaMemcpyHostToDevice, h->s_copy));
                    CK(cudaEventRecord(h->ev_slice[sl], h->s_copy));
                    CK(cudaStreamWaitEvent(h->s_acq, h->ev_slice[sl], 0));
                    if (h->pfa)
                        CK(gb::acq_launch_permute(h->chunk, 0, ~0ull, h->npos, h->N, (int)b0, (int)nb, h->iq_perm, h->s_acq));
                    a.g_lo = g0; a.g_cnt = g1 - g0;
                    CK(gb::acq_launch_forward(h->plan, a, n_fwd, h->s_acq));
                }
                CK(cudaEventRecord(h->ev_chunk_free, h->s_acq));
                h->chunk_free_valid = true;
                a.g_lo = 0; a.g_cnt = 0;   // forward path done: inverse kernel only
                CK(gb::acq_launch_shared(h->plan, a, h->D, h->s_acq));
            } else {
                if (host_iq) CK(cudaMemcpyAsync(h->chunk, host_iq, (size_t)K * h->N * sizeof(float2), cudaMemcpyHostToDevice, h->s_acq));
                if (h->pfa) CK(pfa_inputs(h, a, K));
                if (alias) {
                    a.d_lo = 0;
                    CK(gb::acq_launch_forward(h->plan, a, n_fwd, h->s_acq));
                    a.g_cnt = 0;   // forward path done: inverse kernel only
                    CK(gb::acq_launch_shared(h->plan, a, h->D, h->s_acq));
                } else {
                    for (int d_lo = 0; d_lo < h->D; d_lo += (int)slab) {
                        a.d_lo = d_lo;
                        const int n_d = (h->D - d_lo) < (int)slab ? (h->D - d_lo) : (int)slab;
                        CK(gb::acq_launch_shared(h->plan, a, n_d, h->s_acq));
                    }
                }
            }
            CK(cudaEventRecord(h->ev_s1[slot], h->s_acq));
        } else {
            CK(cudaEventRecord(h->ev_s0[slot], h->s_acq));
            if (h->pfa) CK(pfa_inputs(h, a, K));
            CK(gb::acq_launch_search(h->plan, a, h->s_acq));
            CK(cudaEventRecord(h->ev_s1[slot], h->s_acq));
        }
    }
    CK(cudaMemcpyAsync(h->cells_pin_slot[slot], h->cells_dev, n_cells * sizeof(gb_acq_cell), cudaMemcpyDeviceToHost, h->s_acq));
    CK(cudaEventRecord(h->ev_done[slot], h->s_acq));
    gb_handle::Pending& pd = h->pend[slot];
    pd.active = true; pd.local_tail = local_tail; pd.prn_mask = prn_mask; pd.n_active = n_active; pd.n_cells = n_cells;
    pd.has_enable = enable != nullptr;
    if (enable) pd.enable.assign(enable, enable + h->n_prn);
    return GB_OK;
}

static int search_finish(gb_handle* h, int slot, gb_acq_cell* cells_out)
{
    if (slot < 0 || slot > 1 || !h->pend[slot].active) return GB_ESTATE;
    gb_handle::Pending& pd = h->pend[slot];
    pd.active = false;
    CK(cudaEventSynchronize(h->ev_done[slot]));
    if (pd.n_active > 0) CK(cudaEventElapsedTime(&h->last_acq_ms, h->ev_s0[slot], h->ev_s1[slot]));
    else h->last_acq_ms = 0.f;
    if (cells_out) memcpy(cells_out, h->cells_pin_slot[slot], pd.n_cells * sizeof(gb_acq_cell));
    return GB_OK;
}

static int search_cells(gb_handle* h, const float2* iq_dev, uint64_t start, uint64_t mask, int K, uint32_t prn_mask,
                        const uint8_t* enable, gb_acq_cell* cells_out, const gb_c32* host_iq = nullptr, uint64_t local_tail = 0)
{
    int rc = search_enqueue(h, iq_dev, start, mask, K, prn_mask, enable, host_iq, local_tail, 0);
    if (rc) return rc;
    return search_finish(h, 0, cells_out);
}

// Host sample buffers cross the boundary WITH their length: K * fft_size samples are read
static int host_chunk_ok(gb_handle* h, const gb_c32* iq, uint64_t n_samples, int K)
{
    if (!iq || K < 1) return GB_EINVAL;
    if (h->plan < 0) return GB_ESTATE;
    if (n_samples < (uint64_t)K * (uint64_t)h->N) return GB_ERANGE;
    return GB_OK;
}

extern "C" int gb_acq_search_cells(gb_handle* h, const gb_c32* iq, uint64_t n_samples, int K, uint32_t prn_mask,
                                   const uint8_t* enable, gb_acq_cell* cells_out)
{
    if (!h) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    CK(cudaSetDevice(h->device));
    int rc = host_chunk_ok(h, iq, n_samples, K);
    if (rc) return rc;
    rc = ensure(h, &h->chunk, &h->chunk_cap, (size_t)K * h->N);
    if (rc) return rc;
    return search_cells(h, h->chunk, 0, ~0ull, K, prn_mask, enable, cells_out, iq);
}

static int ring_range_ok(gb_handle* h, uint64_t local_tail, uint64_t n)
{
    if (!h->ring) return GB_ESTATE;
    if (local_tail + n > h->ring_head) return GB_ERANGE;             // not written yet
    if (h->ring_head - local_tail > h->ring_cap) return GB_ERANGE;   // already overwritten
    return GB_OK;
}

extern "C" int gb_acq_search_cells_ring(gb_handle* h, uint64_t local_tail, int K, uint32_t prn_mask,
                                        const uint8_t* enable, gb_acq_cell* cells_out)
{
    if (!h || K < 1) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    if (h->plan < 0) return GB_ESTATE;
    CK(cudaSetDevice(h->device));
    {
        std::lock_guard<std::recursive_mutex> lr(h->mu_ring);   // the writer thread may be inside gb_ring_write
        int rc = ring_range_ok(h, local_tail, (uint64_t)K * h->N);
        if (rc) return rc;
        CK(cudaStreamWaitEvent(h->s_acq, h->ev_copy, 0));
    }
    return search_cells(h, h->ring, local_tail, h->ring_cap - 1, K, prn_mask, enable, cells_out);
}

// is_good_satellite on a cell (do_acquisition.rs:235-237)
static inline float cell_metric(float peak, float sum8, int fft_size)
{
    const float avg = (sum8 - peak) / (float)(fft_size - 1);
    return peak / avg;
}

// search_satellite's control flow (do_acquisition.rs:171-223) on per-bin cells (Q1): running best with strict '>',
// tested after every bin; the first bin at which the running best passes wins.
extern "C" int gb_acq_decide(const gb_acq_cell* cells, const float* carr, int D, int prn, int fft_size, float fs,
                             uint64_t local_tail, float threshold, gb_acq_result* out)
{
    if (!cells || !carr || !out || D < 1 || fft_size < 2) return GB_EINVAL;
    memset(out, 0, sizeof(*out));
    out->prn = (uint8_t)prn;
    out->doppler_bin = -1;
    float gmax = 0.0f, gsum = 0.0f, gfreq = 0.0f, gp2 = 0.0f;
    uint32_t gphase = 0;
    int gbin = -1;
    for (int d = 0; d < D; d++) {
        if (cells[d].peak > gmax) {
            gmax = cells[d].peak; gsum = cells[d].sum8; gphase = cells[d].argmax; gfreq = carr[d]; gp2 = cells[d].peak2;
            gbin = d;
        }
        const float metric = cell_metric(gmax, gsum, fft_size);  // 0/0 = NaN before any record -> false
        if (metric > threshold) {
            out->found = 1;
            out->doppler_bin = (int16_t)gbin;
            out->code_phase_samples = gphase;
            out->code_phase_chips = (float)gphase * kCodeRate / fs;  // :213-214
            out->carrier_freq = gfreq;
            out->fs = fs;
            out->mag_relative = gmax;
            out->sample_global_index = local_tail + gphase;
            out->metric = metric;
            out->peak_ratio = gp2 > 0.f ? sqrtf(gmax / gp2) : 0.f;
            return GB_OK;
        }
    }
    return GB_OK;
}

static int decide_all(gb_handle* h, int slot, uint64_t local_tail, uint32_t prn_mask, const uint8_t* enable, gb_acq_result* results)
{
    for (int p = 0; p < h->n_prn; p++) {
        const bool on = enable ? enable[p] != 0 : (p < 32 ? ((prn_mask >> p) & 1u) != 0 : true);
        if (on) {
            int rc = gb_acq_decide(h->cells_pin_slot[slot] + (size_t)p * h->D, h->carr.data(), h->D, p + 1, h->N, h->fs,
                                   local_tail, h->threshold, &results[p]);
            if (rc) return rc;
        } else {
            memset(&results[p], 0, sizeof(gb_acq_result));
            results[p].prn = (uint8_t)(p + 1);
            results[p].doppler_bin = -1;
        }
    }
    return GB_OK;
}

extern "C" int gb_acq_search(gb_handle* h, const gb_c32* iq, uint64_t n_samples, int K, uint64_t local_tail, uint32_t prn_mask,
                             const uint8_t* enable, gb_acq_result* results)
{
    if (!h || !results) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    int rc = gb_acq_search_cells(h, iq, n_samples, K, prn_mask, enable, nullptr);
    if (rc) return rc;
    return decide_all(h, 0, local_tail, prn_mask, enable, results);
}

extern "C" int gb_acq_search_ring(gb_handle* h, uint64_t local_tail, int K, uint32_t prn_mask, const uint8_t* enable,
                                  gb_acq_result* results)
{
    if (!h || !results) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    int rc = gb_acq_search_cells_ring(h, local_tail, K, prn_mask, enable, nullptr);
    if (rc) return rc;
    return decide_all(h, 0, local_tail, prn_mask, enable, results);
}

// Asynchronous pair on ONE handle (two slots): enqueue returns as soon as the copies and kernels are queued; the upload
// of the next search overlaps the inverse kernel of the one in flight.  iq must stay valid (and should be pinned) until
// the matching wait.
extern "C" int gb_acq_search_enqueue(gb_handle* h, const gb_c32* iq, uint64_t n_samples, int K, uint64_t local_tail,
                                     uint32_t prn_mask, const uint8_t* enable, int slot)
{
    if (!h) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    CK(cudaSetDevice(h->device));
    int rc = host_chunk_ok(h, iq, n_samples, K);
    if (rc) return rc;
    rc = ensure(h, &h->chunk, &h->chunk_cap, (size_t)K * h->N);
    if (rc) return rc;
    return search_enqueue(h, h->chunk, 0, ~0ull, K, prn_mask, enable, iq, local_tail, slot);
}

extern "C" int gb_acq_search_wait(gb_handle* h, int slot, gb_acq_result* results, gb_acq_cell* cells_out)
{
    if (!h || slot < 0 || slot > 1) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    CK(cudaSetDevice(h->device));
    const gb_handle::Pending pd = h->pend[slot];
    int rc = search_finish(h, slot, cells_out);
    if (rc) return rc;
    if (results) return decide_all(h, slot, pd.local_tail, pd.prn_mask, pd.has_enable ? pd.enable.data() : nullptr, results);
    return GB_OK;
}

// n_rec recordings searched back to back with the pair above (recording r + 1 uploads while r is being searched)
extern "C" int gb_acq_search_batch(gb_handle* h, const gb_c32* const* recordings, int n_rec, uint64_t n_samples, int K,
                                   uint64_t local_tail, uint32_t prn_mask, const uint8_t* enable, gb_acq_result* results)
{
    if (!h || !recordings || !results || n_rec < 1) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    int rc = gb_acq_search_enqueue(h, recordings[0], n_samples, K, local_tail, prn_mask, enable, 0);
    if (rc) return rc;
    for (int r = 0; r < n_rec; r++) {
        if (r + 1 < n_rec) {
            rc = gb_acq_search_enqueue(h, recordings[r + 1], n_samples, K, local_tail, prn_mask, enable, (r + 1) & 1);
            if (rc) {
                gb_acq_search_wait(h, r & 1, nullptr, nullptr);
                return rc;
            }
        }
        rc = gb_acq_search_wait(h, r & 1, results + (size_t)r * h->n_prn, nullptr);
        if (rc) return rc;
    }
    return GB_OK;
}

static int stage_chunk(gb_handle* h, const gb_c32* iq, int K)
{
    const size_t n = (size_t)K * h->N;
    int rc = ensure(h, &h->chunk, &h->chunk_cap, n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->chunk, iq, n * sizeof(float2), cudaMemcpyHostToDevice, h->s_acq));
    return GB_OK;
}

extern "C" int gb_acq_bin_power(gb_handle* h, const gb_c32* iq, uint64_t n_samples, int K, int prn, int doppler_bin,
                                float* power_out)
{
    if (!h || !power_out) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    if (h->plan < 0 || h->D == 0) return GB_ESTATE;
    if (prn < 1 || prn > h->n_prn || doppler_bin < 0 || doppler_bin >= h->D || K % h->n_coh != 0) return GB_EINVAL;
    if (h->pend[0].active || h->pend[1].active) return GB_ESTATE;   // not while an enqueued search is in flight
    CK(cudaSetDevice(h->device));
    int rc = host_chunk_ok(h, iq, n_samples, K);
    if (rc) return rc;
    rc = stage_chunk(h, iq, K);
    if (rc) return rc;
    h->rows_pin[0] = prn - 1;
    CK(cudaMemcpyAsync(h->rows_dev, h->rows_pin, sizeof(int), cudaMemcpyHostToDevice, h->s_acq));
    gb::AcqArgs a;
    a.iq = h->chunk; a.iq_start = 0; a.iq_mask = ~0ull;
    a.tables = h->tables; a.code_fft = h->code_fft; a.tw = h->tw;
    a.rot = h->n_coh > 1 ? h->rot : nullptr;
    a.rows = h->rows_dev;
    a.D = h->D; a.K = K; a.n_coh = h->n_coh; a.n_active = 1; a.spc = 0;
    a.cells = nullptr; a.row_out = h->row_dev; a.d0 = doppler_bin; a.spec = nullptr; a.d_lo = 0;
    a.otw = h->otw; a.acc_rows = nullptr; a.npos = h->npos; a.g_lo = 0; a.g_cnt = K / h->n_coh;
    a.plain_inverse = h->mode == GB_ACQ_SHARED_PLAIN;
    a.fwd_bins = nullptr; a.inv_map = nullptr; a.n_prn = h->n_prn;
    a.spec_tc = nullptr; a.code_tc = nullptr; a.tc_n_fwd = 0; a.tc_n_code_sets = 0; a.tc_code_fresh = 0;
    if (h->pfa) {
        rc = ensure(h, &h->iq_perm, &h->iq_perm_cap, (size_t)K * h->N);
        if (rc) return rc;
        CK(pfa_inputs(h, a, K));
    }
    if (h->generic) {
        // one (prn, bin) row through the any-length plan: D = 1 view of the requested bin
        const int M = gb::generic_plan_m(h->gen);
        rc = ensure(h, &h->acc_rows, &h->acc_cap, (size_t)h->N);
        if (!rc) rc = ensure(h, &h->gen_s0, &h->gen_s0_cap, (size_t)M);
        if (!rc) rc = ensure(h, &h->gen_s1, &h->gen_s1_cap, (size_t)M);
        if (rc) return rc;
        a.tables = h->tables + (size_t)doppler_bin * h->N;
        a.rot = h->n_coh > 1 ? h->rot + (size_t)doppler_bin * h->n_coh : nullptr;
        a.D = 1;
        CK(gb::generic_search_slab(h->gen, a, 0, 1, h->acc_rows, h->gen_s0, h->gen_s1, h->s_acq));
        CK(cudaMemcpyAsync(power_out, h->acc_rows, sizeof(float) * h->N, cudaMemcpyDeviceToHost, h->s_acq));
        CK(cudaStreamSynchronize(h->s_acq));
        return GB_OK;
    }
    if (h->cluster) {
        // one (prn, bin) cell row: the cluster kernel leaves the accumulated power row in acc_rows
        if (h->n_coh != 1) return GB_EUNSUPPORTED;
        rc = ensure(h, &h->acc_rows, &h->acc_cap, (size_t)h->N);
        if (rc) return rc;
        const size_t need = (size_t)h->n_prn * h->D;
        if (h->cells_cap < need) {
            if (h->cells_dev) cudaFree(h->cells_dev);
            h->cells_dev = nullptr; h->cells_cap = 0;
            CK(cudaMalloc((void**)&h->cells_dev, need * sizeof(gb_acq_cell)));
            h->cells_cap = need;
        }
        a.tables = h->tables + (size_t)doppler_bin * h->N;
        a.D = 1; a.acc_rows = h->acc_rows; a.cells = h->cells_dev;
        CK(gb::acq_cluster_launch_search(a, h->s_acq));
        CK(cudaMemcpyAsync(power_out, h->acc_rows, sizeof(float) * h->N, cudaMemcpyDeviceToHost, h->s_acq));
        CK(cudaStreamSynchronize(h->s_acq));
        return GB_OK;
    }
    CK(gb::acq_launch_row(h->plan, a, h->s_acq));
    CK(cudaMemcpyAsync(power_out, h->row_dev, sizeof(float) * h->N, cudaMemcpyDeviceToHost, h->s_acq));
    CK(cudaStreamSynchronize(h->s_acq));
    return GB_OK;
}

// ------------------------------------------------------------------ fine Doppler
// finer_doppler, acquisition_bk.rs:215-302.  Device work in fine_doppler.cu; the frequency mapping (:250-253, :282-299)
// is a handful of f32 operations per request and stays on the host, in the reference's order.
static int fine_common(gb_handle* h, const float2* x_dev, uint64_t start, uint64_t mask, uint64_t n_long, float fs,
                       int long_ms, int is_complex, const gb_fine_req* req, int n_req, const int8_t* codes1023,
                       gb_fine_result* out, float* mag_out)
{
    if (!req || !out || n_req < 1 || long_ms < 2 || !(fs > 0.f)) return GB_EINVAL;
    const size_t n_code = f32_as_usize(roundf(fs / (kCodeRate / 1023.0f)));   // :236-239
    const size_t use = (size_t)(long_ms - 1) * n_code;                         // :240
    if (use < 2) return GB_EINVAL;
    int m = 0;
    while (((size_t)1 << m) < use) m++;                                        // next_power_of_two, :249
    if (m < 12 || m > 19) return GB_EUNSUPPORTED;                              // 4096 <= P2 <= 524288 samples
    const int log_a = m / 2 > 8 ? 8 : m / 2, log_b = m - log_a;
    const size_t P2 = (size_t)1 << m, M = P2 * 8;
    std::vector<int8_t> codes((size_t)n_req * 1023);
    std::vector<unsigned long long> cps(n_req);
    for (int i = 0; i < n_req; i++) {
        // the legacy slice samples_iq[code_phase..size_signal_use + code_phase] panics when it runs past the end (:258)
        if ((uint64_t)req[i].code_phase + use > n_long) return GB_ERANGE;
        cps[i] = req[i].code_phase;
        if (codes1023) memcpy(&codes[(size_t)i * 1023], codes1023 + (size_t)i * 1023, 1023);
        else {
            if (req[i].prn < 1 || req[i].prn > 32) return GB_EINVAL;
            ca_chips(req[i].prn, &codes[(size_t)i * 1023]);
        }
    }
    int rc;
    if ((rc = ensure(h, &h->fine_codes, &h->fine_codes_cap, codes.size()))) return rc;
    if ((rc = ensure(h, &h->fine_u64, &h->fine_u64_cap, (size_t)2 * n_req))) return rc;
    if ((rc = ensure(h, &h->fine_y, &h->fine_y_cap, (size_t)n_req * M))) return rc;
    if (!h->fine_mean) CK(cudaMalloc((void**)&h->fine_mean, sizeof(float2)));
    if (mag_out && (rc = ensure(h, &h->fine_mag, &h->fine_mag_cap, (size_t)n_req * M))) return rc;
    CK(cudaMemcpyAsync(h->fine_codes, codes.data(), codes.size(), cudaMemcpyHostToDevice, h->s_acq));
    CK(cudaMemcpyAsync(h->fine_u64, cps.data(), sizeof(unsigned long long) * n_req, cudaMemcpyHostToDevice, h->s_acq));
    gb::FineArgs a;
    a.x = x_dev; a.start = start; a.mask = mask; a.use = (unsigned)use; a.fs = fs; a.log_a = log_a; a.log_b = log_b;
    a.codes = h->fine_codes; a.code_phase = h->fine_u64; a.mean = h->fine_mean; a.Y = h->fine_y;
    a.best = h->fine_u64 + n_req; a.mag_out = mag_out ? h->fine_mag : nullptr;
    CK(cudaEventRecord(h->ev_a0, h->s_acq));
    CK(gb::fine_launch(a, n_req, x_dev, start, mask, n_long, h->s_acq));
    CK(cudaEventRecord(h->ev_a1, h->s_acq));
    std::vector<unsigned long long> best(n_req);
    CK(cudaMemcpyAsync(best.data(), h->fine_u64 + n_req, sizeof(unsigned long long) * n_req, cudaMemcpyDeviceToHost, h->s_acq));
    if (mag_out) CK(cudaMemcpyAsync(mag_out, h->fine_mag, sizeof(float) * (size_t)n_req * M, cudaMemcpyDeviceToHost, h->s_acq));
    CK(cudaStreamSynchronize(h->s_acq));
    cudaEventElapsedTime(&h->last_fine_ms, h->ev_a0, h->ev_a1);
    const size_t one_side = f32_as_usize(ceilf(((float)M + 1.0f) / 2.0f));     // :250
    for (int i = 0; i < n_req; i++) {
        const uint32_t idx = 0xffffffffu - (uint32_t)(best[i] & 0xffffffffull);
        const uint32_t mbits = (uint32_t)(best[i] >> 32);
        gb_fine_result& r = out[i];
        r.fft_size = (uint32_t)M;
        r.idx = idx;
        memcpy(&r.mag, &mbits, 4);
        if (idx < one_side) {
            const float bin = (float)idx * fs / (float)M;                       // :251-253
            r.carrier_freq = (is_complex ? -1.0f : 1.0f) * bin;                 // :296-298
            r.ref_defined = 1;
        } else {
            // the legacy reads fft_freq_bins[one_side] (out of bounds) and panics; this is the value its arithmetic
            // would produce had the Vec been long enough (:283-295): idx' = idx - one_side, carrier = +bins(one_side - idx')
            const size_t i2 = (size_t)idx - one_side;
            r.carrier_freq = (float)(one_side - i2) * fs / (float)M;
            r.ref_defined = 0;
        }
    }
    return GB_OK;
}

extern "C" int gb_acq_fine_doppler(gb_handle* h, const gb_c32* long_samples, uint64_t n_long, float fs, int long_ms,
                                   int is_complex, const gb_fine_req* req, int n_req, const int8_t* codes1023,
                                   gb_fine_result* out, float* mag_out)
{
    if (!h || !long_samples || n_long < 1) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    CK(cudaSetDevice(h->device));
    int rc = ensure(h, &h->fine_x, &h->fine_x_cap, (size_t)n_long);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->fine_x, long_samples, n_long * sizeof(float2), cudaMemcpyHostToDevice, h->s_acq));
    return fine_common(h, h->fine_x, 0, ~0ull, n_long, fs, long_ms, is_complex, req, n_req, codes1023, out, mag_out);
}

extern "C" int gb_acq_fine_doppler_ring(gb_handle* h, uint64_t start, uint64_t n_long, float fs, int long_ms, int is_complex,
                                        const gb_fine_req* req, int n_req, const int8_t* codes1023, gb_fine_result* out)
{
    if (!h || n_long < 1) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    CK(cudaSetDevice(h->device));
    int rc = ring_range_ok(h, start, n_long);
    if (rc) return rc;
    CK(cudaStreamWaitEvent(h->s_acq, h->ev_copy, 0));
    return fine_common(h, h->ring, start, h->ring_cap - 1, n_long, fs, long_ms, is_complex, req, n_req, codes1023, out,
                       nullptr);
}

extern "C" float gb_acq_fine_last_kernel_ms(gb_handle* h) { return h ? h->last_fine_ms : 0.f; }

extern "C" float gb_acq_last_kernel_ms(gb_handle* h) { return h ? h->last_acq_ms : 0.f; }

// Measured FP32 (non-tensor) FMA throughput of this GPU, for the roofline denominator
extern "C" int gb_bench_fp32_tflops(gb_handle* h, float* tflops_out)
{
    if (!h || !tflops_out) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    CK(cudaSetDevice(h->device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    float* d = nullptr;
    CK(cudaMalloc((void**)&d, 4));
    const int blocks = sms * 8, threads = 256, kIters = 1024;  // 8 resident CTAs per SM, ~2 ms per launch
    float best = 0.f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(h->ev_a0, h->s_acq);
        fp32_peak_kernel<<<blocks, threads, 0, h->s_acq>>>(d, 1.0000001f, 1e-9f, kIters);
        cudaEventRecord(h->ev_a1, h->s_acq);
        cudaError_t e = cudaStreamSynchronize(h->s_acq);
        if (e != cudaSuccess) { cudaFree(d); return fail(h, e, "fp32_peak_kernel"); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_a0, h->ev_a1);
        const double flops = 2.0 * 8.0 * 32.0 * (double)kIters * (double)blocks * threads;
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(d);
    *tflops_out = best;
    return GB_OK;
}

// ------------------------------------------------------------------ FFT facade (fft.rs:5-56)
// Any length (and f64): the any-length plan of acq_generic.cu -- FFT<T> / RealFFT<T> accept every N in the reference
// (rustfft / realfft planners, fft.rs:12-15, :39-40) and are generic over f32 / f64.
template <typename T, typename T2>
static int fft_any(gb_handle* h, int n, int inverse, const void* in, void* out, int batch, int real_in, int power_out, int n_out)
{
    if (n > kGenericMaxN || batch > 32768) return GB_EUNSUPPORTED;
    CK(cudaSetDevice(h->device));
    gb::GenericPlan* gp;
    int rc = generic_plan_for(h, n, &gp);
    if (rc) return rc;
    const size_t M = (size_t)gb::generic_plan_m(gp);
    const size_t in_bytes = (size_t)batch * n * (real_in ? sizeof(T) : sizeof(T2));
    const size_t out_bytes = (size_t)batch * n_out * (power_out ? sizeof(T) : sizeof(T2));
    void *din = nullptr, *dout = nullptr;
    T2 *x = nullptr, *s0 = nullptr, *s1 = nullptr;
    cudaError_t e = cudaMalloc(&din, in_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dout, out_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&x, sizeof(T2) * (size_t)batch * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s0, sizeof(T2) * (size_t)batch * M);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s1, sizeof(T2) * (size_t)batch * M);
    if (e == cudaSuccess) e = cudaMemcpyAsync(din, in, in_bytes, cudaMemcpyHostToDevice, h->s_acq);
    if (e == cudaSuccess) {
        const size_t total = (size_t)batch * n;
        if (sizeof(T) == 4) {
            float2* xin = real_in ? (float2*)x : (float2*)din;
            if (real_in) e = gb::generic_real_to_complex((const float*)din, (float2*)x, total, h->s_acq);
            if (e == cudaSuccess) e = gb::generic_dft_f32(gp, inverse, xin, (float2*)x, batch, (float2*)s0, (float2*)s1, h->s_acq);
            if (e == cudaSuccess) e = gb::generic_take_f32((const float2*)x, dout, n, n_out, batch, power_out, h->s_acq);
        } else {
            double2* xin = real_in ? (double2*)x : (double2*)din;
            if (real_in) e = gb::generic_r2c_f64((const double*)din, (double2*)x, total, h->s_acq);
            if (e == cudaSuccess) e = gb::generic_dft_f64(gp, inverse, xin, (double2*)x, batch, (double2*)s0, (double2*)s1, h->s_acq);
            if (e == cudaSuccess) e = gb::generic_take_f64((const double2*)x, dout, n, n_out, batch, power_out, h->s_acq);
        }
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, h->s_acq);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_acq);
    for (void* p : {din, dout, (void*)x, (void*)s0, (void*)s1})
        if (p) cudaFree(p);
    if (e != cudaSuccess) return fail(h, e, "fft (any-length plan)");
    return GB_OK;
}

static int fft_common(gb_handle* h, int n, int inverse, const void* in, void* out, int batch, int real_in, int power_out,
                      int n_out)
{
    if (!h || !in || !out || batch < 1 || n < 2) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    const int plan = gb::acq_plan_index(n);
    if (plan < 0) return fft_any<float, float2>(h, n, inverse, in, out, batch, real_in, power_out, n_out);
    CK(cudaSetDevice(h->device));
    FftRes* fr;
    int rc = fft_resources(h, plan, n, &fr);
    if (rc) return rc;
    const size_t in_bytes = (size_t)batch * n * (real_in ? sizeof(float) : sizeof(float2));
    const size_t out_bytes = (size_t)batch * n_out * (power_out ? sizeof(float) : sizeof(float2));
    void *din = nullptr, *dout = nullptr;
    CK(cudaMalloc(&din, in_bytes));
    cudaError_t e = cudaMalloc(&dout, out_bytes);
    if (e != cudaSuccess) { cudaFree(din); return fail(h, e, "cudaMalloc"); }
    gb::FftArgs a;
    a.in = din; a.out = dout; a.tw = fr->tw; a.freq_of_pos = fr->fop; a.npos = fr->npos;
    a.real_in = real_in; a.power_out = power_out; a.n_out = n_out;
    e = cudaMemcpyAsync(din, in, in_bytes, cudaMemcpyHostToDevice, h->s_acq);
    if (e == cudaSuccess) e = gb::acq_launch_fft(plan, inverse, a, batch, h->s_acq);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, h->s_acq);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_acq);
    cudaFree(din);
    cudaFree(dout);
    if (e != cudaSuccess) return fail(h, e, "fft");
    return GB_OK;
}
extern "C" int gb_fft_c2c(gb_handle* h, int n, int inverse, const gb_c32* in, gb_c32* out, int batch)
{
    return fft_common(h, n, inverse, in, out, batch, 0, 0, n);
}
extern "C" int gb_fft_power_spectrum(gb_handle* h, int n, const gb_c32* in, float* out, int batch)
{
    return fft_common(h, n, 0, in, out, batch, 0, 1, n);
}
extern "C" int gb_rfft(gb_handle* h, int n, const float* in, gb_c32* out, int batch)
{
    return fft_common(h, n, 0, in, out, batch, 1, 0, n / 2 + 1);
}
// RealFFT<f32>::power_spectrum (fft.rs:47-55): norm_sqr of the n/2 + 1 bins
extern "C" int gb_rfft_power_spectrum(gb_handle* h, int n, const float* in, float* out, int batch)
{
    return fft_common(h, n, 0, in, out, batch, 1, 1, n / 2 + 1);
}
// FFT<f64> / RealFFT<f64> (fft.rs is generic over T: Float): always through the any-length plan
static int fft_f64(gb_handle* h, int n, int inverse, const void* in, void* out, int batch, int real_in, int power_out, int n_out)
{
    if (!h || !in || !out || batch < 1 || n < 2) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_acq);
    return fft_any<double, double2>(h, n, inverse, in, out, batch, real_in, power_out, n_out);
}
extern "C" int gb_fft_c2c_f64(gb_handle* h, int n, int inverse, const gb_c64* in, gb_c64* out, int batch)
{
    return fft_f64(h, n, inverse, in, out, batch, 0, 0, n);
}
extern "C" int gb_fft_power_spectrum_f64(gb_handle* h, int n, const gb_c64* in, double* out, int batch)
{
    return fft_f64(h, n, 0, in, out, batch, 0, 1, n);
}
extern "C" int gb_rfft_f64(gb_handle* h, int n, const double* in, gb_c64* out, int batch)
{
    return fft_f64(h, n, 0, in, out, batch, 1, 0, n / 2 + 1);
}
extern "C" int gb_rfft_power_spectrum_f64(gb_handle* h, int n, const double* in, double* out, int batch)
{
    return fft_f64(h, n, 0, in, out, batch, 1, 1, n / 2 + 1);
}

// ------------------------------------------------------------------ tracking: host helpers
// LoopFilter::new (do_tracking.rs:59-64)
extern "C" int gb_loop_filter_new(float noise_bw, float damping, float gain, float* tau1, float* tau2)
{
    if (!tau1 || !tau2) return GB_EINVAL;
    const float w = noise_bw * 8.0f * damping / (4.0f * powf(damping, 2.0f) + 1.0f);
    *tau1 = gain / (w * w);
    *tau2 = (2.0f * damping) / w;
    return GB_OK;
}
// TrackingChannel::new (do_tracking.rs:118-146, constants :16-29)
extern "C" int gb_trk_channel_init(gb_trk_channel* c, uint8_t id, float fs)
{
    if (!c) return GB_EINVAL;
    memset(c, 0, sizeof(*c));
    c->id = id;
    c->state = GB_TRK_IDLE;
    c->fs = fs;
    c->num_samples_per_code = (uint64_t)gb_num_samples_per_code(kCodeRate, fs);
    c->code_rate = kCodeRate;
    gb_loop_filter_new(25.0f, 0.7f, 0.25f, &c->pll_tau1, &c->pll_tau2);
    gb_loop_filter_new(2.0f, 0.7f, 1.0f, &c->dll_tau1, &c->dll_tau2);
    return GB_OK;
}
// TrackingChannel::start (do_tracking.rs:148-154, Q8); code_row = prn reproduces Q6
extern "C" int gb_trk_channel_start(gb_trk_channel* c, const gb_acq_result* r)
{
    if (!c || !r) return GB_EINVAL;
    c->prn = r->prn;
    c->code_row = r->prn;
    c->carrier_freq = r->carrier_freq;
    c->code_phase = r->code_phase_chips;
    c->next_sample_index = r->sample_global_index;
    c->state = GB_TRK_TRACKING;
    return GB_OK;
}
// the same hand-over with the C/A row of the satellite itself (prn - 1), i.e. without the reference's Q6 off-by-one
extern "C" int gb_trk_channel_start_corrected(gb_trk_channel* c, const gb_acq_result* r)
{
    const int rc = gb_trk_channel_start(c, r);
    if (rc) return rc;
    if (r->prn < 1 || r->prn > 32) return GB_EINVAL;
    c->code_row = (uint8_t)(r->prn - 1);
    return GB_OK;
}

// TrackingChannel::reset (do_tracking.rs:311-326, Q9)
extern "C" int gb_trk_channel_reset(gb_trk_channel* c)
{
    if (!c) return GB_EINVAL;
    c->prn = 0; c->code_row = 0; c->state = GB_TRK_IDLE; c->lost_counter = 0; c->next_sample_index = 0;
    c->carrier_freq = 0; c->carrier_phase = 0; c->carrier_error = 0; c->carrier_nco = 0;
    c->code_phase = 0; c->code_error = 0; c->code_nco = 0; c->code_rate = 0;
    c->i_prompt = 0; c->q_prompt = 0;
    return GB_OK;
}

// ------------------------------------------------------------------ tracking: device
static int trk_reserve(gb_handle* h, int n)
{
    if (n <= h->ch_cap) return GB_OK;
    void* olds[] = {h->ch_dev, h->corr_dev, h->ran_dev, h->lost_dev, h->offs_dev};
    for (void* p : olds)
        if (p) cudaFree(p);
    h->ch_dev = nullptr; h->corr_dev = nullptr; h->ran_dev = nullptr; h->lost_dev = nullptr; h->offs_dev = nullptr;
    h->ch_cap = 0;
    CK(cudaMalloc((void**)&h->ch_dev, sizeof(gb_trk_channel) * n));
    CK(cudaMalloc((void**)&h->corr_dev, sizeof(gb_trk_corr) * n));
    CK(cudaMalloc((void**)&h->ran_dev, n));
    CK(cudaMalloc((void**)&h->lost_dev, n));
    CK(cudaMalloc((void**)&h->offs_dev, sizeof(unsigned long long) * n));
    h->ch_cap = n;
    return GB_OK;
}

// Largest sample rate of the batch.  A channel whose code_row is out of the table (the reference's get_ca_chip indexes
// GPS_CA_CODE_32_PRN[prn], so PRN 32 panics there, Q6) does not fail the batch: it alone is idled (`bad[c]` = 1, handled
// by the callers like a reset: state IDLE, lost = 1) and every other channel runs.
static int trk_validate(const gb_trk_channel* ch, int n, float* fs_max, std::vector<uint8_t>* bad)
{
    float m = 0.f;
    if (bad) bad->assign(n, 0);
    for (int c = 0; c < n; c++) {
        if (ch[c].code_row >= 32 && ch[c].state == GB_TRK_TRACKING) {
            if (bad) (*bad)[c] = 1;
            continue;
        }
        if (!(ch[c].fs > 0.f)) return GB_EINVAL;
        if (ch[c].fs > m) m = ch[c].fs;
    }
    *fs_max = m;
    return GB_OK;
}

static int trk_n_max(float fs_max)
{
    const int nominal = gb_num_samples_per_code(kCodeRate, fs_max);
    return nominal + nominal / 64 + 16;
}

extern "C" int gb_trk_upload(gb_handle* h, const gb_trk_channel* ch, int n)
{
    if (!h || !ch || n < 1) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_trk);
    CK(cudaSetDevice(h->device));
    float fs_max;
    int rc = trk_validate(ch, n, &fs_max, &h->trk_bad);
    if (rc) return rc;
    rc = trk_reserve(h, n);
    if (rc) return rc;
    h->trk_stage.assign(ch, ch + n);
    for (int c = 0; c < n; c++)
        if (h->trk_bad[c]) gb_trk_channel_reset(&h->trk_stage[c]);   // the offending channel idles, the others run
    ch = h->trk_stage.data();
    CK(cudaMemcpyAsync(h->ch_dev, ch, sizeof(gb_trk_channel) * n, cudaMemcpyHostToDevice, h->s_trk));
    CK(cudaMemsetAsync(h->corr_dev, 0, sizeof(gb_trk_corr) * n, h->s_trk));
    CK(cudaStreamSynchronize(h->s_trk));
    h->n_ch = n;
    h->trk_fs_max = fs_max;
    return GB_OK;
}

extern "C" int gb_trk_download(gb_handle* h, gb_trk_channel* ch, int n)
{
    if (!h || !ch || n < 1 || n > h->n_ch) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_trk);
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(ch, h->ch_dev, sizeof(gb_trk_channel) * n, cudaMemcpyDeviceToHost, h->s_trk));
    CK(cudaStreamSynchronize(h->s_trk));
    return GB_OK;
}

static int trk_launch_ring(gb_handle* h, int n_epochs, int mode, int filters, float* hist_dev)
{
    gb::TrkArgs a;
    {
        std::lock_guard<std::recursive_mutex> lr(h->mu_ring);   // the writer thread may be inside gb_ring_write
        a.samples = h->ring; a.mask = h->ring_cap - 1; a.head = h->ring_head; a.capacity = h->ring_cap;
        CK(cudaStreamWaitEvent(h->s_trk, h->ev_copy, 0));
    }
    a.offsets = nullptr;
    a.ch = h->ch_dev; a.ca_table = h->ca_table_dev;
    a.n_channels = h->n_ch; a.n_epochs = n_epochs; a.filters = filters; a.n_max = trk_n_max(h->trk_fs_max);
    a.corr = h->corr_dev; a.prompt_hist = hist_dev; a.ran = h->ran_dev; a.lost = h->lost_dev;
    if (mode == GB_TRK_ORDERED) {
        // the in-order sums keep n_max rotated samples + 3 chip rows in shared memory
        int optin = 0;
        CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
        if (gb::trk_ordered_smem_bytes(a.n_max) > (size_t)optin) return GB_EUNSUPPORTED;
    }
    CK(cudaEventRecord(h->ev_t0, h->s_trk));
    CK(gb::trk_launch(a, mode, h->s_trk));
    CK(cudaEventRecord(h->ev_t1, h->s_trk));
    return GB_OK;
}

static int trk_run_common(gb_handle* h, int n_epochs, int mode, float* prompt_hist, bool keep)
{
    if (!h || n_epochs < 1 || (mode != GB_TRK_FAST && mode != GB_TRK_ORDERED)) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_trk);
    if (!h->ring || h->n_ch == 0) return GB_ESTATE;
    CK(cudaSetDevice(h->device));
    float* hist_dev = nullptr;
    const size_t hist_n = (size_t)n_epochs * h->n_ch * 2;
    h->hist_epochs = 0; h->hist_channels = 0;
    if (prompt_hist || keep) {
        int rc = ensure(h, &h->hist_dev, &h->hist_cap, hist_n);
        if (rc) return rc;
        CK(cudaMemsetAsync(h->hist_dev, 0, hist_n * sizeof(float), h->s_trk));
        hist_dev = h->hist_dev;
    }
    int rc = trk_launch_ring(h, n_epochs, mode, 1, hist_dev);
    if (rc) return rc;
    if (prompt_hist) CK(cudaMemcpyAsync(prompt_hist, h->hist_dev, hist_n * sizeof(float), cudaMemcpyDeviceToHost, h->s_trk));
    CK(cudaStreamSynchronize(h->s_trk));
    CK(cudaEventElapsedTime(&h->last_trk_ms, h->ev_t0, h->ev_t1));
    if (hist_dev) { h->hist_epochs = n_epochs; h->hist_channels = h->n_ch; }
    return GB_OK;
}

extern "C" int gb_trk_run(gb_handle* h, int n_epochs, int mode, float* prompt_hist)
{
    return trk_run_common(h, n_epochs, mode, prompt_hist, false);
}

// same run; the prompt history stays on the device for gb_nav_bit_sync(h, NULL, ...) (no D2H -> H2D bounce)
extern "C" int gb_trk_run_keep(gb_handle* h, int n_epochs, int mode) { return trk_run_common(h, n_epochs, mode, nullptr, true); }

extern "C" int gb_trk_epoch(gb_handle* h, gb_trk_channel* ch, int n, int mode, gb_trk_corr* out, uint8_t* ran, uint8_t* lost)
{
    if (!h || !ch || n < 1 || (mode != GB_TRK_FAST && mode != GB_TRK_ORDERED)) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_trk);
    if (!h->ring) return GB_ESTATE;
    int rc = gb_trk_upload(h, ch, n);
    if (rc) return rc;
    rc = trk_launch_ring(h, 1, mode, 1, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ch, h->ch_dev, sizeof(gb_trk_channel) * n, cudaMemcpyDeviceToHost, h->s_trk));
    if (out) CK(cudaMemcpyAsync(out, h->corr_dev, sizeof(gb_trk_corr) * n, cudaMemcpyDeviceToHost, h->s_trk));
    if (ran) CK(cudaMemcpyAsync(ran, h->ran_dev, n, cudaMemcpyDeviceToHost, h->s_trk));
    if (lost) CK(cudaMemcpyAsync(lost, h->lost_dev, n, cudaMemcpyDeviceToHost, h->s_trk));
    CK(cudaStreamSynchronize(h->s_trk));
    CK(cudaEventElapsedTime(&h->last_trk_ms, h->ev_t0, h->ev_t1));
    if (lost)
        for (int c = 0; c < n; c++)
            if (h->trk_bad[c]) lost[c] = 1;   // idled by the upload (code_row out of the table): reported like SatelliteLost
    return GB_OK;
}

extern "C" int gb_trk_correlate(gb_handle* h, gb_trk_channel* ch, int n, const gb_c32* data, uint64_t n_data,
                                const uint64_t* offsets, int mode, gb_trk_corr* out)
{
    if (!h || !ch || !data || !offsets || !out || n < 1 || (mode != GB_TRK_FAST && mode != GB_TRK_ORDERED))
        return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_trk);
    // every channel's segment must lie inside the caller's buffer (the reference slices data_samples[0..n], :176)
    size_t total = 0;
    int n_max = 0;
    for (int c = 0; c < n; c++) {
        const uint64_t spc = ch[c].num_samples_per_code;
        if (spc == 0 || spc > (1u << 22)) return GB_EINVAL;
        if (offsets[c] > n_data || spc > n_data - offsets[c]) return GB_ERANGE;
        const size_t end = (size_t)offsets[c] + (size_t)spc;
        if (end > total) total = end;
        if ((int)spc > n_max) n_max = (int)spc;
    }
    for (int c = 0; c < n; c++)
        if (ch[c].code_row >= 32) return GB_EINVAL;   // open-loop call: nothing to idle, the row does not exist (Q6)
    int rc = gb_trk_upload(h, ch, n);
    if (rc) return rc;
    if (mode == GB_TRK_ORDERED) {
        int optin = 0;
        CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
        if (gb::trk_ordered_smem_bytes(n_max) > (size_t)optin) return GB_EUNSUPPORTED;
    }
    rc = ensure(h, &h->trk_data, &h->trk_data_cap, total);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->trk_data, data, total * sizeof(float2), cudaMemcpyHostToDevice, h->s_trk));
    CK(cudaMemcpyAsync(h->offs_dev, offsets, sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, h->s_trk));
    gb::TrkArgs a;
    a.samples = h->trk_data; a.mask = ~0ull; a.head = total; a.capacity = 0;
    a.offsets = h->offs_dev;
    a.ch = h->ch_dev; a.ca_table = h->ca_table_dev;
    a.n_channels = n; a.n_epochs = 1; a.filters = 0;
    a.n_max = n_max;
    a.corr = h->corr_dev; a.prompt_hist = nullptr; a.ran = h->ran_dev; a.lost = h->lost_dev;
    CK(cudaEventRecord(h->ev_t0, h->s_trk));
    CK(gb::trk_launch(a, mode, h->s_trk));
    CK(cudaEventRecord(h->ev_t1, h->s_trk));
    CK(cudaMemcpyAsync(ch, h->ch_dev, sizeof(gb_trk_channel) * n, cudaMemcpyDeviceToHost, h->s_trk));
    CK(cudaMemcpyAsync(out, h->corr_dev, sizeof(gb_trk_corr) * n, cudaMemcpyDeviceToHost, h->s_trk));
    CK(cudaStreamSynchronize(h->s_trk));
    CK(cudaEventElapsedTime(&h->last_trk_ms, h->ev_t0, h->ev_t1));
    return GB_OK;
}

extern "C" float gb_trk_last_kernel_ms(gb_handle* h) { return h ? h->last_trk_ms : 0.f; }

// N4: bit sync + nav-bit accumulation over a prompt history [n_epochs][n_channels][2] (host memory)
extern "C" int gb_nav_bit_sync(gb_handle* h, const float* prompt_hist, int n_epochs, int n_channels, gb_nav_sync* out,
                               int8_t* bits, int max_bits)
{
    if (!h || !out || !bits || n_epochs < 1 || n_channels < 1 || max_bits < 1) return GB_EINVAL;
    std::lock_guard<std::recursive_mutex> lk(h->mu_trk);
    CK(cudaSetDevice(h->device));
    const size_t hist_n = (size_t)n_epochs * n_channels * 2;
    if (!prompt_hist) {
        // the history gb_trk_run_keep / gb_trk_run left on the device
        if (!h->hist_dev || h->hist_epochs != n_epochs || h->hist_channels != n_channels) return GB_ESTATE;
    } else {
        int rc = ensure(h, &h->hist_dev, &h->hist_cap, hist_n);
        if (rc) return rc;
        h->hist_epochs = 0; h->hist_channels = 0;
    }
    gb_nav_sync* st_dev = nullptr;
    int8_t* bits_dev = nullptr;
    CK(cudaMalloc((void**)&st_dev, sizeof(gb_nav_sync) * n_channels));
    cudaError_t e = cudaMalloc((void**)&bits_dev, (size_t)n_channels * max_bits);
    if (e != cudaSuccess) { cudaFree(st_dev); return fail(h, e, "cudaMalloc"); }
    if (prompt_hist) e = cudaMemcpyAsync(h->hist_dev, prompt_hist, hist_n * sizeof(float), cudaMemcpyHostToDevice, h->s_trk);
    if (e == cudaSuccess) e = cudaMemsetAsync(bits_dev, 0, (size_t)n_channels * max_bits, h->s_trk);
    if (e == cudaSuccess) {
        nav_bit_sync_kernel<<<(n_channels + 63) / 64, 64, 0, h->s_trk>>>(h->hist_dev, n_epochs, n_channels, st_dev, bits_dev, max_bits);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, st_dev, sizeof(gb_nav_sync) * n_channels, cudaMemcpyDeviceToHost, h->s_trk);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bits, bits_dev, (size_t)n_channels * max_bits, cudaMemcpyDeviceToHost, h->s_trk);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_trk);
    cudaFree(st_dev);
    cudaFree(bits_dev);
    if (e != cudaSuccess) return fail(h, e, "nav_bit_sync");
    return GB_OK;
}

// ------------------------------------------------------------------ multi-GPU: sharding + the one collective
// The units of both paths are independent (PRNs / recordings / channels), so ranks never exchange samples or spectra.
// What the host needs from the library is (a) the partition and (b) the final gather of the per-PRN result tables --
// "a final NCCL gather of per-PRN peaks over NVLink".  One process per GPU: rank 0 makes a unique id, the host ships its
// 128 bytes to the peers by whatever transport it already has (the bench uses torch.distributed's store, a Rust host
// would use its own channel), every rank calls gb_group_init.  NCCL is loaded with dlopen at that moment: the library has
// no link-time NCCL dependency and single-GPU users never need it.
#include <dlfcn.h>

namespace {

struct NcclId {
    char internal[128];
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
std::mutex g_nccl_mu;
NcclApi g_nccl;

const NcclApi* nccl_api()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok) return &g_nccl;
    if (!g_nccl.lib) {
        // a process that already carries NCCL (PyTorch bundles one) resolves to that copy by SONAME
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            g_nccl.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
    }
    if (!g_nccl.lib) return nullptr;
    g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.CommDestroy;
    return g_nccl.ok ? &g_nccl : nullptr;
}

}  // namespace

struct gb_group {
    gb_handle* h = nullptr;
    void* comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t s = nullptr;            // its own stream: a gather overlaps the next search
    uint8_t *send_dev = nullptr, *recv_dev = nullptr;
    size_t send_cap = 0, recv_cap = 0;
    uint8_t* send_pin[2] = {nullptr, nullptr};
    uint8_t* recv_pin[2] = {nullptr, nullptr};
    size_t pin_cap[2] = {0, 0};
    size_t bytes[2] = {0, 0};
    bool active[2] = {false, false};
    cudaEvent_t done[2] = {nullptr, nullptr};
    std::mutex mu;
};

extern "C" uint32_t gb_shard_prn_mask(int rank, int world, int n_prn, uint32_t base_mask)
{
    // bit (prn - 1) set for the PRNs this rank searches: the PRNs selected by base_mask dealt round-robin
    // (the reference's mask convention, do_acquisition.rs:307)
    if (world < 1 || rank < 0 || rank >= world || n_prn < 1) return 0;
    uint32_t mask = 0;
    int k = 0;
    for (int p = 0; p < n_prn && p < 32; p++)
        if ((base_mask >> p) & 1u) {
            if (k % world == rank) mask |= 1u << p;
            k++;
        }
    return mask;
}

extern "C" int gb_shard_range(int n_items, int rank, int world, int* first, int* count)
{
    // contiguous block partition of recordings / channels
    if (n_items < 0 || world < 1 || rank < 0 || rank >= world || !first || !count) return GB_EINVAL;
    const int per = (n_items + world - 1) / world;
    const int lo = std::min(n_items, rank * per);
    *first = lo;
    *count = std::min(n_items, lo + per) - lo;
    return GB_OK;
}

extern "C" int gb_group_unique_id(uint8_t* id128)
{
    if (!id128) return GB_EINVAL;
    const NcclApi* api = nccl_api();
    if (!api) return GB_ENCCL;
    NcclId id;
    if (api->GetUniqueId(&id) != 0) return GB_ENCCL;
    memcpy(id128, id.internal, 128);
    return GB_OK;
}

extern "C" int gb_group_allgather(gb_group* g, const void* mine, uint64_t bytes, void* all_out);
extern "C" int gb_group_destroy(gb_group* g);

extern "C" int gb_group_init(gb_handle* h, const uint8_t* id128, int rank, int world, gb_group** out)
{
    if (!h || !id128 || !out || world < 1 || rank < 0 || rank >= world) return GB_EINVAL;
    const NcclApi* api = nccl_api();
    if (!api) return GB_ENCCL;
    CK(cudaSetDevice(h->device));
    gb_group* g = new gb_group();
    g->h = h; g->rank = rank; g->world = world;
    NcclId id;
    memcpy(id.internal, id128, 128);
    const int rc = api->CommInitRank(&g->comm, world, id, rank);
    if (rc != 0) {
        h->last_err = std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error");
        delete g;
        return GB_ENCCL;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&g->s, cudaStreamNonBlocking);
    for (int s = 0; s < 2 && e == cudaSuccess; s++) e = cudaEventCreateWithFlags(&g->done[s], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        api->CommDestroy(g->comm);
        delete g;
        return fail(h, e, "gb_group_init");
    }
    *out = g;
    // NCCL builds its channels during the first collective (hundreds of milliseconds): pay that here, not inside the
    // caller's first timed gather
    uint64_t probe = (uint64_t)rank;
    std::vector<uint64_t> all((size_t)world);
    const int wrc = gb_group_allgather(g, &probe, sizeof(probe), all.data());
    if (wrc) {
        gb_group_destroy(g);
        *out = nullptr;
        return wrc;
    }
    for (int r = 0; r < world; r++)
        if (all[r] != (uint64_t)r) {
            gb_group_destroy(g);
            *out = nullptr;
            return GB_ENCCL;
        }
    return GB_OK;
}

extern "C" int gb_group_rank(gb_group* g) { return g ? g->rank : GB_EINVAL; }
extern "C" int gb_group_world(gb_group* g) { return g ? g->world : GB_EINVAL; }

// every rank contributes `bytes` bytes of host memory; after the matching _end, all_out holds world x bytes, rank-major,
// on every rank.  H2D from pinned staging, ONE ncclAllGather on the group's stream, D2H -- nothing waits until _end.
extern "C" int gb_group_allgather_begin(gb_group* g, const void* mine, uint64_t bytes, int slot)
{
    if (!g || !mine || bytes == 0 || slot < 0 || slot > 1) return GB_EINVAL;
    std::lock_guard<std::mutex> lk(g->mu);
    gb_handle* h = g->h;
    if (g->active[slot]) return GB_ESTATE;
    const NcclApi* api = nccl_api();
    if (!api) return GB_ENCCL;
    CK(cudaSetDevice(h->device));
    const size_t total = (size_t)bytes * g->world;
    if (g->pin_cap[slot] < total) {
        if (g->send_pin[slot]) cudaFreeHost(g->send_pin[slot]);
        if (g->recv_pin[slot]) cudaFreeHost(g->recv_pin[slot]);
        g->send_pin[slot] = g->recv_pin[slot] = nullptr; g->pin_cap[slot] = 0;
        CK(cudaMallocHost((void**)&g->send_pin[slot], bytes));
        CK(cudaMallocHost((void**)&g->recv_pin[slot], total));
        g->pin_cap[slot] = total;
    }
    if (g->send_cap < bytes || g->recv_cap < total) {
        CK(cudaStreamSynchronize(g->s));
        if (g->send_dev) cudaFree(g->send_dev);
        if (g->recv_dev) cudaFree(g->recv_dev);
        g->send_dev = g->recv_dev = nullptr; g->send_cap = g->recv_cap = 0;
        CK(cudaMalloc((void**)&g->send_dev, bytes));
        CK(cudaMalloc((void**)&g->recv_dev, total));
        g->send_cap = bytes; g->recv_cap = total;
    }
    memcpy(g->send_pin[slot], mine, bytes);
    CK(cudaMemcpyAsync(g->send_dev, g->send_pin[slot], bytes, cudaMemcpyHostToDevice, g->s));
    const int rc = api->AllGather(g->send_dev, g->recv_dev, bytes, /* ncclUint8 */ 1, g->comm, g->s);
    if (rc != 0) {
        h->last_err = std::string("ncclAllGather: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error");
        return GB_ENCCL;
    }
    CK(cudaMemcpyAsync(g->recv_pin[slot], g->recv_dev, total, cudaMemcpyDeviceToHost, g->s));
    CK(cudaEventRecord(g->done[slot], g->s));
    g->bytes[slot] = bytes;
    g->active[slot] = true;
    return GB_OK;
}

extern "C" int gb_group_allgather_end(gb_group* g, int slot, void* all_out)
{
    if (!g || !all_out || slot < 0 || slot > 1) return GB_EINVAL;
    std::lock_guard<std::mutex> lk(g->mu);
    gb_handle* h = g->h;
    if (!g->active[slot]) return GB_ESTATE;
    g->active[slot] = false;
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(g->done[slot]));
    memcpy(all_out, g->recv_pin[slot], g->bytes[slot] * g->world);
    return GB_OK;
}

extern "C" int gb_group_allgather(gb_group* g, const void* mine, uint64_t bytes, void* all_out)
{
    int rc = gb_group_allgather_begin(g, mine, bytes, 0);
    if (rc) return rc;
    return gb_group_allgather_end(g, 0, all_out);
}

// the final gather of the path: n result structs per rank -> world x n on every rank
extern "C" int gb_group_gather_results(gb_group* g, const gb_acq_result* mine, int n, gb_acq_result* all)
{
    if (n < 1) return GB_EINVAL;
    return gb_group_allgather(g, mine, (uint64_t)n * sizeof(gb_acq_result), all);
}

extern "C" int gb_group_destroy(gb_group* g)
{
    if (!g) return GB_EINVAL;
    cudaSetDevice(g->h->device);
    if (g->s) cudaStreamSynchronize(g->s);
    const NcclApi* api = nccl_api();
    if (api && g->comm) api->CommDestroy(g->comm);
    if (g->send_dev) cudaFree(g->send_dev);
    if (g->recv_dev) cudaFree(g->recv_dev);
    for (int s = 0; s < 2; s++) {
        if (g->send_pin[s]) cudaFreeHost(g->send_pin[s]);
        if (g->recv_pin[s]) cudaFreeHost(g->recv_pin[s]);
        if (g->done[s]) cudaEventDestroy(g->done[s]);
    }
    if (g->s) cudaStreamDestroy(g->s);
    cudaGetLastError();
    delete g;
    return GB_OK;
}
