/*
 * gnss_oracle.c -- CPU restatement of the gnss-sdr-rs acquisition / tracking hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see gnss_oracle.h for the parity status and the rules on who may
 * call this).  Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared ... -lm -lpthread
 * (oracle/Makefile).  f32 arithmetic follows the reference's operation order; libm calls are the
 * glibc cosf/sinf/atanf/fmodf/floorf/roundf/sqrtf/powf that Rust's std uses on linux-gnu.
 */
#define _GNU_SOURCE
#include "gnss_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define GO_MAX_PRIME 128
#define GO_PI_F 3.14159265358979323846f /* std::f32::consts::PI */

/* ------------------------------------------------------------------ FFT */
#define REAL float
#define CPX go_c32
#define PLAN go_fft_plan
#define FN(x) f32_##x
#include "fft_body.inc"
#undef REAL
#undef CPX
#undef PLAN
#undef FN

#define REAL double
#define CPX go_c64
#define PLAN go_fft64_plan
#define FN(x) f64_##x
#include "fft_body.inc"
#undef REAL
#undef CPX
#undef PLAN
#undef FN

go_fft_plan *go_fft_plan_new(int n, int inverse) { return f32_plan_new(n, inverse); }
void go_fft_plan_free(go_fft_plan *p) { f32_plan_free(p); }
void go_fft_process(const go_fft_plan *p, go_c32 *data) { f32_process(p, data); }
go_fft64_plan *go_fft64_plan_new(int n, int inverse) { return f64_plan_new(n, inverse); }
void go_fft64_plan_free(go_fft64_plan *p) { f64_plan_free(p); }
void go_fft64_process(const go_fft64_plan *p, go_c64 *data) { f64_process(p, data); }

/* fft.rs:20-25 FFT::execute (forward, in place) */
void go_fft_forward(int n, go_c32 *data)
{
    go_fft_plan *p = go_fft_plan_new(n, 0);
    go_fft_process(p, data);
    go_fft_plan_free(p);
}
/* fft.rs:27-29 */
void go_fft_power_spectrum(int n, go_c32 *data, float *out)
{
    go_fft_forward(n, data);
    for (int i = 0; i < n; i++) out[i] = data[i].re * data[i].re + data[i].im * data[i].im;
}
/* fft.rs:47-51 RealFFT::execute: n real -> n/2+1 complex */
void go_rfft_forward(int n, const float *in, go_c32 *out)
{
    go_c32 *tmp = (go_c32 *)malloc(sizeof(go_c32) * n);
    for (int i = 0; i < n; i++) { tmp[i].re = in[i]; tmp[i].im = 0.0f; }
    go_fft_forward(n, tmp);
    memcpy(out, tmp, sizeof(go_c32) * (n / 2 + 1));
    free(tmp);
}

/* ------------------------------------------------------------------ C/A code */
/* G2 phase-selector tap pairs (IS-GPS-200, table 3-Ia) for PRN 1..32. */
static const uint8_t G2_TAPS[32][2] = {
    {2, 6}, {3, 7}, {4, 8}, {5, 9}, {1, 9}, {2, 10}, {1, 8}, {2, 9}, {3, 10}, {2, 3}, {3, 4},
    {5, 6}, {6, 7}, {7, 8}, {8, 9}, {9, 10}, {1, 4}, {2, 5}, {3, 6}, {4, 7}, {5, 8}, {6, 9},
    {1, 3}, {4, 6}, {5, 7}, {6, 8}, {7, 9}, {8, 10}, {1, 6}, {2, 7}, {3, 8}, {4, 9}};

int go_ca_code_chips(int prn, int8_t out[GO_CA_CODE_LEN])
{
    if (prn < 1 || prn > 32) return -1;
    int g1[11], g2[11]; /* stages 1..10 */
    for (int i = 1; i <= 10; i++) g1[i] = g2[i] = 1;
    const int ta = G2_TAPS[prn - 1][0], tb = G2_TAPS[prn - 1][1];
    for (int c = 0; c < GO_CA_CODE_LEN; c++) {
        const int bit = g1[10] ^ g2[ta] ^ g2[tb];
        out[c] = bit ? 1 : -1;
        const int f1 = g1[3] ^ g1[10];
        const int f2 = g2[2] ^ g2[3] ^ g2[6] ^ g2[8] ^ g2[9] ^ g2[10];
        for (int i = 10; i > 1; i--) { g1[i] = g1[i - 1]; g2[i] = g2[i - 1]; }
        g1[1] = f1; g2[1] = f2;
    }
    return 0;
}

static int8_t g_ca_table[32 * GO_CA_CODE_LEN];
static pthread_once_t g_ca_once = PTHREAD_ONCE_INIT;
static void ca_table_init(void)
{
    for (int p = 1; p <= 32; p++) go_ca_code_chips(p, g_ca_table + (p - 1) * GO_CA_CODE_LEN);
}
const int8_t *go_ca_table(void)
{
    pthread_once(&g_ca_once, ca_table_init);
    return g_ca_table;
}

/* Rust `as usize` on f32: saturating, NaN -> 0 */
static inline size_t f32_as_usize(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return (size_t)-1;
    return (size_t)v;
}

/* ca_code.rs:13-17 */
int go_num_samples_per_code(float code_rate, float fs)
{
    return (int)f32_as_usize(roundf(fs / (code_rate / 1023.0f)));
}

/* ca_code.rs:12-27 (Q3: index math stays in f32, in this order) */
int go_generate_ca_code_samples(int prn, float code_rate, float fs, int8_t *out, int cap)
{
    const int n = go_num_samples_per_code(code_rate, fs);
    const int8_t *code = go_ca_table() + (size_t)(prn - 1) * GO_CA_CODE_LEN;
    for (int x = 0; x < n && x < cap; x++) {
        const size_t ind = f32_as_usize(floorf((float)x * code_rate / fs));
        out[x] = code[ind]; /* the reference would panic for ind >= 1023 */
    }
    return n;
}

/* ------------------------------------------------------------------ Doppler wipe-off */
/* doppler_shift.rs:11-21 */
float go_doppler_table(float f_if, float f_d, float fs, int n, go_c32 *table)
{
    const float carr = f_if + f_d;
    const float step = 2.0f * GO_PI_F * carr / fs;
    for (int i = 0; i < n; i++) {
        const float phase = (float)i * step;
        table[i].re = cosf(phase);
        table[i].im = -sinf(phase);
    }
    return carr;
}

/* doppler_shift.rs:25-58: four complex per f32x8; first_part + second_part, lane-wise:
 * re = a*c + ((b*d)*-1), im = a*d + ((b*c)*1); the last len%4 outputs are not written. */
void go_apply_doppler_shift(const go_c32 *s, const go_c32 *t, go_c32 *out, int len)
{
    const int chunks = len / 4;
    for (int i = 0; i < chunks * 4; i++) {
        const float a = s[i].re, b = s[i].im, c = t[i].re, d = t[i].im;
        out[i].re = a * c + (b * d) * -1.0f;
        out[i].im = a * d + (b * c) * 1.0f;
    }
}

/* ------------------------------------------------------------------ acquisition worker */
struct go_acq_worker {
    int prn, n;
    float fs;
    go_fft_plan *fft, *ifft;
    go_c32 *code_fft;   /* ca_code_samples_fft */
    go_c32 *result_buf; /* persists across calls like the reference's (stale tail, A3) */
    go_c32 *coh;        /* n, extension */
    float *acc, *best;
};

go_acq_worker *go_acq_worker_new_code(int prn, int fft_size, float fs, const int8_t *code_samples)
{
    go_acq_worker *w = (go_acq_worker *)calloc(1, sizeof(*w));
    w->prn = prn; w->n = fft_size; w->fs = fs;
    w->fft = go_fft_plan_new(fft_size, 0);
    w->ifft = go_fft_plan_new(fft_size, 1);
    w->code_fft = (go_c32 *)calloc(fft_size, sizeof(go_c32));
    w->result_buf = (go_c32 *)calloc(fft_size, sizeof(go_c32));
    w->coh = (go_c32 *)calloc(fft_size, sizeof(go_c32));
    w->acc = (float *)calloc(fft_size, sizeof(float));
    w->best = (float *)calloc(fft_size, sizeof(float));
    for (int i = 0; i < fft_size; i++) { w->code_fft[i].re = (float)code_samples[i]; w->code_fft[i].im = 0.0f; }
    go_fft_process(w->fft, w->code_fft);
    return w;
}

/* do_acquisition.rs:131-156.  The reference panics inside rustfft if the resampled code length
 * differs from fft_size; here the shorter of the two is used and the rest is zero. */
go_acq_worker *go_acq_worker_new(int prn, int fft_size, float fs)
{
    const int n_code = go_num_samples_per_code(GO_CA_CODE_RATE, fs);
    int8_t *code = (int8_t *)calloc((size_t)(n_code > fft_size ? n_code : fft_size) + 1, 1);
    go_generate_ca_code_samples(prn, GO_CA_CODE_RATE, fs, code, n_code);
    go_acq_worker *w = go_acq_worker_new_code(prn, fft_size, fs, code);
    free(code);
    return w;
}

void go_acq_worker_free(go_acq_worker *w)
{
    if (!w) return;
    go_fft_plan_free(w->fft); go_fft_plan_free(w->ifft);
    free(w->code_fft); free(w->result_buf); free(w->coh); free(w->acc); free(w->best);
    free(w);
}
const go_c32 *go_acq_worker_code_fft(const go_acq_worker *w) { return w->code_fft; }

void go_coh_rotators(float carr, float fs, int n, int n_coh, go_c32 *rot)
{
    for (int c = 0; c < n_coh; c++) {
        const double cyc = (double)carr * (double)c * (double)n / (double)fs;
        const double ang = -2.0 * M_PI * (cyc - floor(cyc));
        rot[c].re = (float)cos(ang);
        rot[c].im = (float)sin(ang);
    }
}

/* one 1 ms block: wipe-off -> FFT -> x conj(code) -> IFFT, left in w->result_buf
 * (do_acquisition.rs:176-188) */
static void acq_correlate_block(go_acq_worker *w, const go_c32 *chunk, const go_c32 *table)
{
    const int n = w->n;
    go_apply_doppler_shift(chunk, table, w->result_buf, n);
    go_fft_process(w->fft, w->result_buf);
    for (int i = 0; i < n; i++) {
        const go_c32 x = w->result_buf[i];
        const float cr = w->code_fft[i].re, ci = -w->code_fft[i].im; /* conj() */
        w->result_buf[i].re = x.re * cr - x.im * ci;
        w->result_buf[i].im = x.re * ci + x.im * cr;
    }
    go_fft_process(w->ifft, w->result_buf);
}

void go_acq_bin_power(go_acq_worker *w, const go_c32 *samples, const go_c32 *table, int K, int n_coh,
                      const go_c32 *rot, int presum, float *acc)
{
    const int n = w->n;
    for (int i = 0; i < n; i++) acc[i] = 0.0f; /* accumulated_power.fill(0.0), :172 */
    if (n_coh <= 1) {
        for (int c = 0; c < K; c++) { /* :174-193 */
            acq_correlate_block(w, samples + (size_t)c * n, table);
            for (int i = 0; i < n; i++) {
                const go_c32 v = w->result_buf[i];
                acc[i] += v.re * v.re + v.im * v.im; /* norm_sqr */
            }
        }
        return;
    }
    /* EXTENSION (not in the reference): n_coh-block coherent sums, then non-coherent */
    for (int g = 0; g + n_coh <= K; g += n_coh) {
        if (presum) {
            go_c32 *s = w->coh;
            for (int i = 0; i < n; i++) { s[i].re = 0.0f; s[i].im = 0.0f; }
            for (int c = 0; c < n_coh; c++) {
                const go_c32 *x = samples + (size_t)(g + c) * n;
                for (int i = 0; i < n; i++) {
                    const float a = x[i].re, b = x[i].im, tc = table[i].re, td = table[i].im;
                    const float re = a * tc - b * td, im = a * td + b * tc;
                    s[i].re += re * rot[c].re - im * rot[c].im;
                    s[i].im += re * rot[c].im + im * rot[c].re;
                }
            }
            /* reuse the block pipeline with an all-ones table: feed s directly */
            memcpy(w->result_buf, s, sizeof(go_c32) * n);
            go_fft_process(w->fft, w->result_buf);
            for (int i = 0; i < n; i++) {
                const go_c32 x = w->result_buf[i];
                const float cr = w->code_fft[i].re, ci = -w->code_fft[i].im;
                w->result_buf[i].re = x.re * cr - x.im * ci;
                w->result_buf[i].im = x.re * ci + x.im * cr;
            }
            go_fft_process(w->ifft, w->result_buf);
            for (int i = 0; i < n; i++) {
                const go_c32 v = w->result_buf[i];
                acc[i] += v.re * v.re + v.im * v.im;
            }
        } else {
            go_c32 *z = w->coh;
            for (int i = 0; i < n; i++) { z[i].re = 0.0f; z[i].im = 0.0f; }
            for (int c = 0; c < n_coh; c++) {
                acq_correlate_block(w, samples + (size_t)(g + c) * n, table);
                for (int i = 0; i < n; i++) {
                    const go_c32 v = w->result_buf[i];
                    z[i].re += v.re * rot[c].re - v.im * rot[c].im;
                    z[i].im += v.re * rot[c].im + v.im * rot[c].re;
                }
            }
            for (int i = 0; i < n; i++) acc[i] += z[i].re * z[i].re + z[i].im * z[i].im;
        }
    }
}

/* do_acquisition.rs:195-202: strict '>' from 0.0 => first index of the maximum */
static void acq_argmax(const float *acc, int n, float *peak, uint32_t *arg)
{
    float local_max = 0.0f;
    uint32_t best = 0;
    for (int i = 0; i < n; i++)
        if (acc[i] > local_max) { local_max = acc[i]; best = (uint32_t)i; }
    *peak = local_max;
    *arg = best;
}

/* do_acquisition.rs:229-234 (Q2): f32x8 lane accumulators over chunks_exact(8), then
 * reduce_sum (ordered, lane 0..7) */
static float acq_sum8(const float *p, int n)
{
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c + 8 <= n; c += 8)
        for (int j = 0; j < 8; j++) lane[j] = lane[j] + p[c + j];
    float s = lane[0];
    for (int j = 1; j < 8; j++) s = s + lane[j];
    return s;
}

/* do_acquisition.rs:235-237 */
int go_is_good_cell(float peak, float sum8, int fft_size, float threshold)
{
    const float avg = (sum8 - peak) / (float)(fft_size - 1);
    return (peak / avg > threshold) ? 1 : 0;
}

static void fill_result(go_acq_result *out, int prn, uint32_t phase, float carr, float fs, float peak,
                        uint64_t local_tail)
{
    out->prn = (uint8_t)prn;
    out->code_phase_samples = phase;
    out->code_phase_chips = (float)phase * GO_CA_CODE_RATE / fs; /* :213-214 */
    out->carrier_freq = carr;
    out->fs = fs;
    out->mag_relative = peak;
    out->sample_global_index = local_tail + phase;
}

/* do_acquisition.rs:158-226 verbatim control flow */
int go_acq_search(go_acq_worker *w, const go_c32 *samples, const go_c32 *tables, const float *carr, int D,
                  uint64_t local_tail, int K, go_acq_result *out)
{
    const int n = w->n;
    float global_max = 0.0f, best_freq = 0.0f;
    uint32_t best_phase = 0;
    for (int i = 0; i < n; i++) w->best[i] = 0.0f;
    for (int d = 0; d < D; d++) {
        go_acq_bin_power(w, samples, tables + (size_t)d * n, K, 1, NULL, 0, w->acc);
        float local_max;
        uint32_t local_phase;
        acq_argmax(w->acc, n, &local_max, &local_phase);
        if (local_max > global_max) {
            global_max = local_max;
            best_freq = carr[d];
            best_phase = local_phase;
            memcpy(w->best, w->acc, sizeof(float) * n);
        }
        if (go_is_good_cell(global_max, acq_sum8(w->best, n), n, 7.0f)) {
            fill_result(out, w->prn, best_phase, best_freq, w->fs, global_max, local_tail);
            return 1;
        }
    }
    return 0;
}

void go_acq_cells(go_acq_worker *w, const go_c32 *samples, const go_c32 *tables, int D, int K, int n_coh,
                  const go_c32 *rot, int presum, go_acq_cell *cells)
{
    const int n = w->n;
    for (int d = 0; d < D; d++) {
        go_acq_bin_power(w, samples, tables + (size_t)d * n, K, n_coh, rot ? rot + (size_t)d * n_coh : NULL,
                         presum, w->acc);
        acq_argmax(w->acc, n, &cells[d].peak, &cells[d].argmax);
        cells[d].sum8 = acq_sum8(w->acc, n);
    }
}

/* Q1: the early-exit loop of search_satellite expressed on per-bin cells: keep the prefix maximum
 * (strict '>'); the answer is the first bin at which the running best passes the test. */
int go_acq_decide(const go_acq_cell *cells, const float *carr, int D, int prn, int fft_size, float fs,
                  uint64_t local_tail, float threshold, go_acq_result *out, int *bin_out)
{
    float gmax = 0.0f, gsum = 0.0f, gfreq = 0.0f;
    uint32_t gphase = 0;
    int gbin = -1;
    for (int d = 0; d < D; d++) {
        if (cells[d].peak > gmax) {
            gmax = cells[d].peak; gsum = cells[d].sum8; gphase = cells[d].argmax; gfreq = carr[d]; gbin = d;
        }
        /* before any record, best_power_results is all zeros: sum 0, max 0 -> NaN -> false */
        if (go_is_good_cell(gmax, gsum, fft_size, threshold)) {
            fill_result(out, prn, gphase, gfreq, fs, gmax, local_tail);
            if (bin_out) *bin_out = gbin;
            return 1;
        }
    }
    return 0;
}

/* acquisition_bk.rs:342-399 restated for one power row (the legacy works on magnitudes; the ordering is the same):
 * first peak = global maximum, second peak = maximum over the legacy's slices (:371-390), bounds verbatim:
 *   left = cp - spc, right = cp + spc
 *   left < 1   : row[right-1 .. N+left)
 *   right >= N : row[right-N-1 .. left)     (the lower bound underflows usize when right == N; clamped to 0 here)
 *   otherwise  : row[0 .. left) ++ row[right .. N)
 * returns first/second as amplitudes; the threshold (1.4) is applied by the caller */
float go_two_peak_ratio(const float *power, int n, int spc, uint32_t *first, uint32_t *second)
{
    float p1, p2 = 0.0f;
    uint32_t i1, i2 = 0;
    acq_argmax(power, n, &p1, &i1);
    const int left = (int)i1 - spc, right = (int)i1 + spc;
    for (int i = 0; i < n; i++) {
        int searched;
        if (left < 1) searched = i >= right - 1 && i < n + left;
        else if (right >= n) searched = i >= (right - n - 1 < 0 ? 0 : right - n - 1) && i < left;
        else searched = i < left || i >= right;
        if (!searched) continue;
        if (power[i] > p2) { p2 = power[i]; i2 = (uint32_t)i; }
    }
    if (first) *first = i1;
    if (second) *second = i2;
    return sqrtf(p1) / sqrtf(p2);
}

/* ------------------------------------------------------------------ threaded drivers */
typedef struct {
    go_acq_worker **workers; int n_workers; const go_c32 *samples, *tables, *rot; const float *carr;
    int D, K, n_coh, presum, early_exit; uint64_t local_tail; int *found; go_acq_result *results;
    go_acq_cell *cells; int next; pthread_mutex_t mu;
} acq_job;

static void *acq_thread(void *arg)
{
    acq_job *j = (acq_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int p = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (p >= j->n_workers) break;
        go_acq_worker *w = j->workers[p];
        if (j->cells && !j->found) {
            go_acq_cells(w, j->samples, j->tables, j->D, j->K, j->n_coh, j->rot, j->presum,
                         j->cells + (size_t)p * j->D);
        } else if (j->early_exit) {
            j->found[p] = go_acq_search(w, j->samples, j->tables, j->carr, j->D, j->local_tail, j->K, &j->results[p]);
        } else {
            go_acq_cell *cells = (go_acq_cell *)malloc(sizeof(go_acq_cell) * j->D);
            go_acq_cells(w, j->samples, j->tables, j->D, j->K, 1, NULL, 0, cells);
            j->found[p] = go_acq_decide(cells, j->carr, j->D, w->prn, w->n, w->fs, j->local_tail, 7.0f,
                                        &j->results[p], NULL);
            free(cells);
        }
    }
    return NULL;
}

static void run_threads(void *(*fn)(void *), void *arg, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, fn, arg);
    fn(arg);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
}

void go_acq_search_all(go_acq_worker **workers, int n_workers, const go_c32 *samples, const go_c32 *tables,
                       const float *carr, int D, uint64_t local_tail, int K, int early_exit, int n_threads,
                       int *found, go_acq_result *results)
{
    acq_job j;
    memset(&j, 0, sizeof(j));
    j.workers = workers; j.n_workers = n_workers; j.samples = samples; j.tables = tables; j.carr = carr;
    j.D = D; j.K = K; j.n_coh = 1; j.early_exit = early_exit; j.local_tail = local_tail;
    j.found = found; j.results = results;
    pthread_mutex_init(&j.mu, NULL);
    run_threads(acq_thread, &j, n_threads);
    pthread_mutex_destroy(&j.mu);
}

void go_acq_cells_all(go_acq_worker **workers, int n_workers, const go_c32 *samples, const go_c32 *tables, int D,
                      int K, int n_coh, const go_c32 *rot, int presum, int n_threads, go_acq_cell *cells)
{
    acq_job j;
    memset(&j, 0, sizeof(j));
    j.workers = workers; j.n_workers = n_workers; j.samples = samples; j.tables = tables; j.rot = rot;
    j.D = D; j.K = K; j.n_coh = n_coh; j.presum = presum; j.cells = cells;
    pthread_mutex_init(&j.mu, NULL);
    run_threads(acq_thread, &j, n_threads);
    pthread_mutex_destroy(&j.mu);
}

/* ------------------------------------------------------------------ acquisition manager */
/* do_acquisition.rs:50-56 */
void go_acq_manager_update_mode(go_acq_manager *m, size_t tracked)
{
    m->mode = tracked == 0 ? 0 : (tracked <= 4 ? 1 : 2);
}
/* do_acquisition.rs:58-73 */
void go_acq_manager_pacing(const go_acq_manager *m, uint32_t active_mask, uint64_t *interval_ms, uint32_t *mask)
{
    static const uint64_t interval[3] = {500, 1000, 2000};
    static const int size[3] = {32, 8, 5};
    uint32_t out = 0;
    int taken = 0;
    for (int prn = 1; prn <= 32 && taken < size[m->mode]; prn++) {
        if ((active_mask >> (prn - 1)) & 1u) continue;
        out |= 1u << (prn - 1);
        taken++;
    }
    *interval_ms = interval[m->mode];
    *mask = out;
}

/* ------------------------------------------------------------------ tracking */
/* do_tracking.rs:59-64 */
go_loop_filter go_loop_filter_new(float noise_bw, float damping, float gain)
{
    go_loop_filter f;
    const float w = noise_bw * 8.0f * damping / (4.0f * powf(damping, 2.0f) + 1.0f);
    f.tau1 = gain / (w * w);
    f.tau2 = (2.0f * damping) / w;
    return f;
}
/* do_tracking.rs:67-70 */
float go_loop_filter_update(const go_loop_filter *f, float d_err, float err, float dt)
{
    return d_err * (dt / f->tau1) + (d_err - err) * (f->tau2 / f->tau1);
}

/* do_tracking.rs:118-146 (constants :16-29) */
void go_trk_channel_init(go_trk_channel *c, uint8_t id, float fs)
{
    memset(c, 0, sizeof(*c));
    c->id = id;
    c->state = GO_IDLE;
    c->fs = fs;
    c->num_samples_per_code = (uint64_t)go_num_samples_per_code(GO_CA_CODE_RATE, fs);
    c->code_rate = GO_CA_CODE_RATE;
    c->pll_filter = go_loop_filter_new(25.0f, 0.7f, 0.25f);
    c->dll_filter = go_loop_filter_new(2.0f, 0.7f, 1.0f);
}

/* do_tracking.rs:148-154 (Q8) */
void go_trk_channel_start(go_trk_channel *c, const go_acq_result *r)
{
    c->prn = r->prn;
    c->code_row = r->prn; /* Q6: get_ca_chip indexes GPS_CA_CODE_32_PRN[prn], not [prn-1] */
    c->carrier_freq = r->carrier_freq;
    c->code_phase = r->code_phase_chips;
    c->next_sample_index = r->sample_global_index;
    c->state = GO_TRACKING;
}

/* do_tracking.rs:311-326 (Q9: code_rate reset to 0.0) */
void go_trk_channel_reset(go_trk_channel *c)
{
    c->prn = 0; c->code_row = 0; c->state = GO_IDLE; c->lost_counter = 0; c->next_sample_index = 0;
    c->carrier_freq = 0; c->carrier_phase = 0; c->carrier_error = 0; c->carrier_nco = 0;
    c->code_phase = 0; c->code_error = 0; c->code_nco = 0; c->code_rate = 0;
    c->i_prompt = 0; c->q_prompt = 0;
}

/* do_tracking.rs:274-277 (Q6 row, Q7 saturating cast).  code_row 32 is out of bounds in the
 * reference (panic); callers must keep code_row in 0..=31. */
float go_trk_get_ca_chip(const go_trk_channel *c, float phase)
{
    const size_t idx = f32_as_usize(floorf(phase)) % 1023;
    return (float)go_ca_table()[(size_t)c->code_row * GO_CA_CODE_LEN + idx];
}

/* do_tracking.rs:231-272 */
void go_trk_early_late(go_trk_channel *c, go_c32 *data, float out6[6])
{
    const size_t n = c->num_samples_per_code;
    for (size_t i = 0; i < n; i++) {
        const float phase = c->carrier_phase + (2.0f * GO_PI_F * c->carrier_freq * (float)i / c->fs);
        const float cos_p = cosf(phase);
        const float sin_p = -sinf(phase);
        const float re = data[i].re * cos_p - data[i].im * sin_p;
        const float im = data[i].re * sin_p + data[i].im * cos_p;
        data[i].re = re;
        data[i].im = im;
    }
    c->carrier_phase =
        fmodf(c->carrier_phase + 2.0f * GO_PI_F * c->carrier_freq * ((float)n / c->fs), 2.0f * GO_PI_F);

    float i_p = 0, q_p = 0, i_e = 0, q_e = 0, i_l = 0, q_l = 0;
    for (size_t i = 0; i < n; i++) {
        const float chip_idx = fmodf(c->code_phase + ((float)i * (c->code_rate / c->fs)), 1023.0f);
        const float p_chip = go_trk_get_ca_chip(c, chip_idx);
        const float e_chip = go_trk_get_ca_chip(c, chip_idx + 0.5f);
        const float l_chip = go_trk_get_ca_chip(c, chip_idx - 0.5f);
        i_p += data[i].re * p_chip;
        q_p += data[i].im * p_chip;
        i_e += data[i].re * e_chip;
        q_e += data[i].im * e_chip;
        i_l += data[i].re * l_chip;
        q_l += data[i].im * l_chip;
    }
    c->code_phase = fmodf(c->code_phase + (c->code_rate / c->fs) * (float)n, 1023.0f);
    c->i_prompt = i_p;
    c->q_prompt = q_p;
    out6[0] = i_p; out6[1] = q_p; out6[2] = i_e; out6[3] = q_e; out6[4] = i_l; out6[5] = q_l;
}

/* do_tracking.rs:279-302 */
void go_trk_run_loop_filters(go_trk_channel *c, const float in6[6])
{
    const float i_p = in6[0], q_p = in6[1], i_e = in6[2], q_e = in6[3], i_l = in6[4], q_l = in6[5];
    const float pll_err = atanf(q_p / i_p) / (2.0f * GO_PI_F);
    c->carrier_nco = go_loop_filter_update(&c->pll_filter, pll_err, c->carrier_error, 0.001f);
    c->carrier_error = pll_err;
    c->carrier_freq += c->carrier_nco;

    const float pow_e = sqrtf(i_e * i_e + q_e * q_e);
    const float pow_l = sqrtf(i_l * i_l + q_l * q_l);
    const float dll_err = ((pow_e + pow_l) != 0.0f) ? (pow_e - pow_l) / (pow_e + pow_l) : 0.0f;
    c->code_nco = go_loop_filter_update(&c->dll_filter, dll_err, c->code_error, 0.001f);
    c->code_error = dll_err;
    c->code_rate += c->code_nco;
}

/* do_tracking.rs:183-210 (Q9, Q10) */
int go_trk_do_work(go_trk_channel *c, go_c32 *data, float out6[6], uint8_t *msg_prn)
{
    go_trk_early_late(c, data, out6);
    const float power = out6[0] * out6[0] + out6[1] * out6[1];
    if (power > 15.0f) {
        c->lost_counter = 0;
        go_trk_run_loop_filters(c, out6);
    } else {
        c->lost_counter += 1;
        if (c->lost_counter >= 20) {
            go_trk_channel_reset(c);
            if (msg_prn) *msg_prn = c->prn; /* prn already zeroed by reset (Q9) */
            return 1;
        }
    }
    c->next_sample_index += c->num_samples_per_code;
    c->num_samples_per_code = f32_as_usize(roundf(c->fs / (c->code_rate / 1023.0f)));
    return 0;
}

/* ------------------------------------------------------------------ ring buffer */
int go_ring_init(go_ring *r, size_t buf_size)
{
    if (buf_size == 0 || (buf_size & (buf_size - 1))) return -1; /* the reference asserts */
    r->buffer = (go_c32 *)calloc(buf_size, sizeof(go_c32));
    r->buf_size = buf_size; r->mask = buf_size - 1; r->head = 0;
    return r->buffer ? 0 : -1;
}
void go_ring_free(go_ring *r) { free(r->buffer); r->buffer = NULL; }
/* multicast_ring_buffer.rs:66-101 */
void go_ring_write(go_ring *r, const go_c32 *s, size_t n)
{
    const size_t start = r->head & r->mask;
    if (start + n <= r->buf_size) {
        memcpy(r->buffer + start, s, n * sizeof(go_c32));
    } else {
        const size_t first = r->buf_size - start;
        memcpy(r->buffer + start, s, first * sizeof(go_c32));
        memcpy(r->buffer, s + first, (n - first) * sizeof(go_c32));
    }
    r->head += n;
}
size_t go_ring_head(const go_ring *r) { return r->head; }
/* multicast_ring_buffer.rs:107-129 */
void go_ring_copy_to_slice(const go_ring *r, size_t start, go_c32 *dest, size_t n)
{
    const size_t ps = start & r->mask;
    if (ps + n <= r->buf_size) {
        memcpy(dest, r->buffer + ps, n * sizeof(go_c32));
    } else {
        const size_t first = r->buf_size - ps;
        memcpy(dest, r->buffer + ps, first * sizeof(go_c32));
        memcpy(dest + first, r->buffer, (n - first) * sizeof(go_c32));
    }
}

/* intended update() (do_tracking.rs:160-180; Q5: as written it slices an empty Vec and panics) */
int go_trk_update(go_trk_channel *c, const go_ring *ring, go_c32 *scratch, float out6[6], int *msg, uint8_t *msg_prn)
{
    if (msg) *msg = 0;
    if (c->state != GO_TRACKING) return 0;
    /* :165-166: length of generate_ca_code_samples(prn, code_rate, fs) */
    c->num_samples_per_code = (uint64_t)go_num_samples_per_code(c->code_rate, c->fs);
    const size_t head = go_ring_head(ring);
    if ((int64_t)(head - (c->next_sample_index + c->num_samples_per_code)) < 0) return 0;
    go_ring_copy_to_slice(ring, c->next_sample_index, scratch, c->num_samples_per_code);
    const int m = go_trk_do_work(c, scratch, out6, msg_prn);
    if (msg) *msg = m;
    return 1;
}

typedef struct {
    go_trk_channel *ch; int n_channels; const go_c32 *stream; size_t stream_len; int n_epochs; float *hist;
    int next; pthread_mutex_t mu;
} trk_job;

static void *trk_thread(void *arg)
{
    trk_job *j = (trk_job *)arg;
    go_c32 *scratch = NULL;
    size_t cap = 0;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int c = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (c >= j->n_channels) break;
        go_trk_channel *ch = &j->ch[c];
        for (int e = 0; e < j->n_epochs; e++) {
            if (ch->state != GO_TRACKING) break;
            const size_t n = ch->num_samples_per_code;
            if (ch->next_sample_index + n > j->stream_len) break;
            if (n > cap) { cap = n * 2; scratch = (go_c32 *)realloc(scratch, cap * sizeof(go_c32)); }
            memcpy(scratch, j->stream + ch->next_sample_index, n * sizeof(go_c32));
            float out6[6];
            uint8_t mp;
            go_trk_do_work(ch, scratch, out6, &mp);
            if (j->hist) {
                j->hist[((size_t)e * j->n_channels + c) * 2 + 0] = out6[0];
                j->hist[((size_t)e * j->n_channels + c) * 2 + 1] = out6[1];
            }
        }
    }
    free(scratch);
    return NULL;
}

void go_trk_run_all(go_trk_channel *ch, int n_channels, const go_c32 *stream, size_t stream_len, int n_epochs,
                    int n_threads, float *hist)
{
    trk_job j;
    memset(&j, 0, sizeof(j));
    j.ch = ch; j.n_channels = n_channels; j.stream = stream; j.stream_len = stream_len; j.n_epochs = n_epochs;
    j.hist = hist;
    go_ca_table();
    pthread_mutex_init(&j.mu, NULL);
    run_threads(trk_thread, &j, n_threads);
    pthread_mutex_destroy(&j.mu);
}

/* ------------------------------------------------------------------ digital front-end */
/* rf/nco_lut.rs:24-42 + rf/dc_remove.rs:11-21 + rf/frontend.rs:18-30 */
void go_frontend_init(go_frontend *f, float f_if, float fs_in)
{
    memset(f, 0, sizeof(*f));
    for (int i = 0; i < GO_LUT_SIZE; i++) {
        const float angle = (2.0f * GO_PI_F * (float)i) / (float)GO_LUT_SIZE;
        f->lut_re[i] = cosf(angle);
        f->lut_im[i] = -sinf(angle);
    }
    f->phase_step = (f_if / fs_in) * (float)GO_LUT_SIZE;
    f->alpha = 0.001f;
    f->con = 1.0f - f->alpha;
}

/* rf/frontend.rs:32-62: 8 complex samples per step; 8 independent DC-bias lanes; sequential f32 phase
 * accumulator; mix_simd (nco_lut.rs:8-15) verbatim: i' = I*re + Q*im, q' = I*im - Q*re with im = -sin. */
void go_frontend_process_block(go_frontend *f, go_c32 *s, size_t n)
{
    for (size_t c = 0; c + 8 <= n; c += 8) {
        float re[8], im[8];
        for (int j = 0; j < 8; j++) {
            f->bias_re[j] = f->bias_re[j] * f->con + s[c + j].re * f->alpha;
            f->bias_im[j] = f->bias_im[j] * f->con + s[c + j].im * f->alpha;
            re[j] = s[c + j].re - f->bias_re[j];
            im[j] = s[c + j].im - f->bias_im[j];
        }
        for (int j = 0; j < 8; j++) {
            const size_t idx = f32_as_usize(f->phase_accumulator) % GO_LUT_SIZE;
            f->phase_accumulator = fmodf(f->phase_accumulator + f->phase_step, (float)GO_LUT_SIZE);
            const float lc = f->lut_re[idx], ls = f->lut_im[idx];
            s[c + j].re = re[j] * lc + im[j] * ls;
            s[c + j].im = re[j] * ls - im[j] * lc;
        }
    }
}

/* ------------------------------------------------------------------ bit sync / nav-bit accumulation (N4) */
void go_nav_bit_sync(const float *prompt_i, int n_epochs, int stride, go_nav_sync *st, int8_t *bits, int max_bits)
{
    memset(st, 0, sizeof(*st));
    st->sync_epoch = -1;
    float old_ip = 0.0f, acc = 0.0f;
    for (int cnt = 0; cnt < n_epochs; cnt++) {
        const float ip = prompt_i[(size_t)cnt * stride];
        const int biti = cnt % 20;
        if (!st->flag_bit_sync && cnt > 1000) {         /* decoding.rs:121-123 */
            if (old_ip * ip < 0.0f) {                   /* check_bit_sync, :164-180 */
                st->bit_sync_buff[biti] += 1;
                int i_max = 0;
                uint32_t v_max = 0;
                for (int i = 0; i < 20; i++)
                    if (st->bit_sync_buff[i] >= v_max) { v_max = st->bit_sync_buff[i]; i_max = i; } /* max_by: last max */
                st->frame_sync_ind = i_max;
                if (v_max == 30) { st->flag_bit_sync = 1; st->sync_epoch = cnt; }
            }
        }
        if (st->flag_bit_sync) {                        /* bit_accumulation, :182-213 */
            if (biti == st->frame_sync_ind) acc = ip; else acc += ip;
            if (biti == (st->frame_sync_ind + 19) % 20) {
                if (st->n_bits < max_bits) bits[st->n_bits] = acc > 0.0f ? 1 : -1;
                st->n_bits++;
            }
        }
        old_ip = ip;
    }
    /* check_preamble_syn, decoding.rs:215-226 (see gnss_oracle.h for the legacy's single test) */
    static const int pre[8] = {1, -1, -1, -1, 1, -1, 1, 1};   /* GPS_CA_PREAMBLE, gps_property_constants.rs:12 */
    st->preamble_bit = -1;
    const int nb = st->n_bits < max_bits ? st->n_bits : max_bits;
    for (int i0 = 0; i0 + 8 <= nb; i0++) {
        int corr = 0;
        for (int x = 0; x < 8; x++) corr += (int)bits[i0 + x] * pre[x % 8];
        const int hit = corr == 8 || corr == -8;
        if (i0 == 0 && hit) { st->ref_frame_sync = 1; st->ref_polarity = corr > 0 ? 1 : -1; }
        if (hit && st->preamble_bit < 0) { st->preamble_bit = i0; st->polarity = corr > 0 ? 1 : -1; }
        if (st->preamble_bit >= 0) break;
    }
}

/* ------------------------------------------------------------------ fine Doppler (N3), acquisition_bk.rs:215-302 */
int go_fine_doppler(const go_c32 *x, size_t n_long, const int8_t *code1023, size_t code_phase, float fs, int long_ms,
                    int is_complex, go_fine_result *out, float *mag_out)
{
    memset(out, 0, sizeof(*out));
    /* :234-235 mean = iter().sum::<Complex32>() / len as f32 (sequential f32) */
    float sre = 0.0f, sim = 0.0f;
    for (size_t i = 0; i < n_long; i++) { sre += x[i].re; sim += x[i].im; }
    const float mre = sre / (float)n_long, mim = sim / (float)n_long;
    const size_t n_code = f32_as_usize(roundf(fs / (GO_CA_CODE_RATE / (float)GO_CA_CODE_LEN)));   /* :236-239 */
    const size_t use = (size_t)(long_ms - 1) * n_code;                                              /* :240 */
    if (use == 0 || code_phase + use > n_long) return -1;
    size_t p2 = 1;
    while (p2 < use) p2 <<= 1;
    const size_t fft_size = 8 * p2;                                                                  /* :249 */
    go_c32 *buf = (go_c32 *)calloc(fft_size, sizeof(go_c32));
    if (!buf) return -1;
    for (size_t i = 0; i < use; i++) {
        const size_t ind = f32_as_usize(floorf((float)i * GO_CA_CODE_RATE / fs)) % GO_CA_CODE_LEN;   /* :241-247 */
        const float c = (float)code1023[ind];
        /* (x - mean) * Complex32::new(c, 0): re = a*c - b*0, im = a*0 + b*c  (:265-268) */
        const float a = x[code_phase + i].re - mre, b = x[code_phase + i].im - mim;
        buf[i].re = a * c - b * 0.0f;
        buf[i].im = a * 0.0f + b * c;
    }
    go_fft_plan *pl = go_fft_plan_new((int)fft_size, 0);
    go_fft_process(pl, buf);
    go_fft_plan_free(pl);
    float mx = -INFINITY;
    for (size_t k = 0; k < fft_size; k++) {
        const float m = hypotf(buf[k].re, buf[k].im);                                                /* :274 */
        if (mag_out) mag_out[k] = m;
        buf[k].re = m;
        mx = fmaxf(mx, m);                                                                           /* :276 */
    }
    size_t idx = 0;
    while (idx < fft_size && buf[idx].re != mx) idx++;                                               /* :277-280 */
    const size_t one_side = f32_as_usize(ceilf(((float)fft_size + 1.0f) / 2.0f));                    /* :250 */
    out->fft_size = (uint32_t)fft_size;
    out->idx = (uint32_t)idx;
    out->mag = mx;
    if (idx < one_side) {
        const float bin = (float)idx * fs / (float)fft_size;                                         /* :251-253 */
        out->carrier_freq = (is_complex ? -1.0f : 1.0f) * bin;                                       /* :296-298 */
        out->ref_defined = 1;
    } else {
        float m2 = -INFINITY;
        for (size_t k = one_side; k < fft_size; k++) m2 = fmaxf(m2, buf[k].re);
        size_t i2 = 0;
        while (buf[one_side + i2].re != m2) i2++;
        const float bin = (float)(one_side - i2) * fs / (float)fft_size;                             /* :283-295 */
        out->carrier_freq = bin;
        out->ref_defined = 0;
    }
    free(buf);
    return 0;
}
