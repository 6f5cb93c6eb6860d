/*
 * gnss_oracle.h -- CPU restatement ("oracle") of the gnss-sdr-rs hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing that ships (gnss-sdr-rs_b200/, the C-ABI
 * library libgnss_b200.so) may include, link or call this.  It is used by
 * tests/, by __graft_entry__.smoke() as the checker, and by bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: the reference (Rust nightly + rustfft 6.1.0) cannot be built
 * in this environment and its bundled IQ recording is missing, so FFT output
 * VALUES are "parity unpinned".  What IS pinned against the reference's own
 * vectors: the C/A table (PRN-1 1023-chip KAT from src/bk/gps_ca_prn.rs:72-123
 * and the sha256 of src/constants/gps_ca_constants.rs), the AcquisitionManager
 * masks (do_acquisition.rs:371-394), the ring-buffer wrap test
 * (multicast_ring_buffer.rs:147-209) and the loop-filter constants.  The FFT is
 * pinned against scipy.fft (pocketfft) and an f64 DFT in tests/.
 *
 * Every function cites the reference file:line it follows.  All arithmetic is
 * IEEE f32 in the reference's operation order; build with -ffp-contract=off.
 */
#ifndef GNSS_ORACLE_H
#define GNSS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } go_c32;
typedef struct { double re, im; } go_c64;

/* ---- constants (src/constants/gps_property_constants.rs:3-5) ---- */
#define GO_CA_CODE_RATE 1.023e6f
#define GO_CA_CODE_LEN 1023

/* ---- C/A code ---- */
/* G1/G2 LFSR with phase-selector taps; chips are +1 for bit 1, -1 for bit 0
 * (the convention of src/constants/gps_ca_constants.rs).  prn in 1..=32. */
int go_ca_code_chips(int prn, int8_t out[GO_CA_CODE_LEN]);
/* the whole 32x1023 table, row r = PRN r+1 (GPS_CA_CODE_32_PRN) */
const int8_t *go_ca_table(void);

/* utilities/ca_code.rs:12-27.  Returns the number of samples
 * round(fs / (code_rate / 1023)); writes min(n, cap) of them. */
int go_generate_ca_code_samples(int prn, float code_rate, float fs, int8_t *out, int cap);
int go_num_samples_per_code(float code_rate, float fs);

/* ---- FFT (stands in for rustfft 6.1.0: unnormalised, any N) ---- */
typedef struct go_fft_plan go_fft_plan;
go_fft_plan *go_fft_plan_new(int n, int inverse);
void go_fft_plan_free(go_fft_plan *p);
void go_fft_process(const go_fft_plan *p, go_c32 *data);          /* in place, f32 */
typedef struct go_fft64_plan go_fft64_plan;
go_fft64_plan *go_fft64_plan_new(int n, int inverse);
void go_fft64_plan_free(go_fft64_plan *p);
void go_fft64_process(const go_fft64_plan *p, go_c64 *data);      /* in place, f64 */
/* src/fft.rs:5-56 facade */
void go_fft_forward(int n, go_c32 *data);
void go_fft_power_spectrum(int n, go_c32 *data, float *out);
void go_rfft_forward(int n, const float *in, go_c32 *out /* n/2+1 */);

/* ---- acquisition ---- */
/* acquisition/doppler_shift.rs:11-21: table[i] = (cos(i*step), -sin(i*step)),
 * step = 2*pi*(f_if+f_d)/fs ; returns the stored doppler_freq_hz = f_if+f_d. */
float go_doppler_table(float f_if, float f_d, float fs, int n, go_c32 *table);
/* acquisition/doppler_shift.rs:25-58 (tail of len%4 samples left untouched) */
void go_apply_doppler_shift(const go_c32 *samples, const go_c32 *table, go_c32 *out, int len);

typedef struct {
    uint8_t prn;
    uint64_t code_phase_samples;
    float code_phase_chips;
    float carrier_freq;
    float fs;
    float mag_relative;
    uint64_t sample_global_index;
} go_acq_result; /* do_acquisition.rs:93-102 */

typedef struct {
    float peak;      /* max accumulated power in the bin            */
    uint32_t argmax; /* first index of the strict maximum (init 0.0) */
    float sum8;      /* 8-lane sum over chunks_exact(8) (Q2)         */
} go_acq_cell;

typedef struct go_acq_worker go_acq_worker;
/* do_acquisition.rs:131-156 */
go_acq_worker *go_acq_worker_new(int prn, int fft_size, float fs);
/* generic-code variant (extension: any +-1 code of length fft_size already resampled) */
go_acq_worker *go_acq_worker_new_code(int prn, int fft_size, float fs, const int8_t *code_samples);
void go_acq_worker_free(go_acq_worker *w);
const go_c32 *go_acq_worker_code_fft(const go_acq_worker *w);

/* do_acquisition.rs:158-226, reference semantics incl. early exit (Q1).
 * tables: D contiguous tables of fft_size entries; carr[D] = stored doppler_freq_hz.
 * Returns 1 and fills *out when a satellite is found, else 0. */
int go_acq_search(go_acq_worker *w, const go_c32 *samples, const go_c32 *tables, const float *carr,
                  int n_doppler, uint64_t local_tail, int num_integrations, go_acq_result *out);

/* Full grid (no early exit): cells[d] for every Doppler bin.  n_coh = 1 is the
 * reference's per-bin arithmetic (do_acquisition.rs:171-202, 229-234).
 * n_coh > 1 is the EXTENSION named by BASELINE config 2: the complex
 * correlations of n_coh consecutive blocks are summed, block c rotated by
 * rot[d*n_coh + c] (see go_coh_rotators), before |.|^2; num_integrations must be a
 * multiple of n_coh.  presum != 0 sums the wiped blocks BEFORE the FFT
 * (mathematically identical by linearity, n_coh x fewer FFTs). */
void go_acq_cells(go_acq_worker *w, const go_c32 *samples, const go_c32 *tables, int n_doppler,
                  int num_integrations, int n_coh, const go_c32 *rot, int presum, go_acq_cell *cells);
/* accumulated power of one bin (for tests), same options */
void go_acq_bin_power(go_acq_worker *w, const go_c32 *samples, const go_c32 *table, int num_integrations,
                      int n_coh, const go_c32 *rot, int presum, float *power);
/* rot[c] = exp(-j*2*pi*carr*(c*n)/fs) evaluated in f64, rounded to f32 */
void go_coh_rotators(float carr, float fs, int n, int n_coh, go_c32 *rot);

/* do_acquisition.rs:229-238 on a cell */
int go_is_good_cell(float peak, float sum8, int fft_size, float threshold);
/* Q1 scan over cells in table order == search_satellite's decision */
int go_acq_decide(const go_acq_cell *cells, const float *carr, int n_doppler, int prn, int fft_size,
                  float fs, uint64_t local_tail, float threshold, go_acq_result *out, int *bin_out);
/* legacy two-peak metric, acquisition_bk.rs:342-399 restated on a power row */
float go_two_peak_ratio(const float *power, int n, int samples_per_chip, uint32_t *first, uint32_t *second);

/* threaded search over PRNs (one worker per PRN, as rayon does at do_acquisition.rs:302-313).
 * found[p] = 1/0, results[p]; early_exit=0 runs the full grid + decide. */
void go_acq_search_all(go_acq_worker **workers, int n_workers, const go_c32 *samples, const go_c32 *tables,
                       const float *carr, int n_doppler, uint64_t local_tail, int num_integrations,
                       int early_exit, int n_threads, int *found, go_acq_result *results);
void go_acq_cells_all(go_acq_worker **workers, int n_workers, const go_c32 *samples, const go_c32 *tables,
                      int n_doppler, int num_integrations, int n_coh, const go_c32 *rot, int presum,
                      int n_threads, go_acq_cell *cells /* n_workers x n_doppler */);

/* do_acquisition.rs:39-74 */
typedef struct { int mode; /* 0 cold, 1 warm, 2 steady */ } go_acq_manager;
void go_acq_manager_update_mode(go_acq_manager *m, size_t tracked);
void go_acq_manager_pacing(const go_acq_manager *m, uint32_t active_mask /* bit prn-1 */, uint64_t *interval_ms,
                           uint32_t *mask);

/* ---- tracking ---- */
typedef struct { float tau1, tau2; } go_loop_filter; /* do_tracking.rs:52-71 */
go_loop_filter go_loop_filter_new(float noise_bw, float damping, float gain);
float go_loop_filter_update(const go_loop_filter *f, float d_err, float err, float dt);

enum { GO_IDLE = 0, GO_TRACKING = 1 };
typedef struct {
    uint8_t id, prn;
    int32_t state;       /* GO_IDLE / GO_TRACKING(prn) */
    int32_t code_row;    /* row of the C/A table used by get_ca_chip: reference = prn (Q6) */
    uint32_t lost_counter;
    float fs;
    uint64_t next_sample_index;
    uint64_t num_samples_per_code;
    float carrier_freq, carrier_phase, carrier_error, carrier_nco;
    float code_phase, code_error, code_nco, code_rate;
    float i_prompt, q_prompt;
    go_loop_filter pll_filter, dll_filter;
} go_trk_channel; /* do_tracking.rs:88-115 */

void go_trk_channel_init(go_trk_channel *c, uint8_t id, float fs);    /* :118-146 */
void go_trk_channel_start(go_trk_channel *c, const go_acq_result *r); /* :148-154 */
void go_trk_channel_reset(go_trk_channel *c);                         /* :311-326 */
float go_trk_get_ca_chip(const go_trk_channel *c, float phase);       /* :274-277 */
/* :231-272; data is rotated in place exactly as the reference does */
void go_trk_early_late(go_trk_channel *c, go_c32 *data, float out6[6]);
void go_trk_run_loop_filters(go_trk_channel *c, const float in6[6]);  /* :279-302 */
/* :183-210; returns 0 = no message, 1 = SatelliteLost(msg_prn) */
int go_trk_do_work(go_trk_channel *c, go_c32 *data, float out6[6], uint8_t *msg_prn);

/* ---- ring buffer (utilities/multicast_ring_buffer.rs:36-130) ---- */
typedef struct {
    go_c32 *buffer;
    size_t buf_size, mask, head;
} go_ring;
int go_ring_init(go_ring *r, size_t buf_size); /* power of two or returns -1 */
void go_ring_free(go_ring *r);
void go_ring_write(go_ring *r, const go_c32 *samples, size_t n);
size_t go_ring_head(const go_ring *r);
void go_ring_copy_to_slice(const go_ring *r, size_t start, go_c32 *dest, size_t n);
/* intended TrackingChannel::update (do_tracking.rs:160-180, Q5): 1 if an epoch ran */
int go_trk_update(go_trk_channel *c, const go_ring *ring, go_c32 *scratch, float out6[6], int *msg, uint8_t *msg_prn);

/* many channels x many epochs over one shared stream (threaded over channels,
 * do_tracking.rs:364-371).  hist (optional) = n_epochs x n_channels x 2 prompt I/Q. */
void go_trk_run_all(go_trk_channel *ch, int n_channels, const go_c32 *stream, size_t stream_len, int n_epochs,
                    int n_threads, float *hist);

/* ---- digital front-end: rf/frontend.rs:32-62, rf/dc_remove.rs:23-29, rf/nco_lut.rs:24-42 ---- */
#define GO_LUT_SIZE 2048
typedef struct {
    float lut_re[GO_LUT_SIZE], lut_im[GO_LUT_SIZE];
    float phase_accumulator, phase_step;
    float bias_re[8], bias_im[8];
    float alpha, con;
} go_frontend;
void go_frontend_init(go_frontend *f, float f_if, float fs_in);
/* in place on n complex samples; processes floor(n/8)*8 samples, the tail is left untouched (chunks_exact_mut(16)) */
void go_frontend_process_block(go_frontend *f, go_c32 *samples, size_t n);

/* ---- bit synchronisation + 20 ms prompt accumulation (orphaned legacy decoding.rs:8,115-127,164-213) ----
 * Per channel over its prompt-I history (one value per 1 ms epoch, cnt = epoch index):
 *   biti = cnt % 20; while not synchronised and cnt > 1000: a sign change between consecutive prompts increments
 *   bit_sync_buff[biti]; frame_sync_ind = index of the maximum (Iterator::max_by keeps the LAST of equal maxima);
 *   synchronised when that maximum reaches BIT_SYNC_THRESHOLD = 30 (decoding.rs:164-180).
 *   Once synchronised: i_p restarts at biti == frame_sync_ind and a bit (+1 if sum > 0 else -1) is emitted at the 20th
 *   epoch of the window.  The legacy compares biti with frame_sync_ind + 19 WITHOUT the modulo (decoding.rs:199-201),
 *   which can only fire for frame_sync_ind == 0; the intended modular condition is used here and stated as a deviation. */
typedef struct {
    int32_t flag_bit_sync;
    int32_t frame_sync_ind;
    int32_t sync_epoch;     /* cnt at which synchronisation was declared, -1 if never */
    int32_t n_bits;
    uint32_t bit_sync_buff[20];
    /* check_preamble_syn (decoding.rs:215-226) on the emitted bits: corr = sum_{x<8} bits[i0+x] * GPS_CA_PREAMBLE[x % 8]
     * (gps_property_constants.rs:12), frame sync iff |corr| == 8, polarity = signum(corr).  The legacy pushes every bit
     * into buff_preamble (:207-209) and tests only while buff_preamble.len() == 8 (:131-136); the VecDeque is never popped,
     * so only the FIRST 8 bits are ever tested: ref_frame_sync / ref_polarity.  preamble_bit / polarity are the intended
     * sliding search (first window that matches). */
    int32_t preamble_bit, polarity, ref_frame_sync, ref_polarity;
} go_nav_sync;
void go_nav_bit_sync(const float *prompt_i, int n_epochs, int stride, go_nav_sync *st, int8_t *bits, int max_bits);

/* ---- fine Doppler: finer_doppler, acquisition_bk.rs:215-302 ----
 * long_samples: LONG_SAMPLES_LENGTH (11) ms of complex samples (the legacy takes i16 I/Q pairs and converts to f32,
 * :223-233).  Steps, in the reference's f32 operation order: complex mean (sequential f32 sum / len) removed (:234-235);
 * N = round(fs / (1.023e6/1023)); size_signal_use = (long_ms-1)*N (:240); code index
 * floor(x as f32 * 1.023e6 / fs) as usize % 1023 (:241-247); fft_size = 8 * next_power_of_two(size_signal_use) (:249);
 * carrier_sig = (samples[code_phase + x] - mean) * code1023[idx[x]] zero-padded (:259-271); forward FFT (rustfft Radix4,
 * restated by go_fft); magnitude = Complex::abs = hypot (:274); max by f32::max then FIRST index equal to it (:276-280).
 * Frequency mapping (:282-299): one_side = ceil((fft_size as f32 + 1)/2); bins[x] = x as f32 * fs / fft_size as f32;
 *   idx <  one_side : carrier = (is_complex ? -1 : +1) * bins[idx]                       -> ref_defined = 1
 *   idx >= one_side : the legacy indexes fft_freq_bins[one_side] (a Vec of length one_side) and PANICS (:283-287, and
 *                     :297 for idx == one_side).  ref_defined = 0 and carrier = the value the code would have produced
 *                     had the Vec been long enough: idx' = argmax over mag[one_side..] (:288-295),
 *                     carrier = +bins(one_side - idx').
 * Returns 0, or -1 if the recording is shorter than code_phase + size_signal_use (the legacy slice would panic). */
typedef struct {
    uint32_t fft_size;
    uint32_t idx;          /* first index of the maximum magnitude over the whole spectrum */
    float mag;             /* that magnitude */
    float carrier_freq;
    int32_t ref_defined;
} go_fine_result;
int go_fine_doppler(const go_c32 *long_samples, size_t n_long, const int8_t *code1023, size_t code_phase, float fs,
                    int long_ms, int is_complex, go_fine_result *out, float *mag_out /* fft_size or NULL */);

#ifdef __cplusplus
}
#endif
#endif
