/*
 * gnss_b200.h -- C-ABI of libgnss_b200: the B200 (sm_100a) implementation of the gnss-sdr-rs
 * acquisition / correlator hot path.
 *
 * Pure C, bindgen-friendly (fixed-width ints, POD structs, opaque handle); it is meant to be added
 * to the reference crate's wrapper.h next to src/include/rtl-sdr.h and linked through build.rs
 * exactly like libconvenience (reference build.rs:11-27).  Every entry point names the reference
 * item it replaces.  Conventions follow the reference's C side (src/include/convenience.h:52,62):
 * int return, 0 = success, negative = error (gb_strerror()); "satellite not found" is NOT an
 * error (found[] = 0), mirroring Option::None at do_acquisition.rs:225.
 *
 * Ownership: the caller owns every host buffer passed in or out; the library owns all device
 * memory behind gb_handle and frees it in gb_destroy().  No callbacks, no global state.
 * Threading: one handle may be used from an acquisition thread, a tracking thread and a sample-writer
 * thread at the same time (separate CUDA streams, like main.rs:205-227).  Calls of the same family
 * (gb_acq_* / gb_trk_* / gb_ring_*) on one handle are serialised inside the library by a per-family lock,
 * so the reference's 32 rayon workers (do_acquisition.rs:302-313) may call into one handle concurrently.
 * Host sample buffers always cross the boundary with their length; a buffer shorter than the call needs
 * is refused with GB_ERANGE, never read past its end.
 * There is NO CPU fallback: every compute entry point returns GB_ENODEVICE without a CUDA device.
 */
#ifndef GNSS_B200_H
#define GNSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_VERSION 120

/* ---- error codes ---- */
#define GB_OK 0
#define GB_EINVAL (-1)       /* bad argument */
#define GB_ENODEVICE (-2)    /* no CUDA device / driver */
#define GB_ECUDA (-3)        /* a CUDA call failed; gb_last_cuda_error() has the text */
#define GB_EUNSUPPORTED (-4) /* fft_size has no sm_100a plan */
#define GB_ESTATE (-5)       /* call order (e.g. search before configure) */
#define GB_ENOMEM (-6)
#define GB_ERANGE (-7)       /* samples requested that the ring (or the caller's buffer) does not hold */
#define GB_ENCCL (-8)        /* NCCL could not be loaded or a collective failed (gb_group_*) */

typedef struct gb_handle gb_handle;

typedef struct { float re, im; } gb_c32;  /* num_complex::Complex32 */
typedef struct { double re, im; } gb_c64; /* num_complex::Complex64 (FFT<f64> facade only) */

typedef struct {
    int32_t device;        /* CUDA ordinal */
    uint64_t ring_capacity; /* device sample ring, complex samples, power of two (0 = none yet) */
    uint32_t flags;        /* reserved, 0 */
} gb_config;

const char *gb_strerror(int code);
const char *gb_last_cuda_error(gb_handle *h);
int gb_version(void);
int gb_device_count(void);

/* Explicit A/B and tuning switches of the kernels (process-wide, default = the shipped organisation).  The library
 * reads NO environment variables.  Keys: "trk_ws" (-1 = generic tracking kernel everywhere, 0 = default, else
 * NSW*100+U*10+ROT), "trk_t" (CTA size of the generic tracking kernel), "acq_variant", "acq_nolw", "acq_nodb",
 * "acq_spec_ldg", "acq_lw_tmem" (default 2: the power accumulators and the per-thread code spectrum of the N = 4092 inverse kernel
 * live in tensor memory; 1 = the accumulators alone; 0 = the register forms selected by "acq_lw_minb" / "acq_lw_db"; cells are identical), "acq_tmem" (the same for the generic
 * inverse kernel of the other plans: -1 = per-plan default, on for the power-of-two plans; 0 / 1 force it), "fe_sequential", and "acq_tc"
 * (1 = the radix-31 stage of the N = 4092 inverse
 * kernel as 3 x TF32 products on the warp-level tensor path: a measured A/B, slower than the FP32 default and outside
 * the north star, results within 2e-6 of the default's).  Unknown keys are stored and ignored. */
int gb_tuning_set(const char *key, int value);
int gb_tuning_get(const char *key, int dflt);

int gb_create(const gb_config *cfg, gb_handle **out);
int gb_destroy(gb_handle *h);
int gb_synchronize(gb_handle *h);

/* ------------------------------------------------------------------ C/A code (host, no device)
 * replaces utilities/ca_code.rs:12-27 and the table constants/gps_ca_constants.rs (G1/G2 LFSR). */
int gb_ca_code_chips(int prn, int8_t *out1023);
int gb_num_samples_per_code(float code_rate, float fs);
int gb_generate_ca_code_samples(int prn, float code_rate, float fs, int8_t *out, int cap);

/* ------------------------------------------------------------------ sample ring (device + pinned staging)
 * replaces MulticastRingBuffer::{new, write_samples, get_head, copy_to_slice}
 * (utilities/multicast_ring_buffer.rs:46-129): monotonically increasing absolute sample index,
 * power-of-two capacity, single writer.  gb_ring_write stages through pinned host memory and
 * issues cudaMemcpyAsync on the copy stream; the acquisition/tracking streams wait on its event. */
int gb_ring_create(gb_handle *h, uint64_t capacity_pow2);
int gb_ring_write(gb_handle *h, const gb_c32 *samples, uint64_t n);
/* signed 8-bit real samples (the reference recordings, do_acquisition.rs:420-424): im = 0 */
int gb_ring_write_i8(gb_handle *h, const int8_t *samples, uint64_t n);
uint64_t gb_ring_head(gb_handle *h);
int gb_ring_copy_to_slice(gb_handle *h, uint64_t start, gb_c32 *dest, uint64_t n);
int gb_ring_reset(gb_handle *h);

/* ------------------------------------------------------------------ digital front-end
 * replaces DigitalFrontend::{new, process_block} (rf/frontend.rs:18-62), DcRemoverSimd (rf/dc_remove.rs:11-29),
 * NcoLut / mix_simd (rf/nco_lut.rs:8-42) and the per-block body of rf_thread (rf/rf_thread.rs:44-48): raw complex
 * samples -> DC removal (alpha 0.001, 8 interleaved lanes) -> 2048-entry NCO LUT mix -> appended to the ring.
 * Bit-exact with the reference arithmetic (sequential f32 phase accumulator and bias recurrences included).
 * n must be a multiple of 8 (the reference processes 16 floats at a time).
 * The phase accumulator (`acc = (acc + step) % 2048.0` per sample, frontend.rs:48-52) does not depend on the samples:
 * gb_frontend_configure computes its orbit from 0 -- a tail of mu states then a cycle of lambda states, found with
 * Brent's algorithm in the reference's f32 arithmetic -- and the kernel looks LUT indices up by sample number, so only
 * the DC-bias recurrences stay sequential.  Orbits longer than 2^24 states (none met in practice) and gb_tuning_set("fe_sequential", 1)
 * fall back to a one-thread sequential accumulator with identical results. */
int gb_frontend_configure(gb_handle *h, float f_if, float fs_in);
int gb_frontend_write(gb_handle *h, const gb_c32 *raw, uint64_t n);
/* GB_FE_EXACT (default): one CTA walks the 16 DC-bias recurrences in the reference's order -- bit-identical output.
 * GB_FE_PARALLEL: the recurrences as a segmented scan over many CTAs (three launches per call).  Every step inside a
 * 256-sample segment is rounded like dc_remove.rs:23-29, the composition across segments is not: samples and bias
 * state agree with the reference to 1e-5 * max|x| (measured 1e-6; tests/test_gpu_ring_fft.py), the NCO phase stays
 * exact.  Returns GB_EUNSUPPORTED when the configured (f_if, fs) has no orbit table (sequential fallback active). */
#define GB_FE_EXACT 0
#define GB_FE_PARALLEL 1
int gb_frontend_set_mode(gb_handle *h, int mode);
int gb_frontend_state(gb_handle *h, float *state17 /* phase_accumulator, bias_re[8], bias_im[8] */);
/* Host-only diagnostic (no device needed): the orbit gb_frontend_configure would build.  mu / lambda receive the tail
 * and cycle lengths; idx_out (may be NULL) receives the LUT index of samples 0 .. n_idx-1.  GB_EUNSUPPORTED if the
 * orbit exceeds the 2^24-state cap. */
int gb_frontend_orbit(float f_if, float fs_in, uint64_t *mu, uint64_t *lambda, uint16_t *idx_out, uint64_t n_idx);

/* ------------------------------------------------------------------ acquisition
 * replaces AcquisitionWorker::{new, search_satellite, is_good_satellite}
 * (do_acquisition.rs:118-239) for ALL PRNs at once (the rayon loop at :302-313),
 * DopplerShiftTable::new / apply_doppler_shift (doppler_shift.rs:11-58). */
typedef struct {
    float peak;      /* max of the accumulated power over code phase (local_max, :195-202) */
    uint32_t argmax; /* first index attaining it (local_best_phase) */
    float sum8;      /* sum over the first 8*floor(N/8) bins (is_good_satellite's SIMD sum, :229-234) */
    float peak2;     /* legacy two-peak metric (acquisition_bk.rs:342-399): largest power over the slices the
                        legacy searches around argmax (:371-390, bounds verbatim: [0,cp-spc) u [cp+spc,N), and the
                        two wrap cases [cp+spc-1, N+cp-spc) / [cp+spc-N-1, cp-spc)); 0 if disabled */
} gb_acq_cell;

typedef struct {
    uint8_t prn;
    uint8_t found; /* 1 = Some(result), 0 = None */
    int16_t doppler_bin; /* index into the Doppler table list, -1 if none */
    uint64_t code_phase_samples;
    float code_phase_chips;
    float carrier_freq; /* = f_if + f_d (doppler_shift.rs:20) */
    float fs;
    float mag_relative;
    uint64_t sample_global_index;
    float metric;       /* peak / ((sum8 - peak)/(N-1)) of the deciding bin */
    float peak_ratio;   /* sqrt(peak/peak2) of the deciding bin (0 if disabled) */
} gb_acq_result; /* AcquisitionResult, do_acquisition.rs:93-102, plus diagnostics */

/* Plan the search: AcquisitionWorker::new for prn = 1..n_prn (do_acquisition.rs:131-156).
 * codes == NULL: GPS C/A resampled per ca_code.rs:12-27 (n_prn <= 32).
 * codes != NULL: n_prn x fft_size +-1 samples (other constellations / the reference's test codes).
 * fft_size: ANY multiple of 4 up to 131072, like the reference's FftPlanner (:132-142).  The lengths listed by
 * gb_acq_supported_sizes() -- every sample rate of the BASELINE configurations -- run tuned shared-memory kernels; all other
 * lengths run the any-length plan (Bluestein chirp-z over a power-of-two Stockham FFT: same results, not tuned).
 * GB_EUNSUPPORTED only for fft_size % 4 != 0: there apply_doppler_shift (doppler_shift.rs:26) leaves the last fft_size % 4
 * samples of result_buf holding the previous block's unnormalised IFFT output, which is fed back N times larger every
 * block -- the reference's own arithmetic reaches inf / NaN within a dozen blocks, so no behaviour exists to reproduce. */
/* Every call re-plans and resets the per-configuration settings (n_coh = 1, aliasing off, no Doppler tables). */
int gb_acq_configure(gb_handle *h, int fft_size, float fs, int n_prn, const int8_t *codes);
/* The per-worker constructor of a drop-in (the reference builds one AcquisitionWorker per PRN, :268-271): plans the
 * built-in GPS C/A configuration unless exactly this one is already planned, in which case it returns GB_OK at once and
 * keeps the Doppler tables and settings -- 32 workers cost one plan, not 32. */
int gb_acq_configure_once(gb_handle *h, int fft_size, float fs, int n_prn);
int gb_acq_supported_sizes(int *sizes, int cap);   /* the lengths with TUNED plans (any multiple of 4 is accepted) */

/* DopplerShiftTable::new for each f_d in dopplers[] (doppler_shift.rs:11-21), evaluated on the
 * device in f32 in the reference's operation order.  carr_out[d] = f_if + f_d (may be NULL). */
int gb_acq_make_doppler_tables(gb_handle *h, float f_if, const float *dopplers, int n_doppler, float *carr_out);
/* Caller-built tables (the pub `table` field): n_doppler x fft_size complex + stored doppler_freq_hz */
int gb_acq_set_doppler_tables(gb_handle *h, const gb_c32 *tables, const float *carr, int n_doppler);
int gb_acq_get_doppler_tables(gb_handle *h, gb_c32 *tables_out, float *carr_out);

/* EXTENSION (BASELINE config 2, not in the reference): n_coh consecutive 1 ms blocks are summed
 * coherently (block c rotated by exp(-j 2 pi carr c N / fs)) before |.|^2.  1 = reference. */
int gb_acq_set_coherent(gb_handle *h, int n_coh);
/* Kernel organisation (results are bit-identical):
 *   GB_ACQ_FUSED : one kernel, one CTA per (PRN, Doppler) does the whole chain in shared memory;
 *   GB_ACQ_SHARED: two-kernel chain -- the PRN-independent forward path (wipe-off, coherent sum,
 *                  forward FFT) once per (Doppler, group), spectra left in L2, then per (PRN, Doppler)
 *                  x conj(code) -> IFFT -> |.|^2 -> cell.  Default.  For fft_size 4092 the inverse kernel
 *                  is the leftover-warp form (acq_lw.cu);
 *   GB_ACQ_SHARED_PLAIN: GB_ACQ_SHARED with the generic inverse kernel for every size (A/B, tests). */
#define GB_ACQ_FUSED 0
#define GB_ACQ_SHARED 1
#define GB_ACQ_SHARED_PLAIN 2
int gb_acq_set_mode(gb_handle *h, int mode);
/* Doppler aliasing (shared chain; OFF by default: the default is the reference's own per-bin wipe-off arithmetic).  When two bins of a grid built by gb_acq_make_doppler_tables lie a
 * whole number m of FFT bins (fs / fft_size) apart, the wiped block of one is the other's times exp(-j 2 pi m n / N):
 * its spectrum is the other's circularly shifted by m.  The forward path (wipe-off, coherent sum, forward FFT) then
 * runs for one bin per class only and the inverse kernel pairs that spectrum with the code spectrum shifted by m
 * (the classic circular-shift Doppler search).  Mathematically identical to doppler_shift.rs:11-58; numerically it
 * replaces the f32 rounding of cos(i * step_d) by that of the class's first bin (~1e-6 relative on the power, well
 * inside the 1e-3 contract).  off = every bin runs its own forward path with its own table (reference arithmetic).
 * BASELINE config 2 (50 Hz steps, 1 kHz FFT bins): 201 bins -> 20 forward spectra.  A re-planning gb_acq_configure
 * switches it back off (like n_coh, it is a per-configuration setting). */
int gb_acq_set_doppler_aliasing(gb_handle *h, int on);
/* number of forward spectra per group the next shared-chain search computes (== n_doppler when nothing is shared) */
int gb_acq_forward_bins(gb_handle *h);
/* samples_per_chip > 0 enables peak2; threshold is is_good_satellite's 7.0 */
int gb_acq_set_detector(gb_handle *h, float threshold, int samples_per_chip);

/* The fused search over the PRN x Doppler grid.  iq = n_samples host samples of which the first
 * num_integrations*fft_size are searched (search_satellite's samples_chunk; fewer -> GB_ERANGE); prn_mask bit
 * (prn-1) selects PRNs as at do_acquisition.rs:307 (for n_prn > 32 pass enable[] instead, NULL = all).
 * cells_out: n_prn x n_doppler, rows of PRNs that were not searched are zero.  May be NULL. */
int gb_acq_search_cells(gb_handle *h, const gb_c32 *iq, uint64_t n_samples, int num_integrations, uint32_t prn_mask,
                        const uint8_t *enable, gb_acq_cell *cells_out);
/* same, reading the chunk from the device ring at absolute index local_tail (do_acquisition.rs:297-301) */
int gb_acq_search_cells_ring(gb_handle *h, uint64_t local_tail, int num_integrations, uint32_t prn_mask,
                             const uint8_t *enable, gb_acq_cell *cells_out);
/* search_satellite's decision (early-exit order, Q1) on one PRN's cells; host-side O(D) scan */
int gb_acq_decide(const gb_acq_cell *cells, const float *carr, int n_doppler, int prn, int fft_size, float fs,
                  uint64_t local_tail, float threshold, gb_acq_result *out);
/* cells + decide for every selected PRN: results[n_prn] */
int gb_acq_search(gb_handle *h, const gb_c32 *iq, uint64_t n_samples, int num_integrations, uint64_t local_tail,
                  uint32_t prn_mask, const uint8_t *enable, gb_acq_result *results);
int gb_acq_search_ring(gb_handle *h, uint64_t local_tail, int num_integrations, uint32_t prn_mask,
                       const uint8_t *enable, gb_acq_result *results);
/* Asynchronous form on ONE handle (two slots, 0 / 1): gb_acq_search_enqueue returns as soon as the copies and kernels
 * are queued, gb_acq_search_wait(slot) delivers that search's results (and, optionally, its cells).  With two searches
 * alternating between the slots the upload of one overlaps the inverse kernel of the other (cudaMemcpyAsync on the copy
 * stream, events to the acquisition stream) -- the receiver loop of do_acquisition.rs:297-320 without a stall per chunk.
 * iq must stay valid and unchanged until the matching wait, and should be pinned host memory (a pageable buffer makes
 * the copy synchronous).  GB_ESTATE: the slot is still occupied / nothing was enqueued on it. */
int gb_acq_search_enqueue(gb_handle *h, const gb_c32 *iq, uint64_t n_samples, int num_integrations, uint64_t local_tail,
                          uint32_t prn_mask, const uint8_t *enable, int slot);
int gb_acq_search_wait(gb_handle *h, int slot, gb_acq_result *results /* n_prn, may be NULL */,
                       gb_acq_cell *cells_out /* n_prn x n_doppler, may be NULL */);
/* n_rec recordings of n_samples each, searched back to back through the pair above; results: n_rec x n_prn */
int gb_acq_search_batch(gb_handle *h, const gb_c32 *const *recordings, int n_rec, uint64_t n_samples, int num_integrations,
                        uint64_t local_tail, uint32_t prn_mask, const uint8_t *enable, gb_acq_result *results);
/* accumulated power row of one (prn, doppler bin) -- diagnostics / tests */
int gb_acq_bin_power(gb_handle *h, const gb_c32 *iq, uint64_t n_samples, int num_integrations, int prn, int doppler_bin,
                     float *power_out);
/* device time of the last search in milliseconds (CUDA events on the acquisition stream).  For the ring-resident
 * searches this is kernel time only; for the host-buffer searches the sliced upload is overlapped with the forward
 * path inside the same pair of events, so the figure includes the part of the H2D copy that could not be hidden. */
float gb_acq_last_kernel_ms(gb_handle *h);

/* ------------------------------------------------------------------ fine Doppler
 * replaces finer_doppler (acquisition_bk.rs:215-302), the legacy sub-bin carrier estimate used at the
 * acquisition -> tracking hand-over: long_ms (LONG_SAMPLES_LENGTH = 11) ms of samples, complex mean removed,
 * (long_ms-1) ms starting at code_phase stripped of the C/A code, zero-padded to 8 * next_power_of_two samples,
 * forward FFT, FIRST index of the largest magnitude.  The zero padding is never materialised (fine_doppler.cu).
 * Frequency mapping as in the legacy (:282-299): idx below one_side = ceil((fft_size+1)/2) gives
 * carrier_freq = (is_complex ? -1 : +1) * idx * fs / fft_size and ref_defined = 1.  For idx >= one_side the legacy
 * indexes its fft_freq_bins Vec out of bounds and panics; ref_defined = 0 and carrier_freq is what its arithmetic
 * would have produced.  code_phase + (long_ms-1)*N > n_long returns GB_ERANGE (the legacy slice panics).
 * codes1023: n_req x 1023 chips (+-1), or NULL for GPS C/A selected by req[i].prn.
 * Supported: 4096 <= next_power_of_two((long_ms-1)*N) <= 524288, else GB_EUNSUPPORTED. */
typedef struct {
    uint8_t prn;
    uint8_t reserved[3];
    uint32_t code_phase; /* samples, AcquisitionResult::code_phase */
} gb_fine_req;
typedef struct {
    uint32_t fft_size;
    uint32_t idx;        /* first index of the maximum magnitude */
    float mag;           /* that magnitude (Complex::abs) */
    float carrier_freq;
    int32_t ref_defined;
} gb_fine_result;
/* mag_out (optional, diagnostics): n_req x fft_size magnitudes */
int gb_acq_fine_doppler(gb_handle *h, const gb_c32 *long_samples, uint64_t n_long, float fs, int long_ms, int is_complex,
                        const gb_fine_req *req, int n_req, const int8_t *codes1023, gb_fine_result *out, float *mag_out);
/* same, on n_long samples of the device ring starting at absolute index start */
int gb_acq_fine_doppler_ring(gb_handle *h, uint64_t start, uint64_t n_long, float fs, int long_ms, int is_complex,
                             const gb_fine_req *req, int n_req, const int8_t *codes1023, gb_fine_result *out);
float gb_acq_fine_last_kernel_ms(gb_handle *h);

/* measured FP32 FMA throughput of the device (TFLOP/s): the roofline denominator for these kernels,
 * which are FP32-pipe / shared-memory bound, not HBM- or tensor-bound */
int gb_bench_fp32_tflops(gb_handle *h, float *tflops_out);

/* ------------------------------------------------------------------ FFT facade
 * replaces FFT<T>::{execute, power_spectrum}, RealFFT<T>::{execute, power_spectrum} (fft.rs:5-56), T = f32 / f64, for ANY
 * length n >= 2 (rustfft / realfft plan any n): natural-order, unnormalised, `batch` transforms of length n.  Lengths with a
 * tuned shared-memory plan (gb_acq_supported_sizes) run it in f32; every other length, and f64, run the any-length plan
 * (Bluestein over a power-of-two Stockham FFT, n <= 131072, batch <= 32768). */
int gb_fft_c2c(gb_handle *h, int n, int inverse, const gb_c32 *in, gb_c32 *out, int batch);
int gb_fft_power_spectrum(gb_handle *h, int n, const gb_c32 *in, float *out, int batch);
int gb_rfft(gb_handle *h, int n, const float *in, gb_c32 *out /* batch x (n/2+1) */, int batch);
int gb_rfft_power_spectrum(gb_handle *h, int n, const float *in, float *out /* batch x (n/2+1) */, int batch);
int gb_fft_c2c_f64(gb_handle *h, int n, int inverse, const gb_c64 *in, gb_c64 *out, int batch);
int gb_fft_power_spectrum_f64(gb_handle *h, int n, const gb_c64 *in, double *out, int batch);
int gb_rfft_f64(gb_handle *h, int n, const double *in, gb_c64 *out /* batch x (n/2+1) */, int batch);
int gb_rfft_power_spectrum_f64(gb_handle *h, int n, const double *in, double *out /* batch x (n/2+1) */, int batch);

/* ------------------------------------------------------------------ tracking
 * replaces TrackingChannel::{early_late_correlation, get_ca_chip, run_loop_filters, do_work, update}
 * and the rayon loop of TrackingManager::process_channels (do_tracking.rs:160-302, 364-371). */
#define GB_TRK_IDLE 0
#define GB_TRK_TRACKING 1

typedef struct {
    uint8_t id, prn;
    uint8_t state;      /* GB_TRK_IDLE / GB_TRK_TRACKING */
    uint8_t code_row;   /* C/A table row used by get_ca_chip; the reference uses prn (Q6), not prn-1 */
    uint32_t lost_counter;
    float fs;
    uint32_t epochs_done;
    uint64_t next_sample_index;
    uint64_t num_samples_per_code;
    float carrier_freq, carrier_phase, carrier_error, carrier_nco;
    float code_phase, code_error, code_nco, code_rate;
    float i_prompt, q_prompt;
    float pll_tau1, pll_tau2, dll_tau1, dll_tau2; /* LoopFilter, do_tracking.rs:52-71 */
} gb_trk_channel; /* TrackingChannel's pub state, do_tracking.rs:88-115 */

typedef struct { float i_p, q_p, i_e, q_e, i_l, q_l; } gb_trk_corr;

/* modes */
#define GB_TRK_FAST 0     /* tree reductions, device sincosf                                   */
#define GB_TRK_ORDERED 1  /* six sums accumulated in sample order, f64-evaluated sin/cos; for
                             closed-loop parity runs (SURVEY note E3)                            */

/* host helpers (no device): TrackingChannel::new / start / reset, LoopFilter::new */
int gb_trk_channel_init(gb_trk_channel *c, uint8_t id, float fs);
/* TrackingChannel::start verbatim (do_tracking.rs:148-154): code_row = prn, the reference's Q6 off-by-one (the
 * channel then correlates with PRN + 1's code; PRN 32 has no row -- the reference panics, here that channel is idled) */
int gb_trk_channel_start(gb_trk_channel *c, const gb_acq_result *r);
/* the same hand-over with the satellite's own C/A row (code_row = prn - 1): what a working receiver wants */
int gb_trk_channel_start_corrected(gb_trk_channel *c, const gb_acq_result *r);
int gb_trk_channel_reset(gb_trk_channel *c);
int gb_loop_filter_new(float noise_bw, float damping, float gain, float *tau1, float *tau2);

/* early_late_correlation for n_channels channels, each on its own n = num_samples_per_code host
 * samples inside data[0 .. n_data) (offsets[c] = start of channel c's samples; a segment that leaves the buffer ->
 * GB_ERANGE).  Updates carrier_phase, code_phase, i_prompt, q_prompt like the reference does; no loop filters. */
int gb_trk_correlate(gb_handle *h, gb_trk_channel *ch, int n_channels, const gb_c32 *data, uint64_t n_data,
                     const uint64_t *offsets, int mode, gb_trk_corr *out);
/* do_work for every active channel whose samples are in the ring (TrackingChannel::update,
 * do_tracking.rs:160-210): one launch, correlators + lock test + loop filters + bookkeeping.
 * ran[c] = 1 if the channel consumed an epoch; lost[c] = 1 if it emitted SatelliteLost.  A channel whose code_row is
 * outside the C/A table (the reference panics) is reset to idle and reported lost; the other channels run.
 * GB_TRK_ORDERED keeps an epoch's samples in shared memory: sample rates above ~20 Msps return GB_EUNSUPPORTED. */
int gb_trk_epoch(gb_handle *h, gb_trk_channel *ch, int n_channels, int mode, gb_trk_corr *out, uint8_t *ran,
                 uint8_t *lost);
/* persistent form: state stays on the device; n_epochs epochs per channel (or until the ring head)
 * in one launch.  prompt_hist (optional): n_epochs x n_channels x {i_p, q_p}. */
int gb_trk_upload(gb_handle *h, const gb_trk_channel *ch, int n_channels);
int gb_trk_run(gb_handle *h, int n_epochs, int mode, float *prompt_hist);
/* same run; the prompt history is kept on the device for gb_nav_bit_sync(h, NULL, ...) -- no copy to the host and back */
int gb_trk_run_keep(gb_handle *h, int n_epochs, int mode);
int gb_trk_download(gb_handle *h, gb_trk_channel *ch, int n_channels);
float gb_trk_last_kernel_ms(gb_handle *h);

/* ------------------------------------------------------------------ bit sync + nav-bit accumulation
 * restates the orphaned legacy check_bit_sync / bit_accumulation (decoding.rs:8, 115-127, 164-213) on the batched
 * prompt history gb_trk_run() returns: per channel, sign changes of prompt-I are histogrammed by epoch % 20 (after
 * epoch 1000) until one slot reaches 30 (BIT_SYNC_THRESHOLD); from then on 20 prompts are summed per bit and its sign
 * emitted.  The legacy end-of-bit test omits the modulo (can only fire for frame_sync_ind == 0); the intended modular
 * condition is used.  bits: n_channels x max_bits (+1 / -1). */
typedef struct {
    int32_t flag_bit_sync;
    int32_t frame_sync_ind;
    int32_t sync_epoch; /* epoch at which synchronisation was declared, -1 if never */
    int32_t n_bits;
    uint32_t bit_sync_buff[20];
    /* preamble search on the bit stream (check_preamble_syn, decoding.rs:215-226): 8 consecutive bits correlated with
     * GPS_CA_PREAMBLE {1,-1,-1,-1,1,-1,1,1}, frame sync when the sum is +-8, polarity = its sign. */
    int32_t preamble_bit;   /* first bit index whose 8-bit window matches (the intended sliding search), -1 if none */
    int32_t polarity;       /* +1 / -1, 0 if none */
    int32_t ref_frame_sync; /* the legacy's literal outcome: its buff_preamble is never popped, so only the FIRST 8 bits
                               are ever tested (decoding.rs:131-136, 207-209) */
    int32_t ref_polarity;
} gb_nav_sync;
/* prompt_hist == NULL: the history the last gb_trk_run / gb_trk_run_keep left on the device (same n_epochs, n_channels) */
int gb_nav_bit_sync(gb_handle *h, const float *prompt_hist /* n_epochs x n_channels x 2 */, int n_epochs, int n_channels,
                    gb_nav_sync *out, int8_t *bits, int max_bits);

/* ------------------------------------------------------------------ multi-GPU: sharding + the final gather
 * replaces the rayon fan-out of do_acquisition.rs:302-313 / do_tracking.rs:364-371 across the GPUs of one box.  The units
 * are independent (PRNs of one recording, recordings of a batch, tracking channels), so there is NO data-path collective:
 * every rank (one process / one gb_handle per GPU) works on its share and the per-PRN result tables are gathered ONCE at
 * the end -- one ncclAllGather over NVLink on the group's own stream.  In a single process that drives several GPUs the
 * handles simply write into the caller's memory and no collective is needed at all.
 * NCCL is loaded with dlopen("libnccl.so.2") by the first gb_group_* call (GB_ENCCL if absent). */
uint32_t gb_shard_prn_mask(int rank, int world, int n_prn, uint32_t base_mask); /* PRNs of base_mask dealt round-robin */
int gb_shard_range(int n_items, int rank, int world, int *first, int *count);     /* contiguous blocks of recordings / channels */
typedef struct gb_group gb_group;
int gb_group_unique_id(uint8_t *id128);   /* rank 0; ship the 128 bytes to the peers by any host transport */
int gb_group_init(gb_handle *h, const uint8_t *id128, int rank, int world, gb_group **out);
int gb_group_rank(gb_group *g);
int gb_group_world(gb_group *g);
/* every rank contributes `bytes` bytes (host); all_out = world x bytes, rank-major, on every rank */
int gb_group_allgather(gb_group *g, const void *mine, uint64_t bytes, void *all_out);
/* asynchronous pair (two slots): the gather of batch k runs while batch k+1 is being searched */
int gb_group_allgather_begin(gb_group *g, const void *mine, uint64_t bytes, int slot);
int gb_group_allgather_end(gb_group *g, int slot, void *all_out);
/* n results per rank -> world x n on every rank */
int gb_group_gather_results(gb_group *g, const gb_acq_result *mine, int n, gb_acq_result *all);
int gb_group_destroy(gb_group *g);

#ifdef __cplusplus
}
#endif
#endif
