// gnss_sdr_rs.hpp -- C++ host-side mirror of the reference crate's acquisition / tracking API over the
// C-ABI (include/gnss_b200.h).  The reference is Rust and no Rust toolchain exists in the build image,
// so the host side above the C-ABI is written in C++ with the crate's own names, argument meaning and
// error behaviour ("not found" = empty optional, errors = AcqError / TrackingError exceptions standing
// in for the Result<_, AcqError> / Result<_, TrackingError> of do_acquisition.rs:76-91 and
// do_tracking.rs:31-45).  Header-only; link with -lgnss_b200.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <optional>
#include <set>
#include <stdexcept>
#include <utility>
#include <array>
#include <vector>

#include "../../include/gnss_b200.h"

namespace gnss_sdr_rs {

using Complex32 = gb_c32;

struct AcqError : std::runtime_error {  // do_acquisition.rs:76-91
    int code;
    explicit AcqError(int c) : std::runtime_error(std::string("Acquisition Error! ") + gb_strerror(c)), code(c) {}
};
struct TrackingError : std::runtime_error {  // do_tracking.rs:31-45
    int code;
    explicit TrackingError(int c) : std::runtime_error(std::string("TrackingError ") + gb_strerror(c)), code(c) {}
};
struct MulticastRingBuffError : std::runtime_error {  // multicast_ring_buffer.rs:9-24
    explicit MulticastRingBuffError(int c) : std::runtime_error(gb_strerror(c)) {}
};

// One GPU.  Shared by the acquisition and tracking stages (separate CUDA streams inside).
class GpuEngine {
  public:
    explicit GpuEngine(int device = 0)
    {
        gb_config cfg{device, 0, 0};
        const int rc = gb_create(&cfg, &h_);
        if (rc) throw AcqError(rc);
    }
    ~GpuEngine() { if (h_) gb_destroy(h_); }
    GpuEngine(const GpuEngine&) = delete;
    GpuEngine& operator=(const GpuEngine&) = delete;
    gb_handle* raw() const { return h_; }

    // Per-engine caches behind the kept per-worker API (AcquisitionWorker::search_satellite is called by 32 rayon
    // threads with the SAME samples_chunk and the SAME doppler_table, do_acquisition.rs:302-313).  One lock; the first
    // caller uploads the tables and runs ONE fused search for all PRNs, the other 31 read its results.
    std::mutex mu;
    uint64_t tables_key = 0;       // identity of the Doppler tables now on the device (0 = none)
    uint64_t search_key = 0;       // identity of the (chunk, tables, local_tail, K) whose results are cached
    std::vector<gb_acq_result> search_results;
    void invalidate_caches() { std::lock_guard<std::mutex> lk(mu); tables_key = 0; search_key = 0; }

  private:
    gb_handle* h_ = nullptr;
};

// identity of a buffer: address, length and a strided sample of its contents (FNV-1a).  A buffer that is rewritten in
// place between calls changes its content sample; GpuEngine::invalidate_caches() is there for the adversarial case.
inline uint64_t buffer_key(const void* p, size_t bytes, uint64_t seed)
{
    uint64_t h = 1469598103934665603ull ^ seed;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    mix((uint64_t)(uintptr_t)p);
    mix((uint64_t)bytes);
    const unsigned char* b = static_cast<const unsigned char*>(p);
    const size_t step = bytes > 4096 ? bytes / 509 : 1;   // ~500 probes spread over the buffer (+ its last bytes)
    for (size_t i = 0; i < bytes; i += step) mix(b[i]);
    for (size_t i = bytes > 64 ? bytes - 64 : 0; i < bytes; i++) mix(b[i]);
    return h ? h : 1;
}

// constants (do_acquisition.rs:20-23, do_tracking.rs:16-29, gps_property_constants.rs:3-5)
constexpr float FREQ_SEARCH_ACQUISITION_HZ = 14e3f;
constexpr uint16_t FREQ_SEARCH_STEP_HZ = 500;
constexpr uint8_t PRN_SEARCH_ACQUISITION_TOTAL = 32;
constexpr size_t LONG_SAMPLES_LENGTH = 10;
constexpr float GPS_L1_CA_CODE_RATE_CHIPS_PER_S = 1.023e6f;
constexpr float GPS_L1_CA_CODE_LENGTH_CHIPS = 1023.0f;
constexpr float LOCK_THRESHOLD = 15.0f;
constexpr uint32_t MAX_LOST_EPOCHS = 20;
constexpr size_t NUM_OF_CHANNELS = 15;

// utilities/ca_code.rs:12-27
inline std::vector<int8_t> generate_ca_code_samples(uint8_t prn, float code_rate, float f_sampling)
{
    const int n = gb_num_samples_per_code(code_rate, f_sampling);
    std::vector<int8_t> out(n);
    const int rc = gb_generate_ca_code_samples(prn, code_rate, f_sampling, out.data(), n);
    if (rc < 0) throw AcqError(rc);
    return out;
}

// utilities/multicast_ring_buffer.rs:36-130 -- the samples live in the GPU's HBM ring
class MulticastRingBuffer {
  public:
    MulticastRingBuffer(std::shared_ptr<GpuEngine> e, size_t buf_size) : e_(std::move(e))
    {
        if (buf_size == 0 || (buf_size & (buf_size - 1))) throw std::invalid_argument("Buffer size must be a power of two");
        const int rc = gb_ring_create(e_->raw(), buf_size);
        if (rc) throw MulticastRingBuffError(rc);
    }
    void write_samples(const std::vector<Complex32>& s)
    {
        const int rc = gb_ring_write(e_->raw(), s.data(), s.size());  // staged through pinned memory
        if (rc) throw MulticastRingBuffError(rc);
    }
    size_t get_head() const { return (size_t)gb_ring_head(e_->raw()); }
    void copy_to_slice(size_t start, Complex32* dest, size_t n) const
    {
        const int rc = gb_ring_copy_to_slice(e_->raw(), start, dest, n);
        if (rc) throw MulticastRingBuffError(rc);
    }
    const std::shared_ptr<GpuEngine>& engine() const { return e_; }

  private:
    std::shared_ptr<GpuEngine> e_;
};

// rf/frontend.rs:6-62 + rf/rf_thread.rs:44-48 -- DC removal + NCO-LUT mix; the processed block goes straight into the ring
// (process_block and write_samples of rf_thread's loop body in one call; the raw block is NOT modified in place).
class DigitalFrontend {
  public:
    DigitalFrontend(std::shared_ptr<GpuEngine> e, float f_if, float fs_in, float /*fs_out*/) : e_(std::move(e))
    {
        const int rc = gb_frontend_configure(e_->raw(), f_if, fs_in);
        if (rc) throw MulticastRingBuffError(rc);
    }
    // false (default): bit-identical to the reference.  true: segmented-scan DC removal, 1e-6 * max|x| from it, 5x the rate.
    void set_parallel(bool on)
    {
        const int rc = gb_frontend_set_mode(e_->raw(), on ? GB_FE_PARALLEL : GB_FE_EXACT);
        if (rc) throw MulticastRingBuffError(rc);
    }
    // raw_floats: interleaved I/Q, a multiple of 16 floats (chunks_exact_mut(16), frontend.rs:34)
    void process_block(const std::vector<float>& raw_floats)
    {
        const int rc = gb_frontend_write(e_->raw(), reinterpret_cast<const gb_c32*>(raw_floats.data()), raw_floats.size() / 2);
        if (rc) throw MulticastRingBuffError(rc);
    }
    // phase_accumulator, bias_re[8], bias_im[8]
    std::array<float, 17> state() const
    {
        std::array<float, 17> st{};
        const int rc = gb_frontend_state(e_->raw(), st.data());
        if (rc) throw MulticastRingBuffError(rc);
        return st;
    }

  private:
    std::shared_ptr<GpuEngine> e_;
};

// acquisition/doppler_shift.rs:5-21 -- built on the host with libm exactly like the reference
struct DopplerShiftTable {
    float doppler_freq_hz;
    std::vector<Complex32> table;
    DopplerShiftTable(float f_if, float doppler_freq, float fs, size_t num_samples)
    {
        const float carr_freq = f_if + doppler_freq;
        const float phase_step = 2.0f * 3.14159265358979323846f * carr_freq / fs;
        table.reserve(num_samples);
        for (size_t i = 0; i < num_samples; i++) {
            const float phase = (float)i * phase_step;
            table.push_back(Complex32{cosf(phase), -sinf(phase)});
        }
        doppler_freq_hz = carr_freq;
    }
};

// do_acquisition.rs:93-116
struct AcquisitionResult {
    uint8_t prn = 0;
    size_t code_phase_samples = 0;
    float code_phase_chips = 0.f, carrier_freq = 0.f, fs = 0.f, mag_relative = 0.f;
    size_t sample_global_index = 0;
};

enum class SearchMode { ColdStart, WarmStart, SteadyState };

// do_acquisition.rs:39-74
class AcquisitionManager {
  public:
    SearchMode mode = SearchMode::ColdStart;
    void update_mode(size_t tracked)
    {
        mode = tracked == 0 ? SearchMode::ColdStart : (tracked <= 4 ? SearchMode::WarmStart : SearchMode::SteadyState);
    }
    std::pair<uint64_t, uint32_t> get_pacing_and_list(const std::set<uint8_t>& active_prns) const
    {
        const uint64_t interval = mode == SearchMode::ColdStart ? 500 : (mode == SearchMode::WarmStart ? 1000 : 2000);
        const size_t size = mode == SearchMode::ColdStart ? PRN_SEARCH_ACQUISITION_TOTAL : (mode == SearchMode::WarmStart ? 8 : 5);
        uint32_t mask = 0;
        size_t taken = 0;
        for (uint8_t prn = 1; prn <= PRN_SEARCH_ACQUISITION_TOTAL && taken < size; prn++) {
            if (active_prns.count(prn)) continue;
            mask |= 1u << (prn - 1);
            taken++;
        }
        return {interval, mask};
    }
};

// do_acquisition.rs:118-239.  One object per PRN keeps the reference's signature; all workers of a
// receiver share one engine, and search_all() is the batched form of the rayon loop at :302-313.
class AcquisitionWorker {
  public:
    AcquisitionWorker(std::shared_ptr<GpuEngine> e, uint8_t prn, size_t fft_size, float freq_sampling_hz)
        : e_(std::move(e)), prn_(prn), fft_size_(fft_size), fs_(freq_sampling_hz)
    {
        // planned once per (fft_size, fs): the other 31 constructors of the reference's worker set return at once
        const int rc = gb_acq_configure_once(e_->raw(), (int)fft_size, freq_sampling_hz, PRN_SEARCH_ACQUISITION_TOTAL);
        if (rc) throw AcqError(rc);
    }

    // Uploads the tables unless these very tables are already on the device.  Returns their identity.
    static uint64_t upload_tables(GpuEngine& e, const std::vector<DopplerShiftTable>& t)
    {
        std::lock_guard<std::mutex> lk(e.mu);
        return upload_tables_locked(e, t);
    }

    // The reference signature.  Thread-safe: the 32 workers of one receiver may call it concurrently on one engine.
    // The first call for a (samples_chunk, doppler_table, local_tail, num_integrations) runs ONE fused search for all
    // 32 PRNs; every later call with the same arguments -- the other workers of the rayon loop -- reads that result.
    std::optional<AcquisitionResult> search_satellite(const std::vector<Complex32>& samples_chunk,
                                                      const std::vector<DopplerShiftTable>& doppler_table, size_t local_tail,
                                                      size_t num_integrations)
    {
        GpuEngine& e = *e_;
        std::lock_guard<std::mutex> lk(e.mu);
        const uint64_t tkey = upload_tables_locked(e, doppler_table);
        uint64_t key = buffer_key(samples_chunk.data(), samples_chunk.size() * sizeof(Complex32), tkey);
        key = buffer_key(&local_tail, 0, key ^ (uint64_t)local_tail * 0x9E3779B97F4A7C15ull ^ (uint64_t)num_integrations << 48);
        if (key != e.search_key || e.search_results.size() != PRN_SEARCH_ACQUISITION_TOTAL) {
            e.search_key = 0;
            e.search_results.assign(PRN_SEARCH_ACQUISITION_TOTAL, gb_acq_result{});
            const int rc = gb_acq_search(e.raw(), samples_chunk.data(), samples_chunk.size(), (int)num_integrations, local_tail,
                                         0xFFFFFFFFu, nullptr, e.search_results.data());
            if (rc) throw AcqError(rc);
            e.search_key = key;
        }
        return convert(e.search_results[prn_ - 1]);
    }

    // all PRNs selected by `mask` in one fused launch (the rayon loop)
    static std::vector<AcquisitionResult> search_all(GpuEngine& e, const std::vector<Complex32>& samples_chunk, size_t local_tail,
                                                     size_t num_integrations, uint32_t mask)
    {
        std::vector<gb_acq_result> res(PRN_SEARCH_ACQUISITION_TOTAL);
        const int rc = gb_acq_search(e.raw(), samples_chunk.data(), samples_chunk.size(), (int)num_integrations, local_tail, mask,
                                     nullptr, res.data());
        if (rc) throw AcqError(rc);
        std::vector<AcquisitionResult> out;
        for (const auto& r : res)
            if (auto a = convert(r)) out.push_back(*a);
        return out;
    }

  private:
    static uint64_t upload_tables_locked(GpuEngine& e, const std::vector<DopplerShiftTable>& t)
    {
        if (t.empty()) throw AcqError(GB_EINVAL);
        uint64_t key = buffer_key(t.data(), 0, (uint64_t)t.size());
        for (const auto& d : t) {
            key = buffer_key(d.table.data(), d.table.size() * sizeof(Complex32), key);
            key = buffer_key(&d.doppler_freq_hz, sizeof(float), key);
        }
        if (key == e.tables_key) return key;
        e.tables_key = 0;
        e.search_key = 0;
        std::vector<Complex32> flat;
        std::vector<float> carr;
        flat.reserve(t.size() * t[0].table.size());
        for (const auto& d : t) {
            flat.insert(flat.end(), d.table.begin(), d.table.end());
            carr.push_back(d.doppler_freq_hz);
        }
        const int rc = gb_acq_set_doppler_tables(e.raw(), flat.data(), carr.data(), (int)carr.size());
        if (rc) throw AcqError(rc);
        e.tables_key = key;
        return key;
    }
    static std::optional<AcquisitionResult> convert(const gb_acq_result& r)
    {
        if (!r.found) return std::nullopt;
        AcquisitionResult a;
        a.prn = r.prn; a.code_phase_samples = r.code_phase_samples; a.code_phase_chips = r.code_phase_chips;
        a.carrier_freq = r.carrier_freq; a.fs = r.fs; a.mag_relative = r.mag_relative;
        a.sample_global_index = r.sample_global_index;
        return a;
    }
    std::shared_ptr<GpuEngine> e_;
    uint8_t prn_;
    size_t fft_size_;
    float fs_;
};

// legacy finer_doppler (acquisition_bk.rs:215-302): refines result.carrier_freq in place from LONG_SAMPLES_LENGTH (11) ms
// of samples; `is_complex` as in the legacy call.  Returns false where the legacy returns None / would panic
// (recording too short, or the peak in the half of the spectrum whose bin table the legacy indexes out of bounds).
inline bool finer_doppler(GpuEngine& e, const std::vector<Complex32>& long_samples, bool is_complex, AcquisitionResult& result,
                          float freq_sampling)
{
    gb_fine_req req{};
    req.prn = result.prn;
    req.code_phase = (uint32_t)result.code_phase_samples;
    gb_fine_result out{};
    const int rc = gb_acq_fine_doppler(e.raw(), long_samples.data(), long_samples.size(), freq_sampling, 11, is_complex ? 1 : 0,
                                       &req, 1, nullptr, &out, nullptr);
    if (rc == GB_ERANGE) return false;
    if (rc) throw AcqError(rc);
    if (!out.ref_defined) return false;
    result.carrier_freq = out.carrier_freq;
    return true;
}

// do_tracking.rs:52-71
struct LoopFilter {
    float tau1, tau2;
    LoopFilter(float noise_bw, float dumping_ratio, float gain) { gb_loop_filter_new(noise_bw, dumping_ratio, gain, &tau1, &tau2); }
    float update(float d_err, float err, float dt) const { return d_err * (dt / tau1) + (d_err - err) * (tau2 / tau1); }
};

enum class TrackingMessageKind { SatelliteLost, SatelliteLocked };
struct TrackingMessage {
    TrackingMessageKind kind;
    uint8_t prn;
};

// do_tracking.rs:88-326.  The state is the POD the kernels read and write (same field names, all public).
class TrackingChannel {
  public:
    gb_trk_channel s;
    TrackingChannel(uint8_t id, float fs) { gb_trk_channel_init(&s, id, fs); }
    // start(AcquisitionResult) (do_tracking.rs:148-154).  The reference's get_ca_chip indexes the C/A table with `prn`
    // instead of `prn - 1` (Q6): a channel started that way correlates with the NEXT satellite's code and never locks.
    // The default here is the satellite's own row; reference_code_row = true reproduces the reference verbatim.
    void start(const AcquisitionResult& r, bool reference_code_row = false)
    {
        gb_acq_result g{};
        g.prn = r.prn; g.found = 1; g.carrier_freq = r.carrier_freq; g.code_phase_chips = r.code_phase_chips;
        g.sample_global_index = r.sample_global_index; g.fs = r.fs;
        const int rc = reference_code_row ? gb_trk_channel_start(&s, &g) : gb_trk_channel_start_corrected(&s, &g);
        if (rc) throw TrackingError(rc);
    }
    bool is_active() const { return s.state == GB_TRK_TRACKING; }
    void reset() { gb_trk_channel_reset(&s); }

    // early_late_correlation on `data_samples` (do_tracking.rs:231-272); returns (i_p,q_p,i_e,q_e,i_l,q_l)
    gb_trk_corr early_late_correlation(GpuEngine& e, const std::vector<Complex32>& data_samples, int mode = GB_TRK_FAST)
    {
        s.num_samples_per_code = data_samples.size();
        const uint64_t off = 0;
        gb_trk_corr out{};
        const int rc = gb_trk_correlate(e.raw(), &s, 1, data_samples.data(), data_samples.size(), &off, mode, &out);
        if (rc) throw TrackingError(rc);
        return out;
    }
    // update(): one do_work if the ring holds the epoch (do_tracking.rs:160-210)
    std::optional<TrackingMessage> update(const MulticastRingBuffer& buff, int mode = GB_TRK_FAST)
    {
        uint8_t ran = 0, lost = 0;
        const int rc = gb_trk_epoch(buff.engine()->raw(), &s, 1, mode, nullptr, &ran, &lost);
        if (rc) throw TrackingError(rc);
        if (lost) return TrackingMessage{TrackingMessageKind::SatelliteLost, s.prn};
        return std::nullopt;
    }
};

// do_tracking.rs:329-382: first-idle assignment + ONE launch for all active channels
class TrackingManager {
  public:
    std::vector<TrackingChannel> channels;
    explicit TrackingManager(float fs, size_t n = NUM_OF_CHANNELS)
    {
        for (size_t id = 0; id < n; id++) channels.emplace_back((uint8_t)id, fs);
    }
    std::vector<TrackingMessage> process_channels(const MulticastRingBuffer& ring, std::vector<AcquisitionResult>& acq_to_trk,
                                                  int mode = GB_TRK_FAST)
    {
        std::vector<TrackingMessage> msgs;
        for (const auto& m : acq_to_trk) {
            for (auto& c : channels)
                if (c.s.state == GB_TRK_IDLE) {
                    msgs.push_back({TrackingMessageKind::SatelliteLocked, m.prn});
                    c.start(m);
                    break;
                }
        }
        acq_to_trk.clear();
        std::vector<gb_trk_channel> pod;
        for (auto& c : channels) pod.push_back(c.s);
        std::vector<uint8_t> ran(pod.size()), lost(pod.size());
        const int rc = gb_trk_epoch(ring.engine()->raw(), pod.data(), (int)pod.size(), mode, nullptr, ran.data(), lost.data());
        if (rc) throw TrackingError(rc);
        for (size_t i = 0; i < pod.size(); i++) {
            channels[i].s = pod[i];
            if (lost[i]) msgs.push_back({TrackingMessageKind::SatelliteLost, pod[i].prn});
        }
        return msgs;
    }
    size_t next_tracking_index() const
    {
        size_t best = 0;
        bool any = false;
        for (const auto& c : channels)
            if (c.is_active()) {
                const size_t v = c.s.next_sample_index + c.s.num_samples_per_code;
                if (!any || v < best) best = v;
                any = true;
            }
        return any ? best : 0;
    }
};

// fft.rs:5-56
class FFT {
  public:
    FFT(std::shared_ptr<GpuEngine> e, size_t len) : e_(std::move(e)), len_(len) {}
    std::vector<Complex32> execute(std::vector<Complex32>& input)
    {
        std::vector<Complex32> out(input.size());
        const int rc = gb_fft_c2c(e_->raw(), (int)len_, 0, input.data(), out.data(), (int)(input.size() / len_));
        if (rc) throw AcqError(rc);
        input = out;  // the reference transforms in place and returns a clone
        return out;
    }
    std::vector<float> power_spectrum(std::vector<Complex32>& input)
    {
        std::vector<float> out(input.size());
        const int rc = gb_fft_power_spectrum(e_->raw(), (int)len_, input.data(), out.data(), (int)(input.size() / len_));
        if (rc) throw AcqError(rc);
        return out;
    }

  private:
    std::shared_ptr<GpuEngine> e_;
    size_t len_;
};

class RealFFT {
  public:
    RealFFT(std::shared_ptr<GpuEngine> e, size_t len) : e_(std::move(e)), len_(len) {}
    std::vector<Complex32> execute(const std::vector<float>& input)
    {
        std::vector<Complex32> out(len_ / 2 + 1);
        const int rc = gb_rfft(e_->raw(), (int)len_, input.data(), out.data(), 1);
        if (rc) throw AcqError(rc);
        return out;
    }
    // fft.rs:47-55: norm_sqr of the len/2 + 1 bins
    std::vector<float> power_spectrum(const std::vector<float>& input)
    {
        std::vector<float> out(len_ / 2 + 1);
        const int rc = gb_rfft_power_spectrum(e_->raw(), (int)len_, input.data(), out.data(), 1);
        if (rc) throw AcqError(rc);
        return out;
    }

  private:
    std::shared_ptr<GpuEngine> e_;
    size_t len_;
};

}  // namespace gnss_sdr_rs
