"""fine Doppler (finer_doppler, acquisition_bk.rs:215-302) and the acquisition -> tracking hand-over.
GPU (through the C-ABI) vs the oracle on the same seeded recordings: FFT index bit-exact (reported when the oracle's
two best magnitudes tie within f32 FFT noise), magnitudes within 1e-3 relative (achieved ~1e-6), carrier frequency
bit-exact given the index (same f32 formula)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL = 1e-3


def _recording(fs, sats, n_ms=11, seed=1, noise=1.0):
    from gnss_sdr_rs_b200 import sdr_mock
    return sdr_mock.baseband(fs, n_ms, sats, seed=seed, noise_sigma=noise)


def _check(gpu, oracle, fs, x, reqs, is_complex=True):
    from gnss_sdr_rs_b200 import acquisition, sdr_mock
    out, mag = acquisition.finer_doppler(gpu, x, reqs, fs, is_complex=is_complex, want_mag=True)
    for i, (prn, cp) in enumerate(reqs):
        ref, rmag = oracle.fine_doppler(x, sdr_mock.ca_code(prn), cp, fs, is_complex=is_complex, want_mag=True)
        assert out[i]["fft_size"] == ref.fft_size
        scale = float(rmag.max())
        assert np.abs(mag[i] - rmag).max() <= REL * scale, np.abs(mag[i] - rmag).max() / scale
        if out[i]["idx"] != ref.idx:
            # only acceptable if the two candidates tie in the oracle to within f32 FFT rounding
            assert abs(float(rmag[out[i]["idx"]]) - scale) <= 4e-6 * scale, (out[i]["idx"], ref.idx)
        else:
            assert out[i]["carrier_freq"] == np.float32(ref.carrier_freq)
            assert out[i]["ref_defined"] == ref.ref_defined
        np.testing.assert_allclose(out[i]["mag"], ref.mag, rtol=REL)
    return out


@pytest.mark.parametrize("fs", [2.048e6, 4.092e6, 16.3676e6])
def test_fine_doppler_matches_oracle(gpu, oracle, fs):
    n = int(round(fs / 1000.0))
    # the legacy convention for complex input: carrier_freq = -(signal frequency).  Positive signal frequencies land in
    # the lower half of the spectrum (ref_defined), negative ones in the half where the legacy panics (still reported)
    sats = [{"prn": 5, "doppler": 1234.5, "code_phase": 321, "cn0_dbhz": 50.0},
            {"prn": 17, "doppler": 2771.0, "code_phase": n - 7, "cn0_dbhz": 47.0},
            {"prn": 9, "doppler": -3020.25, "code_phase": 1000, "cn0_dbhz": 48.0}]
    x = _recording(fs, sats)
    out = _check(gpu, oracle, fs, x, [(s["prn"], s["code_phase"]) for s in sats])
    df = fs / out[0]["fft_size"]
    # lower-half results are the signal frequency to within one fine bin, sign flipped (is_complex)
    assert out[0]["ref_defined"] == 1 and abs(out[0]["carrier_freq"] + 1234.5) <= 2 * df
    assert out[1]["ref_defined"] == 1 and abs(out[1]["carrier_freq"] + 2771.0) <= 2 * df
    # upper half: flagged, value follows the legacy arithmetic (two bins high, :283-295)
    assert out[2]["ref_defined"] == 0 and abs(out[2]["carrier_freq"] - 3020.25) <= 4 * df


def test_fine_doppler_real_if_recording(gpu, oracle):
    """The reference recording's shape (real int8 IF samples, fs 16.3676 MHz, IF 4.1304 MHz): is_complex = false keeps
    the positive sign; the code-stripped carrier sits at the satellite's IF carrier from config.txt:8-17."""
    from gnss_sdr_rs_b200 import sdr_mock
    raw, truth = sdr_mock.if_recording(n_ms=11, prns=[2, 3, 19])
    x = sdr_mock.i8_to_c32(raw)
    reqs = [(t["prn"], t["code_phase"]) for t in truth]
    out = _check(gpu, oracle, sdr_mock.CONFIG_FS, x, reqs, is_complex=False)
    df = sdr_mock.CONFIG_FS / out[0]["fft_size"]
    for o, t in zip(out, truth):
        # a real signal's spectrum is mirror-symmetric: whichever half holds the first maximum, |carrier| is the IF carrier
        assert abs(o["carrier_freq"] - t["carrier"]) <= (2 if o["ref_defined"] else 4) * df, (o, t)


def test_fine_doppler_from_ring_and_errors(gpu, oracle, ffi):
    from gnss_sdr_rs_b200 import acquisition, ring
    fs, n = 2.048e6, 2048
    sats = [{"prn": 3, "doppler": -800.0, "code_phase": 100, "cn0_dbhz": 50.0}]
    x = _recording(fs, sats, n_ms=14)
    r = ring.MulticastRingBuffer(gpu, 1 << 15)     # 32768 samples = 16 ms: the 11 ms window below straddles the wrap
    r.write_samples(np.zeros(10 * n, np.complex64))
    r.write_samples(x)
    start = 10 * n + 3 * n
    got = acquisition.finer_doppler(gpu, start, [(3, 100)], fs)
    want = acquisition.finer_doppler(gpu, x[3 * n:14 * n], [(3, 100)], fs)
    assert got[0]["idx"] == want[0]["idx"] and got[0]["mag"] == want[0]["mag"]
    # the legacy slice [code_phase .. code_phase + 10 N) must fit in the 11 N samples
    with pytest.raises(ffi.GnssB200Error) as e:
        acquisition.finer_doppler(gpu, x[:11 * n], [(3, n + 1)], fs)
    assert e.value.code == ffi.GB_ERANGE
    with pytest.raises(ffi.GnssB200Error) as e:
        acquisition.finer_doppler(gpu, x[:11 * n], [(40, 0)], fs)
    assert e.value.code == ffi.GB_EINVAL


def test_handover_fine_doppler_pulls_in_faster(gpu, oracle):
    """Acquisition (500 Hz bins) -> finer_doppler -> TrackingChannel::start: with the refined carrier the PLL starts
    within a few Hz instead of up to 250 Hz off, so the first epochs already hold most of the prompt energy."""
    from gnss_sdr_rs_b200 import acquisition, ring, tracking
    fs, n = 2.048e6, 2048
    true_dopp = 1730.0
    sats = [{"prn": 7, "doppler": true_dopp, "code_phase": 600, "cn0_dbhz": 52.0}]
    x = _recording(fs, sats, n_ms=300, seed=5)
    eng = acquisition.AcquisitionEngine(gpu, n, fs)
    eng.make_doppler_tables(0.0, np.arange(-5000, 5001, 500, dtype=np.float32))
    res = eng.search(x[:4 * n], 4)[6]
    assert res is not None and res["code_phase_samples"] == 600
    fine = acquisition.finer_doppler(gpu, x[:11 * n], [(7, res["code_phase_samples"])], fs)
    # the legacy's sign convention for complex input: carrier_freq = -(signal frequency)
    assert fine[0]["ref_defined"] == 1
    refined = -float(fine[0]["carrier_freq"])
    assert abs(refined - true_dopp) < 2.0 * fs / fine[0]["fft_size"] + 1.0
    # Q1: the early-exit search reports the FIRST bin whose running best passes the threshold, here far below the truth
    assert abs(res["carrier_freq"] - true_dopp) > 100.0

    def first_epochs(carrier):
        r = ring.MulticastRingBuffer(gpu, 1 << 20)
        r.write_samples(x)
        ch = tracking.channel_array(1, fs)
        # the reference's start() takes BOTH sample_global_index (already at the code start) and
        # code_phase_chips; a consistent hand-over is chips = 0 at that index (and code_row = prn - 1, Q6)
        tracking.start(ch[0], 7, carrier, 0.0, res["sample_global_index"], fs, code_row=6)
        te = tracking.TrackingEngine(gpu)
        te.upload(ch)
        hist = te.run(20, mode=0, want_hist=True)
        te.download(ch)
        i2, q2 = float((hist[:, 0, 0] ** 2).sum()), float((hist[:, 0, 1] ** 2).sum())
        return (i2 + q2) / 20.0, i2 / (i2 + q2), float(ch[0].carrier_freq)

    p_fine, ifrac_fine, f_fine = first_epochs(refined)
    p_coarse, ifrac_coarse, _ = first_epochs(res["carrier_freq"])
    # a coarse start hundreds of Hz off loses most of the 1 ms prompt power and leaves the Costas loop unlocked (energy
    # split between I and Q); the refined start holds the carrier from the first epoch
    assert p_fine > 1.5 * p_coarse, (p_fine, p_coarse)
    assert ifrac_fine > 0.9 and ifrac_fine > ifrac_coarse + 0.2, (ifrac_fine, ifrac_coarse)
    assert abs(f_fine - true_dopp) < 10.0, f_fine
