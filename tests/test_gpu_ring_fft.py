"""Device sample ring (multicast_ring_buffer.rs:147-209 restated) and the FFT facade (fft.rs:5-56)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_multicast_ring_buffer_reference_case(gpu):
    """The reference's own test: 1024-entry ring, writes of 500 / 530 / 20 samples, wrap at 1024."""
    from gnss_sdr_rs_b200 import ring
    rb = ring.MulticastRingBuffer(gpu, 1024)
    rb.write_samples(np.arange(0, 500, dtype=np.float32).astype(np.complex64))
    assert rb.get_head() == 500
    rb.write_samples(np.arange(500, 1030, dtype=np.float32).astype(np.complex64))
    assert rb.get_head() == 1030
    assert (rb.copy_to_slice(1020, 4).real == np.arange(1020, 1024)).all()
    assert (rb.copy_to_slice(1024, 6).real == np.arange(1024, 1030)).all()      # physical 0..6
    assert (rb.copy_to_slice(1020, 10).real == np.arange(1020, 1030)).all()     # across the wrap
    rb.write_samples(np.arange(1030, 1050, dtype=np.float32).astype(np.complex64))
    assert rb.get_head() == 1050
    assert (rb.copy_to_slice(1030, 10).real == np.arange(1030, 1040)).all()
    with pytest.raises(AssertionError):
        ring.MulticastRingBuffer(gpu, 1000)  # not a power of two


def test_ring_i8_and_acquisition_from_ring(gpu, oracle):
    """run()'s data path: head -> local_tail = head - 10*N -> search (do_acquisition.rs:297-313), int8 ingest."""
    from gnss_sdr_rs_b200 import acquisition, ring, sdr_mock
    n, fs, f_if, K = 16368, 16367600.0, 4130400.0, 4
    raw, _ = sdr_mock.if_recording(K + 1, prns={2, 3, 19})
    rb = ring.MulticastRingBuffer(gpu, 1 << 16)   # smaller than the recording: forces a wrap
    rb.write_samples(raw[:30000])
    rb.write_samples(raw[30000:])
    head = rb.get_head()
    assert head == len(raw)
    local_tail = head - K * n
    x = sdr_mock.i8_to_c32(raw)[local_tail:]
    assert (rb.copy_to_slice(local_tail, 100) == x[:100]).all()
    eng = acquisition.AcquisitionEngine(gpu, n, fs)
    d = np.array(acquisition.reference_doppler_grid(), np.float32)
    carr, tabs = oracle.doppler_tables(f_if, d, fs, n)
    eng.set_doppler_tables(tabs, carr)
    a = eng.search_cells_ring(local_tail, K, prn_mask=0b110)
    b = eng.search_cells(x, K, prn_mask=0b110)
    assert a.tobytes() == b.tobytes()
    r = eng.search_ring(local_tail, K, prn_mask=0b110)
    assert r[1] is not None and r[1]["sample_global_index"] == local_tail + r[1]["code_phase_samples"]
    import gnss_sdr_rs_b200._ffi as ffi
    with pytest.raises(ffi.GnssB200Error) as e:
        eng.search_cells_ring(head, K)          # not written yet
    assert e.value.code == ffi.GB_ERANGE
    with pytest.raises(ffi.GnssB200Error) as e:
        eng.search_cells_ring(0, K)             # overwritten
    assert e.value.code == ffi.GB_ERANGE


@pytest.mark.parametrize("n", [1024, 2048, 4092, 4096, 8184, 16368, 20000])
def test_fft_facade(gpu, n):
    from gnss_sdr_rs_b200 import acquisition
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(np.complex64)
    f = acquisition.FFT(gpu, n)
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    got = f.execute(x)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 5e-7
    inv = f.execute(got, inverse=True) / n
    assert np.abs(inv - x).max() < 5e-6
    np.testing.assert_allclose(f.power_spectrum(x), np.abs(ref) ** 2, rtol=2e-5, atol=1e-3)
    r = rng.standard_normal(n).astype(np.float32)
    rf = acquisition.RealFFT(gpu, n)
    ref_r = np.fft.rfft(r.astype(np.float64))
    assert np.abs(rf.execute(r) - ref_r).max() / np.abs(ref_r).max() < 5e-7


@pytest.mark.parametrize("n", [2, 6, 31, 100, 1000, 2046, 5000, 8192, 10230, 65536, 100003])
def test_fft_facade_any_length_f32_and_f64(gpu, n):
    """FFT<T> / RealFFT<T> accept any length and are generic over f32 / f64 (fft.rs:5-56): lengths without a tuned plan
    (and every f64 transform) run the any-length plan -- Bluestein over a power-of-two Stockham FFT -- including primes,
    odd lengths and powers of two above the tuned ones."""
    from gnss_sdr_rs_b200 import acquisition
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
    ref = np.fft.fft(x, axis=1)
    f = acquisition.FFT(gpu, n)
    for dt, tol in ((np.complex64, 3e-6), (np.complex128, 1e-13)):
        xx = x.astype(dt)
        got = f.execute(xx)
        assert got.dtype == dt
        assert np.abs(got - ref).max() / np.abs(ref).max() < tol, (dt, np.abs(got - ref).max() / np.abs(ref).max())
        inv = f.execute(got, inverse=True) / n
        assert np.abs(inv - xx).max() < tol * 30
        ps = f.power_spectrum(xx)
        np.testing.assert_allclose(ps, np.abs(ref) ** 2, rtol=max(30 * tol, 1e-12), atol=np.abs(ref).max() ** 2 * tol)
    r = rng.standard_normal(n)
    ref_r = np.fft.rfft(r)
    rf = acquisition.RealFFT(gpu, n)
    for dt, tol in ((np.float32, 3e-6), (np.float64, 1e-13)):
        got = rf.execute(r.astype(dt))
        assert got.shape == (n // 2 + 1,)
        assert np.abs(got - ref_r).max() / np.abs(ref_r).max() < tol
        np.testing.assert_allclose(rf.power_spectrum(r.astype(dt)), np.abs(ref_r) ** 2, rtol=max(30 * tol, 1e-12),
                                   atol=np.abs(ref_r).max() ** 2 * tol)


@pytest.mark.parametrize("sequential", [False, True])
@pytest.mark.parametrize("f_if,fs", [(4130400.0, 16367600.0), (4092000.0, 16368000.0), (-420000.0, 2048000.0)])
def test_digital_frontend_bit_exact(gpu, oracle, ffi, sequential, f_if, fs):
    """rf/frontend.rs process_block restated -- DC removal + NCO LUT mix, bit-exact incl. the sequential
    f32 phase accumulator, across several rf_thread-sized blocks (2048) and one odd-sized write.  Both NCO forms: the
    phase-orbit table (default; lambda = 6313323 / 4 / 512 for these steps, so the short cycles wrap many times inside
    one write) and the one-thread sequential accumulator (gb_tuning_set("fe_sequential", 1))."""
    from gnss_sdr_rs_b200 import ring
    ffi.tuning_set("fe_sequential", 1 if sequential else 0)
    rng = np.random.default_rng(3)
    raw = ((rng.standard_normal(5 * 2048 + 1000) * 20 + 3.0) + 1j * (rng.standard_normal(5 * 2048 + 1000) * 20 - 2.0)).astype(np.complex64)
    rb = ring.MulticastRingBuffer(gpu, 1 << 15)
    fe = ring.DigitalFrontend(gpu, f_if, fs)
    of = oracle.frontend(f_if, fs)
    ref = []
    pos = 0
    for blk in (2048, 2048, 2048, 1000, 2048, 2048):
        fe.process_block_into_ring(raw[pos:pos + blk])
        ref.append(oracle.frontend_process(of, raw[pos:pos + blk]))
        pos += blk
    ref = np.concatenate(ref)
    got = rb.copy_to_slice(0, pos)
    assert rb.get_head() == pos
    assert got.tobytes() == ref.tobytes()
    st = fe.state()
    assert st["phase_accumulator"] == of.phase_accumulator
    assert (st["bias_re"] == np.array(of.bias_re[:], np.float32)).all() and (st["bias_im"] == np.array(of.bias_im[:], np.float32)).all()
    import gnss_sdr_rs_b200._ffi as ffi
    with pytest.raises(ffi.GnssB200Error):
        fe.process_block_into_ring(raw[:13])  # not a multiple of 8


def test_digital_frontend_long_write_and_reconfigure(gpu, oracle):
    """One 64-tile write (131072 samples, the bench's call size) followed by a ragged one, then a reconfigure that must
    restart the phase orbit and the DC state from zero."""
    from gnss_sdr_rs_b200 import ring
    f_if, fs = 4130400.0, 16367600.0
    rng = np.random.default_rng(9)
    n = 131072 + 8 * 333
    raw = ((rng.standard_normal(n) * 30 - 5.0) + 1j * (rng.standard_normal(n) * 30 + 7.0)).astype(np.complex64)
    rb = ring.MulticastRingBuffer(gpu, 1 << 18)
    for _ in range(2):
        fe = ring.DigitalFrontend(gpu, f_if, fs)
        of = oracle.frontend(f_if, fs)
        start = rb.get_head()
        fe.process_block_into_ring(raw[:131072])
        fe.process_block_into_ring(raw[131072:])
        ref = np.concatenate([oracle.frontend_process(of, raw[:131072]), oracle.frontend_process(of, raw[131072:])])
        assert rb.copy_to_slice(start, n).tobytes() == ref.tobytes()
        assert fe.state()["phase_accumulator"] == of.phase_accumulator


@pytest.mark.parametrize("f_if,fs", [(4130400.0, 16367600.0), (4092000.0, 16368000.0)])
def test_digital_frontend_parallel_mode_within_tolerance(gpu, oracle, ffi, f_if, fs):
    """GB_FE_PARALLEL (segmented-scan DC removal, three multi-CTA launches): NOT bit-exact by construction -- samples
    and bias state within 1e-5 * max|x| of rf/frontend.rs (float32 rounding of the segment composition), the NCO phase
    exact.  Call sizes: one step, a ragged segment, several segments + ragged tail, the bench's 131072, and a ring wrap."""
    from gnss_sdr_rs_b200 import ring
    ffi.tuning_set("fe_sequential", 0)
    rng = np.random.default_rng(17)
    sizes = (8, 8 * 5, 256, 8 * 100, 2048, 131072, 8 * 4097, 131072)
    n = sum(sizes)
    raw = ((rng.standard_normal(n) * 25 + 40.0) + 1j * (rng.standard_normal(n) * 25 - 60.0)).astype(np.complex64)
    scale = float(np.abs(raw.view(np.float32)).max())
    rb = ring.MulticastRingBuffer(gpu, 1 << 18)     # 262144 < n: the last writes wrap
    fe = ring.DigitalFrontend(gpu, f_if, fs, parallel=True)
    of = oracle.frontend(f_if, fs)
    pos = 0
    worst = 0.0
    for blk in sizes:
        fe.process_block_into_ring(raw[pos:pos + blk])
        ref = oracle.frontend_process(of, raw[pos:pos + blk])
        got = rb.copy_to_slice(pos, blk)
        worst = max(worst, float(np.abs(got.view(np.float32) - ref.view(np.float32)).max()))
        pos += blk
        st = fe.state()
        assert st["phase_accumulator"] == of.phase_accumulator
        np.testing.assert_allclose(st["bias_re"], np.array(of.bias_re[:], np.float32), rtol=0, atol=1e-5 * scale)
        np.testing.assert_allclose(st["bias_im"], np.array(of.bias_im[:], np.float32), rtol=0, atol=1e-5 * scale)
    assert worst <= 1e-5 * scale, (worst, scale)
    print("parallel front-end: worst |delta| = %.3g (%.3g of max|x|)" % (worst, worst / scale))
    # back to the default: bit-exact again from a fresh configure
    fe = ring.DigitalFrontend(gpu, f_if, fs)
    of = oracle.frontend(f_if, fs)
    start = rb.get_head()
    fe.process_block_into_ring(raw[:4096])
    assert rb.copy_to_slice(start, 4096).tobytes() == oracle.frontend_process(of, raw[:4096]).tobytes()
    with pytest.raises(ffi.GnssB200Error):
        gpu.call("gb_frontend_set_mode", 7)


def test_fresh_ring_survives_a_busy_default_stream(gpu):
    """Regression: gb_ring_create used to clear the ring with cudaMemset, which runs asynchronously on the legacy default
    stream; the handle's copy stream is non-blocking and does not wait for it, so with the default stream busy (here: a
    few large torch matmuls, torch's current stream IS the legacy default stream) the memset landed after the first
    write_samples and wiped them.  The same holds for the front-end state cleared by gb_frontend_configure."""
    import torch
    from gnss_sdr_rs_b200 import ring
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(1 << 16) + 1j * rng.standard_normal(1 << 16)).astype(np.complex64)
    a = torch.randn(8192, 8192, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        b = a
        for _ in range(4):
            b = b @ a                                   # ~40 ms of queued work on the default stream
        rb = ring.MulticastRingBuffer(gpu, 1 << 22)     # 32 MB to clear
        rb.write_samples(x)
        fe = ring.DigitalFrontend(gpu, 0.0, 2.048e6)
        fe.process_block_into_ring(x[:4096])
        st = fe.state()
        got = rb.copy_to_slice(0, x.size)
        assert got.tobytes() == x.tobytes()
        assert np.abs(st["bias_re"]).max() > 0.0        # the state the kernel left, not a late memset's zeros
        del b
    torch.cuda.synchronize()
