"""The oracle against every golden vector / known-answer test the reference holds for this path
 and against independent NumPy/SciPy restatements.  CPU only."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest
import scipy.fft

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_prn1_known_answer(oracle):
    """src/bk/gps_ca_prn.rs:72-123: the full 1023-chip PRN-1 vector."""
    gold = json.load(open(os.path.join(GOLD, "prn1_ca_code.json")))["chips"]
    assert oracle.ca_table()[0].tolist() == gold


def test_ca_table_sha256_and_heads(oracle):
    """constants/gps_ca_constants.rs: sha256 of the whole 32 x 1023 table + first 16 chips of each row."""
    gold = json.load(open(os.path.join(GOLD, "ca_table.json")))
    t = oracle.ca_table()
    assert hashlib.sha256(t.tobytes()).hexdigest() == gold["sha256"]
    assert t[:, :16].tolist() == gold["first16"]
    assert set(np.unique(t)) == {-1, 1}


@pytest.mark.skipif(not os.path.exists("/root/reference/src/constants/gps_ca_constants.rs"), reason="authoring container only")
def test_ca_table_equals_reference_file(oracle):
    import re
    src = open("/root/reference/src/constants/gps_ca_constants.rs").read()
    vals = [int(v) for v in re.findall(r"-?\d+", src[src.index("= [") + 3:])]
    assert oracle.ca_table().ravel().tolist() == vals


def test_prn_out_of_range(oracle):
    buf = np.zeros(1023, np.int8)
    L = oracle.lib()
    assert L.go_ca_code_chips(40, buf.ctypes.data_as(C.c_void_p)) != 0  # should_panic at gps_ca_prn.rs:65-70
    assert L.go_ca_code_chips(0, buf.ctypes.data_as(C.c_void_p)) != 0


def test_acquisition_manager_reference_tests(oracle):
    """do_acquisition.rs:339-395 restated."""
    L = oracle.lib()
    m = oracle.AcqManager(0)
    iv, mask = C.c_uint64(), C.c_uint32()
    L.go_acq_manager_pacing(C.byref(m), 0, C.byref(iv), C.byref(mask))
    assert (iv.value, mask.value) == (500, 0xFFFFFFFF)          # cold start
    L.go_acq_manager_update_mode(C.byref(m), 3)
    assert m.mode == 1
    L.go_acq_manager_pacing(C.byref(m), 0b111, C.byref(iv), C.byref(mask))
    assert (iv.value, mask.value) == (1000, 2040)                # warm start, PRN 1-3 active
    L.go_acq_manager_update_mode(C.byref(m), 5)
    assert m.mode == 2
    L.go_acq_manager_pacing(C.byref(m), 0, C.byref(iv), C.byref(mask))
    assert (iv.value, mask.value) == (2000, 0b11111)
    L.go_acq_manager_update_mode(C.byref(m), 0)
    assert m.mode == 0


def test_ring_buffer_reference_test(oracle):
    """utilities/multicast_ring_buffer.rs:147-209 restated."""
    L = oracle.lib()
    r = oracle.Ring()
    assert L.go_ring_init(C.byref(r), 1000) != 0  # power of two required
    assert L.go_ring_init(C.byref(r), 1024) == 0

    def write(lo, hi):
        a = np.arange(lo, hi, dtype=np.float32).astype(np.complex64)
        L.go_ring_write(C.byref(r), a.ctypes.data_as(C.c_void_p), len(a))

    def phys(lo, hi):
        return np.ctypeslib.as_array(C.cast(r.buffer, C.POINTER(C.c_float)), shape=(1024, 2))[lo:hi, 0].copy()

    write(0, 500)
    assert L.go_ring_head(C.byref(r)) == 500
    write(500, 1030)
    assert L.go_ring_head(C.byref(r)) == 1030
    assert (phys(1020, 1024) == np.arange(1020, 1024)).all()
    assert (phys(0, 6) == np.arange(1024, 1030)).all()
    dest = np.zeros(10, np.complex64)
    L.go_ring_copy_to_slice(C.byref(r), 1020, dest.ctypes.data_as(C.c_void_p), 10)
    assert (dest.real == np.arange(1020, 1030)).all()
    write(1030, 1050)
    assert L.go_ring_head(C.byref(r)) == 1050
    assert (phys(6, 16) == np.arange(1030, 1040)).all()
    L.go_ring_free(C.byref(r))


def test_loop_filter_constants(oracle):
    """do_tracking.rs:16-29, 59-70: tau1/tau2 of the PLL (25 Hz, 0.7, 0.25) and DLL (2 Hz, 0.7, 1.0)."""
    L = oracle.lib()
    pll = L.go_loop_filter_new(25.0, 0.7, 0.25)
    dll = L.go_loop_filter_new(2.0, 0.7, 1.0)
    f = np.float32
    for flt, bw, z, g in ((pll, 25.0, 0.7, 0.25), (dll, 2.0, 0.7, 1.0)):
        w = f(bw) * f(8.0) * f(z) / (f(4.0) * f(z) * f(z) + f(1.0))
        assert flt.tau1 == f(g) / (w * w) and flt.tau2 == (f(2.0) * f(z)) / w
    assert abs(pll.tau1 - 1.117551e-4) < 1e-9 and abs(pll.tau2 - 0.0296) < 1e-6
    assert abs(dll.tau1 - 0.06984694) < 1e-7 and abs(dll.tau2 - 0.37) < 1e-6
    assert abs(0.001 / pll.tau1 - 8.948138) < 1e-4 and abs(pll.tau2 / pll.tau1 - 264.86487) < 1e-2
    out = L.go_loop_filter_update(C.byref(pll), 0.01, 0.0, 0.001)
    assert out == f(0.01) * (f(0.001) / f(pll.tau1)) + (f(0.01) - f(0.0)) * (f(pll.tau2) / f(pll.tau1))


def test_ca_code_resampling_q3(oracle):
    """ca_code.rs:12-27.  Q3: at 4.092 MHz the f32 index arithmetic differs from exact arithmetic at 93 samples."""
    t = oracle.ca_table()
    for fs, n, mism in ((2.048e6, 2048, 0), (4.096e6, 4096, 0), (16367600.0, 16368, 0), (20e6, 20000, 0), (4.092e6, 4092, 93)):
        s = oracle.ca_code_samples(7, 1.023e6, fs)
        assert len(s) == n
        x = np.arange(n, dtype=np.float32)
        idx32 = np.floor((x * np.float32(1.023e6)) / np.float32(fs)).astype(np.int64)
        assert (s == t[6][idx32]).all()
        exact = (np.arange(n, dtype=np.int64) * 1023000) // int(fs)
        assert int((idx32 != exact).sum()) == mism


@pytest.mark.parametrize("n", [4, 12, 31, 64, 1023, 2048, 4092, 4096, 16368, 20000, 2039, 789])
def test_fft_against_scipy_and_f64(oracle, n):
    """The FFT that stands in for rustfft 6.1.0: unnormalised forward / inverse, any length."""
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ref = scipy.fft.fft(x.astype(np.complex128))
    refi = scipy.fft.ifft(x.astype(np.complex128)) * n
    assert np.abs(oracle.fft(x) - ref).max() / np.abs(ref).max() < 5e-7
    assert np.abs(oracle.fft(x, True) - refi).max() / np.abs(refi).max() < 5e-7
    assert np.abs(oracle.fft64(x) - ref).max() / np.abs(ref).max() < 1e-14
    assert np.abs(oracle.fft(x) - scipy.fft.fft(x)).max() / np.abs(ref).max() < 5e-7  # pocketfft f32


def test_fft_facade(oracle):
    """fft.rs:5-56: forward c2c, power spectrum, real FFT with n/2+1 bins."""
    L = oracle.lib()
    n = 1000
    rng = np.random.default_rng(0)
    r = rng.standard_normal(n).astype(np.float32)
    out = np.zeros(n // 2 + 1, np.complex64)
    L.go_rfft_forward(n, r.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    ref = np.fft.rfft(r.astype(np.float64))
    assert np.abs(out - ref).max() / np.abs(ref).max() < 5e-7
    x = r.astype(np.complex64)
    p = np.zeros(n, np.float32)
    L.go_fft_power_spectrum(n, x.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
    np.testing.assert_allclose(p, np.abs(np.fft.fft(r.astype(np.float64))) ** 2, rtol=1e-4, atol=1e-3)


def test_doppler_table_and_apply(oracle):
    """doppler_shift.rs:11-58: table = (cos, -sin)(i*step), stored freq includes IF; 4-sample chunks, stale tail."""
    f_if, fd, fs, n = np.float32(4130400.0), np.float32(-2500.0), np.float32(16367600.0), 16368
    carr, tab = oracle.doppler_table(f_if, fd, fs, n)
    assert carr == f_if + fd
    step = np.float32(2.0) * np.float32(np.pi) * (f_if + fd) / fs
    phase = np.arange(n, dtype=np.float32) * step
    assert phase.dtype == np.float32
    ref = np.cos(phase.astype(np.float64)) - 1j * np.sin(phase.astype(np.float64))
    assert np.abs(tab - ref).max() < 1.2e-7  # correctly-rounded libm on the f32 phase
    rng = np.random.default_rng(1)
    s = (rng.standard_normal(10) + 1j * rng.standard_normal(10)).astype(np.complex64)
    out = np.full(10, 99 + 99j, np.complex64)
    oracle.apply_doppler_shift(s, tab[:10], out)
    a, b, c, d = s.real[:8], s.imag[:8], tab.real[:8], tab.imag[:8]
    assert (out.real[:8] == a * c + (b * d) * np.float32(-1)).all() and (out.imag[:8] == a * d + b * c).all()
    assert (out[8:] == 99 + 99j).all()  # A3: len % 4 tail is not written


def _numpy_cells(x, tabs, code, K, n):
    """Independent f64 restatement of search_satellite's per-bin arithmetic (do_acquisition.rs:171-202, 229-234)."""
    cf = np.conj(scipy.fft.fft(code.astype(np.float64)))
    peaks, args, sums = [], [], []
    for t in tabs.astype(np.complex128):
        acc = np.zeros(n)
        for k in range(K):
            y = scipy.fft.ifft(scipy.fft.fft(x[k * n:(k + 1) * n].astype(np.complex128) * t) * cf) * n
            acc += np.abs(y) ** 2
        peaks.append(acc.max()); args.append(int(acc.argmax())); sums.append(acc[:(n // 8) * 8].sum())
    return np.array(peaks), np.array(args), np.array(sums)


@pytest.mark.parametrize("fs,n", [(2.048e6, 2048), (4.092e6, 4092)])
def test_acquisition_cells_against_numpy_f64(oracle, fs, n):
    from gnss_sdr_rs_b200 import sdr_mock
    K = 3
    sats = [{"prn": 4, "doppler": 900.0, "code_phase": 1234, "cn0_dbhz": 50.0}]
    x = sdr_mock.baseband(fs, K, sats, seed=3)
    carr, tabs = oracle.doppler_tables(0.0, np.array([0.0, 500.0, 1000.0, 1500.0], np.float32), fs, n)
    w = oracle.AcqWorker(4, n, fs)
    cells = w.cells(x, tabs, K)
    pk, ag, sm = _numpy_cells(x, tabs, oracle.ca_code_samples(4, 1.023e6, fs), K, n)
    assert (cells["argmax"] == ag).all()
    np.testing.assert_allclose(cells["peak"], pk, rtol=5e-6)
    np.testing.assert_allclose(cells["sum8"], sm, rtol=5e-6)   # Q2: the last n % 8 bins are excluded
    assert cells["argmax"][2] == 1234


def test_search_early_exit_equals_decide_and_config_txt(oracle):
    """Q1 on the stand-in for the bundled recording (config.txt:8-17): the early-exit search, the full-grid
    decide(), and the table of PRNs / code phases agree; detections are a subset of the true satellites."""
    from gnss_sdr_rs_b200 import acquisition, sdr_mock
    gold = json.load(open(os.path.join(GOLD, "config_txt.json")))
    assert [(r["prn"], round(r["carrier_mhz"] * 1e6), r["code_phase"]) for r in gold["rows"]] == \
        [(p, round(c), ph) for p, c, ph in sdr_mock.CONFIG_TXT]
    assert gold["fs"] == sdr_mock.CONFIG_FS and gold["if"] == sdr_mock.CONFIG_IF
    n, fs, f_if, K = 16368, 16367600.0, 4130400.0, 10
    raw, truth = sdr_mock.if_recording(K, prns={2, 3, 19, 14})
    x = sdr_mock.i8_to_c32(raw)
    d = np.array(acquisition.reference_doppler_grid(), np.float32)
    assert len(d) == 29 and d[0] == -7000 and d[-1] == 7000
    carr, tabs = oracle.doppler_tables(f_if, d, fs, n)
    prns = [1, 2, 3, 14, 19, 22]
    workers = [oracle.AcqWorker(p, n, fs) for p in prns]
    early = oracle.acq_search_all(workers, x, tabs, carr, 77, K, early_exit=True)
    cells = oracle.acq_cells_all(workers, x, tabs, K)
    for i, p in enumerate(prns):
        dec = oracle.acq_decide(cells[i], carr, p, n, fs, local_tail=77)
        assert (dec is None) == (early[i] is None)
        if dec:
            for k in ("code_phase_samples", "carrier_freq", "mag_relative", "sample_global_index", "code_phase_chips"):
                assert dec[k] == early[i][k]
            assert dec["sample_global_index"] == 77 + dec["code_phase_samples"]
            assert dec["code_phase_chips"] == np.float32(dec["code_phase_samples"]) * np.float32(1.023e6) / np.float32(fs)
            t = next(t for t in truth if t["prn"] == p)
            assert abs(dec["code_phase_samples"] - t["code_phase"]) <= 3
            assert dec["bin"] <= int(cells[i]["peak"].argmax())  # first passing record, not the global max
    found = {p for i, p in enumerate(prns) if early[i]}
    assert found == {2, 3, 14, 19}


def test_coherent_presum_equals_postsum(oracle):
    from gnss_sdr_rs_b200 import sdr_mock
    fs, n, K, n_coh = 2.048e6, 2048, 8, 4
    x = sdr_mock.baseband(fs, K, [{"prn": 9, "doppler": 333.0, "code_phase": 50, "cn0_dbhz": 40.0}], seed=8)
    carr, tabs = oracle.doppler_tables(0.0, np.arange(0, 700, 125, dtype=np.float32), fs, n)
    rot = oracle.coh_rotators(carr, fs, n, n_coh)
    w = oracle.AcqWorker(9, n, fs)
    a = w.cells(x, tabs, K, n_coh=n_coh, rot=rot, presum=0)
    b = w.cells(x, tabs, K, n_coh=n_coh, rot=rot, presum=1)
    np.testing.assert_allclose(a["peak"], b["peak"], rtol=1e-5)
    assert (a["argmax"] == b["argmax"]).all()
    one = w.cells(x, tabs, K)
    best = int(a["peak"].argmax())
    assert abs(float(carr[best]) - 333.0) <= 125 and a["argmax"][best] == 50
    assert a["peak"][best] / np.median(a["peak"]) > one["peak"][best] / np.median(one["peak"])  # coherent gain


def test_tracking_quirks_and_loops(oracle):
    """do_tracking.rs:148-154, 183-210, 231-302: Q6 (row = prn), Q7 (saturating cast), Q8, Q9, Q10 and a closed loop
    on a noise-free signal pulling a 50 Hz carrier error in."""
    from gnss_sdr_rs_b200 import sdr_mock
    fs, n = 4.096e6, 4096
    ch = oracle.trk_channel(3, fs)
    assert ch.num_samples_per_code == n and ch.code_rate == np.float32(1.023e6) and ch.state == 0
    oracle.trk_start(ch, 2, 2950.0, 0.25, 1000, fs)
    assert (ch.prn, ch.code_row, ch.state, ch.next_sample_index) == (2, 2, 1, 1000) and ch.code_phase == np.float32(0.25)
    L = oracle.lib()
    t = oracle.ca_table()
    assert L.go_trk_get_ca_chip(C.byref(ch), -0.5) == t[2][0]      # Q7
    assert L.go_trk_get_ca_chip(C.byref(ch), 1023.25) == t[2][0]   # 1023 % 1023
    assert L.go_trk_get_ca_chip(C.byref(ch), 5.99) == t[2][5]      # Q6: row prn, i.e. PRN 3's code
    # the reference's own synthetic test signal (do_tracking.rs:464-520): positive error -> positive NCO
    code = oracle.ca_code_samples(2, 1.023e6, fs)
    sig = sdr_mock.reference_test_signal(code, 3000.0, 0.0, 0.0, fs)
    assert len(sig) == n
    ch = oracle.trk_channel(0, fs)
    oracle.trk_start(ch, 2, 2950.0, 0.0, 0, fs)
    ch.code_row = 1  # correlate with the code the signal was built from (the reference test relies on Q6 by accident)
    out, msg, _ = oracle.trk_do_work(ch, sig)
    assert msg == 0 and ch.next_sample_index == n
    assert ch.carrier_error > 0 and ch.carrier_nco > 0 and ch.carrier_freq > 2950.0
    # loss of lock: 20 epochs of zeros (Q9: reset zeroes prn and code_rate; Q10: phases advance meanwhile)
    ch2 = oracle.trk_channel(1, fs)
    oracle.trk_start(ch2, 5, 100.0, 0.0, 0, fs)
    z = np.zeros(n, np.complex64)
    for e in range(19):
        _, msg, _ = oracle.trk_do_work(ch2, z)
        assert msg == 0 and ch2.lost_counter == e + 1 and ch2.carrier_phase != 0.0
    _, msg, mp = oracle.trk_do_work(ch2, z)
    assert msg == 1 and mp == 0 and ch2.state == 0 and ch2.code_rate == 0.0 and ch2.prn == 0


def test_tracking_closed_loop_pull_in(oracle):
    from gnss_sdr_rs_b200 import sdr_mock
    fs, n_ms = 2.048e6, 400
    x = sdr_mock.baseband(fs, n_ms, [{"prn": 8, "doppler": 1000.0, "code_phase": 200, "cn0_dbhz": 50.0}], seed=4)
    chs = (oracle.TrkChannel * 1)()
    c = oracle.trk_channel(0, fs)
    oracle.trk_start(c, 8, 1000.0 - 30.0, 0.1, 200, fs)
    c.code_row = 7
    chs[0] = c
    hist = oracle.trk_run_all(chs, x, 380)
    assert chs[0].state == 1 and abs(chs[0].carrier_freq - 1000.0) < 5.0
    p = np.hypot(hist[100:, 0, 0], hist[100:, 0, 1])
    assert p.min() > 4.0  # prompt power stays far above LOCK_THRESHOLD = 15 (do_tracking.rs:741)


def test_digital_frontend_against_numpy(oracle):
    """rf/frontend.rs:32-62 / dc_remove.rs:23-29 / nco_lut.rs:8-42 restated independently in NumPy f32."""
    f_if, fs = np.float32(4130400.0), np.float32(16367600.0)
    f = oracle.frontend(f_if, fs)
    i = np.arange(2048, dtype=np.float32)
    ang = (np.float32(2.0) * np.float32(np.pi) * i) / np.float32(2048.0)
    assert np.abs(np.array(f.lut_re[:]) - np.cos(ang.astype(np.float64))).max() < 1e-7
    assert np.abs(np.array(f.lut_im[:]) + np.sin(ang.astype(np.float64))).max() < 1e-7
    assert f.phase_step == (f_if / fs) * np.float32(2048.0) and f.alpha == np.float32(0.001)
    rng = np.random.default_rng(0)
    x = ((rng.standard_normal(4096 + 5) + 0.5) + 1j * (rng.standard_normal(4096 + 5) - 0.25)).astype(np.complex64)
    y = oracle.frontend_process(f, x)
    # independent restatement
    lut_re, lut_im = np.array(f.lut_re[:], np.float32), np.array(f.lut_im[:], np.float32)
    alpha, con = np.float32(0.001), np.float32(1.0) - np.float32(0.001)
    bre, bim = np.zeros(8, np.float32), np.zeros(8, np.float32)
    acc, step = np.float32(0.0), f.phase_step
    out = x.copy()
    for c in range(0, (len(x) // 8) * 8, 8):
        re, im = x.real[c:c + 8].copy(), x.imag[c:c + 8].copy()
        bre = bre * con + re * alpha
        bim = bim * con + im * alpha
        re, im = re - bre, im - bim
        for j in range(8):
            k = int(acc) % 2048
            acc = np.float32(np.fmod(np.float32(acc + step), np.float32(2048.0)))
            out[c + j] = complex(np.float32(re[j] * lut_re[k]) + np.float32(im[j] * lut_im[k]),
                                 np.float32(re[j] * lut_im[k]) - np.float32(im[j] * lut_re[k]))
    assert y.tobytes() == out.astype(np.complex64).tobytes()
    assert (y[-5:] == x[-5:]).all()  # chunks_exact_mut(16 floats): the tail is left raw
    assert f.phase_accumulator == acc


def test_nav_bit_sync_against_python(oracle):
    """SURVEY 8f N4 (legacy decoding.rs:115-127, 164-213): independent restatement on a synthetic prompt sequence."""
    rng = np.random.default_rng(2)
    n, offset = 4000, 7                      # bit edges at epochs == 7 (mod 20)
    bits = rng.integers(0, 2, n // 20 + 2) * 2 - 1
    ip = np.array([bits[(e - offset) // 20 + 1] * 1000.0 for e in range(n)], np.float32) + rng.standard_normal(n).astype(np.float32) * 50
    st, out = oracle.nav_bit_sync(ip, 512)
    buff = [0] * 20
    sync, ind, acc, got, old = False, 0, 0.0, [], np.float32(0)
    for cnt in range(n):
        biti = cnt % 20
        if not sync and cnt > 1000 and old * ip[cnt] < 0:
            buff[biti] += 1
            vmax = max(buff)
            ind = max(i for i in range(20) if buff[i] == vmax)
            sync = vmax == 30
        if sync:
            acc = ip[cnt] if biti == ind else np.float32(acc + ip[cnt])
            if biti == (ind + 19) % 20:
                got.append(1 if acc > 0 else -1)
        old = ip[cnt]
    assert st.flag_bit_sync == 1 and st.frame_sync_ind == ind == offset
    assert out.tolist() == got and list(st.bit_sync_buff) == buff
    first = (st.sync_epoch - offset) // 20 + 1   # synchronisation is declared ON a bit edge: that bit is the first one out
    assert out.tolist()[:20] == bits[first:first + 20].tolist()


def test_preamble_search_against_python(oracle):
    """check_preamble_syn (decoding.rs:215-226): |sum_{x<8} bits[i0+x] * GPS_CA_PREAMBLE[x]| == 8; the sliding search and
    the legacy's literal single test of the first 8 bits (its buff_preamble is pushed to but never popped)."""
    pre = [1, -1, -1, -1, 1, -1, 1, 1]
    rng = np.random.default_rng(9)
    for planted_at, pol in ((0, 1), (13, -1), (None, 0)):
        nav = (rng.integers(0, 2, 120) * 2 - 1).tolist()
        for i0 in range(0, 112):     # remove accidental matches
            c = sum(nav[i0 + x] * pre[x] for x in range(8))
            if abs(c) == 8:
                nav[i0] = -nav[i0]
        if planted_at is not None:
            nav[planted_at:planted_at + 8] = [pol * v for v in pre]
        # prompt sequence whose bit edges sit at epoch % 20 == 0, long enough to synchronise (after epoch 1000), then nav
        lead = (rng.integers(0, 2, 70) * 2 - 1).tolist()
        lead = [(-1) ** k for k in range(70)]                       # alternating: an edge every 20 ms -> sync at once
        seq = lead + nav
        ip = np.repeat(np.array(seq, np.float32) * 1000.0, 20)
        st, bits = oracle.nav_bit_sync(ip, 512)
        assert st.flag_bit_sync == 1 and st.frame_sync_ind == 0
        first_bit = (st.sync_epoch // 20)                            # bits emitted from the bit that contains sync_epoch
        emitted = seq[first_bit:first_bit + st.n_bits]
        assert bits.tolist() == emitted[:len(bits)]
        # python restatement of both searches on the emitted bits
        hits = [i0 for i0 in range(len(emitted) - 7) if abs(sum(emitted[i0 + x] * pre[x] for x in range(8))) == 8]
        exp_first = hits[0] if hits else -1
        assert st.preamble_bit == exp_first
        if exp_first >= 0:
            assert st.polarity == (1 if sum(emitted[exp_first + x] * pre[x] for x in range(8)) > 0 else -1)
        assert st.ref_frame_sync == (1 if (hits and hits[0] == 0) else 0)


def test_reference_arithmetic_diverges_when_fft_size_is_not_a_multiple_of_4(oracle):
    """apply_doppler_shift writes 4 * floor(len / 4) samples (doppler_shift.rs:26); for len % 4 != 0 the tail of the
    worker's result_buf keeps the previous block's UNNORMALISED inverse FFT, which is fed back N times larger each block.
    The oracle restates that faithfully: at N = 2046 (2 samples per chip) the accumulated powers are inf / NaN after one
    10-block search -- which is why gb_acq_configure refuses these lengths instead of "reproducing" them."""
    from gnss_sdr_rs_b200 import sdr_mock
    n, fs, K = 2046, 2.046e6, 10
    x = sdr_mock.baseband(fs, K, [{"prn": 5, "doppler": 500.0, "code_phase": 100, "cn0_dbhz": 50.0}], seed=1)
    carr, tabs = oracle.doppler_tables(0.0, np.arange(-1000, 1001, 500, dtype=np.float32), fs, n)
    cells = oracle.AcqWorker(5, n, fs).cells(x, tabs, K)
    assert not np.isfinite(cells["peak"]).all() or cells["peak"].max() > 1e30
    # a multiple of 4 right next to it is perfectly fine
    n2, fs2 = 2048, 2.048e6
    x2 = sdr_mock.baseband(fs2, K, [{"prn": 5, "doppler": 500.0, "code_phase": 100, "cn0_dbhz": 50.0}], seed=1)
    carr2, tabs2 = oracle.doppler_tables(0.0, np.arange(-1000, 1001, 500, dtype=np.float32), fs2, n2)
    c2 = oracle.AcqWorker(5, n2, fs2).cells(x2, tabs2, K)
    assert np.isfinite(c2["peak"]).all() and int(c2["argmax"][3]) == 100


# ------------------------------------------------------------------ N3: finer_doppler (acquisition_bk.rs:215-302)
@pytest.mark.parametrize("fs,dopp", [(2.048e6, 1234.5), (4.092e6, -2771.0)])
def test_fine_doppler_oracle_vs_numpy_f64(oracle, fs, dopp):
    """The oracle's restatement against an independent NumPy f64 evaluation of the same steps (mean removal, f32 code
    index, zero-padded 8x FFT, first arg-max) and the legacy frequency mapping, including the half where it panics."""
    from gnss_sdr_rs_b200 import sdr_mock
    n = int(round(fs / 1000.0))
    cp = 345
    x = sdr_mock.baseband(fs, 11, [{"prn": 12, "doppler": dopp, "code_phase": cp, "cn0_dbhz": 50.0}], seed=3)
    code = sdr_mock.ca_code(12)
    res, mag = oracle.fine_doppler(x, code, cp, fs, want_mag=True)
    use = 10 * n
    p2 = 1 << int(np.ceil(np.log2(use)))
    assert res.fft_size == 8 * p2
    idx = np.floor(np.arange(use, dtype=np.float32) * np.float32(1.023e6) / np.float32(fs)).astype(np.int64) % 1023
    buf = np.zeros(8 * p2, np.complex128)
    buf[:use] = (x.astype(np.complex128) - x.astype(np.complex128).mean())[cp:cp + use] * code[idx]
    M = np.abs(np.fft.fft(buf))
    assert np.abs(M - mag).max() <= 2e-6 * M.max()
    assert int(M.argmax()) == res.idx
    df = fs / res.fft_size
    one_side = res.fft_size // 2 + 1
    if dopp > 0:      # positive signal frequency: lower half, defined, sign flipped for complex input (:296-298)
        assert res.ref_defined == 1 and res.idx < one_side
        assert res.carrier_freq == -np.float32(np.float32(res.idx) * np.float32(fs) / np.float32(res.fft_size))
        assert abs(res.carrier_freq + dopp) <= df
    else:             # upper half: the legacy indexes fft_freq_bins out of bounds (:283-287)
        assert res.ref_defined == 0 and res.idx >= one_side
        assert abs(res.carrier_freq - (res.fft_size - res.idx + 2) * df) < 1e-3 * df + 1e-2
    # too short a recording: the legacy slice would panic -> None
    assert oracle.fine_doppler(x[:10 * n + cp - 1], code, cp, fs)[0] is None


# ------------------------------------------------------------------ A7: two-peak window (acquisition_bk.rs:342-399)
def _legacy_two_peaks(row, spc):
    """Literal slicing of satellite_detection_two_peaks for one Doppler row (acquisition_bk.rs:367-395)."""
    n = len(row)
    cp = int(np.argmax(row))
    left, right = cp - spc, cp + spc
    if left < 1:
        new = row[right - 1:n + left]
    elif right >= n:
        new = row[right - n - 1:left]
    else:
        new = np.concatenate([row[0:left], row[right:n]])
    return cp, float(row[cp]), float(new.max())


@pytest.mark.parametrize("cp", [0, 1, 3, 4, 5, 500, 1017, 1018, 1020, 1022])
def test_two_peak_window_matches_legacy_slices(oracle, cp):
    # cp + spc == N (here 1019) is left out: the legacy computes `right_index - N - 1` in usize and panics there
    import ctypes as C
    n, spc = 1023, 4
    rng = np.random.default_rng(cp)
    row = rng.uniform(1.0, 2.0, n).astype(np.float32)
    row[cp] = 50.0
    # plant decoys on both edges of the exclusion window so an off-by-one changes the answer
    for off, v in ((-spc, 9.0), (spc - 1, 8.0), (spc, 7.0), (-spc - 1, 6.0), (spc - 2, 10.0)):
        row[(cp + off) % n] = v
    _, p1, p2 = _legacy_two_peaks(row, spc)
    first, second = C.c_uint32(), C.c_uint32()
    ratio = oracle.lib().go_two_peak_ratio(row.ctypes.data_as(C.c_void_p), n, spc, C.byref(first), C.byref(second))
    assert first.value == cp
    assert row[second.value] == np.float32(p2)
    assert abs(ratio - np.sqrt(p1) / np.sqrt(p2)) < 1e-5


def test_doppler_aliasing_identity_f64():
    """The algebra behind gb_acq_set_doppler_aliasing, checked in NumPy f64 on the reference's own
    definitions (doppler_shift.rs:11-21 tables, do_acquisition.rs:176-192 chain): for carr_d = carr_b + m fs/N the
    correlation power of bin d equals the power obtained from bin b's spectrum paired with the code spectrum shifted by m,
    |IFFT(X_d conj C)|^2 == |IFFT(X_b conj C_m)|^2 with C_m[j] = C[j - m], for positive, negative and wrapping m."""
    rng = np.random.default_rng(5)
    n, fs = 4092, 4.092e6
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    code = np.sign(rng.standard_normal(n))
    C = np.fft.fft(code)
    i = np.arange(n)
    carr_b = 150.0
    Xb = np.fft.fft(x * np.exp(-2j * np.pi * carr_b * i / fs))
    for m in (1, -1, 5, -4, 37):
        carr_d = carr_b + m * fs / n
        Xd = np.fft.fft(x * np.exp(-2j * np.pi * carr_d * i / fs))
        np.testing.assert_allclose(Xd, np.roll(Xb, -m), atol=1e-7 * np.abs(Xb).max())      # X_d[k] = X_b[k + m]
        direct = np.abs(np.fft.ifft(Xd * np.conj(C))) ** 2
        aliased = np.abs(np.fft.ifft(Xb * np.conj(np.roll(C, m)))) ** 2                     # C_m[j] = C[j - m]
        np.testing.assert_allclose(aliased, direct, rtol=1e-9, atol=1e-12 * direct.max())
