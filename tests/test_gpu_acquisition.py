"""Parity of the fused CUDA acquisition (through the C-ABI) against the CPU oracle.

Contract (BASELINE.json north_star / SURVEY 8c): identical detected PRN set, code-phase sample index
and Doppler bin; correlation magnitudes and metrics within 1e-3 relative (tolerances below are the
stated ones; achieved error is ~1e-6)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-3  # north_star: "correlation magnitudes and acquisition metrics within 1e-3 relative"


def _sats(n, seed):
    rng = np.random.default_rng(seed)
    prns = rng.choice(np.arange(1, 33), size=4, replace=False)
    return [{"prn": int(p), "doppler": float(rng.uniform(-2500, 2500)), "code_phase": int(rng.integers(0, n)),
             "cn0_dbhz": float(c)} for p, c in zip(prns, (52.0, 49.0, 47.0, 45.0))]


def _engine(gpu, n, fs, **kw):
    from gnss_sdr_rs_b200 import acquisition
    return acquisition.AcquisitionEngine(gpu, n, fs, **kw)


def _assert_same_cells(a, b, exact_sum=True):
    """peak / argmax / peak2 are order-independent reductions and must agree bit for bit; sum8 is a float sum whose
    association follows the CTA geometry, so kernels with different working-thread counts agree to rounding only."""
    for f in ("peak", "argmax", "peak2"):
        assert a[f].tobytes() == b[f].tobytes(), f
    if exact_sum:
        assert a["sum8"].tobytes() == b["sum8"].tobytes()
    else:
        np.testing.assert_allclose(a["sum8"], b["sum8"], rtol=2e-6)


@pytest.mark.parametrize("n", [1024, 2048, 4092, 4096, 8184, 16368, 20000])
def test_fused_and_shared_chains_are_bit_identical(gpu, oracle, ffi, n):
    """GB_ACQ_FUSED (one kernel) and GB_ACQ_SHARED (forward path shared by all PRNs) run the same arithmetic."""
    from gnss_sdr_rs_b200 import sdr_mock
    fs = float(n) * 1000.0
    K = 4
    x = sdr_mock.baseband(fs, K, _sats(n, 7 * n), seed=n + 1)
    eng = _engine(gpu, n, fs)
    eng.make_doppler_tables(0.0, np.arange(-1000, 1001, 250, dtype=np.float32))
    eng.set_doppler_aliasing(False)   # same wipe-off table per bin in both chains (aliasing: test_doppler_aliasing_*)
    eng.set_detector(7.0, 2)
    out = {}
    for n_coh in (1, 2):
        eng.set_coherent(n_coh)
        for mode in (ffi.GB_ACQ_FUSED, ffi.GB_ACQ_SHARED):
            eng.set_mode(mode)
            out[(n_coh, mode)] = eng.search_cells(x, K, prn_mask=0x80000F0F)
        # n = 4092: the default shared chain is the leftover-warp kernel (128 working threads, the fused one has 160)
        _assert_same_cells(out[(n_coh, 0)], out[(n_coh, 1)], exact_sum=(n != 4092))
    assert out[(1, 0)]["peak"][0].max() > 0 and out[(1, 0)]["peak"][4].max() == 0


@pytest.mark.parametrize("K,n_coh", [(1, 1), (7, 1), (8, 1), (9, 1), (20, 1), (40, 2), (33, 1)])
def test_leftover_warp_kernel_is_bit_identical(gpu, ffi, K, n_coh):
    """N = 4092: the default inverse kernel batches the 4 ragged radix-31 butterflies of 8 consecutive groups into one
    pass of a fifth warp (acq_lw.cu).  Every group count around the batch boundaries must give the power rows of the
    generic kernel (GB_ACQ_SHARED_PLAIN) and of the single fused kernel: peak, arg-max and second peak byte for byte
    (they see every row element), the 8-lane sum to rounding (its association follows the CTA geometry)."""
    from gnss_sdr_rs_b200 import sdr_mock
    n, fs = 4092, 4.092e6
    x = sdr_mock.baseband(fs, K, _sats(n, 5), seed=K)
    eng = _engine(gpu, n, fs)
    eng.make_doppler_tables(0.0, np.arange(-1500, 1501, 250, dtype=np.float32))
    eng.set_doppler_aliasing(False)
    eng.set_detector(7.0, 4)
    eng.set_coherent(n_coh)
    out = []
    for mode in (ffi.GB_ACQ_SHARED, ffi.GB_ACQ_SHARED_PLAIN, ffi.GB_ACQ_FUSED):
        eng.set_mode(mode)
        out.append(eng.search_cells(x, K).copy())
    assert out[1].tobytes() == out[2].tobytes()
    _assert_same_cells(out[0], out[1], exact_sum=False)
    assert out[0]["peak"].min() > 0 and out[0]["peak2"].min() > 0


@pytest.mark.parametrize("K,n_coh,alias", [(1, 1, False), (9, 1, False), (20, 1, True), (40, 2, True)])
def test_tmem_accumulators_are_bit_identical(gpu, ffi, K, n_coh, alias):
    """The default N = 4092 inverse kernel keeps the 36 power accumulators of a working thread -- and its 31 code-spectrum
    values, the same for every group -- in tensor memory (tcgen05.alloc / ld / st: 128 columns per CTA) instead of
    registers / L1, so it fits 96 registers and four CTAs per SM; gb_tuning_set("acq_lw_tmem", 0) selects the register
    form, 1 the accumulators alone.  Same arithmetic in the same order: every cell byte for
    byte, also with a sparse PRN mask and over repeated searches (TMEM is allocated and freed by every CTA)."""
    from gnss_sdr_rs_b200 import sdr_mock
    n, fs = 4092, 4.092e6
    x = sdr_mock.baseband(fs, K, _sats(n, 5), seed=200 + K)
    eng = _engine(gpu, n, fs)
    eng.make_doppler_tables(0.0, np.arange(-1500, 1501, 250, dtype=np.float32))
    eng.set_doppler_aliasing(alias)
    eng.set_detector(7.0, 4)
    eng.set_coherent(n_coh)
    ffi.tuning_set("acq_lw_tmem", 0)
    try:
        ref = eng.search_cells(x, K).copy()
        ref_sub = eng.search_cells(x, K, prn_mask=(1 << 4) | (1 << 30)).copy()
    finally:
        ffi.tuning_set("acq_lw_tmem", 2)
    assert ffi.lib().gb_tuning_get(b"acq_lw_tmem", 2) == 2
    got = [eng.search_cells(x, K).copy() for _ in range(3)]      # the default: accumulators + code spectrum in tensor memory
    ffi.tuning_set("acq_lw_tmem", 1)        # accumulators only
    try:
        got.append(eng.search_cells(x, K).copy())
    finally:
        ffi.tuning_set("acq_lw_tmem", 2)
    sub = eng.search_cells(x, K, prn_mask=(1 << 4) | (1 << 30)).copy()
    for g in got:
        assert g.tobytes() == ref.tobytes()
    assert sub.tobytes() == ref_sub.tobytes()
    assert ref["peak"].min() > 0


@pytest.mark.parametrize("n", [1024, 2048, 4092, 4096, 8184, 16368, 20000])
def test_generic_kernel_tmem_accumulators_are_bit_identical(gpu, ffi, n):
    """gb_tuning_set("acq_tmem", 1): the generic inverse kernel of every plan with its power accumulators in tensor memory
    (warps that share a TMEM lane quarter stack their column ranges; 32 ... 256 columns per CTA, up to 8 CTAs per SM; the
    default for the power-of-two plans, whose register form spills them): cells byte for byte those of the register form
    ("acq_tmem", 0), Doppler aliasing on and off, with and without coherent pre-sum."""
    from gnss_sdr_rs_b200 import sdr_mock
    fs = float(n) * 1000.0
    K = 6
    x = sdr_mock.baseband(fs, K, _sats(n, n + 3), seed=n + 9)
    eng = _engine(gpu, n, fs)
    eng.make_doppler_tables(0.0, np.arange(-1500, 1501, 250, dtype=np.float32))
    eng.set_detector(7.0, 2)
    eng.set_mode(ffi.GB_ACQ_SHARED_PLAIN)     # N = 4092: the generic kernel, not the leftover-warp form
    for alias, n_coh in ((False, 1), (True, 1), (True, 2)):
        eng.set_doppler_aliasing(alias)
        eng.set_coherent(n_coh)
        try:
            ffi.tuning_set("acq_tmem", 0)
            ref = eng.search_cells(x, K).copy()
            for tmode in (2, 3, 1):     # code spectrum | both | accumulators in tensor memory (a mode that does not fit runs the default)
                ffi.tuning_set("acq_tmem", tmode)
                got = eng.search_cells(x, K).copy()
                assert got.tobytes() == ref.tobytes(), (alias, n_coh, tmode)
            got2 = eng.search_cells(x, K, prn_mask=0x80000001).copy()
        finally:
            ffi.tuning_set("acq_tmem", -1)    # per-plan default: on for the power-of-two plans
        assert got.tobytes() == ref.tobytes(), (alias, n_coh)
        assert eng.search_cells(x, K).tobytes() == ref.tobytes()
        assert got2[0].tobytes() == ref[0].tobytes() and got2[31].tobytes() == ref[31].tobytes()
    assert ref["peak"].min() > 0


@pytest.mark.parametrize("K,n_coh,alias", [(1, 1, False), (9, 1, False), (20, 1, True), (40, 2, True), (40, 2, False)])
def test_tensor_stage_matches_fp32_kernel(gpu, oracle, ffi, K, n_coh, alias):
    """A/B switch gb_tuning_set("acq_tc", 1): the radix-31 stage of the N = 4092 inverse kernel as 3 x TF32 products on the
    warp-level tensor path (acq_lw.cu; not the default, outside the north star).  Against the FP32 leftover-warp kernel
    on the same input: every peak, second peak and 8-lane sum within 2e-6 relative (stated bound of the split: 1e-6 of a
    butterfly's largest input), arg-max identical on every cell with a clear peak, decisions identical; and against the
    oracle within the 1e-3 contract.  The switch is per process, so it is always put back."""
    from gnss_sdr_rs_b200 import sdr_mock
    n, fs = 4092, 4.092e6
    x = sdr_mock.baseband(fs, K, _sats(n, 5), seed=100 + K)
    dopplers = np.arange(-1500, 1501, 250, dtype=np.float32)
    eng = _engine(gpu, n, fs)
    eng.make_doppler_tables(0.0, dopplers)
    eng.set_doppler_aliasing(alias)
    eng.set_detector(7.0, 4)
    eng.set_coherent(n_coh)
    ref = eng.search_cells(x, K).copy()
    dec_ref = eng.search(x, K)
    ffi.tuning_set("acq_tc", 1)
    try:
        got = eng.search_cells(x, K).copy()
        got2 = eng.search_cells(x, K).copy()     # second search: the cached code spectra in fragment order
        dec = eng.search(x, K)
        sub = eng.search_cells(x, K, prn_mask=(1 << 4) | (1 << 30)).copy()   # sparse PRN mask
    finally:
        ffi.tuning_set("acq_tc", 0)
    assert got.tobytes() == got2.tobytes()
    for f in ("peak", "peak2", "sum8"):
        err = np.abs(got[f].astype(np.float64) - ref[f]) / np.maximum(ref[f], 1e-30)
        assert err.max() < 2e-6, (f, err.max())
    clear = ref["peak"] > 1.5 * ref["peak2"]
    assert clear.sum() >= 5 and (got["argmax"][clear] == ref["argmax"][clear]).all()
    assert (got["argmax"] == ref["argmax"]).mean() > 0.99
    key = lambda r: None if r is None else (r["prn"], r["code_phase_samples"], r["carrier_freq"])
    assert [key(r) for r in dec] == [key(r) for r in dec_ref]
    assert sub[4].tobytes() == got[4].tobytes() and sub[30].tobytes() == got[30].tobytes() and sub["peak"][0].max() == 0
    carr, tabs = oracle.doppler_tables(0.0, dopplers, fs, n)
    for prn in (1, 7, 32) if n_coh == 1 else ():
        o = oracle.AcqWorker(prn, n, fs).cells(x, tabs, K)
        assert np.allclose(o["peak"], got[prn - 1]["peak"], rtol=1e-3)


@pytest.mark.parametrize("n", [1024, 2048, 4092, 4096, 8184, 16368, 20000])
def test_doppler_aliasing_matches_per_bin_tables(gpu, oracle, ffi, n):
    """Bins a whole number of FFT bins apart share one forward spectrum (gb_acq_set_doppler_aliasing; off by default,
    the default being the reference's per-bin tables):
    13 bins at 250 Hz over 1 kHz FFT bins -> 4 forward spectra, shifts -1..+2.  Against the same search with every
    bin's own reference table: identical arg-max on every cell with a clear peak, powers within 2e-5 (the f32
    rounding of the table phases is all that differs), through the ring path and the sliced host-pointer path,
    with and without coherent pre-sum; and against the oracle within the 1e-3 contract."""
    from gnss_sdr_rs_b200 import ring, sdr_mock
    fs = float(n) * 1000.0
    K = 4
    x = sdr_mock.baseband(fs, K, _sats(n, 3 * n + 1), seed=n + 5)
    dopplers = np.arange(-1500, 1501, 250, dtype=np.float32)
    eng = _engine(gpu, n, fs)
    eng.make_doppler_tables(0.0, dopplers)
    eng.set_detector(7.0, 2)
    assert eng.forward_bins() == len(dopplers)        # default: every bin its own forward path (reference arithmetic)
    eng.set_doppler_aliasing(True)
    assert eng.forward_bins() == 4
    rb = ring.MulticastRingBuffer(gpu, 1 << 17)
    rb.write_samples(x)
    modes = (ffi.GB_ACQ_SHARED, ffi.GB_ACQ_SHARED_PLAIN) if n == 4092 else (ffi.GB_ACQ_SHARED,)
    for n_coh in (1, 2):
        eng.set_coherent(n_coh)
        eng.set_doppler_aliasing(False)
        assert eng.forward_bins() == len(dopplers)
        ref = eng.search_cells(x, K).copy()
        eng.set_doppler_aliasing(True)
        for mode in modes:
            eng.set_mode(mode)
            for got in (eng.search_cells(x, K).copy(), eng.search_cells_ring(0, K).copy()):
                np.testing.assert_allclose(got["peak"], ref["peak"], rtol=2e-5)
                np.testing.assert_allclose(got["sum8"], ref["sum8"], rtol=2e-5)
                np.testing.assert_allclose(got["peak2"], ref["peak2"], rtol=2e-5)
                clear = ref["peak"] > 1.05 * ref["peak2"]
                assert clear.mean() > 0.5
                assert (got["argmax"][clear] == ref["argmax"][clear]).all()
        eng.set_mode(ffi.GB_ACQ_SHARED)
    # reference arithmetic (oracle) vs the aliased search, n_coh = 1
    eng.set_coherent(1)
    cells = eng.search_cells(x, K)
    carr, tabs = oracle.doppler_tables(0.0, dopplers, fs, n)
    for s in _sats(n, 3 * n + 1)[:2]:
        w = oracle.AcqWorker(s["prn"], n, fs)
        o = w.cells(x, tabs, K)
        np.testing.assert_allclose(cells[s["prn"] - 1]["peak"], o["peak"], rtol=REL)
        best = int(o["peak"].argmax())
        assert int(cells[s["prn"] - 1]["argmax"][best]) == int(o["argmax"][best])
        assert abs(int(o["argmax"][best]) - s["code_phase"]) <= max(1, n // 2046)   # within half a chip of the truth
    # decisions through the public search are the oracle's
    found = {r["prn"]: r for r in eng.search(x, K) if r}
    for prn in sorted(found)[:3]:
        r0 = oracle.AcqWorker(prn, n, fs).search_satellite(x, tabs, carr, 0, K)
        assert r0 is not None and r0["code_phase_samples"] == found[prn]["code_phase_samples"]
        assert r0["carrier_freq"] == found[prn]["carrier_freq"]


def test_doppler_aliasing_needs_exact_bin_multiples(gpu):
    """fs / fft_size of the reference recording is 999.97 Hz: bins 1 kHz apart are NOT whole FFT bins apart and every
    bin keeps its own forward path; caller-supplied tables are never analysed."""
    from gnss_sdr_rs_b200 import acquisition
    eng = _engine(gpu, 16368, 16367600.0)
    grid = np.array(acquisition.reference_doppler_grid(), np.float32)
    eng.make_doppler_tables(4130400.0, grid)
    eng.set_doppler_aliasing(True)
    assert eng.forward_bins() == len(grid)
    eng2 = _engine(gpu, 4092, 4.092e6)
    carr = eng2.make_doppler_tables(0.0, np.array([0.0, 1000.0, 2000.0, 500.0], np.float32))
    eng2.set_doppler_aliasing(True)
    assert eng2.forward_bins() == 2
    eng2.set_doppler_tables(eng2.get_doppler_tables(), carr)
    assert eng2.forward_bins() == 4


@pytest.mark.parametrize("n", [1024, 2048, 4092, 4096, 8184, 16368, 20000])
def test_cells_match_oracle_every_plan(gpu, oracle, n):
    from gnss_sdr_rs_b200 import sdr_mock
    fs = float(n) * 1000.0
    K = 2
    sats = _sats(n, n)
    x = sdr_mock.baseband(fs, K, sats, seed=n)
    dopplers = np.arange(-2500, 2501, 500, dtype=np.float32)
    carr, tabs = oracle.doppler_tables(0.0, dopplers, fs, n)
    eng = _engine(gpu, n, fs)
    eng.set_doppler_tables(tabs, carr)
    eng.set_detector(7.0, max(1, int(round(fs / 1.023e6))))
    cells = eng.search_cells(x, K)
    check = sorted({s["prn"] for s in sats} | {1, 32})
    for prn in check:
        w = oracle.AcqWorker(prn, n, fs)
        ref = w.cells(x, tabs, K)
        got = cells[prn - 1]
        assert (ref["argmax"] == got["argmax"]).all(), (prn, ref["argmax"], got["argmax"])
        np.testing.assert_allclose(got["peak"], ref["peak"], rtol=REL)
        np.testing.assert_allclose(got["sum8"], ref["sum8"], rtol=REL)
        # achieved accuracy is far inside the contract; keep a canary at 2e-5
        assert np.abs(got["peak"] / ref["peak"] - 1).max() < 2e-5
    # every planted satellite peaks where its code period starts, in the bin nearest its Doppler
    # (the early-exit DECISION is compared with the oracle only: threshold 7.0 fires on noise at K=2)
    for s in sats[:2]:
        row = cells[s["prn"] - 1]
        best = int(row["peak"].argmax())
        assert abs(float(dopplers[best]) - s["doppler"]) <= 500.0
        assert abs(int(row["argmax"][best]) - s["code_phase"] % n) <= 1
    res = eng.search(x, K)
    for prn in check:
        ref = oracle.AcqWorker(prn, n, fs).search_satellite(x, tabs, carr, 0, K)
        got = res[prn - 1]
        assert (ref is None) == (got is None)
        if ref:
            assert ref["code_phase_samples"] == got["code_phase_samples"] and ref["carrier_freq"] == got["carrier_freq"]


@pytest.mark.parametrize("n", [2048, 4092, 16368])
def test_power_row_matches_oracle(gpu, oracle, n):
    from gnss_sdr_rs_b200 import sdr_mock
    fs = float(n) * 1000.0
    K = 3
    sats = _sats(n, 3 * n)
    x = sdr_mock.baseband(fs, K, sats, seed=5)
    dopplers = np.array([sats[0]["doppler"] - 100.0, 0.0], np.float32)
    carr, tabs = oracle.doppler_tables(0.0, dopplers, fs, n)
    eng = _engine(gpu, n, fs)
    eng.set_doppler_tables(tabs, carr)
    prn = sats[0]["prn"]
    got = eng.bin_power(x, K, prn, 0)
    ref = oracle.AcqWorker(prn, n, fs).bin_power(x, tabs[0], K)
    assert np.abs(got - ref).max() <= 1e-5 * ref.max()
    assert int(got.argmax()) == int(ref.argmax())


def test_device_doppler_tables_match_libm(gpu, oracle):
    n, fs, f_if = 16368, 16367600.0, 4130400.0
    from gnss_sdr_rs_b200 import acquisition
    eng = _engine(gpu, n, fs)
    d = np.array(acquisition.reference_doppler_grid(), np.float32)
    carr = eng.make_doppler_tables(f_if, d)
    rc, rt = oracle.doppler_tables(f_if, d, fs, n)
    assert (carr == rc).all()
    got = eng.get_doppler_tables()
    # phase = i*step is bit-identical; device cosf/sinf differ from glibc by <= 2 ulp
    assert np.abs(got - rt).max() < 5e-7


def test_reference_recording_standin_config1(gpu, oracle):
    """BASELINE config 1 / do_acquisition.rs:399-466: int8 IF recording, fs 16.3676 MHz, IF 4.1304 MHz, 29 bins,
    10 x 1 ms, PRN 1..32 -- detections, code phases and Doppler bins identical to the oracle's
    search_satellite (early exit included)."""
    from gnss_sdr_rs_b200 import acquisition, sdr_mock
    n, fs, f_if, K = 16368, 16367600.0, 4130400.0, 10
    raw, truth = sdr_mock.if_recording(K)
    x = sdr_mock.i8_to_c32(raw)
    d = np.array(acquisition.reference_doppler_grid(), np.float32)
    carr, tabs = oracle.doppler_tables(f_if, d, fs, n)
    eng = _engine(gpu, n, fs)
    eng.set_doppler_tables(tabs, carr)
    got = eng.search(x, K, local_tail=0)
    workers = [oracle.AcqWorker(p, n, fs) for p in range(1, 33)]
    ref = oracle.acq_search_all(workers, x, tabs, carr, 0, K, early_exit=True)
    truth_prns = {t["prn"] for t in truth}
    for p in range(32):
        assert (ref[p] is None) == (got[p] is None), "PRN %d detection differs" % (p + 1)
        if ref[p] is None:
            continue
        assert got[p]["code_phase_samples"] == ref[p]["code_phase_samples"]
        assert got[p]["carrier_freq"] == ref[p]["carrier_freq"]
        assert got[p]["sample_global_index"] == ref[p]["sample_global_index"]
        assert got[p]["code_phase_chips"] == ref[p]["code_phase_chips"]
        assert abs(got[p]["mag_relative"] / ref[p]["mag_relative"] - 1) < REL
        assert (p + 1) in truth_prns  # the reference's one-sided assertion (do_acquisition.rs:454)
    # the strong half of the table must be acquired
    assert {2, 3, 19, 14, 18}.issubset({p + 1 for p in range(32) if got[p]})
    # BASELINE configs[0] as worded: +-5 kHz / 500 Hz (D = 21), ONE 1 ms block (K = 1) -- same comparison, plus the
    # margin of every decision to the 7.0 threshold: cells agree to ~1e-6, so a decision can only differ
    # when the metric sits that close to 7.0
    d21 = np.arange(-5000.0, 5000.0 + 1.0, 500.0, dtype=np.float32)
    carr21, tabs21 = oracle.doppler_tables(f_if, d21, fs, n)
    assert len(d21) == 21
    eng.set_doppler_tables(tabs21, carr21)
    got1 = eng.search(x, 1, local_tail=7)
    cells1 = eng.search_cells(x, 1)
    ref1 = oracle.acq_search_all(workers, x, tabs21, carr21, 7, 1, early_exit=True)
    margins = []
    for p in range(32):
        o = workers[p].cells(x, tabs21, 1)
        np.testing.assert_allclose(cells1[p]["peak"], o["peak"], rtol=REL)
        assert (cells1[p]["argmax"] == o["argmax"]).all()
        gmax = gsum = 0.0
        for b in range(21):                       # the early-exit scan (Q1) on the oracle's cells
            if o["peak"][b] > gmax:
                gmax, gsum = float(o["peak"][b]), float(o["sum8"][b])
            metric = gmax / ((gsum - gmax) / (n - 1))
            margins.append(abs(metric - 7.0))
            if metric > 7.0:
                break
        assert (ref1[p] is None) == (got1[p] is None), "PRN %d detection differs (K = 1)" % (p + 1)
        if ref1[p]:
            for k in ("code_phase_samples", "carrier_freq", "sample_global_index", "code_phase_chips"):
                assert got1[p][k] == ref1[p][k]
    print("config 1 (D = 21, K = 1): detected %s, smallest |metric - 7.0| met by any scan: %.4f" % (
        [p + 1 for p in range(32) if got1[p]], min(margins)))
    assert min(margins) > 1e-3


def test_prn_mask_and_decide(gpu, oracle):
    from gnss_sdr_rs_b200 import acquisition, sdr_mock
    n, fs, K = 2048, 2.048e6, 4
    sats = [{"prn": 3, "doppler": 700.0, "code_phase": 100, "cn0_dbhz": 50.0},
            {"prn": 20, "doppler": -1900.0, "code_phase": 1900, "cn0_dbhz": 50.0}]
    x = sdr_mock.baseband(fs, K, sats, seed=11)
    d = np.arange(-3000, 3001, 250, dtype=np.float32)
    carr, tabs = oracle.doppler_tables(0.0, d, fs, n)
    eng = _engine(gpu, n, fs)
    eng.set_doppler_tables(tabs, carr)
    mask = (1 << 2) | (1 << 7)  # PRN 3 and 8 only (do_acquisition.rs:307)
    cells = eng.search_cells(x, K, prn_mask=mask)
    assert cells[19]["peak"].max() == 0.0 and cells[2]["peak"].max() > 0.0 and cells[7]["peak"].max() > 0.0
    res = eng.search(x, K, local_tail=12345, prn_mask=mask)
    assert res[2] is not None and res[19] is None and res[7] is None
    assert res[2]["sample_global_index"] == 12345 + res[2]["code_phase_samples"]
    # Q1: the decision is the first record-setting bin that passes, not the global maximum
    full = eng.search_cells(x, K)
    for prn in (3, 20, 8):
        a = acquisition.decide(full[prn - 1], carr, prn, n, fs)
        b = oracle.acq_decide(full[prn - 1][["peak", "argmax", "sum8"]].astype(oracle.CELL_DTYPE), carr, prn, n, fs)
        assert (a is None) == (b is None)
        if a:
            assert a["doppler_bin"] == b["bin"] and a["code_phase_samples"] == b["code_phase_samples"]
            w = oracle.AcqWorker(prn, n, fs)
            c = w.search_satellite(x, tabs, carr, 0, K)
            assert c["carrier_freq"] == a["carrier_freq"] and c["code_phase_samples"] == a["code_phase_samples"]


@pytest.mark.parametrize("n,n_coh,K", [(4092, 2, 4), (2048, 5, 10)])
def test_coherent_extension_matches_oracle(gpu, oracle, n, n_coh, K):
    """EXTENSION (BASELINE config 2): coherent sums of n_coh blocks.  The GPU pre-sums the wiped blocks before one
    FFT pair; the oracle's definitional form sums the n_coh correlations after the IFFT."""
    from gnss_sdr_rs_b200 import sdr_mock
    fs = float(n) * 1000.0
    sats = [{"prn": 7, "doppler": 1234.0, "code_phase": 777, "cn0_dbhz": 42.0},
            {"prn": 30, "doppler": -420.0, "code_phase": 99, "cn0_dbhz": 44.0}]
    x = sdr_mock.baseband(fs, K, sats, seed=21)
    step = 1000.0 / n_coh / 2.0
    d = np.arange(-1500, 1501, step, dtype=np.float32)
    carr, tabs = oracle.doppler_tables(0.0, d, fs, n)
    rot = oracle.coh_rotators(carr, fs, n, n_coh)
    eng = _engine(gpu, n, fs)
    eng.set_doppler_tables(tabs, carr)
    eng.set_coherent(n_coh)
    cells = eng.search_cells(x, K)
    for prn in (7, 30, 12):
        w = oracle.AcqWorker(prn, n, fs)
        post = w.cells(x, tabs, K, n_coh=n_coh, rot=rot, presum=0)
        pre = w.cells(x, tabs, K, n_coh=n_coh, rot=rot, presum=1)
        got = cells[prn - 1]
        np.testing.assert_allclose(pre["peak"], post["peak"], rtol=1e-4)
        np.testing.assert_allclose(got["peak"], post["peak"], rtol=REL)
        np.testing.assert_allclose(got["sum8"], post["sum8"], rtol=REL)
        strong = post["peak"] > 3.0 * np.median(post["peak"])
        assert (got["argmax"][strong] == post["argmax"][strong]).all()
    # the coherent gain puts the planted satellites at the right Doppler bin
    for s in sats:
        best = int(cells[s["prn"] - 1]["peak"].argmax())
        assert abs(float(d[best]) - s["doppler"]) <= step


@pytest.mark.parametrize("code_phase", [1000, 2, 4, 4093, 4091])
def test_two_peak_metric(gpu, oracle, code_phase):
    """peak2 follows the legacy's slice bounds verbatim (acquisition_bk.rs:371-390), wrap cases included."""
    from gnss_sdr_rs_b200 import sdr_mock
    n, fs, K = 4096, 4.096e6, 2
    sats = [{"prn": 9, "doppler": 0.0, "code_phase": code_phase, "cn0_dbhz": 50.0}]
    x = sdr_mock.baseband(fs, K, sats, seed=3)
    carr, tabs = oracle.doppler_tables(0.0, np.array([0.0, 500.0], np.float32), fs, n)
    eng = _engine(gpu, n, fs)
    eng.set_doppler_tables(tabs, carr)
    spc = 4
    eng.set_detector(7.0, spc)
    cells = eng.search_cells(x, K)
    for prn in (9, 10):
        w = oracle.AcqWorker(prn, n, fs)
        for b in range(2):
            row = w.bin_power(x, tabs[b], K)
            L = oracle.lib()
            import ctypes as C
            ratio = L.go_two_peak_ratio(row.ctypes.data_as(C.c_void_p), n, spc, None, None)
            got = np.sqrt(cells[prn - 1]["peak"][b] / cells[prn - 1]["peak2"][b])
            assert abs(got / ratio - 1) < REL
    assert cells[8]["argmax"][0] == code_phase
    assert np.sqrt(cells[8]["peak"][0] / cells[8]["peak2"][0]) > 1.4  # acquisition_bk.rs threshold


@pytest.mark.parametrize("n,n_coh", [(5000, 1), (8192, 1), (10000, 2), (12000, 1), (2044, 1), (25000, 1)])
def test_any_length_plan_matches_oracle(gpu, oracle, ffi, n, n_coh):
    """AcquisitionWorker::new accepts ANY fft_size (rustfft planner, do_acquisition.rs:131-142).  Sample rates without a
    tuned shared-memory plan run the any-length plan (acq_generic.cu): cells against the oracle within the 1e-3 contract
    (achieved ~1e-5), arg-max identical on every cell with a clear peak, decisions (early exit, Q1) identical, the
    accumulated power row of one bin, the ring path, a sparse PRN mask, and the coherent extension."""
    from gnss_sdr_rs_b200 import ring, sdr_mock
    fs = float(n) * 1000.0
    K = 4
    sats = _sats(n, 3 * n + 7)
    x = sdr_mock.baseband(fs, K, sats, seed=n + 3)
    dopplers = np.arange(-1500, 1501, 500, dtype=np.float32)
    assert n not in ffi_supported(ffi)
    eng = _engine(gpu, n, fs)
    carr = eng.make_doppler_tables(0.0, dopplers)
    eng.set_detector(7.0, max(1, n // 1023))
    eng.set_coherent(n_coh)
    cells = eng.search_cells(x, K)
    _, tabs = oracle.doppler_tables(0.0, dopplers, fs, n)
    rot = oracle.coh_rotators(carr, fs, n, n_coh) if n_coh > 1 else None
    prns = sorted({s["prn"] for s in sats[:3]} | {1, 32})
    for prn in prns:
        w = oracle.AcqWorker(prn, n, fs)
        o = w.cells(x, tabs, K, n_coh=n_coh, rot=rot, presum=1)
        np.testing.assert_allclose(cells[prn - 1]["peak"], o["peak"], rtol=REL)
        np.testing.assert_allclose(cells[prn - 1]["sum8"], o["sum8"], rtol=REL)
        assert np.abs(cells[prn - 1]["peak"] / o["peak"] - 1).max() < 5e-5
        clear = o["peak"] > 2.0 * np.median(o["peak"])
        assert (cells[prn - 1]["argmax"][clear] == o["argmax"][clear]).all()
        if n_coh == 1:
            dec = oracle.acq_decide(o, carr, prn, n, fs, local_tail=5)
            got = eng.search(x, K, local_tail=5, prn_mask=1 << (prn - 1))[prn - 1]
            assert (dec is None) == (got is None)
            if dec:
                assert dec["code_phase_samples"] == got["code_phase_samples"] and dec["carrier_freq"] == got["carrier_freq"]
                assert dec["sample_global_index"] == got["sample_global_index"] == 5 + dec["code_phase_samples"]
    s0 = sats[0]
    best = int(cells[s0["prn"] - 1]["peak"].argmax())
    assert abs(int(cells[s0["prn"] - 1]["argmax"][best]) - s0["code_phase"]) <= max(1, n // 2046)
    # one bin's accumulated power row
    row = eng.bin_power(x, K, s0["prn"], best)
    ref_row = oracle.AcqWorker(s0["prn"], n, fs).bin_power(x, tabs[best], K, n_coh=n_coh,
                                                          rot=None if rot is None else rot[best], presum=1)
    np.testing.assert_allclose(row, ref_row, rtol=2e-4, atol=ref_row.max() * 2e-6)
    assert int(row.argmax()) == int(ref_row.argmax())
    # ring path + sparse mask: identical cells, masked rows zero
    rb = ring.MulticastRingBuffer(gpu, 1 << 17)
    rb.write_samples(x)
    mask = (1 << (s0["prn"] - 1)) | 1
    cr = eng.search_cells_ring(0, K, prn_mask=mask)
    for p in range(32):
        if (mask >> p) & 1:
            assert cr[p].tobytes() == cells[p].tobytes()
        else:
            assert (cr[p]["peak"] == 0).all()


def ffi_supported(ffi):
    import ctypes as C
    buf = (C.c_int * 32)()
    k = ffi.lib().gb_acq_supported_sizes(buf, 32)
    return set(buf[:k])


def test_errors_and_edge_cases(gpu, ffi):
    from gnss_sdr_rs_b200 import acquisition
    for bad_n, bad_fs in ((2046, 2.046e6), (10230, 10.23e6), (4101, 4.101e6)):
        with pytest.raises(ffi.GnssB200Error) as e:   # fft_size % 4 != 0: the reference's own arithmetic diverges (header)
            acquisition.AcquisitionEngine(gpu, bad_n, bad_fs)
        assert e.value.code == ffi.GB_EUNSUPPORTED
    eng = acquisition.AcquisitionEngine(gpu, 2048, 2.048e6)
    with pytest.raises(ffi.GnssB200Error) as e:  # search before tables
        eng.carr = np.zeros(1, np.float32)
        eng.search_cells(np.zeros(2048, np.complex64), 1)
    assert e.value.code == ffi.GB_ESTATE
    eng.make_doppler_tables(0.0, [0.0])
    # all-zero input: no bin is ever > 0.0 -> argmax 0, peak 0, metric NaN -> None (Q1)
    cells = eng.search_cells(np.zeros(2048, np.complex64), 1)
    assert (cells["peak"] == 0).all() and (cells["argmax"] == 0).all()
    assert all(r is None for r in eng.search(np.zeros(2048, np.complex64), 1))
    eng.set_coherent(2)
    with pytest.raises(ffi.GnssB200Error) as e:  # K not a multiple of n_coh
        eng.search_cells(np.zeros(3 * 2048, np.complex64), 3)
    assert e.value.code == ffi.GB_EINVAL
    eng.set_coherent(1)
    # host buffers cross the boundary with their length: a chunk shorter than K * fft_size is refused, not read past
    for call in (lambda: eng.search_cells(np.zeros(2 * 2048 - 1, np.complex64), 2),
                 lambda: eng.search(np.zeros(2047, np.complex64), 1),
                 lambda: eng.bin_power(np.zeros(2048, np.complex64), 2, 1, 0),
                 lambda: eng.search_enqueue(np.zeros(100, np.complex64), 1, 0)):
        with pytest.raises(ffi.GnssB200Error) as e:
            call()
        assert e.value.code == ffi.GB_ERANGE


def test_two_handles_on_one_device(ffi):
    """Two handles (own streams and buffers) searched from two host threads give the single-handle results, and
    destroying one leaves no stale CUDA error behind for the other's next launch (gb_destroy used to free the plan's
    twiddles twice)."""
    import threading
    from gnss_sdr_rs_b200 import acquisition, sdr_mock
    n, fs, K = 4092, 4.092e6, 4
    x = sdr_mock.baseband(fs, K, _sats(n, 11), seed=3)
    hs = [ffi.Handle(0), ffi.Handle(0)]
    engs = []
    for h in hs:
        e = acquisition.AcquisitionEngine(h, n, fs)
        e.make_doppler_tables(0.0, np.arange(-2000, 2001, 250, dtype=np.float32))
        engs.append(e)
    ref = engs[0].search_cells(x, K).copy()
    out = [None, None]

    def work(i):
        for _ in range(5):
            out[i] = engs[i].search_cells(x, K).copy()

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert out[0].tobytes() == ref.tobytes() and out[1].tobytes() == ref.tobytes()
    hs[1].close()
    assert engs[0].search_cells(x, K).tobytes() == ref.tobytes()
    hs[0].close()


def test_tensor_memory_is_shared_between_concurrent_searches(ffi):
    """Tensor memory is a per-SM resource the inverse kernels allocate per CTA (N = 4092: 128 columns, four CTAs fill the
    SM's 512; power-of-two plans: 32 columns; the cluster plan: 256): searches of three handles running at the same time
    on one GPU -- different plans, different allocation sizes, more CTAs than fit -- must wait for each other's columns
    (tcgen05.alloc blocks, every CTA frees what it took) and give the results of the same searches run alone."""
    import threading
    from gnss_sdr_rs_b200 import acquisition, sdr_mock
    dopplers = np.arange(-5000, 5001, 250, dtype=np.float32)     # 41 bins x 32 PRNs = 1312 CTAs per search
    jobs = []
    for n, K in ((4092, 8), (2048, 8), (4092, 6)):
        fs = float(n) * 1000.0
        h = ffi.Handle(0)
        e = acquisition.AcquisitionEngine(h, n, fs)
        e.make_doppler_tables(0.0, dopplers)
        x = sdr_mock.baseband(fs, K, _sats(n, n + K), seed=n + K)
        jobs.append((h, e, x, K, e.search_cells(x, K).copy()))
    out = [None] * len(jobs)

    def work(i):
        _, e, x, K, ref = jobs[i]
        ok = True
        for _ in range(6):
            ok &= e.search_cells(x, K).tobytes() == ref.tobytes()
        out[i] = ok

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert all(not t.is_alive() for t in th), "a search did not finish: tensor-memory allocation must never deadlock"
    assert out == [True] * len(jobs)
    for h, *_ in jobs:
        h.close()


def test_headline_kernel_edge_cases(gpu, oracle, ffi):
    """N = 4092 through the default chain (leftover-warp kernel + Doppler aliasing): an IF offset with negative Dopplers,
    a sparse PRN mask (masked rows stay zero), all-zero input (peak 0, arg-max 0, metric NaN -> None, Q1) and a grid
    whose alias classes have different sizes (5 bins: shifts 0, +1, +2 for one class, 0 for the other two)."""
    from gnss_sdr_rs_b200 import sdr_mock
    n, fs, K = 4092, 4.092e6, 3
    f_if = 1250.0
    sats = [{"prn": 4, "doppler": f_if - 750.0, "code_phase": 4091, "cn0_dbhz": 52.0},
            {"prn": 32, "doppler": f_if + 1250.0, "code_phase": 0, "cn0_dbhz": 50.0}]
    x = sdr_mock.baseband(fs, K, sats, seed=77)
    dopplers = np.array([-750.0, 250.0, 1250.0, -500.0, 100.0], np.float32)   # carriers 500, 1500, 2500 | 750 | 1350
    eng = _engine(gpu, n, fs)
    carr = eng.make_doppler_tables(f_if, dopplers)
    eng.set_doppler_aliasing(True)
    assert eng.forward_bins() == 3
    eng.set_detector(7.0, 4)
    mask = (1 << 3) | (1 << 31) | (1 << 10)
    cells = eng.search_cells(x, K, prn_mask=mask)
    _, tabs = oracle.doppler_tables(f_if, dopplers, fs, n)
    for prn in (4, 32, 11):
        o = oracle.AcqWorker(prn, n, fs).cells(x, tabs, K)
        np.testing.assert_allclose(cells[prn - 1]["peak"], o["peak"], rtol=REL)
        np.testing.assert_allclose(cells[prn - 1]["sum8"], o["sum8"], rtol=REL)
    assert int(cells[3]["argmax"][0]) == 4091 and int(cells[31]["argmax"][2]) == 0
    untouched = np.ones(32, bool)
    untouched[[3, 31, 10]] = False
    assert (cells["peak"][untouched] == 0).all()
    found = {r["prn"]: r for r in eng.search(x, K, prn_mask=mask) if r}
    assert set(found) == {4, 32}
    assert found[4]["carrier_freq"] == np.float32(carr[0]) and found[32]["carrier_freq"] == np.float32(carr[2])
    z = np.zeros(K * n, np.complex64)
    cz = eng.search_cells(z, K)
    assert (cz["peak"] == 0).all() and (cz["argmax"] == 0).all() and (cz["sum8"] == 0).all()
    assert all(r is None for r in eng.search(z, K))


def test_config2_full_size_properties(gpu):
    """BASELINE configs[1] at full size (N = 4092, 32 PRNs, 201 Doppler bins, 10 ms coherent x 20 non-coherent, 200 ms of
    signal) through properties that need no oracle: determinism; exact homogeneity (power-of-two scaling of the input is
    exact in f32, so every cell scales by exactly 4 and no arg-max moves); shift equivariance (dropping s samples moves
    every strong cell's code phase by -s mod N and leaves its Doppler bin); the eight satellites of the scene are found at
    their code phases and nearest Doppler bins and nothing else passes the two-peak test."""
    import bench
    from gnss_sdr_rs_b200 import acquisition
    n, K, s = bench.N_FFT, bench.K_MS, 1000
    from gnss_sdr_rs_b200 import sdr_mock
    sats = [{"prn": p, "doppler": d, "code_phase": c, "cn0_dbhz": cn} for p, d, c, cn in bench.SATS]
    x = sdr_mock.baseband(bench.FS, K + 1, sats, seed=0x6E56, nav=False)
    eng = acquisition.AcquisitionEngine(gpu, n, bench.FS)
    eng.make_doppler_tables(0.0, bench.DOPPLERS)
    eng.set_coherent(bench.N_COH)
    eng.set_detector(7.0, 4)
    a = eng.search_cells(x[:K * n], K)
    b = eng.search_cells(x[:K * n], K)
    assert a.tobytes() == b.tobytes()                                   # deterministic
    c = eng.search_cells((x[:K * n] * np.complex64(2.0)), K)
    assert (c["argmax"] == a["argmax"]).all()
    assert (c["peak"] == a["peak"] * np.float32(4.0)).all() and (c["sum8"] == a["sum8"] * np.float32(4.0)).all()
    d = eng.search_cells(x[s:s + K * n], K)
    best = a["peak"].argmax(axis=1)
    truth = {p: (dp, cp) for p, dp, cp, _ in bench.SATS}
    found = []
    for p in range(32):
        bb = int(best[p])
        ratio = np.sqrt(a["peak"][p, bb] / a["peak2"][p, bb])
        if ratio > 1.4:                                                  # acquisition_bk.rs two-peak threshold
            found.append(p + 1)
            dp, cp = truth[p + 1]
            assert abs(float(bench.DOPPLERS[bb]) - dp) <= 25.0 + 1e-3
            assert min((int(a["argmax"][p, bb]) - cp) % n, (cp - int(a["argmax"][p, bb])) % n) <= 1
            # shift equivariance on the detected cell
            assert int(d["peak"][p].argmax()) == bb
            assert (int(a["argmax"][p, bb]) - int(d["argmax"][p, bb])) % n in (s % n, (s - 1) % n, (s + 1) % n)
            assert abs(d["peak"][p, bb] / a["peak"][p, bb] - 1) < 0.2
    assert sorted(found) == sorted(truth), found


def test_async_enqueue_wait_and_batch(gpu, oracle, ffi):
    """gb_acq_search_enqueue / gb_acq_search_wait (two slots on one handle) and gb_acq_search_batch: the same results as
    the synchronous call for every recording, in order, whatever is in flight; slot misuse is refused (GB_ESTATE)."""
    import ctypes as C
    from gnss_sdr_rs_b200 import sdr_mock
    n, fs, K = 4092, 4.092e6, 4
    recs = [sdr_mock.baseband(fs, K, _sats(n, 100 + i), seed=200 + i) for i in range(5)]
    eng = _engine(gpu, n, fs)
    eng.make_doppler_tables(0.0, np.arange(-2500, 2501, 250, dtype=np.float32))
    eng.set_detector(7.0, 4)
    for alias in (False, True):
        eng.set_doppler_aliasing(alias)
        sync = [eng.search(r, K, local_tail=10 * i) for i, r in enumerate(recs)]
        sync_cells = [eng.search_cells(r, K).copy() for r in recs]
        got, got_cells = [], []
        eng.search_enqueue(recs[0], K, 0, local_tail=0)
        for i in range(len(recs)):
            if i + 1 < len(recs):
                eng.search_enqueue(recs[i + 1], K, (i + 1) & 1, local_tail=10 * (i + 1))
            res, cells = eng.search_wait(i & 1, want_cells=True)
            got.append(res)
            got_cells.append(cells)
        for i in range(len(recs)):
            assert got_cells[i].tobytes() == sync_cells[i].tobytes()
            for a, b in zip(got[i], sync[i]):
                assert (a is None) == (b is None)
                if a:
                    assert a == b
        assert any(r is not None for res in got for r in res)
    # batch form: n_rec recordings back to back
    ptrs = (C.c_void_p * len(recs))(*[r.ctypes.data for r in recs])
    out = (ffi.AcqResult * (len(recs) * 32))()
    gpu.call("gb_acq_search_batch", ptrs, len(recs), recs[0].size, K, 0, 0xFFFFFFFF, None, out)
    for i in range(len(recs)):
        ref = eng.search(recs[i], K, local_tail=0)
        for p in range(32):
            r = out[i * 32 + p]
            assert bool(r.found) == (ref[p] is not None)
            if r.found:
                assert r.code_phase_samples == ref[p]["code_phase_samples"] and r.carrier_freq == ref[p]["carrier_freq"]
    # misuse: waiting on an empty slot, enqueueing twice on one slot
    with pytest.raises(ffi.GnssB200Error) as e:
        eng.search_wait(1)
    assert e.value.code == ffi.GB_ESTATE
    eng.search_enqueue(recs[0], K, 0)
    with pytest.raises(ffi.GnssB200Error) as e:
        eng.search_enqueue(recs[1], K, 0)
    assert e.value.code == ffi.GB_ESTATE
    with pytest.raises(ffi.GnssB200Error) as e:      # no re-planning / single-bin diagnostics while a search is in flight
        eng.bin_power(recs[0], K, 1, 0)
    assert e.value.code == ffi.GB_ESTATE
    eng.search_wait(0)


def test_group_collective_single_rank(gpu, ffi):
    """gb_group_* with a world of one (NCCL is loaded inside the library and set up through the same calls the multi-GPU
    bench makes): unique id, init + warm-up gather, all-gather of bytes, the asynchronous slots, gather of result structs."""
    import ctypes as C
    L = ffi.lib()
    ids = (C.c_uint8 * 128)()
    ffi.check(L.gb_group_unique_id(ids), "gb_group_unique_id")
    g = C.c_void_p()
    ffi.check(L.gb_group_init(gpu.h, ids, 0, 1, C.byref(g)), "gb_group_init", gpu.h)
    try:
        assert L.gb_group_rank(g) == 0 and L.gb_group_world(g) == 1
        mine = np.arange(1000, dtype=np.uint8)
        out = np.zeros(1000, np.uint8)
        ffi.check(L.gb_group_allgather(g, ffi.ptr(mine), mine.size, ffi.ptr(out)), "gb_group_allgather", gpu.h)
        assert (out == mine).all()
        a, b = np.full(64, 7, np.uint8), np.full(4096, 9, np.uint8)
        oa, ob = np.zeros_like(a), np.zeros_like(b)
        ffi.check(L.gb_group_allgather_begin(g, ffi.ptr(a), a.size, 0), "begin0", gpu.h)
        ffi.check(L.gb_group_allgather_begin(g, ffi.ptr(b), b.size, 1), "begin1", gpu.h)
        assert L.gb_group_allgather_begin(g, ffi.ptr(a), a.size, 0) == ffi.GB_ESTATE      # slot still occupied
        ffi.check(L.gb_group_allgather_end(g, 1, ffi.ptr(ob)), "end1", gpu.h)
        ffi.check(L.gb_group_allgather_end(g, 0, ffi.ptr(oa)), "end0", gpu.h)
        assert (oa == 7).all() and (ob == 9).all()
        assert L.gb_group_allgather_end(g, 0, ffi.ptr(oa)) == ffi.GB_ESTATE               # nothing pending
        res = (ffi.AcqResult * 32)()
        for p in range(32):
            res[p].prn, res[p].found, res[p].code_phase_samples = p + 1, p % 2, 100 * p
        allr = (ffi.AcqResult * 32)()
        ffi.check(L.gb_group_gather_results(g, res, 32, allr), "gb_group_gather_results", gpu.h)
        assert bytes(allr) == bytes(res)
    finally:
        L.gb_group_destroy(g)
