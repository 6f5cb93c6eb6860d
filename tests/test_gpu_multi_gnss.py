"""BASELINE config 4 shape: GPS L1 C/A + BeiDou B1I (N = 20000) + Galileo-E1-like BOC(1,1) 4 ms (N = 80000, the
thread-block-cluster / DSMEM plan) at 20 Msps.  The reference implements GPS only, so parity here is
oracle-vs-GPU self-consistency on the same codes ("parity unpinned" for B1I / E1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FS = 20e6
REL = 1e-3


def _signal(n_ms):
    from gnss_sdr_rs_b200 import sdr_mock
    sats = [{"system": "G", "prn": 5, "doppler": 1250.0, "code_phase": 12345, "cn0_dbhz": 50.0},
            {"system": "C", "prn": 8, "doppler": -750.0, "code_phase": 4321, "cn0_dbhz": 50.0},
            {"system": "E", "prn": 3, "doppler": 500.0, "code_phase": 55555, "cn0_dbhz": 50.0}]
    return sdr_mock.multi_gnss(FS, n_ms, sats), sats


def test_beidou_b1i_custom_codes_n20000(gpu, oracle):
    from gnss_sdr_rs_b200 import acquisition, sdr_mock
    n, K = 20000, 3
    x, sats = _signal(K)
    prns = [1, 8, 20]
    codes = np.stack([sdr_mock.resample_code(sdr_mock.b1i_code(p), 2.046e6, FS, n) for p in prns])
    eng = acquisition.AcquisitionEngine(gpu, n, FS, n_prn=len(prns), codes=codes)
    d = np.arange(-1500, 1501, 250, dtype=np.float32)
    carr, tabs = oracle.doppler_tables(0.0, d, FS, n)
    eng.set_doppler_tables(tabs, carr)
    cells = eng.search_cells(x, K)
    for i, p in enumerate(prns):
        ref = oracle.AcqWorker(p, n, FS, code_samples=codes[i]).cells(x, tabs, K)
        assert (ref["argmax"] == cells[i]["argmax"]).all()
        np.testing.assert_allclose(cells[i]["peak"], ref["peak"], rtol=REL)
        np.testing.assert_allclose(cells[i]["sum8"], ref["sum8"], rtol=REL)
    best = int(cells[1]["peak"].argmax())
    assert d[best] == -750.0 and cells[1]["argmax"][best] == 4321
    assert cells[1]["peak"][best] > 10 * np.median(cells[0]["peak"])


def test_galileo_e1_cluster_plan_n80000(gpu, oracle, ffi):
    """N = 80000 does not fit one CTA: radix-4 outer stage over a cluster of 4 CTAs, sub-blocks exchanged through
    distributed shared memory (csrc/acq_cluster.cu)."""
    from gnss_sdr_rs_b200 import acquisition, sdr_mock
    n, K = 80000, 2          # two 4 ms blocks
    x, sats = _signal(4 * K)
    prns = [3, 11]
    codes = np.stack([sdr_mock.resample_code(sdr_mock.e1_surrogate_code(p), 1.023e6, FS, n, boc11=True) for p in prns])
    assert 80000 in _sizes(ffi)
    eng = acquisition.AcquisitionEngine(gpu, n, FS, n_prn=len(prns), codes=codes)
    d = np.arange(0, 1001, 125, dtype=np.float32)   # 1/T_coh-scaled bin width for 4 ms
    carr, tabs = oracle.doppler_tables(0.0, d, FS, n)
    eng.set_doppler_tables(tabs, carr)
    eng.set_detector(7.0, 20)
    cells = eng.search_cells(x, K)
    for i, p in enumerate(prns):
        w = oracle.AcqWorker(p, n, FS, code_samples=codes[i])
        ref = w.cells(x, tabs, K)
        assert (ref["argmax"] == cells[i]["argmax"]).all(), (ref["argmax"], cells[i]["argmax"])
        np.testing.assert_allclose(cells[i]["peak"], ref["peak"], rtol=REL)
        np.testing.assert_allclose(cells[i]["sum8"], ref["sum8"], rtol=REL)
    best = int(cells[0]["peak"].argmax())
    assert d[best] == 500.0 and cells[0]["argmax"][best] == 55555
    row = eng.bin_power(x, K, 1, best)
    ref_row = oracle.AcqWorker(3, n, FS, code_samples=codes[0]).bin_power(x, tabs[best], K)
    assert np.abs(row - ref_row).max() <= 2e-5 * ref_row.max()
    # the early-exit decision (threshold 7.0 fires on noise at N = 80000, K = 2) is compared with the oracle's
    res = eng.search(x, K, local_tail=9)
    for i, p in enumerate(prns):
        ref = oracle.AcqWorker(p, n, FS, code_samples=codes[i]).search_satellite(x, tabs, carr, 9, K)
        assert (ref is None) == (res[i] is None)
        if ref:
            assert ref["code_phase_samples"] == res[i]["code_phase_samples"] and ref["carrier_freq"] == res[i]["carrier_freq"]
            assert res[i]["sample_global_index"] == 9 + res[i]["code_phase_samples"]
    eng.set_coherent(2)
    with pytest.raises(ffi.GnssB200Error) as e:
        eng.search_cells(x, K)
    assert e.value.code == ffi.GB_EUNSUPPORTED


def _sizes(ffi):
    buf = np.zeros(32, np.int32)
    k = ffi.lib().gb_acq_supported_sizes(ffi.ptr(buf), 32)
    return set(buf[:k].tolist())


def test_gps_at_20msps(gpu, oracle):
    from gnss_sdr_rs_b200 import acquisition
    n, K = 20000, 2
    x, sats = _signal(K)
    eng = acquisition.AcquisitionEngine(gpu, n, FS)
    d = np.arange(0, 2001, 250, dtype=np.float32)
    carr, tabs = oracle.doppler_tables(0.0, d, FS, n)
    eng.set_doppler_tables(tabs, carr)
    cells = eng.search_cells(x, K, prn_mask=1 << 4)
    ref = oracle.AcqWorker(5, n, FS).cells(x, tabs, K)
    assert (ref["argmax"] == cells[4]["argmax"]).all()
    np.testing.assert_allclose(cells[4]["peak"], ref["peak"], rtol=REL)
    best = int(cells[4]["peak"].argmax())
    assert d[best] == 1250.0 and cells[4]["argmax"][best] == 12345


def test_batch_snapshot_acquisition_config5_shape(gpu, oracle):
    """BASELINE configs[4] shape on one rank: a batch of short snapshot recordings with different satellites in view,
    searched one after the other through `sharding.search_batch`; every recording's decisions equal the oracle's
    search_satellite (early-exit order included) for each PRN."""
    from gnss_sdr_rs_b200 import acquisition, sdr_mock, sharding
    fs, n, K = 4.092e6, 4092, 4
    dopp = np.arange(-4000, 4001, 500, dtype=np.float32)
    views = [[(3, 1500.0, 100), (17, -2500.0, 3000)], [(8, 0.0, 2222)], [], [(30, 3500.0, 4000), (3, -500.0, 17), (21, 1000.0, 999)]]
    recs = [sdr_mock.baseband(fs, K, [{"prn": p, "doppler": d, "code_phase": c, "cn0_dbhz": 52.0} for p, d, c in v],
                              seed=100 + i) for i, v in enumerate(views)]
    eng = acquisition.AcquisitionEngine(gpu, n, fs)
    carr, tabs = oracle.doppler_tables(0.0, dopp, fs, n)
    eng.set_doppler_tables(tabs, carr)
    table = sharding.search_batch(eng, recs, K)
    assert table.shape == (len(recs), 32, 6)
    workers = {p: oracle.AcqWorker(p, n, fs) for p in range(1, 33)}
    for i, v in enumerate(views):
        found = {p + 1 for p in range(32) if table[i, p, 0] == 1}
        assert {p for p, _, _ in v} <= found
        for p in range(1, 33):
            ref = workers[p].search_satellite(recs[i], tabs, carr, 0, K)
            assert (ref is not None) == (p in found), (i, p)
            if ref:
                assert table[i, p - 1, 2] == ref["code_phase_samples"] and table[i, p - 1, 3] == ref["carrier_freq"]
