#!/usr/bin/env python3
"""A/B of the N = 4092 inverse kernels (leftover-warp vs generic vs fused): prints how many cells differ."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnss_sdr_rs_b200._ffi as ffi  # noqa: E402
from gnss_sdr_rs_b200 import acquisition, sdr_mock  # noqa: E402

n, fs = 4092, 4.092e6
hd = ffi.Handle(0)
for K, n_coh in ((1, 1), (4, 1), (7, 1), (8, 1), (9, 1), (20, 1), (40, 2)):
    x = sdr_mock.baseband(fs, K, [{"prn": 5, "doppler": 700.0, "code_phase": 321, "cn0_dbhz": 50.0}], seed=K)
    eng = acquisition.AcquisitionEngine(hd, n, fs)
    eng.make_doppler_tables(0.0, np.arange(-1500, 1501, 250, dtype=np.float32))
    eng.set_coherent(n_coh)
    eng.set_detector(7.0, 4)
    out = {}
    for name, mode in (("lw", ffi.GB_ACQ_SHARED), ("plain", ffi.GB_ACQ_SHARED_PLAIN)):   # Doppler aliasing on in both
        eng.set_mode(mode)
        out[name] = eng.search_cells(x, K).copy()
    eng.set_doppler_aliasing(False)   # the fused kernel has no aliasing: compare it with every bin's own table
    for name, mode in (("plain_noalias", ffi.GB_ACQ_SHARED_PLAIN), ("fused", ffi.GB_ACQ_FUSED)):
        eng.set_mode(mode)
        out[name] = eng.search_cells(x, K).copy()
    a, b = out["lw"], out["plain"]
    bad = (a["peak"] != b["peak"]) | (a["argmax"] != b["argmax"]) | (a["peak2"] != b["peak2"])
    rel = np.abs(a["sum8"] - b["sum8"]) / np.abs(b["sum8"])
    print("K=%d n_coh=%d: lw!=plain (peak, argmax, peak2) in %d of %d cells (max rel sum8 diff %.3g); plain==fused (no aliasing) %s" % (
        K, n_coh, int(bad.sum()), bad.size, float(rel.max()), out["plain_noalias"].tobytes() == out["fused"].tobytes()))
hd.close()
