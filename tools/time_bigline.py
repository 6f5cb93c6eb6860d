#!/usr/bin/env python3
"""A/B of the big-line acquisition plans (GB_TUNING builds): config 1 (N = 16368) and the 20 Msps code period (N = 20000)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnss_sdr_rs_b200._ffi as ffi
from gnss_sdr_rs_b200 import acquisition, sdr_mock

hd = ffi.Handle(0)
rng = np.random.default_rng(1)
def run(n, fs, n_prn, codes, bins, K, variants):
    x = (rng.standard_normal(K * n) + 1j * rng.standard_normal(K * n)).astype(np.complex64)
    ref = None
    for v in variants:
        ffi.tuning_set("acq_variant", v)
        eng = acquisition.AcquisitionEngine(hd, n, fs, n_prn=n_prn, codes=codes)
        eng.make_doppler_tables(0.0, bins)
        ms = []
        for _ in range(5):
            cells = eng.search_cells(x, K)
            ms.append(eng.last_kernel_ms())
        if ref is None:
            ref = cells
        same = bool((cells["arg"] == ref["arg"]).all()) if "arg" in cells.dtype.names else None
        rel = float(np.abs(cells["peak"] - ref["peak"]).max() / ref["peak"].max())
        print("N=%d variant %d: kernel %.3f ms (min of 5: %s)  argmax same %s, peak rel diff %.2e" % (n, v, min(ms), " ".join("%.3f" % m for m in ms), same, rel), flush=True)
    ffi.tuning_set("acq_variant", 0)
print(ffi.CELL_DTYPE)
run(16368, 16.3676e6, 32, None, np.arange(-7000, 7001, 500, dtype=np.float32), 10, (0, 3, 4, 2))
fs, n = 20.0e6, 20000
codes = np.stack([sdr_mock.resample_code(sdr_mock.ca_code(p), 1.023e6, fs, n) for p in range(1, 33)] * 2).astype(np.int8)
run(n, fs, 64, codes, np.arange(-5000, 5001, 250, dtype=np.float32), 20, (0, 1, 2))
hd.close()
