#!/usr/bin/env python3
"""Times the cluster plan (N = 80000 = 4 x 20000: Galileo-E1-like 4 ms codes at 20 Msps), 8 codes x 41 bins x 5 periods."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnss_sdr_rs_b200._ffi as ffi  # noqa: E402
from gnss_sdr_rs_b200 import acquisition, sdr_mock  # noqa: E402

for kv in filter(None, os.environ.get("TUNE", "").split(",")):
    k, v = kv.split("=")
    ffi.tuning_set(k, int(v))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = 80000
rng = np.random.default_rng(1)
hd = ffi.Handle(0)
codes = np.stack([sdr_mock.resample_code(sdr_mock.e1_surrogate_code(p), 1.023e6, 20e6, n, boc11=True) for p in range(1, 9)])
x = (rng.standard_normal(K * n) + 1j * rng.standard_normal(K * n)).astype(np.complex64)
eng = acquisition.AcquisitionEngine(hd, n, 20e6, n_prn=8, codes=codes)
eng.make_doppler_tables(0.0, np.arange(-2500, 2501, 125, dtype=np.float32))
ms = []
for _ in range(4):
    cells = eng.search_cells(x, K)
    ms.append(eng.last_kernel_ms())
print("cluster N=80000 8 codes x 41 bins x %d periods: kernel_ms min %.3f  (%.1f us per cluster-period step)  checksum %.6e" % (
    K, min(ms), min(ms) * 1e3 / (K * np.ceil(8 * 41 * 4 / 148.0)), float(cells["peak"].astype(np.float64).sum())))
hd.close()
