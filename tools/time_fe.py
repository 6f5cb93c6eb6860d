#!/usr/bin/env python3
"""Digital front-end: exact (one CTA) against tolerance mode (segmented scan), wall clock per call incl. H2D."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gnss_sdr_rs_b200._ffi as ffi
from gnss_sdr_rs_b200 import ring

hd = ffi.Handle(0)
rng = np.random.default_rng(0)
rb = ring.MulticastRingBuffer(hd, 1 << 24)
for n in (2048, 131072, 1 << 20, 1 << 24):
    blk = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    pin = torch.from_numpy(blk.view(np.float32)).pin_memory()
    for par in (False, True):
        if not par and n > (1 << 20):
            continue
        fe = ring.DigitalFrontend(hd, 4130400.0, 16367600.0, parallel=par)
        f = lambda: hd.call("gb_frontend_write", pin.data_ptr(), n)
        for _ in range(3): f()
        hd.call("gb_synchronize")
        reps = 20 if n <= (1 << 20) else 5
        t0 = time.perf_counter()
        for _ in range(reps): f()
        hd.call("gb_synchronize")
        dt = (time.perf_counter() - t0) / reps
        print("n=%9d %-8s %9.3f us per call  %9.1f Msamples/s  (%.1f GB/s in+out)" % (n, "parallel" if par else "exact", dt * 1e6, n / dt / 1e6, 16 * n / dt / 1e9))
hd.close()
