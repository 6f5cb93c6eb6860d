#!/usr/bin/env python3
"""Times the acquisition kernels (CUDA events inside the library) for one workload; tuning helper.
usage: time_acq.py [config2|config1|n=<N>,k=<K>,d=<D>,coh=<n_coh>] [shared|plain|fused] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gnss_sdr_rs_b200._ffi as ffi  # noqa: E402
if os.environ.get("GB_LIB"):   # tool-side only: an A/B build of the library (tools/build_alt.sh)
    ffi.LIB_PATH = os.path.abspath(os.environ["GB_LIB"])
from gnss_sdr_rs_b200 import acquisition, ring  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    mode = sys.argv[2] if len(sys.argv) > 2 else "shared"
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    if wl == "config2":
        n, K, D, coh = 4092, 200, 201, 10
    elif wl == "config1":
        n, K, D, coh = 16368, 10, 29, 1
    else:
        kv = dict(p.split("=") for p in wl.split(","))
        n, K, D, coh = int(kv["n"]), int(kv["k"]), int(kv["d"]), int(kv.get("coh", 1))
    fs = n * 1000.0
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(n * K) + 1j * rng.standard_normal(n * K)).astype(np.complex64)
    for kv in filter(None, os.environ.get("TUNE", "").split(",")):   # TUNE=acq_nolw=1,acq_variant=2 (tool-side only)
        k, v = kv.split("=")
        ffi.tuning_set(k, int(v))
    hd = ffi.Handle(0)
    eng = acquisition.AcquisitionEngine(hd, n, fs)
    eng.make_doppler_tables(0.0, np.linspace(-5000, 5000, D).astype(np.float32))
    eng.set_coherent(coh)
    eng.set_doppler_aliasing(os.environ.get('ALIAS', '1') != '0')
    eng.set_mode({"fused": ffi.GB_ACQ_FUSED, "plain": ffi.GB_ACQ_SHARED_PLAIN}.get(mode, ffi.GB_ACQ_SHARED))
    # samples resident in the device ring: the timed region holds kernels only (the host-pointer call overlaps its
    # sliced upload with the forward path inside the same events)
    cap = 1
    while cap < n * K:
        cap <<= 1
    rb = ring.MulticastRingBuffer(hd, cap)
    rb.write_samples(x)
    ms = []
    for r in range(reps + 2):
        eng.search_cells_ring(0, K, want_cells=False)
        ms.append(eng.last_kernel_ms())
    ms = ms[2:]
    cells = 32 * D * n
    print("%s %s variant=%s  N=%d K=%d D=%d coh=%d  kernel_ms min %.3f med %.3f  cells/s %.3e" % (
        wl, mode, ffi.lib().gb_tuning_get(b"acq_variant", 0), n, K, D, coh, min(ms), float(np.median(ms)), cells / (min(ms) * 1e-3)))
    hd.close()


if __name__ == "__main__":
    main()
