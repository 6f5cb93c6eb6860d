#!/usr/bin/env python3
"""Breaks the end-to-end acquisition call down: H2D bandwidth, ring-resident search, host-buffer search."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import gnss_sdr_rs_b200._ffi as ffi
from gnss_sdr_rs_b200 import acquisition, ring

hd = ffi.Handle(0)
x = bench.make_recording(1)
x_pin = torch.from_numpy(x.view(np.float32).copy()).pin_memory()
dev = torch.empty_like(x_pin, device="cuda")
for _ in range(3):
    dev.copy_(x_pin, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    dev.copy_(x_pin, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("torch pinned H2D %.3f ms for %.1f MB = %.1f GB/s" % (dt * 1e3, x.nbytes / 1e6, x.nbytes / dt / 1e9))
# pinned H2D against transfer size: the per-copy fixed cost (~8-10 us of submission + DMA start) and the link's asymptote
big = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
bigd = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for sz in (1 << 16, 1 << 18, 1 << 20, 6547200, 1 << 24, 1 << 26, 1 << 28):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        bigd[:sz].copy_(big[:sz], non_blocking=True)
    torch.cuda.synchronize()
    reps = 20 if sz <= (1 << 24) else 4
    ev0.record()
    for _ in range(reps):
        bigd[:sz].copy_(big[:sz], non_blocking=True)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print("pinned H2D %10d B: %8.1f us  %6.1f GB/s" % (sz, ms * 1e3, sz / ms / 1e6))
del big, bigd
rb = ring.MulticastRingBuffer(hd, 1 << 20)
rb.write_samples(x)
eng = acquisition.AcquisitionEngine(hd, bench.N_FFT, bench.FS, 32)
eng.make_doppler_tables(0.0, bench.DOPPLERS)
eng.set_coherent(bench.N_COH)
eng.set_doppler_aliasing(True)
n_x = int(x.size)
def timeit(f, n=20):
    for _ in range(3): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("search_ring (results)      %.3f ms wall, kernel %.3f" % (timeit(lambda: eng.search_ring(0, bench.K_MS)), eng.last_kernel_ms()))
print("search pinned host         %.3f ms wall, kernel %.3f" % (timeit(lambda: eng.search(x_pin.data_ptr(), bench.K_MS, n_samples=n_x)), eng.last_kernel_ms()))
print("search pageable host       %.3f ms wall, kernel %.3f" % (timeit(lambda: eng.search(x, bench.K_MS)), eng.last_kernel_ms()))
print("search_cells_ring no cells %.3f ms wall" % timeit(lambda: eng.search_cells_ring(0, bench.K_MS, want_cells=False)))
def pipe(n):
    eng.search_enqueue(x_pin.data_ptr(), bench.K_MS, 0, n_samples=n_x)
    raw = (ffi.AcqResult * 32)()
    for k in range(n):
        if k + 1 < n:
            eng.search_enqueue(x_pin.data_ptr(), bench.K_MS, (k + 1) & 1, n_samples=n_x)
        eng.search_wait(k & 1, raw=raw)
for n in (10, 20, 50, 200):
    pipe(4)
    t0 = time.perf_counter(); pipe(n); dt = (time.perf_counter() - t0) / n * 1e3
    print("enqueue/wait pipeline, %3d steps: %.3f ms per step (kernel %.3f)" % (n, dt, eng.last_kernel_ms()))
for n in (20, 100):
    print("search pinned host x%d      %.3f ms wall" % (n, timeit(lambda: eng.search(x_pin.data_ptr(), bench.K_MS, n_samples=n_x), n)))
hd.close()
