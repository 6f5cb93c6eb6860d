#!/usr/bin/env python3
"""bench.py -- headline benchmark of the acquisition / correlator hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload acq|trk]
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Workload "acq" (default) = BASELINE.json configs[1]: synthetic GPS L1 C/A at 4.092 Msps, 32 PRNs,
10 ms coherent x 20 non-coherent (200 ms of signal), 50 Hz Doppler step over +-5 kHz (D = 201).
A step = one full search of one recording; metric = PRN x Doppler x code-phase cells per second
(P*D*N = 26 319 744 cells per step).  With N GPUs every rank searches its own recording (sharding by
recording) and the per-PRN results are all-gathered over NCCL each step: weak scaling.

`value`  : device time (CUDA events on the library's acquisition stream) with the IQ already in the
           HBM sample ring; L2 is flushed between timed steps.
`e2e`    : wall clock through the C-ABI call gb_acq_search() with PINNED HOST IQ: H2D copy, kernel,
           D2H of the cells and the host decision scan inside the timed region, every step.  The headline
           figure runs the call from two host threads on two handles of the GPU (steps alternate, so one
           search's upload overlaps the other's inverse kernel); `serial_*` is one thread, one call at a time.
`roofline`: the fused kernel is FP32-pipe / shared-memory bound, so the denominator is the
           FP32 FMA rate measured live by gb_bench_fp32_tflops(); the HBM view is reported beside it.
`cpu_baseline` / --impl reference: the CPU oracle (oracle/, the C restatement of the reference's
           algorithm -- the Rust reference cannot be built here) on all host cores, on a bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS, N_FFT, N_PRN = 4.092e6, 4092, 32
N_COH, N_NONCOH = 10, 20
K_MS = N_COH * N_NONCOH
DOPPLERS = np.arange(-5000.0, 5000.0 + 1e-3, 50.0, dtype=np.float32)
SATS = [(3, 1230.0, 100, 45.0), (7, -2210.0, 2000, 40.0), (11, 3370.0, 3100, 42.0), (14, -440.0, 777, 50.0),
        (19, 4120.0, 1500, 38.0), (22, -3900.0, 4000, 36.0), (28, 60.0, 2500, 47.0), (31, 2780.0, 300, 35.0)]


# filled in from the ncu capture of the same command (profiles/round1_v9_final.txt)
KERNEL_SHARES_NOTE = ("acq_inverse_lw_kernel ~98% / acq_forward_kernel (20 of 201 bins, Doppler aliasing) ~1% / permute < 1% "
                      "of the chain (profiles/round1_v9_final.txt)")


def make_recording(seed):
    from gnss_sdr_rs_b200 import sdr_mock
    sats = [{"prn": p, "doppler": d, "code_phase": c, "cn0_dbhz": cn} for p, d, c, cn in SATS]
    return sdr_mock.baseband(FS, K_MS, sats, seed=seed, nav=True)


def acq_flops(n_forward=None):
    """Algorithmic FLOPs of one step (conventions of SURVEY 8d: FFT = 5 N log2 N, cmul 6, |.|^2 3, add 1,
    wipe-off cmul 6, rotate-accumulate 8).  Returns (minimal, fused, as_run): `minimal` is SURVEY 8d's structure (one
    forward path per Doppler bin, shared by the 32 PRNs) and stays the roofline numerator; `as_run` counts the forward
    path only for the n_forward bins the shared chain really transforms (Doppler aliasing: 20 of 201)."""
    P, D, N, K, G = N_PRN, len(DOPPLERS), N_FFT, K_MS, N_NONCOH
    lg = math.log2(N)
    shared_per_d = K * N * 14 + G * 5 * N * lg          # wipe-off + coherent pre-sum + forward FFT
    per_pd = G * N * (6 + 5 * lg + 3 + 1)               # x conj(code), IFFT, |.|^2, accumulate
    minimal = D * shared_per_d + P * D * per_pd         # forward path shared by the 32 PRNs
    fused = P * D * (shared_per_d + per_pd)             # the single-kernel form repeats the forward path per PRN
    as_run = (D if n_forward is None else n_forward) * shared_per_d + P * D * per_pd
    return minimal, fused, as_run


def acq_bytes(shared=True, n_forward=None, n_shift=1):
    """Algorithmic HBM bytes of one search: IQ + code spectra (one set per distinct alias shift) + wipe-off tables of
    the transformed bins + cells, plus (shared chain) the forward spectra written once and read once."""
    P, D, N, K, G = N_PRN, len(DOPPLERS), N_FFT, K_MS, N_NONCOH
    F = D if n_forward is None else n_forward
    return 8 * N * K + 8 * N * P * n_shift + 8 * N * F + 16 * P * D + (2 * 8 * N * F * G if shared else 0)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (NVML, 10 ms period; nvidia-smi CLI as fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self.stop_flag = index, [], 0, set(), False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
                nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self.stop_flag:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = get_reasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.01)
        except Exception:
            self._cli()

    def _cli(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            while not self.stop_flag:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
                r = [c.strip() for c in out.strip().split(",")]
                self.sm.append(float(r[0]))
                self.mx = max(self.mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": float(self.mx) or None,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_baseline_acq(x, n_threads, max_prns=None):
    """The oracle (kind "port") on a bounded sample: one worker per PRN on all host threads, all 201 bins,
    200 ms, same coherent/non-coherent plan (pre-summed, the cheaper of the oracle's two forms)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as orc
    n_prn = min(N_PRN, max_prns or max(8, n_threads))
    carr, tabs = orc.doppler_tables(0.0, DOPPLERS, FS, N_FFT)
    rot = orc.coh_rotators(carr, FS, N_FFT, N_COH)
    workers = [orc.AcqWorker(p, N_FFT, FS) for p in range(1, n_prn + 1)]
    t0 = time.perf_counter()
    cells = orc.acq_cells_all(workers, x, tabs, K_MS, n_coh=N_COH, rot=rot, presum=1, n_threads=n_threads)
    dt = time.perf_counter() - t0
    n_cells = n_prn * len(DOPPLERS) * N_FFT
    return n_cells / dt, dt, n_prn, cells


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    x = make_recording(0x6E56)
    vals = []
    for s in range(args.warmup + args.steps):
        v, dt, n_prn, _ = cpu_baseline_acq(x, cores, max_prns=max(8, cores) if cores <= 32 else 32)
        if s >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    sample = "%d of 32 PRNs x %d Doppler bins x %d ms (pre-summed %d ms coherent x %d)" % (
        n_prn, len(DOPPLERS), K_MS, N_COH, N_NONCOH)
    line = {"impl": "reference", "metric": "acq_cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean([d for _, d in vals])) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU oracle (C restatement of the reference algorithm; the Rust/rustfft reference cannot be built "
                    "here). rustfft+AVX is expected to be 2-4x faster than this scalar mixed-radix FFT."}
    print(json.dumps(line), flush=True)


def workload_config():
    return {"workload": "BASELINE configs[1]: GPS L1 C/A 4.092 Msps, 32 PRNs, 10 ms coherent x 20 non-coherent, "
                        "50 Hz Doppler step (+-5 kHz, D=201), 200 ms recording per step",
            "fft_size": N_FFT, "n_prn": N_PRN, "n_doppler": len(DOPPLERS), "n_coherent": N_COH,
            "n_noncoherent": N_NONCOH, "cells_per_step": N_PRN * len(DOPPLERS) * N_FFT,
            "l2": "flushed between timed steps (256 MiB write)", "sharding": "one recording per GPU + NCCL all_gather"}


def tracking_numbers(hd, ffi, n_channels=1024, n_epochs=1000, want_state=False):
    """BASELINE configs[2] shape, shortened: 1024 channels (32 PRN-slots x 32 hand-over perturbations) on one shared
    2.048 Msps stream, persistent kernel, FAST mode.  Returns channel-epochs/s from the kernel's CUDA events."""
    from gnss_sdr_rs_b200 import ring, sdr_mock, tracking
    fs, n = 2.048e6, 2048
    prns = [2, 5, 9, 12, 17, 21, 25, 30]
    sats = [{"prn": p, "doppler": float(-2000 + 500 * i), "code_phase": 137 * (i + 1), "cn0_dbhz": 48.0}
            for i, p in enumerate(prns)]
    base_ms = 100
    # periodic construction: Dopplers are multiples of 10 Hz and there is no code Doppler, so a 100 ms tile repeats
    tile = np.zeros(n * base_ms, np.complex64)
    rng = np.random.default_rng(0x6E57)
    t = np.arange(n * base_ms, dtype=np.float64)
    for s in sats:
        code = sdr_mock.ca_code(s["prn"]).astype(np.float64)
        chip = ((t - s["code_phase"]) * 1.023e6 / fs) % 1023.0
        amp = np.sqrt(10.0 ** (s["cn0_dbhz"] / 10.0) / fs)
        tile += (amp * code[np.floor(chip).astype(np.int64) % 1023]
                 * np.exp(2j * np.pi * ((s["doppler"] * t / fs) % 1.0))).astype(np.complex64)
    reps = (n_epochs + 2 + base_ms - 1) // base_ms + 1
    rb = ring.MulticastRingBuffer(hd, 1 << int(math.ceil(math.log2(n * base_ms * reps))))
    for r in range(reps):
        noise = (rng.standard_normal(len(tile)) + 1j * rng.standard_normal(len(tile))).astype(np.complex64) * np.float32(
            1 / np.sqrt(2))
        rb.write_samples(tile + noise)
    ch = tracking.channel_array(n_channels, fs)
    for c in range(n_channels):
        s = sats[c % len(sats)]
        tracking.start(ch[c], s["prn"], s["doppler"] + float(rng.uniform(-50, 50)), float(rng.uniform(0, 0.3)),
                       s["code_phase"], fs, code_row=s["prn"] - 1)
    eng = tracking.TrackingEngine(hd)
    eng.upload(ch)
    eng.run(20)  # warm-up epochs
    eng.run(n_epochs)
    ms = eng.last_kernel_ms()
    eng.download(ch)
    locked = sum(1 for c in range(n_channels) if ch[c].state == 1)
    done = sum(int(ch[c].epochs_done) for c in range(n_channels)) - 20 * n_channels
    out = {"metric": "tracking_channel_epochs_per_sec", "value": done / (ms * 1e-3), "unit": "channel-epochs/s",
           "channels": n_channels, "epochs": n_epochs, "kernel_ms": ms, "locked_channels": locked,
           "x_realtime": (n_epochs * 1e-3) / (ms * 1e-3), "mode": "fast (persistent kernel, on-device loop filters)",
           "fs": fs}
    if want_state:
        out["state"], out["sats"] = ch, sats
    return out


def extra_numbers(hd, ffi):
    """Secondary measurements on the other BASELINE configs / SURVEY 8f rows (kernel time from CUDA events)."""
    import ctypes as C
    from gnss_sdr_rs_b200 import acquisition, ring, sdr_mock
    out = {}
    rng = np.random.default_rng(1)
    # config 1 (reference grid): N = 16368, 29 bins, 10 x 1 ms, 32 PRNs, int8 IF recording through the ring
    raw, _ = sdr_mock.if_recording(10)
    rb = ring.MulticastRingBuffer(hd, 1 << 18)
    rb.write_samples(raw)
    eng = acquisition.AcquisitionEngine(hd, 16368, 16367600.0)
    eng.make_doppler_tables(4130400.0, np.array(acquisition.reference_doppler_grid(), np.float32))
    ms = []
    for _ in range(5):
        res = eng.search_ring(0, 10)
        ms.append(eng.last_kernel_ms())
    out["config1_reference_grid"] = {"fft_size": 16368, "n_doppler": 29, "num_integrations": 10, "kernel_ms": min(ms),
                                     "cells_per_sec": 32 * 29 * 16368 / (min(ms) * 1e-3),
                                     "detected_prns": sorted(r["prn"] for r in res if r),
                                     "x_realtime_vs_10ms_dwell": 10.0 / min(ms)}
    # Galileo-E1-like 4 ms code at 20 Msps: cluster / DSMEM plan, 8 PRNs x 41 bins x 5 blocks (20 ms)
    n = 80000
    codes = np.stack([sdr_mock.resample_code(sdr_mock.e1_surrogate_code(p), 1.023e6, 20e6, n, boc11=True) for p in range(1, 9)])
    x = (rng.standard_normal(5 * n) + 1j * rng.standard_normal(5 * n)).astype(np.complex64)
    eng = acquisition.AcquisitionEngine(hd, n, 20e6, n_prn=8, codes=codes)
    eng.make_doppler_tables(0.0, np.arange(-2500, 2501, 125, dtype=np.float32))
    ms = []
    for _ in range(4):
        eng.search_cells(x, 5)
        ms.append(eng.last_kernel_ms())
    out["galileo_e1_like_n80000_cluster"] = {"fft_size": n, "n_prn": 8, "n_doppler": 41, "num_integrations": 5,
                                             "kernel_ms": min(ms), "cells_per_sec": 8 * 41 * n / (min(ms) * 1e-3)}
    # fine Doppler (N3): the ten satellites of the config-1 stand-in, 11 ms at 16.3676 Msps, 2^21-point zero-padded
    # spectrum each (never materialised)
    raw11, truth = sdr_mock.if_recording(11)
    x11 = sdr_mock.i8_to_c32(raw11)
    reqs = [(t["prn"], t["code_phase"]) for t in truth]
    fine = acquisition.finer_doppler(hd, x11, reqs, sdr_mock.CONFIG_FS, is_complex=False)
    ms = []
    for _ in range(3):
        fine = acquisition.finer_doppler(hd, x11, reqs, sdr_mock.CONFIG_FS, is_complex=False)
        ms.append(float(hd.L.gb_acq_fine_last_kernel_ms(hd.h)))
    out["fine_doppler_config1"] = {"n_satellites": len(reqs), "fft_size": int(fine[0]["fft_size"]), "kernel_ms": min(ms),
                                   "max_abs_err_hz": float(max(abs(float(f["carrier_freq"]) - t["carrier"])
                                                               for f, t in zip(fine, truth))),
                                   "bin_hz": sdr_mock.CONFIG_FS / float(fine[0]["fft_size"])}
    # digital front-end (N2): 64 rf_thread blocks of 2048 raw samples into the ring
    fe = ring.DigitalFrontend(hd, 4130400.0, 16367600.0)
    rb = ring.MulticastRingBuffer(hd, 1 << 20)
    blk = (rng.standard_normal(64 * 2048) + 1j * rng.standard_normal(64 * 2048)).astype(np.complex64)
    fe.process_block_into_ring(blk)
    hd.call("gb_synchronize")
    t0 = time.perf_counter()
    for _ in range(4):
        fe.process_block_into_ring(blk)
    hd.call("gb_synchronize")
    dt = (time.perf_counter() - t0) / 4
    out["digital_frontend"] = {"samples_per_call": int(blk.size), "wall_ms_per_call": dt * 1e3,
                               "msamples_per_sec": blk.size / dt / 1e6,
                               "note": "bit-exact: NCO indices from the precomputed f32 phase orbit, the 16 DC-bias recurrences "
                                       "sequential (FMUL -> FADD per 8 samples): one CTA per stream, bound by that chain; "
                                       "wall clock includes the pinned-less H2D of the block"}
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import gnss_sdr_rs_b200._ffi as ffi
    from gnss_sdr_rs_b200 import acquisition, ring

    if not torch.cuda.is_available() or ffi.lib().gb_device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libgnss_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout during the first collective; keep stdout to the ONE JSON line
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)

    from gnss_sdr_rs_b200 import sharding
    by_prn = args.shard == "prn" and world > 1
    prn_mask = sharding.prn_mask_for_rank(rank, world, N_PRN) if by_prn else 0xFFFFFFFF
    hd = ffi.Handle(local_rank)
    x = make_recording(0x6E56 + (0 if by_prn else rank))
    rb = ring.MulticastRingBuffer(hd, 1 << 20)
    rb.write_samples(x)
    x_pin = torch.from_numpy(x.view(np.float32).copy()).pin_memory()
    x_pin_ptr = x_pin.data_ptr()

    eng = acquisition.AcquisitionEngine(hd, N_FFT, FS, N_PRN)
    carr = eng.make_doppler_tables(0.0, DOPPLERS)
    eng.set_coherent(N_COH)
    eng.set_detector(7.0, 4)
    eng.set_mode(ffi.GB_ACQ_FUSED if args.acq_mode == "fused" else ffi.GB_ACQ_SHARED)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cuda_dev = torch.device("cuda", local_rank)
    gatherer = sharding.ResultGatherer(dist, cuda_dev, N_PRN) if dist is not None else None
    raw_gatherer = sharding.RawResultGatherer(dist, cuda_dev, N_PRN, ffi.AcqResult) if dist is not None else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.zero_()
        # re-align the ranks after the (untimed) L2 flush so the timed gather measures the collective, not flush skew
        barrier()
        t0 = time.perf_counter()
        g_ms = 0.0
        if dist is None:
            res = eng.search_ring(0, K_MS, prn_mask=prn_mask)
            ms = eng.last_kernel_ms()
        else:
            # results land in the gatherer's pinned buffer; the one collective of the path ships the raw structs of
            # every rank, NCCL over NVLink
            eng.search_ring_raw(0, K_MS, prn_mask=prn_mask, out=raw_gatherer.results)
            ms = eng.last_kernel_ms()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            per_rank = raw_gatherer.gather()
            e1.record()
            torch.cuda.synchronize()
            g_ms = e0.elapsed_time(e1)
            if by_prn:   # every PRN is owned by exactly one rank
                res = [None] * N_PRN
                for arr in per_rank:
                    for r in arr:
                        if r.found:
                            res[r.prn - 1] = r.as_dict()
            else:
                res = [r.as_dict() if r.found else None for r in per_rank[rank]]
        return ms + g_ms, (time.perf_counter() - t0) * 1e3, res

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms, res = 0.0, None
    for _ in range(args.steps):
        ms, _, res = step_device()
        dev_ms += ms
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    kernel_ms_avg = dev_ms / args.steps

    # e2e: pinned host IQ -> gb_acq_search (H2D + kernel + D2H + decision) per step
    for _ in range(2):
        eng.search(x_pin_ptr, K_MS, prn_mask=prn_mask)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = eng.search(x_pin_ptr, K_MS, prn_mask=prn_mask)
        if dist is not None:
            gatherer.gather(res_e2e)
    barrier()
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    # e2e, pipelined: the same public call from TWO host threads on two handles of this GPU (each with its own streams
    # and buffers), steps dealt alternately -- the upload of one search overlaps the inverse kernel of the other, as a
    # receiver that acquires recording after recording would run it.  Every step still copies its 6.5 MB of pinned host
    # IQ to the device and reads its cells / results back inside the timed region; with N GPUs the per-step gather is
    # issued by the main thread in step order.
    e2e_ms, e2e_api = e2e_serial_ms, "gb_acq_search (pinned host IQ -> results)"
    if not args.no_pipeline:
        hd2 = ffi.Handle(local_rank)
        eng2 = acquisition.AcquisitionEngine(hd2, N_FFT, FS, N_PRN)
        eng2.make_doppler_tables(0.0, DOPPLERS)
        eng2.set_coherent(N_COH)
        eng2.set_detector(7.0, 4)
        eng2.set_mode(ffi.GB_ACQ_FUSED if args.acq_mode == "fused" else ffi.GB_ACQ_SHARED)
        engines = (eng, eng2)

        def pipelined(n_steps):
            out = [None] * n_steps
            done = [threading.Event() for _ in range(n_steps)]
            errs = []

            def worker(i):
                try:
                    for k in range(i, n_steps, 2):
                        out[k] = engines[i].search(x_pin_ptr, K_MS, prn_mask=prn_mask)
                        done[k].set()
                except Exception as e:  # surface in the main thread
                    errs.append(e)
                    for ev in done:
                        ev.set()

            th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
            for t_ in th:
                t_.start()
            for k in range(n_steps):
                done[k].wait()
                if errs:
                    break
                if dist is not None:
                    gatherer.gather(out[k])
            for t_ in th:
                t_.join()
            if errs:
                raise errs[0]
            return out

        pipelined(4)
        barrier()
        t0 = time.perf_counter()
        res_pipe = pipelined(args.steps)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        e2e_api = "gb_acq_search (pinned host IQ -> results), two host threads / two handles per GPU, steps alternating"
        same = all((a is None) == (b is None) and (a is None or (a["code_phase_samples"] == b["code_phase_samples"] and
                                                                   a["carrier_freq"] == b["carrier_freq"]))
                   for a, b in zip(res_pipe[-1], res_e2e))
        if not same:
            raise RuntimeError("pipelined e2e search disagrees with the serial one")
        hd2.close()
    clocks = sampler.finish()

    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms, wall_ms, e2e_serial_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, wall_ms, e2e_serial_ms = [float(v) for v in t.cpu()]
    cells = N_PRN * len(DOPPLERS) * N_FFT
    ms_per_step = dev_ms / args.steps
    units = 1 if by_prn else world   # recordings searched per step by the whole job
    value = units * cells / (ms_per_step * 1e-3)
    e2e_value = units * cells / (e2e_ms * 1e-3)

    # BASELINE configs[2] on N GPUs: the 1024 channels are sharded by channel (no exchange between epochs);
    # device time of the persistent kernel, max over ranks
    trk_sharded = None
    if dist is not None and 1024 % world == 0:
        tr, err = None, None
        try:
            tr = tracking_numbers(hd, ffi, 1024 // world, 1000)
        except Exception as e:  # report, never hide; every rank still takes part in the reductions below
            err = repr(e)
        t = torch.tensor([tr["kernel_ms"] if tr else float("inf")], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t2 = torch.tensor([float(tr["locked_channels"]) if tr else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        ms_max = float(t[0])
        if err or not math.isfinite(ms_max):
            trk_sharded = {"error": err or "a rank failed"}
        else:
            trk_sharded = {"metric": "tracking_channel_epochs_per_sec", "value": 1024 * 1000 / (ms_max * 1e-3),
                           "unit": "channel-epochs/s", "channels": 1024, "channels_per_gpu": 1024 // world, "epochs": 1000,
                           "kernel_ms_max_over_ranks": ms_max, "locked_channels": int(t2[0]),
                           "x_realtime": 1.0 / (ms_max * 1e-3), "mode": tr["mode"], "fs": tr["fs"],
                           "sharding": "by channel, no collective between epochs"}

    if rank == 0:
        peak_tf = ctypes_float(hd, "gb_bench_fp32_tflops")
        n_fwd = eng.forward_bins() if args.acq_mode != "fused" else len(DOPPLERS)
        n_shift = int(math.ceil(len(DOPPLERS) / n_fwd)) if n_fwd < len(DOPPLERS) else 1
        minimal, fused_flops, shared_flops = acq_flops(n_fwd)
        as_run = fused_flops if args.acq_mode == "fused" else shared_flops
        abytes = acq_bytes(args.acq_mode != "fused", n_fwd, n_shift)
        achieved = minimal / (kernel_ms_avg * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                "acq_fused_4092_bytes_per_launch" if args.acq_mode == "fused" else "acq_chain_4092_bytes_per_launch")
        except Exception:
            pass
        found = sorted(r["prn"] for r in res if r)
        line = {"metric": "acq_cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if by_prn else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
                "wall_ms_per_step_incl_l2_flush": wall_ms / args.steps,
                "x_realtime": units * (K_MS * 1e-3) / (ms_per_step * 1e-3),
                "e2e": {"value": e2e_value, "unit": "cells/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(N_PRN * len(DOPPLERS) * 16),
                        "api": e2e_api, "serial_ms_per_step": e2e_serial_ms,
                        "serial_value": units * cells / (e2e_serial_ms * 1e-3),
                        "serial_api": "one host thread, one gb_acq_search call at a time"},
                # per step: line-order permutation of the IQ blocks (prime-factor plan) + forward + inverse kernels
                "gpu_launches": args.steps * (2 if args.acq_mode == "fused" else 3),
                "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic,
                             "peak_source": "measured live: gb_bench_fp32_tflops FMA probe (MEASURED_PEAKS.json has no "
                                            "FP32 figure; theoretical 148*128*2*1.965 GHz = 74.5)",
                             "flops_per_launch_minimal": minimal, "flops_per_launch_as_run": as_run,
                             "forward_spectra_per_group": n_fwd,
                             "kernel": ("acq_fused_kernel<PfaPlan<4092,160,4,12,11,31>>" if args.acq_mode == "fused" else
                                        "permute_blocks_kernel + acq_forward_kernel<PfaPlan<4092,160,4,12,11,31>> + "
                                        "acq_inverse_lw_kernel<PfaPlan<4092,128,4,12,11,31>> (128 working threads + leftover warp)"),
                             "kernel_ms": kernel_ms_avg,
                             "kernel_shares_ncu": KERNEL_SHARES_NOTE,
                             "hbm_view": {"bound": "hbm", "algorithmic_bytes": abytes,
                                          "achieved": abytes / (kernel_ms_avg * 1e-3) / 1e9,
                                          "peak": hbm_peak, "unit": "GB/s",
                                          "frac": abytes / (kernel_ms_avg * 1e-3) / 1e9 / hbm_peak,
                                          "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"}},
                "clocks": clocks, "detected_prns": found}
        if trk_sharded is not None:
            line["tracking_sharded"] = trk_sharded
        if world == 1:
            cores = os.cpu_count() or 1
            v, dt, n_prn, ocells = cpu_baseline_acq(x, cores)
            line["cpu_baseline"] = {"value": v, "unit": "cells/s", "cores": cores, "kind": "port",
                                    "sample": "%d of 32 PRNs x 201 bins x 200 ms, %.1f s" % (n_prn, dt)}
            # the bench doubles as a full-size parity check on the sampled PRNs
            gcells = eng.search_cells_ring(0, K_MS)
            strong = ocells["peak"] > 4.0 * np.median(ocells["peak"], axis=1, keepdims=True)
            line["parity_vs_oracle"] = {
                "prns": n_prn, "max_rel_peak_err": float(np.abs(gcells["peak"][:n_prn] / ocells["peak"] - 1).max()),
                "argmax_equal_frac": float((gcells["argmax"][:n_prn] == ocells["argmax"]).mean()),
                "argmax_equal_strong_cells": bool((gcells["argmax"][:n_prn][strong] == ocells["argmax"][strong]).all())}
            try:
                line["tracking"] = tracking_numbers(hd, ffi)
                line["tracking_128ch"] = tracking_numbers(hd, ffi, 128, 1000)
                # BASELINE configs[2] at full length: 1024 channels x 60 s = 61.44 M channel-epochs in one launch
                line["tracking_config3_full_60s"] = tracking_numbers(hd, ffi, 1024, 60000)
            except Exception as e:  # report, never hide
                line["tracking"] = {"error": repr(e)}
            try:
                line["extras"] = extra_numbers(hd, ffi)
            except Exception as e:
                line["extras"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    hd.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def ctypes_float(hd, name):
    import ctypes as C
    v = C.c_float(0)
    hd.call(name, C.byref(v))
    return float(v.value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--acq-mode", default="shared", choices=["shared", "fused"])
    ap.add_argument("--no-pipeline", action="store_true", help="e2e from one host thread only")
    ap.add_argument("--shard", default="recording", choices=["recording", "prn"],
                    help="N>1: one recording per GPU (weak scaling, default) or the PRNs of ONE recording dealt to the GPUs (strong)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
