#!/usr/bin/env python3
"""Times the persistent tracking kernel: time_trk.py [channels] [epochs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gnss_sdr_rs_b200._ffi as ffi  # noqa: E402

ch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ep = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for kv in filter(None, os.environ.get("TUNE", "").split(",")):   # TUNE=trk_ws=880 (tool-side only)
    k, v = kv.split("=")
    ffi.tuning_set(k, int(v))
hd = ffi.Handle(0)
r = bench.tracking_numbers(hd, ffi, ch, ep)
print(os.environ.get("TUNE", ""), "channels %d epochs %d kernel_ms %.3f  ch-epochs/s %.3e  x_realtime %.1f  locked %d  us/epoch %.3f" % (
    ch, ep, r["kernel_ms"], r["value"], r["x_realtime"], r["locked_channels"], r["kernel_ms"] * 1e3 / ep))
hd.close()
