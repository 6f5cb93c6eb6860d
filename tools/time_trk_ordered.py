#!/usr/bin/env python3
"""Times the ORDERED (parity) tracking mode: time_trk_ordered.py [channels] [epochs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import gnss_sdr_rs_b200._ffi as ffi
ch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ep = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for kv in filter(None, os.environ.get("TUNE", "").split(",")):
    k, v = kv.split("=")
    ffi.tuning_set(k, int(v))
hd = ffi.Handle(0)
stream = bench.tracking_stream(ep + 100)
r = bench.tracking_numbers(hd, ffi, ch, ep, mode=1, stream=stream)
print("ORDERED channels %d epochs %d: %.2f us/epoch, x%.1f" % (ch, ep, r["kernel_ms"] * 1e3 / ep, r["x_realtime"]))
hd.close()
