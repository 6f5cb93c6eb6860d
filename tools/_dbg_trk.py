import sys; sys.path.insert(0,'/root/repo')
import bench, gnss_sdr_rs_b200._ffi as ffi
hd = ffi.Handle(0)
stream = bench.tracking_stream(2100)
for (ch, ep) in ((1024, 2000), (1024, 1000), (128, 2000), (128, 2000), (1024, 2000)):
    r = bench.tracking_numbers(hd, ffi, ch, ep, stream=stream)
    print("channels %d epochs %d kernel_ms %.3f locked %d" % (ch, ep, r["kernel_ms"], r["locked_channels"]), flush=True)
hd.close()
