python -m pytest tests/test_gpu_tracking.py -x -q 2>&1 | tail -3
for t in trk_ws=888 trk_ws=884 trk_ws=882 trk_ws=881 trk_ws=1644 trk_ws=1642; do TUNE=$t python tools/time_trk.py 128 2000 2>&1 | tail -1; done
for t in trk_ws=-1 trk_ws=481 trk_ws=482 trk_ws=488; do TUNE=$t python tools/time_trk.py 1024 1000 2>&1 | tail -1; done
for t in trk_ws=881 trk_ws=882 trk_ws=481 trk_ws=482; do TUNE=$t python tools/time_trk.py 256 2000 2>&1 | tail -1; TUNE=$t python tools/time_trk.py 512 2000 2>&1 | tail -1; done
TUNE=trk_ws=881 ncu --set full --clock-control none --import-source on -k regex:trk_ws -s 1 -c 1 -o gpurun_out/trk_ws881d_128 -f python tools/time_trk.py 128 2000 > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu6.log
