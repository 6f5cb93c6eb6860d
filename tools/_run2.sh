TUNE=trk_ws=880 python tools/time_trk.py 128 2000 2>&1 | tail -1
TUNE=trk_ws=880 ncu --set full --clock-control none --import-source on -k regex:trk_ws -c 2 -o gpurun_out/trk_ws880_128 -f python tools/time_trk.py 128 2000 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
