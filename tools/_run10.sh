python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench10_ref.json 2>> gpurun_out/bench10.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu10a.log 2>&1; echo ncu-list rc=$?
TUNE=acq_dummy=0 ncu --set full --clock-control none --import-source on -k regex:acq_inverse_lw -s 3 -c 1 -o gpurun_out/r2_acq_lw -f python tools/time_acq.py config2 shared 3 > gpurun_out/ncu10b.log 2>&1; echo ncu-acq rc=$?
TUNE=trk_ws=481 ncu --set full --clock-control none --import-source on -k regex:trk_ws -s 1 -c 1 -o gpurun_out/r2_trk_ws481_1024 -f python tools/time_trk.py 1024 1000 > gpurun_out/ncu10c.log 2>&1; echo ncu-trk rc=$?
