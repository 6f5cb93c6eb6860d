python bench.py --steps 20 --warmup 3 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_acq.csv python bench.py --steps 5 --warmup 3 --acq-only --no-pipeline > gpurun_out/ncu12a.log 2>&1; echo ncu-list rc=$?
python tools/time_e2e.py 2>&1 | tail -8
