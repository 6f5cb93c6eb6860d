python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 10 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo bench rc=$?; tail -3 gpurun_out/bench7.err
