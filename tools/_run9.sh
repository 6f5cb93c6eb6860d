python -m pytest tests/test_gpu_tracking.py -m gpu -x -q -k "full_run" 2>&1 | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -5 gpurun_out/bench_n2.err
