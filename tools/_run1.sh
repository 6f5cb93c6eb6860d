set -x
python -m pytest tests/test_gpu_tracking.py -x -q 2>&1 | tail -15
for t in trk_ws=-1 trk_ws=880 trk_ws=881 trk_ws=1640 trk_ws=1641; do TUNE=$t python tools/time_trk.py 128 2000 2>&1 | tail -1; done
for t in trk_ws=-1 trk_ws=481 trk_ws=480 trk_ws=441 trk_ws=281 trk_ws=881; do TUNE=$t python tools/time_trk.py 1024 1000 2>&1 | tail -1; done
for t in trk_ws=-1 trk_ws=880 trk_ws=481; do TUNE=$t python tools/time_trk.py 256 2000 2>&1 | tail -1; TUNE=$t python tools/time_trk.py 512 2000 2>&1 | tail -1; done
