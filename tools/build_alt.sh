#!/bin/bash
# usage: tools/build_alt.sh <tag> <extra nvcc defines...>   -- A/B build of the library with other compile-time plans:
# gnss-sdr-rs_b200/build_<tag>/libgnss_b200.so (git-ignored; select it tool-side with GB_LIB=<path> tools/time_acq.py ...)
set -e
cd "$(dirname "$0")/../gnss-sdr-rs_b200/csrc"
TAG=$1; shift
OUT=../build_$TAG
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $*"
for f in acq_kernels acq_cluster acq_lw; do nvcc $FLAGS -Xptxas -v -c $f.cu -o $OUT/$f.o > $OUT/$f.ptxas 2>&1 & done
wait
nvcc -shared -o $OUT/libgnss_b200.so $OUT/acq_kernels.o $OUT/acq_cluster.o $OUT/acq_lw.o ../build/acq_generic.o ../build/frontend.o ../build/trk_kernels.o ../build/trk_ws.o ../build/fine_doppler.o ../build/gnss_b200.o
echo "built $OUT/libgnss_b200.so"
