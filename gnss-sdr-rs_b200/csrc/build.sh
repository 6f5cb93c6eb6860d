#!/bin/bash
# Builds gnss-sdr-rs_b200/libgnss_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libgnss_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
# GB_TUNING=1 ./build.sh -f adds the A/B tuning variants of the acquisition plans (not part of the shipped library)
if [ -n "$GB_TUNING" ]; then FLAGS="$FLAGS -DGB_TUNING"; fi
mkdir -p ../build
need=0
for f in acq_kernels acq_lw acq_generic frontend acq_cluster trk_kernels trk_ws fine_doppler gnss_b200; do
  if [ ! -f ../build/$f.o ] || [ -n "$(find . ../../include -newer ../build/$f.o \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then need=1; fi
done
if [ $need -eq 0 ] && [ -f $OUT ] && [ -f ../libgnss_b200.a ] && [ "$1" != "-f" ]; then exit 0; fi
nvcc $FLAGS -c acq_kernels.cu -o ../build/acq_kernels.o & p1=$!
nvcc $FLAGS -fmad=false -c trk_kernels.cu -o ../build/trk_kernels.o & p2=$!
nvcc $FLAGS -c acq_cluster.cu -o ../build/acq_cluster.o & p4=$!
nvcc $FLAGS -c gnss_b200.cu -o ../build/gnss_b200.o & p3=$!
nvcc $FLAGS -c fine_doppler.cu -o ../build/fine_doppler.o & p5=$!
nvcc $FLAGS -c acq_lw.cu -o ../build/acq_lw.o & p6=$!
nvcc $FLAGS -fmad=false -c frontend.cu -o ../build/frontend.o & p7=$!
nvcc $FLAGS -fmad=false -c trk_ws.cu -o ../build/trk_ws.o & p8=$!
nvcc $FLAGS -c acq_generic.cu -o ../build/acq_generic.o & p9=$!
wait $p1; wait $p2; wait $p3; wait $p4; wait $p5; wait $p6; wait $p7; wait $p8; wait $p9
nvcc -shared -o $OUT ../build/acq_kernels.o ../build/acq_lw.o ../build/acq_generic.o ../build/frontend.o ../build/acq_cluster.o ../build/trk_kernels.o ../build/trk_ws.o ../build/fine_doppler.o ../build/gnss_b200.o
# the same objects as a static archive for the reference's build.rs / src/c_lib path (links like libconvenience.a:
# -lgnss_b200 -lcudart_static -ldl -lrt -lpthread -lstdc++)
rm -f ../libgnss_b200.a
ar rcs ../libgnss_b200.a ../build/acq_kernels.o ../build/acq_lw.o ../build/acq_generic.o ../build/frontend.o ../build/acq_cluster.o ../build/trk_kernels.o ../build/trk_ws.o ../build/fine_doppler.o ../build/gnss_b200.o
echo "built $OUT"
