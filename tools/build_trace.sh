#!/bin/bash
# libatvs_trace.so = libatvs.so with the plane-ring kernels compiled -DATVS_RING_TRACE (clock64 stamps of CTA 0 per
# role and plane, printed at kernel end).  Use: ATVS_LIB=$PWD/a-tvsnet_b200/libatvs_trace.so python tools/...
set -e
cd "$(dirname "$0")/../a-tvsnet_b200"
python _build.py > /dev/null
mkdir -p build_trace
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DATVS_RING_TRACE"
for f in conv_ring conv_deconv_ring; do nvcc $F -c csrc/$f.cu -o build_trace/$f.o & done
wait
OBJS=""
for o in build/*.o; do b=$(basename $o); if [ -f build_trace/$b ]; then OBJS="$OBJS build_trace/$b"; else OBJS="$OBJS $o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o libatvs_trace.so $OBJS
echo built libatvs_trace.so
