#!/bin/bash
# which pipeline role bounds the big convolution kernels: conv_probe with the ring debug knobs
# (1 no loads, 2 no MMAs, 4 no stores, 8 no accumulator zeroing) and dispatch knobs
out=${1:-gpurun_out/probe_sweep.log}
: > $out
run() { echo "## $*" >> $out; env "${@:1:$#-1}" python tools/conv_probe.py ${@: -1} >> $out 2>&1; }
for cfg in "32 8" "8 8" "8 16" "16 16 1 0 64 64 80"; do
  for dbg in 0 1 2 4 5 6 7 15; do run ATVS_RING_DEBUG=$dbg "$cfg"; done
  run ATVS_RING_MINB=1 "$cfg"
  for zs in 8 16 32 64; do run ATVS_RING_ZS=$zs "$cfg"; done
  for r in 3 4 6; do run ATVS_RING_R=$r "$cfg"; done
done
run X=1 "8 16 2"
run ATVS_NO_RING_S2=1 "8 16 2"
run X=1 "32 16 2"
run X=1 "16 8 2 1 64 64 80"
run X=1 "32 16 2 1 32 32 40"
run X=1 "16 32 2 0 64 64 80"
