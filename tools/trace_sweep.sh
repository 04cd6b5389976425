#!/bin/bash
# per-role timelines (clock64 stamps of CTA 0) of the ring kernels: needs libatvs_trace.so (-DATVS_RING_TRACE)
export ATVS_LIB=$PWD/a-tvsnet_b200/libatvs_trace.so
for cfg in "8 8" "32 8" "8 16"; do
  for dbg in 0 7; do
    f=gpurun_out/trace_$(echo $cfg | tr ' ' '_')_dbg$dbg.txt
    ATVS_RING_DEBUG=$dbg python tools/conv_probe.py $cfg 1 0 128 128 160 1 > $f 2>&1
    python tools/ring_trace.py $f
  done
done
