#!/bin/bash
# stall-sample hot spots (SASS + source line) of one ring kernel in steady state:  bash tools/hot_ring.sh "8,8,1,0" tag
CASE=${1:-8,8,1,0}; TAG=${2:-hot}
ncu --set full --clock-control none --import-source on -k regex:'k_conv3d_ring|k_deconv3d_ring' -s 2 -c 1 -f -o /tmp/$TAG python tools/steady_probe.py $CASE > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_hot.py /tmp/$TAG.ncu-rep 45 > gpurun_out/${TAG}_hot.txt 2>&1
python tools/ncu_summary.py /tmp/$TAG.ncu-rep > gpurun_out/${TAG}_summary.txt 2>&1
