#!/bin/bash
# ncu --set full of one head kernel: tools/ncu_kernel.sh <kernel-regex> <skip> <tag> "<variant spec>"
mkdir -p gpurun_out
HV_TRACKLETS=296 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c 1 -f \
    -o gpurun_out/prof_$3 python tools/head_variants.py 296 "$4" > gpurun_out/ncu_$3.log 2>&1
tail -n 2 gpurun_out/ncu_$3.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/prof_$3.ncu-rep
python tools/ncu_hot.py gpurun_out/prof_$3.ncu-rep 0.015
