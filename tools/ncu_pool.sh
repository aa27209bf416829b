#!/bin/bash
# ncu --set full of the pooling kernels under a few configurations (one small pool, one head call per pass)
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  i=$((i+1))
  HV_TRACKLETS=296 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool" -s 3 -c 1 -f \
      -o gpurun_out/prof_pool_$i python tools/head_variants.py 296 "$spec" > gpurun_out/ncu_pool_$i.log 2>&1
  echo "==== $spec"; tail -n 2 gpurun_out/ncu_pool_$i.log | cut -c1-300
  python tools/ncu_summary.py gpurun_out/prof_pool_$i.ncu-rep
done
