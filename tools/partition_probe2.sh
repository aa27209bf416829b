#!/bin/bash
# Second partition experiment: (1) the wide pooling kernels ALONE (model without graph layers: the main stream only
# runs the attention kernel) at several partition widths -> per-SM ingest rate of the bulk-copy ring vs register loads;
# (2) the register-load flavour inside the full pipeline; (3) ncu --set full of one launch of each wide kernel.
# usage (on the GPU box): bash tools/partition_probe2.sh
set -u
mkdir -p gpurun_out
P=1764
timeout 120 python tools/head_variants.py $P \
  "gb=0" \
  "gb=0,sub=126,mode=0,psms=16,stages=6" \
  "gb=0,sub=126,mode=0,psms=32,stages=6" \
  "gb=0,sub=126,mode=0,psms=44,stages=6" \
  "gb=0,sub=126,mode=0,psms=64,stages=6" \
  "gb=0,sub=126,mode=0,psms=96,stages=6" \
  "gb=0,sub=126,mode=0,psms=148,stages=6" \
  "gb=0,sub=126,mode=0,psms=44,stages=3" \
  "gb=0,sub=126,mode=0,psms=16,wldg=1" \
  "gb=0,sub=126,mode=0,psms=32,wldg=1" \
  "gb=0,sub=126,mode=0,psms=44,wldg=1" \
  "gb=0,sub=126,mode=0,psms=64,wldg=1" \
  "gb=0,sub=126,mode=0,psms=96,wldg=1" \
  "gb=0,sub=126,mode=0,psms=148,wldg=1" \
  "split=2" \
  "split=2,sub=294,mode=0,psms=36,wldg=1" \
  "split=2,sub=294,mode=0,psms=44,wldg=1" \
  "split=2,sub=294,mode=0,psms=52,wldg=1" \
  "split=2,sub=294,mode=0,psms=60,wldg=1" \
  "split=1,sub=294,mode=0,psms=52,wldg=1" \
  "split=1,sub=294,mode=0,psms=64,wldg=1" \
  "split=1,sub=294,mode=0,psms=76,wldg=1" \
  "split=1,sub=294,mode=0,psms=76,stages=6" \
  > gpurun_out/partition2.log 2> gpurun_out/partition2.err
echo "rc=$?" >> gpurun_out/partition2.err
for K in pool_tma_wide pool_ldg_wide; do
  W=0; [ $K = pool_ldg_wide ] && W=1
  HV_TRACKLETS=504 timeout 70 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 2 -c 1 -f \
    -o gpurun_out/prof_$K python tools/head_variants.py 504 "gb=0,sub=126,mode=0,psms=44,stages=6,wldg=$W" \
    > gpurun_out/ncu_$K.log 2>&1
  echo "rc=$?" >> gpurun_out/ncu_$K.log
done
