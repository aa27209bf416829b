#!/bin/bash
# head ms per pass with the low-rank first layer (option head_lowrank), one pass and partitioned, both GEMM modes
set -u
mkdir -p gpurun_out
HV_CLOCKS=1 HV_REPS=10 timeout 100 python tools/head_variants.py 1764 \
  "split=2" "split=2,lr=1" "split=1" "split=1,lr=1" \
  "split=2,lr=1,sub=294,mode=0,psms=44,stages=6" \
  "split=1,lr=1,sub=294,mode=0,psms=64,stages=6" \
  "split=1,lr=1,sub=294,mode=0,psms=72,stages=6" \
  "split=2,lr=1,call=882" "split=1,lr=1,call=882" \
  > gpurun_out/lowrank.log 2> gpurun_out/lowrank.err
echo "rc=$?" >> gpurun_out/lowrank.err
