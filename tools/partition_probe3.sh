#!/bin/bash
# Third partition experiment: CTA-pair GEMMs (a third less L2 -> SM traffic) inside the spatially partitioned pipeline,
# where the L2 throughput is the contended resource; pooling CTAs launched as clusters of two (whole TPCs).
set -u
mkdir -p gpurun_out
timeout 100 python tools/head_variants.py 1764 \
  "split=2" \
  "split=2,pair=1" \
  "split=2,sub=294,mode=0,psms=44,stages=6" \
  "split=2,sub=294,mode=0,psms=36,stages=6,pair=1" \
  "split=2,sub=294,mode=0,psms=44,stages=6,pair=1" \
  "split=2,sub=294,mode=0,psms=52,stages=6,pair=1" \
  "split=2,sub=294,mode=0,psms=60,stages=6,pair=1" \
  "split=2,sub=224,mode=0,psms=44,stages=6,pair=1" \
  "split=2,sub=441,mode=0,psms=44,stages=6,pair=1" \
  "split=1,sub=294,mode=0,psms=64,stages=6" \
  "split=1,sub=294,mode=0,psms=68,stages=6" \
  "split=1,sub=441,mode=0,psms=64,stages=6" \
  "split=1,sub=588,mode=0,psms=64,stages=6" \
  > gpurun_out/partition3.log 2> gpurun_out/partition3.err
echo "rc=$?" >> gpurun_out/partition3.err
