#!/bin/bash
# Spatial-partition experiment (options pool_sms / gemm_sms): head ms per 11310 tracklets on a
# 1764-tracklet resident pool, one agrl_head_forward_dev call per 1764 tracklets (6 sub-batches of 294 etc.).
# usage (on the GPU box): bash tools/partition_probe.sh > gpurun_out/partition.log
set -u
P=1764
python tools/head_variants.py $P \
  "split=2" \
  "split=2,sub=294,mode=0,stages=4" \
  "split=2,sub=294,mode=0,psms=36,stages=6" \
  "split=2,sub=294,mode=0,psms=44,stages=6" \
  "split=2,sub=294,mode=0,psms=52,stages=6" \
  "split=2,sub=294,mode=0,psms=60,stages=6" \
  "split=2,sub=224,mode=0,psms=36,stages=6" \
  "split=2,sub=224,mode=0,psms=44,stages=6" \
  "split=2,sub=208,mode=0,psms=44,stages=6" \
  "split=2,sub=192,mode=0,psms=52,stages=6" \
  "split=2,sub=441,mode=0,psms=44,stages=6" \
  "split=2,sub=294,mode=0,psms=44,gsms=148,stages=6" \
  "split=2,sub=294,mode=0,psms=44,stages=6,hint=0" \
  "split=2,sub=294,mode=0,psms=44,stages=5" \
  "split=1" \
  "split=1,sub=294,mode=0,psms=52,stages=6" \
  "split=1,sub=294,mode=0,psms=64,stages=6" \
  "split=1,sub=224,mode=0,psms=60,stages=6" \
  "split=2,call=882,sub=147,mode=0,psms=44,stages=6" \
  "split=2,call=882,sub=224,mode=0,psms=44,stages=6"
