#!/bin/bash
# SM clock / board power while the head runs in its different modes (is the overlap bounded by the power cap?):
# pooling alone, one pass, spatially partitioned, fast mode one pass / partitioned.  ~2 s of passes per mode.
set -u
mkdir -p gpurun_out
HV_CLOCKS=1 HV_REPS=30 timeout 150 python tools/head_variants.py 1764 \
  "gb=0" \
  "gb=0,sub=126,mode=0,psms=44,stages=6" \
  "split=2" \
  "split=2,sub=294,mode=0,psms=44,stages=6" \
  "split=2,sub=294,mode=0,psms=48,stages=6" \
  "split=1" \
  "split=1,sub=294,mode=0,psms=64,stages=6" \
  > gpurun_out/clock_probe.log 2> gpurun_out/clock_probe.err
echo "rc=$?" >> gpurun_out/clock_probe.err
