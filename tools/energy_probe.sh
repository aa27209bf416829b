#!/bin/bash
# NVML energy counter (joules per 11310-tracklet pass, average watts) and nvidia-smi clocks / throttle reasons for the head's
# configurations: pooling alone, one layer, operand modes, low-rank on / off, CTA pairs on / off.
# usage: gpurun -- 'bash tools/energy_probe.sh'   -> gpurun_out/energy_probe.log (committed as profiles/r2/energy_probe*.log)
mkdir -p gpurun_out
HV_ENERGY=1.5 HV_REPS=5 HV_CLOCKS=1 timeout 400 python tools/head_variants.py 882 \
  "gb=0" "gb=1" "split=2,lr=0" "split=2" "split=4,pair=0" "split=4" "split=4,lr=0" "split=1" "split=3" \
  > gpurun_out/energy_probe.log 2> gpurun_out/energy_probe.err
cat gpurun_out/energy_probe.log; tail -n 3 gpurun_out/energy_probe.err
