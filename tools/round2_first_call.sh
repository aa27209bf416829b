#!/bin/bash
# First GPU call of the next round (one B200, ~4 min): validates what round 1 left unmeasured.
#   1. the whole GPU suite with the low-rank first layer as the DEFAULT (AGRL_HEAD_LOWRANK=1 only changes the option's
#      default) -> if green, flip the default in csrc/api.cu
#   2. the experimental flavours under a timeout (a deadlock must not take the box): graph_mix2_kernel (head_lowrank = 2),
#      CTA pairs with direct signalling (gemm_pair = 2)
#   3. timings: one pass / low-rank / low-rank + mix2 / pairs (relay, direct), both GEMM modes
# usage: gpurun --timeout 420 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
(AGRL_HEAD_LOWRANK=1 timeout 150 python -m pytest tests -m gpu -q > gpurun_out/pytest_lowrank_default.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lowrank_default.log)
(AGRL_EXPERIMENTAL=mix2 timeout 90 python -m pytest tests/test_gpu_head.py -q -k "lowrank" > gpurun_out/pytest_experimental.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_experimental.log)
HV_CLOCKS=1 HV_REPS=10 timeout 120 python tools/head_variants.py 1764 \
  "split=2" "split=2,lr=1" "split=2,lr=2" "split=2,pair=1" "split=2,lr=1,call=882" "split=2,lr=2,call=882" \
  "split=1" "split=1,lr=1" "split=1,lr=2" "split=1,pair=1" \
  > gpurun_out/round2_first.log 2> gpurun_out/round2_first.err
echo "rc=$?" >> gpurun_out/round2_first.err
# the direct-signalling pairs last and alone: if they deadlock, everything above is already on disk
HV_REPS=5 timeout 40 python tools/head_variants.py 882 "split=2,pair=2" "split=1,pair=2" \
  > gpurun_out/round2_pair_direct.log 2> gpurun_out/round2_pair_direct.err
echo "rc=$?" >> gpurun_out/round2_pair_direct.err
(AGRL_EXPERIMENTAL=pair2 timeout 60 python -m pytest tests/test_gpu_head.py -q -k "spatially" > gpurun_out/pytest_pair_direct.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pair_direct.log)
tail -n 5 gpurun_out/pytest_lowrank_default.log gpurun_out/pytest_experimental.log gpurun_out/round2_first.err gpurun_out/round2_pair_direct.err gpurun_out/pytest_pair_direct.log
cat gpurun_out/round2_first.log gpurun_out/round2_pair_direct.log
