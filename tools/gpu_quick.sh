#!/bin/bash
# quick GPU visit: parity tests + short bench variants (no captures).  usage: gpu_quick.sh [pytest-target] [pools...]
mkdir -p gpurun_out
TARGET=${1:-tests}; shift
POOLS=${@:-252}
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider --timeout 180 --timeout-method=thread"
timeout 900 $PT $TARGET > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
for pool in $POOLS; do
timeout 600 python bench.py --steps 3 --warmup 3 --pool $pool --no-e2e --no-cpu-baseline > gpurun_out/bench_pool$pool.json 2> gpurun_out/bench_pool$pool.err; echo "rc=$?" >> gpurun_out/bench_pool$pool.err
done
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench_pool*.err
