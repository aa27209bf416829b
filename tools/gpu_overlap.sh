#!/bin/bash
# GPU visit for the pooling-overlap work: head parity tests, then pipeline variants on one resident pool.
mkdir -p gpurun_out
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider --timeout 180 --timeout-method=thread -x"
timeout 600 $PT tests/test_gpu_head.py > gpurun_out/pytest_head.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_head.log
tail -n 15 gpurun_out/pytest_head.log
timeout 900 python tools/head_variants.py 882 "$@" > gpurun_out/variants.log 2> gpurun_out/variants.err
cat gpurun_out/variants.log; tail -n 5 gpurun_out/variants.err
