#!/bin/bash
# CTA-pair GEMM bring-up: the GEMM-backed parity tests with the pair kernels switched on, then head variants
mkdir -p gpurun_out
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider --timeout 120 --timeout-method=thread -x"
AGRL_GEMM_PAIR=1 timeout 900 $PT tests/test_gpu_distance.py tests/test_gpu_head.py > gpurun_out/pytest_pair.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pair.log
tail -n 30 gpurun_out/pytest_pair.log
timeout 600 python tools/head_variants.py 882 "$@" > gpurun_out/variants.log 2> gpurun_out/variants.err
cat gpurun_out/variants.log; tail -n 5 gpurun_out/variants.err
