#!/bin/bash
# run a set of GPU test files: tools/gpu_tests.sh tests/test_gpu_pose.py ...
mkdir -p gpurun_out
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider --timeout 180 --timeout-method=thread -x"
timeout 900 $PT "$@" > gpurun_out/pytest_sel.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_sel.log
tail -n 40 gpurun_out/pytest_sel.log
