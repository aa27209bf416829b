#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider --timeout 180 --timeout-method=thread"
timeout 900 $PT tests -s 2>&1 | grep -E "clustered:|passed|failed|FAILED|Error" > gpurun_out/pytest_gpu.log; echo "rc=$?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for v in 0 1 2 3; do
AGRL_GRAPH_VARIANT=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/variant$v.json 2> gpurun_out/variant$v.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/variant$v.json').read().strip().splitlines()[-1])
    print('variant $v value %.0f head_ms %.2f eval_ms %.3f' % (d['value'], d['head_ms'], d['eval_ms']), {k:v['ms'] for k,v in d['kernels'].items() if v['share']>0.002})
except Exception as e:
    print('variant $v FAILED', e); print(open('gpurun_out/variant$v.err').read()[-600:])
PY
done
