#!/bin/bash
mkdir -p gpurun_out
run() { # sub pool poolctas
  AGRL_HEAD_SUB=$1 AGRL_POOL_CTAS=$3 timeout 600 python bench.py --steps 3 --warmup 3 --pool $2 --no-e2e --no-cpu-baseline > gpurun_out/sweep_$1_$2_$3.json 2> gpurun_out/sweep_$1_$2_$3.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_$1_$2_$3.json').read().strip().splitlines()[-1])
    print('sub $1 pool $2 ctas $3 value %.0f head_ms %.2f' % (d['value'], d['head_ms']), {k:v['ms'] for k,v in d['kernels'].items() if v['share']>0.02})
except Exception as e:
    print('sub $1 pool $2 ctas $3 FAILED', e); print(open('gpurun_out/sweep_$1_$2_$3.err').read()[-800:])
PY
}
for a in "$@"; do IFS=: read s p c <<< "$a"; run $s $p ${c:-1}; done
