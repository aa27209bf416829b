for i in 1 2; do
  for lib in "" tools/_alt/libagrl_b200_olddrain.so; do
    HV_LIB=$lib HV_REPS=5 timeout 100 python tools/head_variants.py 882 "split=4" "split=4,lr=0" 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$lib', d['spec'], d['head_ms'], d['checksum'], d['kernels'])"
  done
done
