"""rank_mars_merge in isolation (no collectives ahead of it on the stream): ms per launch for (queries, shards).
HV_LIB=path selects a diagnostic build (tools/build_alt.sh), e.g. -DAGRL_MERGE_SORT for the bitonic-sort merge."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from agrl.pytorch_b200 import _lib
if os.environ.get('HV_LIB'):
    _lib.LIB_PATH = os.path.abspath(os.environ['HV_LIB'])
from agrl.pytorch_b200 import sharded
ops = sharded.CudaOps()
g = torch.Generator(device='cuda').manual_seed(0)
for nq, parts in ((1980, 8), (10000, 8), (10000, 2)):
    K = 50
    d = torch.rand(parts, nq, K, generator=g, device='cuda')
    idx = torch.arange(parts * nq * K, device='cuda').view(parts, nq, K) % 1000000
    keys = ((d.sort(dim=2).values.view(torch.int32).to(torch.int64) | 0x80000000) << 32) | idx
    keys = keys.sort(dim=2).values.contiguous()
    cls = (torch.rand(parts, nq, K, generator=g, device='cuda') < 0.05).to(torch.uint8)
    ngood = torch.full((nq,), 5, dtype=torch.int32, device='cuda')
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    for _ in range(3): r = ops.merge(keys, cls, ngood, K, st)
    with _lib.profile(torch.cuda.current_stream().cuda_stream) as p:
        for _ in range(5): r = ops.merge(keys, cls, ngood, K, st)
    t = p.totals()
    print(os.environ.get('HV_LIB', 'product'), nq, parts, {k: round(v[1] / v[0], 4) for k, v in t.items()}, float(r[1]))
