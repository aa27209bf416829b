"""Distance GEMM variants on one GPU: error against fp64 and time per call for each operand split.
usage: python tools/distance_variants.py [nq ng d]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from agrl.pytorch_b200 import _lib
if os.environ.get('HV_LIB'):                      # diagnostic A/B build (tools/build_alt.sh)
    _lib.LIB_PATH = os.path.abspath(os.environ['HV_LIB'])
from agrl.pytorch_b200.metrics.distance import compute_distance_matrix, PreparedOperand, distance_prepared

nq, ng, d = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1980, 9330, 4096)
_lib.require_device()
g = torch.Generator(device='cuda').manual_seed(0)
for kind in ('normal', 'clustered', 'scaled_rows'):
    q = torch.randn(nq, d, device='cuda', generator=g); x = torch.randn(ng, d, device='cuda', generator=g)
    if kind == 'clustered':
        c = torch.randn(64, d, device='cuda', generator=g)
        q = c[torch.arange(nq, device='cuda') % 64] + 0.05 * q; x = c[torch.arange(ng, device='cuda') % 64] + 0.05 * x
    if kind == 'scaled_rows':
        q = q * torch.logspace(-6, 6, nq, device='cuda')[:, None]; x = x * torch.logspace(-3, 3, ng, device='cuda')[:, None]
    for metric in ('euclidean', 'cosine'):
        q64, x64 = q.double(), x.double()
        if metric == 'euclidean':
            ref = (q64 * q64).sum(1)[:, None] + (x64 * x64).sum(1)[None] - 2 * q64 @ x64.T
            scale = (q64 * q64).sum(1)[:, None] + (x64 * x64).sum(1)[None]
        else:
            ref = 1 - torch.nn.functional.normalize(q64, dim=1) @ torch.nn.functional.normalize(x64, dim=1).T
            scale = torch.ones_like(ref)
        t32 = (q * q).sum(1)[:, None] + (x * x).sum(1)[None] - 2 * q @ x.T if metric == 'euclidean' else \
            1 - torch.nn.functional.normalize(q, dim=1) @ torch.nn.functional.normalize(x, dim=1).T
        row = {'kind': kind, 'metric': metric, 'torch_fp32_err': float(((t32.double() - ref).abs() / scale).max())}
        for split in (_lib.SPLIT_BF16X3, _lib.SPLIT_FP16X2, _lib.SPLIT_BF16X2):
            out = compute_distance_matrix(q, x, metric, split=split)
            err = float(((out.double() - ref).abs() / scale).max())
            qo, go = PreparedOperand(q, metric, split), PreparedOperand(x, metric, split)
            for _ in range(3):
                distance_prepared(qo, go, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                distance_prepared(qo, go, out=out)
            e1.record(); torch.cuda.synchronize()
            row['split%d' % split] = {'err': err, 'gemm_ms': round(e0.elapsed_time(e1) / 10, 4)}
        print(json.dumps(row), flush=True)
