"""Diagnostic for the tcgen05 GEMM: structured inputs whose wrong answers localise a descriptor /
swizzle / pipeline bug.  Run on the GPU box; prints a compact report."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from agrl.pytorch_b200 import _lib
from agrl.pytorch_b200.metrics import compute_distance_matrix as cdm


def report(name, out, ref):
    err = np.abs(out - ref)
    bad = err > 1e-3 * max(1.0, np.abs(ref).max())
    print('%-28s max_err %.3e  bad %d/%d' % (name, err.max(), bad.sum(), bad.size), flush=True)
    if bad.any():
        rows, cols = np.nonzero(bad)
        print('   bad rows: n=%d first %s | bad cols: n=%d first %s' % (
            len(set(rows)), sorted(set(rows))[:12], len(set(cols)), sorted(set(cols))[:12]))
        i, j = rows[0], cols[0]
        print('   e.g. out[%d,%d]=%.6g ref=%.6g' % (i, j, out[i, j], ref[i, j]))


def dot_via_cosine_free(a, b, split):
    # euclid: (|a|^2+|b|^2) - 2ab  ->  recover ab
    out = cdm(a, b, 'euclidean', split=split).cpu().double().numpy()
    na = (a.double() ** 2).sum(1).cpu().numpy()[:, None]
    nb = (b.double() ** 2).sum(1).cpu().numpy()[None, :]
    return (na + nb - out) / 2


torch.manual_seed(0)
for split in (_lib.SPLIT_BF16X2, _lib.SPLIT_BF16X3):
    print('== split', split)
    for (m, n, k) in [(128, 128, 64), (128, 128, 128), (128, 128, 512), (256, 384, 64), (100, 200, 72), (1980, 9330, 2048)]:
        # small integers: every product exact in bf16, any error is structural
        a = torch.randint(-3, 4, (m, k)).float().cuda()
        b = torch.randint(-3, 4, (n, k)).float().cuda()
        ref = (a.double() @ b.double().t()).cpu().numpy()
        got = dot_via_cosine_free(a, b, split)
        report('int  %dx%dx%d' % (m, n, k), got, ref)
    a = torch.randn(256, 2048).cuda(); b = torch.randn(512, 2048).cuda()
    ref = (a.double() @ b.double().t()).cpu().numpy()
    got = dot_via_cosine_free(a, b, split)
    print('randn 256x512x2048: max abs err of dot %.3e (fp32 torch: %.3e)' % (
        np.abs(got - ref).max(), np.abs((a @ b.t()).double().cpu().numpy() - ref).max()))
print('launches', _lib.launch_count())
