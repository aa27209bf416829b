"""bench.py's same-box comparator (stock torch.nn modules laid out as vmgn.py:296-321 / distance.py:59-73) must
compute what the oracle computes, otherwise the reported PyTorch-on-B200 time belongs to a different function."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench
from oracle import distance as odist
from oracle import head as ohead
from agrl.pytorch_b200 import synthetic as synth


def test_comparator_matches_the_oracle():
    n = 2
    w = bench.make_head_weights(seed=5)
    g = torch.Generator().manual_seed(11)
    x1 = torch.randn(n * bench.S, bench.C, bench.H, bench.W, generator=g).clamp_(min=0)
    x2 = torch.randn(n * bench.S, bench.C, bench.H, bench.W, generator=g).clamp_(min=0)
    adj = synth.pose_adjacency(n, bench.S, 7, seed=11)
    head, distance = bench.build_eager(torch.device('cpu'), w)
    with torch.no_grad():
        got = head(x1, x2, adj)
        ref = ohead.head_forward(x1, x2, adj, w, dtype=torch.float64).float()
    assert got.shape == (n, 2 * bench.C)
    assert (got - ref).norm() / ref.norm() < 2e-6
    q, gal = torch.randn(5, 64, generator=g), torch.randn(7, 64, generator=g)
    d = distance(q, gal)
    assert torch.allclose(d, odist.distance_matrix(q, gal, 'euclidean'), rtol=1e-5, atol=1e-4)
