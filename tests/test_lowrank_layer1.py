"""Groundwork for DESIGN.md section 10 item 1, pinned on the CPU with the oracle: the pooled nodes of a frame are
linear combinations of its four quarter strips (vmgn.py:249-251,304-308: the half and whole strips are means of two
and four quarter strips), so X = T.Q with a fixed (7S x 4S) matrix T, and the first graph layer's product
G.X.W^T (vmgn.py:148,168) equals (G.T).(Q.W^T) -- 32 GEMM rows per tracklet instead of 56.  This test states the
identity and measures its rounding against the reference goldens, so that a kernel built on it has a checked spec."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_files
from oracle import head as ohead
from test_oracle_head import regenerate

S, P, Q = 8, 7, 4


def strip_matrix(S=S):
    """T (7S x 4S): node s*7+p from the quarter strips s*4+k of the same frame"""
    per_frame = torch.zeros(P, Q, dtype=torch.float64)
    for k in range(4):
        per_frame[k, k] = 1.0                       # q0..q3
    per_frame[4, 0:2] = 0.5                         # h0 = (q0 + q1) / 2
    per_frame[5, 2:4] = 0.5                         # h1 = (q2 + q3) / 2
    per_frame[6, :] = 0.25                          # whole = mean of the four
    return torch.block_diag(*[per_frame] * S)


def layer1_lowrank(x, adj, w, prefix, dtype):
    """graph_layer (oracle/head.py) with the GEMM on the quarter rows only"""
    B, V, C = x.shape
    T = strip_matrix().to(dtype)
    q = x.reshape(B, S, P, C)[:, :, :Q].reshape(B, S * Q, C)              # the quarter rows ARE rows of X
    z = q @ w[prefix + '.linear.weight'].to(dtype).t()                    # (B, 32, C): the only big product
    g = (ohead._l1_rows(adj.to(dtype)) + ohead._l1_rows(ohead.affinity(x))) / 2
    hp = torch.bmm(torch.matmul(g, T), z)                                 # (G.T).(Q.W^T)
    hp = ohead._bn_eval(hp.reshape(B * V, C), {k: v.to(dtype) for k, v in w.items()}, prefix + '.bn').reshape(B, V, C)
    hp = torch.where(hp >= 0, hp, hp * ohead.LEAKY)
    return (1 - ohead.GAMMA) * x + ohead.GAMMA * hp


def test_strip_matrix_shape_and_rows():
    T = strip_matrix()
    assert tuple(T.shape) == (56, 32) and torch.allclose(T.sum(1), torch.ones(56, dtype=torch.float64))
    assert int(torch.linalg.matrix_rank(T)) == 32


@pytest.mark.parametrize('fname', golden_files('head_'))
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_layer1_on_quarter_rows_matches_the_reference(fname, dtype):
    g = np.load(os.path.join(GOLDEN, fname))
    x1, x2, adj, wts = regenerate(g)
    B = int(g['B'])
    x = ohead.pool_nodes(x2.to(dtype), B, S, [4, 2, 1])
    # the pooled nodes themselves: X = T.Q up to the rounding of a 64- / 128-element mean vs a mean of quarter means
    q = x.reshape(B, S, P, -1)[:, :, :Q].reshape(B, S * Q, -1)
    recon = torch.matmul(strip_matrix().to(dtype), q)
    assert float((recon - x).abs().max() / x.abs().max()) < (3e-7 if dtype == torch.float32 else 1e-15)
    # first layer
    w = {k: v.to(dtype) for k, v in wts.items()}
    full = ohead.graph_layer(x, adj.to(dtype), w, 'graph_layers.0')
    low = layer1_lowrank(x, adj, w, 'graph_layers.0', dtype)
    tol = 1e-6 if dtype == torch.float32 else 1e-13
    assert float((low - full).abs().max() / full.abs().max()) < tol
    assert float((low - full).norm() / full.norm()) < tol
    # whole head with that first layer: still on the reference golden within the oracle's own 2e-6
    f = ohead.graph_layer(low, adj.to(dtype), w, 'graph_layers.1')
    fused = ohead.attention_fuse(f.reshape(B, S, P, -1))
    att = ohead._bn_eval(fused.mean(dim=1), w, 'att_bottleneck').float()
    ref = torch.from_numpy(g['out'])[:, 2048:]
    assert float((att - ref).abs().max() / ref.abs().max()) < 2e-6
    assert float((att - ref).norm() / ref.norm()) < 2e-6
