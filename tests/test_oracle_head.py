"""Pin oracle/head.py to outputs of the reference's GSTA.forward (vmgn.py:296-321) run on the same
seeded maps (tests/golden/head_*.npz; inputs are regenerated from oracle.synth, guarded by a checksum)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_files
from oracle import head as ohead
from oracle import synth


def regenerate(g):
    B, w = int(g['B']), int(g['w'])
    x1, x2 = synth.feature_maps(B, 8, 2048, 16, w, seed=int(g['maps_seed']), scale=float(g['scale']))
    adj = synth.pose_adjacency(B, 8, 7, seed=int(g['maps_seed']), mode=str(g['adj_mode']))
    wts = synth.head_weights(2048, 2, seed=int(g['weights_seed']), randomise_bn=bool(g['randomise_bn']))
    chk = float(x1.double().sum() + 3 * x2.double().sum() + 7 * adj.double().sum()
                + sum(v.double().sum() for v in wts.values()))
    # summation order of .sum() may differ between CPUs: compare to 1e-12, not bit for bit
    assert abs(chk - float(g['checksum'])) <= 1e-12 * abs(chk), 'torch RNG stream drifted: regenerate tests/golden'
    return x1, x2, adj, wts


@pytest.mark.parametrize('fname', golden_files('head_'))
def test_head_oracle_matches_reference_golden(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    x1, x2, adj, wts = regenerate(g)
    ref = torch.from_numpy(g['out'])
    for dtype, tol in ((torch.float32, 2e-6), (torch.float64, 2e-6)):
        got = ohead.head_forward(x1, x2, adj, wts, dtype=dtype).float()
        assert got.shape == ref.shape
        err_max = (got - ref).abs().max() / ref.abs().max()
        err_nrm = (got - ref).norm() / ref.norm()
        assert err_max < tol and err_nrm < tol, (fname, dtype, float(err_max), float(err_nrm))


def test_split_list_is_calc_splits():
    assert ohead.split_list(4, True) == [4, 2, 1]
    assert ohead.split_list(8, True) == [8, 4, 2, 1]
    assert ohead.split_list(4, False) == [4]
