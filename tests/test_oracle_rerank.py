"""Pin oracle/rerank.py to the reference's own re_ranking() (utils/re_ranking.py:30-94): tests/golden/rerank_*.npz
hold the three distance matrices (from the reference's compute_distance_matrix) and the re-ranked distances."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_files
from oracle import rerank as orr


@pytest.mark.parametrize('fname', golden_files('rerank_'))
def test_rerank_oracle_matches_reference_golden(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    out = orr.re_ranking(g['q_g'], g['q_q'], g['g_g'], k1=int(g['k1']), k2=int(g['k2']), lambda_value=float(g['lam']))
    assert out.shape == g['out'].shape and out.dtype == np.float32
    assert np.array_equal(out, g['out']), (fname, float(np.abs(out - g['out']).max()))


def test_normalised_distance_definition():
    rng = np.random.RandomState(0)
    qg, qq, gg = rng.rand(3, 5).astype(np.float32), rng.rand(3, 3).astype(np.float32), rng.rand(5, 5).astype(np.float32)
    D = orr.normalised_dist(qg, qq, gg)
    orig = np.block([[qq, qg], [qg.T, gg]]).astype(np.float32) ** 2
    assert D.shape == (8, 8) and D.dtype == np.float32
    for i in range(8):
        assert np.array_equal(D[i], orig[:, i] / orig[:, i].max())
