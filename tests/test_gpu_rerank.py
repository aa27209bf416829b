"""GPU parity of the k-reciprocal re-ranking (csrc/rerank.cu): the reference's own outputs (tests/golden/rerank_*.npz)
and the oracle on larger problems.  Index work is exact; values differ only through expf vs numpy's float32 exp,
so the bar is 1e-5 of the output range (the outputs live in [0, 1])."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_files
from oracle import distance as odist
from oracle import rerank as orr
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize('fname', golden_files('rerank_'))
def test_rerank_reference_golden(fname):
    from agrl.pytorch_b200.utils import re_ranking
    g = np.load(os.path.join(GOLDEN, fname))
    out = re_ranking(g['q_g'], torch.from_numpy(g['q_q']), g['g_g'], k1=int(g['k1']), k2=int(g['k2']),
                     lambda_value=float(g['lam']))
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == g['out'].shape
    assert float(np.abs(out - g['out']).max()) < TOL, fname


@pytest.mark.parametrize('nq,ng,metric', [(150, 150, 'euclidean'), (89, 89, 'cosine'), (123, 1000, 'euclidean')])
def test_rerank_vs_oracle_on_the_libraries_own_distances(nq, ng, metric):
    """the call sequence of test() with --re-rank: three distance matrices from compute_distance_matrix, re-ranking,
    evaluate_rank on the result"""
    from agrl.pytorch_b200 import metrics
    from agrl.pytorch_b200.utils.re_ranking import re_ranking_dev
    qp, qc, gp, gc = synth.eval_labels((nq, ng, 20, 4), seed=nq)
    qf, gf = synth.eval_features(qp, gp, 256, seed=nq + 1, clustered=True)
    qg = metrics.compute_distance_matrix(qf.cuda(), gf.cuda(), metric)
    qq = metrics.compute_distance_matrix(qf.cuda(), qf.cuda(), metric)
    gg = metrics.compute_distance_matrix(gf.cuda(), gf.cuda(), metric)
    out = re_ranking_dev(qg, qq, gg)
    want = orr.re_ranking(qg.cpu().numpy(), qq.cpu().numpy(), gg.cpu().numpy())
    assert float(np.abs(out.cpu().numpy() - want).max()) < TOL
    cmc, mAP = metrics.evaluate_rank(out, qp, gp, qc, gc, use_metric_mars=True)
    cmc0, mAP0 = metrics.evaluate_rank(qg, qp, gp, qc, gc, use_metric_mars=True)
    assert mAP >= mAP0 - 0.02                                   # re-ranking helps (or at least does not hurt) on clustered data


def test_rerank_strided_inputs_and_limits():
    from agrl.pytorch_b200 import _lib
    from agrl.pytorch_b200.utils.re_ranking import re_ranking_dev
    g = np.load(os.path.join(GOLDEN, 'rerank_clustered.npz'))
    big = torch.zeros(40, 300, device='cuda')
    big[:, :260] = torch.from_numpy(g['q_g']).cuda()
    out = re_ranking_dev(big[:, :260], torch.from_numpy(g['q_q']).cuda(), torch.from_numpy(g['g_g']).cuda())
    assert float(np.abs(out.cpu().numpy() - g['out']).max()) < TOL
    with pytest.raises(_lib.AgrlError):
        re_ranking_dev(g['q_g'], g['q_q'], g['g_g'], k1=64)
