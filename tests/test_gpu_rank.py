"""GPU parity of the ranking kernels (csrc/rank.cu) through the reference-facing API
(agrl.pytorch_b200.metrics.evaluate_rank -> C ABI): bit-exact vs the reference's golden vectors and
vs the pinned oracle on larger seeded inputs, incl. the reference's edge cases."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_files
from oracle import rank as orank
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def metrics():
    from agrl.pytorch_b200 import metrics as m
    return m


def _bits32(x):
    return np.asarray(x, np.float32).view(np.uint32)


def _bits64(x):
    return np.asarray(x, np.float64).view(np.uint64)


@pytest.mark.parametrize('device_input', [False, True])
@pytest.mark.parametrize('fname', golden_files('rank_'))
def test_market1501_golden(metrics, fname, device_input):
    g = dict(np.load(os.path.join(GOLDEN, fname)))
    d = torch.from_numpy(g['distmat']).cuda() if device_input else g['distmat']
    args = (d, g['q_pids'], g['g_pids'], g['q_camids'], g['g_camids'])
    kw = dict(max_rank=int(g['max_rank']), use_metric_market1501=True)
    if 'cy_error' in g:
        with pytest.raises(AssertionError, match='all query identities do not appear in gallery'):
            metrics.evaluate_rank(*args, **kw)
        return
    cmc, mAP = metrics.evaluate_rank(*args, **kw)
    assert cmc.dtype == np.float32 and isinstance(mAP, float)
    assert cmc.shape == g['cy_cmc'].shape
    assert np.array_equal(_bits32(cmc), _bits32(g['cy_cmc']))
    assert mAP == float(g['cy_mAP_f64'])


@pytest.mark.parametrize('device_input', [False, True])
@pytest.mark.parametrize('fname', golden_files('rank_'))
def test_mars_golden(metrics, fname, device_input):
    g = dict(np.load(os.path.join(GOLDEN, fname)))
    d = torch.from_numpy(g['distmat']).cuda() if device_input else g['distmat']
    args = (d, g['q_pids'], g['g_pids'], g['q_camids'], g['g_camids'])
    kw = dict(max_rank=int(g['max_rank']), use_metric_mars=True)
    if 'mars_error' in g:
        exc = ZeroDivisionError if int(g['mars_error']) == 2 else ValueError
        with pytest.raises(exc):
            metrics.evaluate_rank(*args, **kw)
        return
    cmc, mAP = metrics.evaluate_rank(*args, **kw)
    assert cmc.dtype == np.float64 and isinstance(mAP, np.float64)
    assert np.array_equal(_bits64(cmc), _bits64(g['mars_cmc']))
    assert _bits64(mAP) == _bits64(g['mars_mAP'])


CASES = [
    # shape, seed, kind of distance matrix, max_rank
    ('ilidsvid', 0, 'randn', 50),
    ('prid2011', 1, 'ties', 50),
    ('dukev', 2, 'randn', 50),
    ('dukev', 3, 'ties', 20),
    ('mars', 4, 'randn', 50),
    ('mars', 5, 'ties', 50),
    ((257, 5000, 9, 3), 6, 'randn', 50),      # ~550 items per identity: shared-histogram path
    ((64, 3000, 1, 2), 7, 'ties', 50),        # one identity, 3000 items: brute-force overflow path
    ((33, 70, 5, 2), 8, 'randn', 70),         # max_rank == num_g
    ((100, 2049, 40, 4), 9, 'randn', 100),    # odd row stride: unaligned rows
]


def _distmat(kind, nq, ng, seed):
    if kind == 'ties':
        return synth.quantised_distmat(nq, ng, seed=seed)
    return np.random.RandomState(seed).randn(nq, ng).astype(np.float32)


@pytest.mark.parametrize('shape,seed,kind,max_rank', CASES)
def test_market1501_vs_oracle(metrics, shape, seed, kind, max_rank):
    qp, qc, gp, gc = synth.eval_labels(shape, seed=seed)
    d = _distmat(kind, len(qp), len(gp), seed)
    ref_cmc, ref_map, ref_ap, ref_nv = orank.market1501_port(d, qp, gp, qc, gc, max_rank, return_ap=True)
    from agrl.pytorch_b200.metrics.rank_cylib.rank_cy import evaluate_cy
    for dist in (d, torch.from_numpy(d).cuda()):
        cmc, mAP, ap, nv = evaluate_cy(dist, qp, gp, qc, gc, max_rank, return_all_ap=True)
        assert nv == ref_nv
        assert np.array_equal(_bits32(ap), _bits32(ref_ap))
        assert np.array_equal(_bits32(cmc), _bits32(ref_cmc))
        assert mAP == ref_map


@pytest.mark.parametrize('shape,seed,kind,max_rank', CASES)
def test_mars_vs_oracle(metrics, shape, seed, kind, max_rank):
    qp, qc, gp, gc = synth.eval_labels(shape, seed=seed)
    d = _distmat(kind, len(qp), len(gp), seed)
    ref_cmc, ref_map, ref_ap = orank.mars_port(d, qp, gp, qc, gc, max_rank, return_ap=True)
    from agrl.pytorch_b200.metrics.rank import evaluate_mars
    for dist in (d, torch.from_numpy(d).cuda()):
        cmc, mAP, ap = evaluate_mars(dist, qp, gp, qc, gc, max_rank, return_all_ap=True)
        assert np.array_equal(_bits64(ap), _bits64(ref_ap))
        assert np.array_equal(_bits64(cmc), _bits64(ref_cmc))
        assert _bits64(mAP) == _bits64(ref_map)


def test_clustered_features_end_to_end(metrics):
    """distance (GPU) -> ranking (GPU) on clustered features: CMC/mAP equal to the oracle chain."""
    from oracle import distance as odist
    qp, qc, gp, gc = synth.eval_labels('dukev', seed=11)
    qf, gf = synth.eval_features(qp, gp, 512, seed=11, clustered=True)
    d_ref = odist.distance_matrix(qf, gf, 'cosine').numpy()
    ref = orank.mars_port(d_ref, qp, gp, qc, gc, 50)
    d_gpu = metrics.compute_distance_matrix(qf.cuda(), gf.cuda(), 'cosine')
    got = metrics.evaluate_rank(d_gpu, qp, gp, qc, gc, use_metric_mars=True)
    assert np.array_equal(got[0], ref[0]) and got[1] == ref[1]
    assert got[1] > 0.5            # non-trivial retrieval


def test_evaluate_rank_dispatch(metrics):
    d = np.random.rand(4, 8).astype(np.float32)
    ids = np.arange(4), np.arange(8) % 4, np.zeros(4, int), np.ones(8, int)
    assert metrics.evaluate_rank(d, *ids) is None                       # no metric flag (rank.py:232-238)
    with pytest.raises(NotImplementedError):
        metrics.evaluate_rank(d, *ids, use_metric_cuhk03=True)


@pytest.mark.parametrize('shape,seed,kind,parts', [('dukev', 21, 'randn', 3), ((100, 2049, 40, 4), 22, 'ties', 2),
                                                   ('mars', 23, 'randn', 8), ((33, 70, 5, 2), 24, 'ties', 4)])
def test_sharded_partial_and_merge_kernels(shape, seed, kind, parts):
    """the gallery-sharded MARS kernels (partial per shard + merge), single process: shards are
    column slices of one distance matrix; result must be bit-identical to the unsharded oracle"""
    from agrl.pytorch_b200 import sharded
    qp, qc, gp, gc = synth.eval_labels(shape, seed=seed)
    d = _distmat(kind, len(qp), len(gp), seed)
    K = 50
    ref_cmc, ref_map = orank.mars_port(d, qp, gp, qc, gc, K)
    ops = sharded.CudaOps()
    dev = torch.device('cuda')
    dd = torch.from_numpy(d).to(dev)
    tq, tqc = torch.as_tensor(qp).to(dev), torch.as_tensor(qc).to(dev)
    keys, cls, ngood = [], [], 0
    for lo, hi in sharded.shard_bounds(len(gp), parts):
        k, c, n, st = ops.partial(dd[:, lo:hi], tq, torch.as_tensor(gp[lo:hi]).to(dev), tqc,
                                  torch.as_tensor(gc[lo:hi]).to(dev), K, lo)
        keys.append(k); cls.append(c); ngood = ngood + n
    cmc, mAP = ops.merge(torch.stack(keys), torch.stack(cls), ngood, K, st)
    assert np.array_equal(_bits64(cmc), _bits64(ref_cmc))
    assert _bits64(mAP) == _bits64(ref_map)
    # lists that are NOT ascending (another producer): the merge notices and sorts instead of rank-merging
    perm = torch.randperm(K, generator=torch.Generator().manual_seed(seed)).to(dev)
    cmc2, mAP2 = ops.merge(torch.stack(keys)[:, :, perm].contiguous(), torch.stack(cls)[:, :, perm].contiguous(), ngood, K, st)
    assert np.array_equal(_bits64(cmc2), _bits64(ref_cmc)) and _bits64(mAP2) == _bits64(ref_map)


@pytest.mark.parametrize('shape,seed,kind,parts', [('dukev', 31, 'randn', 3), ((100, 2049, 40, 4), 32, 'ties', 2),
                                                   ('mars', 33, 'randn', 8), ((25, 12, 4, 2), 34, 'randn', 2),
                                                   ((257, 5000, 9, 3), 35, 'ties', 4)])
def test_sharded_market1501_kernels(shape, seed, kind, parts):
    """count / gather / bin / finalize kernels of the gallery-sharded market1501 metric, single process:
    shards are column slices of one distance matrix, the 'collectives' are plain sums / stacks"""
    from agrl.pytorch_b200 import sharded
    qp, qc, gp, gc = synth.eval_labels(shape, seed=seed)
    d = _distmat(kind, len(qp), len(gp), seed)
    K = 50
    ref_cmc, ref_map = orank.market1501_port(d, qp, gp, qc, gc, K)
    ops = sharded.CudaOps()
    dev = torch.device('cuda')
    dd = torch.from_numpy(d).to(dev)
    tq, tqc = torch.as_tensor(qp).to(dev), torch.as_tensor(qc).to(dev)
    bounds = sharded.shard_bounds(len(gp), parts)
    lab = [(torch.as_tensor(gp[lo:hi]).to(dev), torch.as_tensor(gc[lo:hi]).to(dev)) for lo, hi in bounds]
    cap = 0
    for (lo, hi), (tg, tgc) in zip(bounds, lab):
        if hi > lo:
            cap = max(cap, int(ops.market_count(tq, tg, tqc, tgc)[0].cpu()))
    cap = max(8, (cap + 7) // 8 * 8)
    keys, counts = [], 0
    for (lo, hi), (tg, tgc) in zip(bounds, lab):
        k, c, st = ops.market_gather(dd[:, lo:hi].contiguous(), tq, tg, tqc, tgc, lo, cap)
        keys.append(k); counts = counts + c
    keys_all = torch.stack(keys)
    cnt = 0
    for (lo, hi), (tg, tgc) in zip(bounds, lab):
        c, srt = ops.market_bin(dd[:, lo:hi].contiguous(), lo, keys_all)
        cnt = cnt + c
    cmc, mAP = ops.market_finalize(cnt, srt, counts, len(gp), parts, cap, K, st)
    assert np.array_equal(_bits32(cmc), _bits32(ref_cmc))
    assert mAP == ref_map
