"""GPU parity of the fused distance -> top-k (agrl_distance_topk_dev + agrl_rank_mars_classify_dev, csrc/distance.cu,
gemm_sm100.cuh EpiTopK, csrc/rank.cu classify_*): keys, class bytes and good counts must be IDENTICAL to what the unfused
route produces from the materialised matrix (agrl_distance_prepared_dev + agrl_rank_mars_partial_dev), which the other
tests pin to `np.argsort(kind='stable')[:max_rank]` of rank.py:171-172.  Covers ragged shapes, both metrics, ties and
duplicates, distractors, galleries shorter than max_rank, a shard offset, the overflow -> fallback route (a gallery
ordered by distance to a query) and the sharded evaluator with the fused route switched on."""
import numpy as np
import pytest
import torch

from oracle import rank as orank
from oracle import synth

pytestmark = pytest.mark.gpu


def _ops():
    from agrl.pytorch_b200 import sharded
    return sharded.CudaOps()


def _operands(qf, gf, metric):
    from agrl.pytorch_b200.metrics.distance import PreparedOperand
    return PreparedOperand(qf, metric), PreparedOperand(gf, metric)


def _labels(nq, ng, seed, nid=40):
    g = torch.Generator(device='cuda').manual_seed(seed)
    qp = torch.randint(0, nid, (nq,), generator=g, device='cuda')
    qc = torch.randint(0, 4, (nq,), generator=g, device='cuda')
    gp = torch.randint(0, nid, (ng,), generator=g, device='cuda')
    gc = torch.randint(0, 4, (ng,), generator=g, device='cuda')
    if ng > 10:
        gp[::7] = -1                                       # distractors (junk, rank.py:167)
    return qp, qc, gp, gc


def _both(qf, gf, metric, K, offset, seed):
    from agrl.pytorch_b200.metrics.distance import distance_prepared
    ops = _ops()
    qop, gop = _operands(qf, gf, metric)
    qp, qc, gp, gc = _labels(qf.shape[0], gf.shape[0], seed)
    fused = ops.topk_fused(qop, gop, qp, gp, qc, gc, K, offset, allow_fallback=False)
    dm = distance_prepared(qop, gop)
    ref = ops.partial(dm, qp, gp, qc, gc, K, offset)
    return fused, ref, dm


@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
@pytest.mark.parametrize('nq,ng,d,K,offset', [(300, 5000, 512, 50, 0), (129, 3001, 200, 50, 123456), (1, 385, 64, 50, 7),
                                              (257, 20000, 256, 5, 0), (64, 9330, 4096, 200, 1 << 20), (1000, 60000, 128, 50, 0)])
def test_fused_topk_equals_the_unfused_route(metric, nq, ng, d, K, offset):
    g = torch.Generator(device='cuda').manual_seed(nq + ng)
    qf = torch.randn(nq, d, generator=g, device='cuda')
    gf = torch.randn(ng, d, generator=g, device='cuda')
    if ng >= 3000:
        gf[1000:1100] = gf[2000:2100]                      # exact duplicates: ties broken by gallery index
        qf[:min(nq, 50)] = gf[2000:2000 + min(nq, 50)] + 0.05 * torch.randn(min(nq, 50), d, generator=g, device='cuda')
    (keys, cls, ngood, st), (rkeys, rcls, rngood, rst), dm = _both(qf, gf, metric, K, offset, seed=nq)
    assert int(st.cpu()) == 0 and int(rst.cpu()) == 0
    assert torch.equal(keys, rkeys)
    assert torch.equal(cls, rcls)
    assert torch.equal(ngood, rngood)
    # and the unfused keys are the stable argsort of the matrix (spot check, ties by index)
    rows = np.unique(np.linspace(0, nq - 1, 8).round().astype(np.int64))
    blk = dm[torch.as_tensor(rows, device='cuda')].cpu().numpy()
    order = np.argsort(blk, axis=1, kind='stable')[:, :K]
    got = (keys[torch.as_tensor(rows, device='cuda')].cpu().numpy().view(np.uint64) & np.uint64(0xffffffff)).astype(np.int64)
    assert np.array_equal(got[:, :min(K, ng)], order + offset)


def test_quantised_features_many_ties():
    """features on a coarse grid: thousands of exactly equal distances per row"""
    g = torch.Generator(device='cuda').manual_seed(5)
    qf = torch.randint(-1, 2, (200, 64), generator=g, device='cuda').float()
    gf = torch.randint(-1, 2, (30000, 64), generator=g, device='cuda').float()
    (keys, cls, ngood, st), (rkeys, rcls, rngood, rst), _ = _both(qf, gf, 'euclidean', 50, 0, seed=9)
    if int(st.cpu()) & 8:
        pytest.skip('tie-heavy rows overflowed the candidate lists (the fallback route is covered below)')
    assert torch.equal(keys, rkeys) and torch.equal(cls, rcls) and torch.equal(ngood, rngood)


@pytest.mark.parametrize('ng', [1, 30, 49])
def test_gallery_shorter_than_max_rank(ng):
    g = torch.Generator(device='cuda').manual_seed(ng)
    qf = torch.randn(70, 96, generator=g, device='cuda')
    gf = torch.randn(ng, 96, generator=g, device='cuda')
    (keys, cls, ngood, st), (rkeys, rcls, rngood, rst), _ = _both(qf, gf, 'euclidean', 50, 1000, seed=3)
    assert torch.equal(keys, rkeys) and torch.equal(cls, rcls) and torch.equal(ngood, rngood)
    assert int((keys[:, ng:] != -1).sum()) == 0          # empty slots: all-ones


def test_overflow_is_flagged_and_the_fallback_gives_the_same_answer():
    """gallery ordered by DEcreasing distance to query 0: every later column beats the threshold, the list overflows"""
    d, ng, nq, K = 64, 12000, 40, 50
    g = torch.Generator(device='cuda').manual_seed(11)
    qf = torch.randn(nq, d, generator=g, device='cuda')
    u = torch.nn.functional.normalize(torch.randn(d, generator=g, device='cuda'), dim=0)
    gf = qf[0][None] + (ng - torch.arange(ng, device='cuda', dtype=torch.float32))[:, None] * 0.01 * u[None]
    ops = _ops()
    qop, gop = _operands(qf, gf, 'euclidean')
    qp, qc, gp, gc = _labels(nq, ng, seed=12)
    keys, cls, ngood, st = ops.topk_fused(qop, gop, qp, gp, qc, gc, K, 0, allow_fallback=False)
    assert int(st.cpu()) & 8
    keys, cls, ngood, st = ops.topk_fused(qop, gop, qp, gp, qc, gc, K, 0)             # falls back
    from agrl.pytorch_b200.metrics.distance import distance_prepared
    rkeys, rcls, rngood, rst = ops.partial(distance_prepared(qop, gop), qp, gp, qc, gc, K, 0)
    assert torch.equal(keys, rkeys) and torch.equal(cls, rcls) and torch.equal(ngood, rngood)
    # the blocked fallback with small blocks (several query blocks) too
    k2, c2, n2, _ = ops.topk_unfused(qop, gop, qp, gp, qc, gc, K, 0, block_bytes=4 * ng * 16)
    assert torch.equal(k2, rkeys) and torch.equal(c2, rcls) and torch.equal(n2, rngood)


@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
def test_sharded_evaluator_fused_equals_unfused_and_the_oracle(metric):
    from agrl.pytorch_b200 import metrics, sharded
    qp, qc, gp, gc = synth.eval_labels((300, 40000, 80, 6), seed=21)
    qf, gf = synth.eval_features(qp, gp, 256, seed=22, clustered=True, num_ids=80)
    qf, gf = qf.cuda(), gf.cuda()
    a = sharded.evaluate_mars_sharded(qf, gf, qp, gp, qc, gc, metric=metric, fused=True)
    b = sharded.evaluate_mars_sharded(qf, gf, qp, gp, qc, gc, metric=metric, fused=False)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1]
    dn = metrics.compute_distance_matrix(qf, gf, metric).cpu().numpy()
    rcmc, rmap = orank.mars_port(dn, qp, gp, qc, gc, 50)
    assert np.array_equal(a[0], rcmc) and a[1] == rmap


@pytest.mark.parametrize('split,nq', [(2, 150), (3, 150), (3, 100), (5, 100), (5, 300)])
def test_fused_topk_with_every_operand_split(split, nq):
    """bf16 x 2 / bf16 x 3 / fp16 x 2 operands, one CTA per tile (nq <= 128) and CTA pairs: the fused route and the matrix
    route agree key for key"""
    from agrl.pytorch_b200 import sharded
    from agrl.pytorch_b200.metrics.distance import PreparedOperand, distance_prepared
    g = torch.Generator(device='cuda').manual_seed(77)
    qf = torch.randn(nq, 320, generator=g, device='cuda')
    gf = torch.randn(7000, 320, generator=g, device='cuda')
    ops = sharded.CudaOps(split=split)
    qop, gop = PreparedOperand(qf, 'euclidean', split), PreparedOperand(gf, 'euclidean', split)
    qp, qc, gp, gc = _labels(nq, 7000, seed=78)
    keys, cls, ngood, st = ops.topk_fused(qop, gop, qp, gp, qc, gc, 50, 0, allow_fallback=False)
    rkeys, rcls, rngood, _ = ops.partial(distance_prepared(qop, gop), qp, gp, qc, gc, 50, 0)
    assert int(st.cpu()) == 0
    assert torch.equal(keys, rkeys) and torch.equal(cls, rcls) and torch.equal(ngood, rngood)
