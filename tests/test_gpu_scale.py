"""GPU parity AT BENCH SCALE: the persistent tcgen05 GEMMs with many tiles per CTA (the `tile += tile_step` loops,
smem-ring phase wraps across tiles, TMEM double-buffer phase flips, the residual prefetched a round ahead), which the
small-batch tests never reach.  bench.py runs 882-tracklet head calls (3088 / 1764 GEMM tiles for 148 CTAs) and a
1980 x 9330 x 4096 distance matrix (1168 tiles); here the same shapes are checked against the oracle:

  * head at B = 128 / 256 / 882 -- fp64 oracle on a 64-tracklet subsample spread over the batch incl. the first and the
    last tracklet (vmgn.py:296-321); bf16x2, bf16x3, fp16 plane, fp16 + e4m3 corrections, low-rank first layer on / off;
  * distance at (1980, 9330, 4096) and (1980, 9330, 2048), both metrics, fp64 oracle (distance.py:59-89);
  * one MARS-config chain head(2048 tracklets) -> distance -> evaluate_rank: features vs the oracle on a subsample, the
    distance matrix vs the fp64 oracle on the GPU features, CMC/mAP bit-exact vs the oracle on the GPU distance matrix
    (rank.py:160-212, rank_cy.pyx:154-241);
  * a (10000 x 100000) block of the retrieval sweep: per-query top-50 keys vs a stable argsort of the same block, the
    distances vs the fp64 oracle on a query subsample.
Maps are generated on the device (882 tracklets = 14.8 GB); only the subsample is copied to the host for the oracle."""
import numpy as np
import pytest
import torch

from oracle import distance as odist
from oracle import head as ohead
from oracle import rank as orank
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4               # north-star bar
S, C, H, W = 8, 2048, 16, 8


def _model(wts, split=None):
    from agrl.pytorch_b200 import models
    kw = {} if split is None else {'head_split': split}
    m = models.init_model('vmgn', num_classes=8, loss={'xent', 'htri'}, last_stride=1, num_split=4, num_gb=2, num_scale=1,
                          pyramid_part=True, use_pose=True, learn_graph=True, pretrained=False, **kw)
    sd = m.state_dict()
    for k, v in wts.items():
        sd[k].copy_(v)
    for name in ('graph_layers', 'global_bottleneck', 'att_bottleneck'):      # the backbone is not part of this path
        getattr(m, name).cuda()
    return m.eval()


def _device_maps(B, seed, scale=2.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x1 = torch.randn(B * S, C, H, W, generator=g, device='cuda').clamp_(min=0).mul_(scale)
    x2 = torch.randn(B * S, C, H, W, generator=g, device='cuda').clamp_(min=0).mul_(scale)
    return x1, x2


def _subsample(B, n=64):
    idx = np.unique(np.concatenate([np.linspace(0, B - 1, n).round().astype(np.int64), [0, B - 1]]))
    return idx


def _oracle_on(idx, x1, x2, adj, wts, chunk=16):
    outs, nodes = [], []
    for i in range(0, len(idx), chunk):
        sel = idx[i:i + chunk]
        fr = torch.as_tensor((sel[:, None] * S + np.arange(S)[None]).reshape(-1), device=x1.device)
        o, _, n = ohead.head_forward(x1[fr].cpu(), x2[fr].cpu(), adj[torch.as_tensor(sel)].cpu(), wts, dtype=torch.float64,
                                     return_nodes=True)
        outs.append(o); nodes.append(n)
    return torch.cat(outs), torch.cat(nodes)


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).abs().max() / ref.abs().max()), float((got - ref).norm() / ref.norm())


@pytest.mark.parametrize('B,split,lowrank', [(128, 2, True), (256, 2, True), (882, 2, True), (882, 3, True), (882, 1, True),
                                             (882, 4, True), (128, 4, True), (882, 2, False), (882, 3, False), (882, 4, False),
                                             (300, 1, False)])
def test_head_at_bench_scale(B, split, lowrank):
    wts = synth.head_weights(C, 2, seed=200 + B, randomise_bn=True)
    model = _model(wts, split)
    model.head_lowrank = lowrank
    x1, x2 = _device_maps(B, seed=201 + B)
    adj = synth.pose_adjacency(B, S, 7, seed=202 + B)
    with torch.no_grad():
        out, nodes = model.head(x1, x2, adj.cuda(), S, return_nodes=True)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    idx = _subsample(B)
    ref, nodes_ref = _oracle_on(idx, x1, x2, adj, wts)
    sel = torch.as_tensor(idx, device='cuda')
    emax, enrm = _rel(out[sel].cpu(), ref)
    nmax, nnrm = _rel(nodes[sel].cpu(), nodes_ref)
    assert emax < TOL and enrm < TOL, (B, split, lowrank, emax, enrm)
    assert nmax < TOL and nnrm < TOL, (B, split, lowrank, nmax, nnrm)
    if split != 1:                                    # the fp32-accurate modes sit two orders below the bar
        assert enrm < 1e-5 and nnrm < 1e-5, (B, split, lowrank, enrm, nnrm)
    # a second call on the same inputs is bit-identical (no race between tiles / rounds of the persistent kernels)
    with torch.no_grad():
        again = model.head(x1, x2, adj.cuda(), S)
    assert torch.equal(again, out)


@pytest.mark.parametrize('d', [4096, 2048])
@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
def test_distance_at_mars_shape(metric, d):
    from agrl.pytorch_b200.metrics import compute_distance_matrix
    qp, qc, gp, gc = synth.eval_labels('mars', seed=31)
    qf, gf = synth.eval_features(qp, gp, d, seed=32, clustered=True, num_ids=626)
    out = compute_distance_matrix(qf.cuda(), gf.cuda(), metric).cpu().numpy()
    ref64 = odist.distance_matrix(qf, gf, metric, dtype=torch.float64).numpy()
    ref32 = odist.distance_matrix(qf, gf, metric).numpy()
    scale = np.abs(ref64).max()
    assert np.abs(out - ref32).max() / scale < TOL
    e_ours, e_ref = np.abs(out - ref64).max(), np.abs(ref32 - ref64).max()
    assert e_ours <= max(4 * e_ref, 2e-6 * scale), (e_ours, e_ref, scale)


def test_mars_config_chain_head_distance_rank():
    """head over 2048 tracklets (pool of 512 maps cycled with different pose graphs and per-call scales, so that the 2048
    features differ) -> 512 x 1536 distance -> both rankers.  Stage-wise parity from identical inputs."""
    from agrl.pytorch_b200 import metrics
    B, calls = 512, 4
    wts = synth.head_weights(C, 2, seed=300, randomise_bn=True)
    model = _model(wts)
    x1, x2 = _device_maps(B, seed=301)
    feats = torch.empty(calls * B, 2 * C, device='cuda')
    adjs = []
    for c in range(calls):
        adj = synth.pose_adjacency(B, S, 7, seed=310 + c)
        adjs.append(adj)
        sh = c * 37 * S                                     # every call sees the pool rotated by 37 tracklets
        with torch.no_grad():
            model.head(torch.roll(x1, sh, 0), torch.roll(x2, sh, 0), adj.cuda(), S, out=feats[c * B:(c + 1) * B])
    torch.cuda.synchronize()
    # features vs the oracle on a subsample of every call
    for c in (0, calls - 1):
        idx = _subsample(B, 16)
        sh = c * 37 * S
        r1, r2 = torch.roll(x1, sh, 0), torch.roll(x2, sh, 0)
        ref, _ = _oracle_on(idx, r1, r2, adjs[c], wts)
        emax, enrm = _rel(feats[c * B:(c + 1) * B][torch.as_tensor(idx, device='cuda')].cpu(), ref)
        assert emax < TOL and enrm < 1e-5, (c, emax, enrm)
        del r1, r2
    nq = 512
    qp, qc, gp, gc = synth.eval_labels((nq, calls * B - nq, 120, 6), seed=320)
    for metric in ('euclidean', 'cosine'):
        d = metrics.compute_distance_matrix(feats[:nq], feats[nq:], metric)
        qc_, gc_ = feats[:nq].cpu(), feats[nq:].cpu()
        ref64 = odist.distance_matrix(qc_, gc_, metric, dtype=torch.float64).numpy()
        ref32 = odist.distance_matrix(qc_, gc_, metric).numpy()
        dn = d.cpu().numpy()
        # head features of one model are close to each other: d = |q|^2 + |g|^2 - 2 q.g cancels ~3 digits, for the
        # reference's own fp32 arithmetic too -- the scale of the rounding error is the norms, not the distance
        scale = float((qc_.double() ** 2).sum(1).max() + (gc_.double() ** 2).sum(1).max()) if metric == 'euclidean' else 1.0
        e_ours, e_ref = np.abs(dn - ref64).max(), np.abs(ref32 - ref64).max()
        assert e_ours <= max(4 * e_ref, 2e-6 * scale), (metric, e_ours, e_ref, scale)
        assert np.abs(dn - ref32).max() <= TOL * scale
        cmc, mAP = metrics.evaluate_rank(d, qp, gp, qc, gc, use_metric_mars=True)
        rcmc, rmAP = orank.mars_port(dn, qp, gp, qc, gc, 50)
        assert np.array_equal(cmc, rcmc) and mAP == rmAP
        cmc, mAP = metrics.evaluate_rank(dn, qp, gp, qc, gc, use_metric_market1501=True)
        rcmc, rmAP = orank.market1501_port(dn, qp, gp, qc, gc, 50)
        assert np.array_equal(cmc, rcmc) and mAP == rmAP


def test_rank_at_mars_shape_on_gpu_distances():
    """CMC/mAP at 1980 x 9330 on the distance matrix the GPU produced from clustered features: both metrics, bit-exact."""
    from agrl.pytorch_b200 import metrics
    qp, qc, gp, gc = synth.eval_labels('mars', seed=41)
    qf, gf = synth.eval_features(qp, gp, 4096, seed=42, clustered=True, num_ids=626)
    for metric in ('euclidean', 'cosine'):
        d = metrics.compute_distance_matrix(qf.cuda(), gf.cuda(), metric)
        dn = d.cpu().numpy()
        cmc, mAP = metrics.evaluate_rank(d, qp, gp, qc, gc, use_metric_mars=True)
        rcmc, rmAP = orank.mars_port(dn, qp, gp, qc, gc, 50)
        assert np.array_equal(cmc, rcmc) and mAP == rmAP, metric
        cmc, mAP = metrics.evaluate_rank(dn, qp, gp, qc, gc, use_metric_market1501=True)
        rcmc, rmAP = orank.market1501_port(dn, qp, gp, qc, gc, 50)
        assert np.array_equal(cmc, rcmc) and mAP == rmAP, metric


def test_sweep_block_topk_vs_oracle():
    """one (10000 x 100000 x 2048) block of the retrieval sweep as bench.py runs it: prepared operands, 2000-query chunks,
    per-shard top-50 keys.  Keys vs a stable argsort of the same fp32 block (ties by gallery index), distances vs fp64."""
    from agrl.pytorch_b200 import sharded
    from agrl.pytorch_b200.metrics.distance import PreparedOperand, distance_prepared
    nq, ng, d, K, qchunk, offset = 10000, 100000, 2048, 50, 2000, 300000
    g = torch.Generator(device='cuda').manual_seed(7)
    qf = torch.randn(nq, d, generator=g, device='cuda')
    gf = torch.randn(ng, d, generator=g, device='cuda')
    gf[5000:5100] = gf[4000:4100]                           # exact duplicates: ties inside the top-k
    qf[::50] = gf[4000:4200] + 0.01 * torch.randn(200, d, generator=g, device='cuda')
    qp = torch.randint(0, 5000, (nq,), generator=g, device='cuda')
    qc = torch.randint(0, 6, (nq,), generator=g, device='cuda')
    gp = torch.randint(0, 5000, (ng,), generator=g, device='cuda')
    gp[::97] = -1                                           # distractors (junk, rank.py:167)
    gc = torch.randint(0, 6, (ng,), generator=g, device='cuda')
    ops = sharded.CudaOps()
    gop = PreparedOperand(gf, 'euclidean')
    dbuf = torch.empty(qchunk, ng, device='cuda')
    check_rows = np.unique(np.concatenate([np.arange(0, nq, 50)[:40], np.linspace(0, nq - 1, 88).round().astype(np.int64)]))
    gf64 = gf.cpu().double()
    for q0 in range(0, nq, qchunk):
        q1 = q0 + qchunk
        dm = distance_prepared(PreparedOperand(qf[q0:q1], 'euclidean'), gop, out=dbuf)
        keys, cls, ngood, st = ops.partial(dm, qp[q0:q1], gp, qc[q0:q1], gc, K, offset)
        rows = check_rows[(check_rows >= q0) & (check_rows < q1)] - q0
        rt = torch.as_tensor(rows, device='cuda')
        blk = dm[rt].cpu().numpy()
        k = keys[rt].cpu().numpy().view(np.uint64)
        order = np.argsort(blk, axis=1, kind='stable')[:, :K]
        assert np.array_equal((k & np.uint64(0xffffffff)).astype(np.int64), order + offset), q0
        # good-image counts and class bytes of the listed items (rank.py:166-169)
        gpn, gcn = gp.cpu().numpy(), gc.cpu().numpy()
        qpn, qcn = qp[q0:q1][rt].cpu().numpy(), qc[q0:q1][rt].cpu().numpy()
        good = (gpn[None] == qpn[:, None]) & (gcn[None] != qcn[:, None])
        junk = (gpn[None] == -1) | ((gpn[None] == qpn[:, None]) & (gcn[None] == qcn[:, None]))
        assert np.array_equal(ngood[rt].cpu().numpy(), good.sum(1)), q0
        c = cls[rt].cpu().numpy()
        assert np.array_equal(c & 1, np.take_along_axis(good, order, 1).astype(np.uint8)), q0
        assert np.array_equal((c >> 1) & 1, np.take_along_axis(junk, order, 1).astype(np.uint8)), q0
        ref64 = odist.distance_matrix(qf[q0:q1][rt].cpu(), gf64, 'euclidean', dtype=torch.float64).numpy()
        assert np.abs(blk - ref64).max() / np.abs(ref64).max() < 2e-6, q0
        assert int(st.cpu()) == 0
