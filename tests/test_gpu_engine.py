"""agrl.pytorch_b200.engine.test -- the device-resident mirror of the reference's test()
(train_vidreid_xent_htri.py:450-542) -- on a B200: the real VMGN (stock cuDNN backbone + the CUDA head) over small
query / gallery loaders, compared with (a) the reference's literal call sequence through this package's host-buffer
API (features to the host per batch, CPU-tensor distance, numpy ranking) and (b) the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import distance as odist
from oracle import rank as orank
from oracle import rerank as orerank
from oracle import synth

pytestmark = pytest.mark.gpu

NQ, NG, S = 6, 22, 8


def make_model():
    from agrl.pytorch_b200 import models
    torch.manual_seed(5)
    m = models.init_model('vmgn', num_classes=8, loss={'xent', 'htri'}, last_stride=1, num_split=4, num_gb=2,
                          num_scale=1, pyramid_part=True, use_pose=True, learn_graph=True, pretrained=False)
    return m.cuda().eval()


def make_loaders(clips=None, batch=4, seed=0):
    qp, qc, gp, gc = synth.eval_labels((NQ, NG, 4, 3), seed=seed)
    g = torch.Generator().manual_seed(seed)

    def make(pids, cams, off):
        n = len(pids)
        lead = (n,) if clips is None else (n, clips)
        imgs = torch.randn(lead + (S, 3, 256, 32), generator=g)          # 16 x 2 layer4 maps
        adj = synth.pose_adjacency(n * (clips or 1), S, 7, seed=seed + off).view(lead + (7 * S, 7 * S))
        return [(imgs[o:o + batch], torch.as_tensor(pids[o:o + batch]), torch.as_tensor(cams[o:o + batch]),
                 adj[o:o + batch]) for o in range(0, n, batch)]
    return make(qp, qc, 1), make(gp, gc, 2), (qp, qc, gp, gc)


def host_sequence(model, ql, gl, labels, metric, clips=None, pool='avg'):
    """test() as the reference writes it, on this package's host-buffer API: features to the host per batch
    (:477), CPU-tensor distance (:520-521), numpy ranking (:531)"""
    from agrl.pytorch_b200 import metrics

    def feats(loader):
        out = []
        for imgs, _, _, adj in loader:
            imgs, adj = imgs.cuda(), adj.cuda()
            if clips is not None:
                b = imgs.size(0)
                f = model(imgs.view((b * clips,) + imgs.shape[2:]), adj.view(b * clips, adj.size(-1), adj.size(-1)))
                f = f.view(b, clips, -1)
                f = torch.mean(f, 1) if pool == 'avg' else torch.max(f, 1)[0]
            else:
                f = model(imgs, adj)
            out.append(f.data.cpu())
        return torch.cat(out, 0)
    with torch.no_grad():
        qf, gf = feats(ql), feats(gl)
    d = metrics.compute_distance_matrix(qf, gf, metric).numpy()
    return qf, gf, d


def norm_scale(qf, gf, metric):
    """what a distance error is measured against: ||q||^2 + ||g||^2 (euclidean), 1 (cosine)"""
    if metric == 'cosine':
        return 1.0
    return float(qf.double().pow(2).sum(1).max() + gf.double().pow(2).sum(1).max())


@pytest.mark.parametrize('metric,re_rank', [('euclidean', False), ('cosine', False), ('euclidean', True)])
def test_engine_matches_host_sequence_and_oracle(metric, re_rank):
    from agrl.pytorch_b200 import engine, metrics
    model = make_model()
    ql, gl, labels = make_loaders()
    qp, qc, gp, gc = labels
    qf, gf, d_host = host_sequence(model, ql, gl, labels, metric)
    assert tuple(qf.shape) == (NQ, 4096) and tuple(gf.shape) == (NG, 4096) and bool(torch.isfinite(qf).all())
    r1, mAP = engine.test(model, ql, gl, 'avg', True, dist_metric=metric, re_rank=re_rank, max_rank=20, verbose=False)
    d = engine.test(model, ql, gl, 'avg', True, return_distmat=True, dist_metric=metric, re_rank=re_rank,
                    max_rank=20, verbose=False)
    assert isinstance(d, np.ndarray) and d.dtype == np.float32 and d.shape == (NQ, NG)
    scale = norm_scale(qf, gf, metric)
    if not re_rank:
        assert float(np.abs(d - d_host).max()) <= 1e-5 * scale                  # same kernels, host-buffer entry
        d_or = odist.distance_matrix(qf, gf, metric, dtype=torch.float64).numpy()
        assert float(np.abs(d - d_or).max()) <= 1e-4 * scale                    # north-star bar, vs the fp64 oracle
    else:
        # re-ranking is discrete in its inputs (top-k sets): the oracle gets the very matrices the engine made
        qd, gd = qf.cuda(), gf.cuda()
        mats = [metrics.compute_distance_matrix(a, b, metric).cpu().numpy() for a, b in ((qd, gd), (qd, qd), (gd, gd))]
        d_or = orerank.re_ranking(*mats)
        assert float(np.abs(d - d_or).max()) < 1e-5
    cmc_or, map_or = orank.mars_port(d, qp, gp, qc, gc, 20)
    assert r1 == cmc_or[0] and mAP == map_or                                    # ranking: bit-exact on the same matrix


@pytest.mark.parametrize('pool', ['avg', 'max'])
def test_engine_dense_sampling(pool):
    from agrl.pytorch_b200 import engine
    model = make_model()
    ql, gl, labels = make_loaders(clips=2, batch=3, seed=1)
    qp, qc, gp, gc = labels
    qf, gf, d_host = host_sequence(model, ql, gl, labels, 'euclidean', clips=2, pool=pool)
    d = engine.test(model, ql, gl, pool, True, return_distmat=True, test_sample='dense', max_rank=20, verbose=False)
    assert d.shape == (NQ, NG)
    assert float(np.abs(d - d_host).max()) <= 1e-5 * norm_scale(qf, gf, 'euclidean')
    r1, mAP = engine.test(model, ql, gl, pool, True, test_sample='dense', max_rank=20, verbose=False)
    cmc_or, map_or = orank.mars_port(d, qp, gp, qc, gc, 20)
    assert r1 == cmc_or[0] and mAP == map_or


def test_engine_has_no_cpu_path():
    from agrl.pytorch_b200 import engine
    model = make_model()
    ql, gl, _ = make_loaders()
    with pytest.raises(RuntimeError):
        engine.test(model, ql, gl, 'avg', False)
    with pytest.raises(RuntimeError):
        engine.test(model.cpu(), ql, gl, 'avg', True, verbose=False)
