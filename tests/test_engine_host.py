"""Host logic of agrl.pytorch_b200.engine.test -- the mirror of the reference's test()
(train_vidreid_xent_htri.py:450-542) -- exercised without a GPU: the device entry points the engine calls are
replaced by the CPU oracle (tests may do that; the product never does), the model by a small torch module, and the
result is compared with the reference's literal call sequence written out in this file."""
import numpy as np
import pytest
import torch
from torch import nn

from oracle import distance as odist
from oracle import rank as orank
from oracle import rerank as orerank
from agrl.pytorch_b200 import engine, synthetic as synth


class TinyModel(nn.Module):
    """(n, s, c, h, w) frames + (n, V, V) graph -> (n, 24) features; stands in for VMGN.forward"""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.proj = nn.Parameter(torch.randn(3 * 4 * 2, 24, generator=g))
        self.calls = 0

    def forward(self, x, adj):
        self.calls += 1
        assert x.dim() == 5 and adj.dim() == 3 and adj.size(0) == x.size(0)
        # (row-wise multiply + sum instead of a matmul: every row's feature is then independent of how the rows are batched)
        f = (x.mean(dim=1).flatten(1).unsqueeze(2) * self.proj.unsqueeze(0)).sum(dim=1)
        return f + adj.mean(dim=(1, 2)).unsqueeze(1)


def loaders(nq=12, ng=40, clips=None, batch=5, seed=0):
    qp, qc, gp, gc = synth.eval_labels((nq, ng, 6, 3), seed=seed)
    g = torch.Generator().manual_seed(seed)

    def make(pids, cams):
        n = len(pids)
        shape = (n, 2, 3, 4, 2) if clips is None else (n, clips, 2, 3, 4, 2)
        imgs = torch.randn(shape, generator=g) + torch.as_tensor(pids, dtype=torch.float32).view(-1, *([1] * (len(shape) - 1)))
        adj = torch.rand((n, 14, 14) if clips is None else (n, clips, 14, 14), generator=g)
        out = []
        for o in range(0, n, batch):
            out.append((imgs[o:o + batch], torch.as_tensor(pids[o:o + batch]), torch.as_tensor(cams[o:o + batch]),
                        adj[o:o + batch]))
        return out
    return make(qp, qc), make(gp, gc), (qp, qc, gp, gc)


@pytest.fixture
def cpu_entries(monkeypatch):
    """the three device entries of the engine -> CPU oracle"""
    class M(object):
        @staticmethod
        def compute_distance_matrix(a, b, metric='euclidean'):
            if metric not in ('euclidean', 'cosine'):
                raise ValueError('Unknown distance metric: {}. Please choose either "euclidean" or "cosine"'.format(metric))
            return odist.distance_matrix(a, b, metric)

        @staticmethod
        def evaluate_rank(d, qp, gp, qc, gc, max_rank=50, use_metric_mars=False, **kw):
            assert use_metric_mars
            return orank.mars_port(np.asarray(d), qp, gp, qc, gc, max_rank)

    def pool_clips(f, n, pool):
        f = f.view(-1, n, f.size(1))
        return f.mean(1) if pool == 'avg' else f.max(1)[0]

    monkeypatch.setattr(engine, 'metrics', M)
    monkeypatch.setattr(engine, 'pool_clips', pool_clips)
    monkeypatch.setattr(engine, 're_ranking_dev',
                        lambda a, b, c: torch.from_numpy(orerank.re_ranking(np.asarray(a), np.asarray(b), np.asarray(c))))


def reference_sequence(model, ql, gl, labels, metric, re_rank=False, clips=None, pool='avg'):
    """the reference's test() body, literally (per-batch features to the host, then distance, [re-ranking], rank)"""
    def feats(loader):
        out = []
        for imgs, _, _, adj in loader:
            if clips is not None:
                for t in range(imgs.size(0)):                    # the reference folds ONE tracklet's clips per batch
                    f = model(imgs[t], adj[t]).view(clips, 1, -1)
                    out.append(torch.mean(f, 0) if pool == 'avg' else torch.max(f, 0)[0])
            else:
                out.append(model(imgs, adj))
        return torch.cat(out, 0)
    with torch.no_grad():
        qf, gf = feats(ql), feats(gl)
    qp, qc, gp, gc = labels
    d = odist.distance_matrix(qf, gf, metric).numpy()
    if re_rank:
        d = orerank.re_ranking(d, odist.distance_matrix(qf, qf, metric).numpy(), odist.distance_matrix(gf, gf, metric).numpy())
    return d, orank.mars_port(d, qp, gp, qc, gc, 20)


@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
@pytest.mark.parametrize('re_rank', [False, True])
def test_engine_follows_the_reference_sequence(cpu_entries, capsys, metric, re_rank):
    ql, gl, labels = loaders()
    model = TinyModel()
    d_ref, (cmc_ref, map_ref) = reference_sequence(model, ql, gl, labels, metric, re_rank)
    model.train()
    r1, mAP = engine.test(model, ql, gl, 'avg', True, ranks=(1, 5), dist_metric=metric, re_rank=re_rank, max_rank=20)
    assert not model.training                                         # test() switches to eval mode (:454)
    assert r1 == cmc_ref[0] and mAP == map_ref
    out = capsys.readouterr().out
    assert 'Extracted features for query set, obtained 12-by-24 matrix' in out
    assert 'Extracted features for gallery set, obtained 40-by-24 matrix' in out
    assert 'Computing distance matrix with metric={} ...'.format(metric) in out
    assert ('Applying person re-ranking ...' in out) == re_rank
    assert 'mAP: {:.2%}'.format(map_ref) in out and 'Rank-5  : {:.2%}'.format(cmc_ref[4]) in out
    d = engine.test(model, ql, gl, 'avg', True, return_distmat=True, dist_metric=metric, re_rank=re_rank,
                    max_rank=20, verbose=False)
    assert isinstance(d, np.ndarray) and d.shape == (12, 40) and np.array_equal(d, d_ref)
    assert capsys.readouterr().out == ''


@pytest.mark.parametrize('pool', ['avg', 'max'])
def test_engine_dense_sampling_pools_the_clips_of_each_tracklet(cpu_entries, pool):
    import types
    ql, gl, labels = loaders(nq=7, ng=30, clips=3, batch=4, seed=2)
    model = TinyModel()
    d_ref, (cmc_ref, map_ref) = reference_sequence(model, ql, gl, labels, 'euclidean', clips=3, pool=pool)
    args = types.SimpleNamespace(test_sample='dense', dist_metric='euclidean', re_rank=False)
    model.calls = 0
    r1, mAP = engine.test(model, ql, gl, pool, True, args=args, max_rank=20, verbose=False)
    assert model.calls == len(ql) + len(gl)                            # one model call per loader batch
    assert abs(mAP - map_ref) < 1e-12 and r1 == cmc_ref[0]


def test_engine_errors(cpu_entries):
    ql, gl, labels = loaders()
    model = TinyModel()
    with pytest.raises(RuntimeError, match='no CPU path'):
        engine.test(model, ql, gl, 'avg', False)
    with pytest.raises(ValueError, match='Unknown distance metric'):
        engine.test(model, ql, gl, 'avg', True, dist_metric='manhattan', verbose=False)
    # a query whose identity never appears under another camera: evaluate_mars divides by zero (rank.py:203)
    bad = [(i, torch.full_like(p, 999), c, a) for i, p, c, a in ql]
    with pytest.raises(ZeroDivisionError):
        engine.test(model, bad, gl, 'avg', True, max_rank=20, verbose=False)
