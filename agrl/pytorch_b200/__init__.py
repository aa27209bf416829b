"""agrl.pytorch_b200 -- AGRL's test-time hot path on one or more B200s.

Host-side mirror of the reference's interface for that path, over the C ABI of libagrl_b200.so
(include/agrl_b200.h):

    agrl.pytorch_b200.models.init_model('vmgn', ...)        <- torchreid/models/__init__.py:34
    agrl.pytorch_b200.metrics.compute_distance_matrix       <- torchreid/metrics/distance.py:11
    agrl.pytorch_b200.metrics.evaluate_rank                 <- torchreid/metrics/rank.py:215
    agrl.pytorch_b200.metrics.rank_cylib.rank_cy.evaluate_cy <- rank_cylib/rank_cy.pyx:24
    agrl.pytorch_b200.utils.re_ranking.re_ranking           <- torchreid/utils/re_ranking.py:30   (section 8f)
    agrl.pytorch_b200.pose.generate_graph                   <- torchreid/dataset_loader.py:218    (section 8f)
    agrl.pytorch_b200.models.pool_clips                     <- train_vidreid_xent_htri.py:471-476 (section 8f)
    agrl.pytorch_b200.engine.test                           <- train_vidreid_xent_htri.py:450-542 (the path's caller)

``install_as_torchreid()`` registers these under the reference's own dotted names so an unmodified
caller (``from torchreid import metrics, models``) picks them up.  There is no CPU fallback.
"""
import sys

from . import engine, metrics, models, pose, utils

__all__ = ['install_as_torchreid']


def install_as_torchreid(force=False):
    """Alias this package's mirrors as ``torchreid.models`` / ``torchreid.metrics`` (and the
    ``torchreid.metrics.rank_cylib.rank_cy`` module the reference's rank.py:12 imports)."""
    import types
    from .metrics import rank, distance, rank_cylib
    from .metrics.rank_cylib import rank_cy
    if 'torchreid' in sys.modules and not force:
        raise RuntimeError('a torchreid package is already imported; pass force=True to shadow it')
    root = types.ModuleType('torchreid')
    root.__path__ = []
    root.metrics = metrics
    sys.modules['torchreid'] = root
    sys.modules['torchreid.metrics'] = metrics
    sys.modules['torchreid.metrics.rank'] = rank
    sys.modules['torchreid.metrics.distance'] = distance
    sys.modules['torchreid.metrics.rank_cylib'] = rank_cylib
    sys.modules['torchreid.metrics.rank_cylib.rank_cy'] = rank_cy
    root.models = models
    sys.modules['torchreid.models'] = models
    sys.modules['torchreid.models.vmgn'] = models.vmgn_module
    # section 8(f): re-ranking (train_vidreid_xent_htri.py:26) and the pose-graph builder of the loader
    root.utils = utils
    sys.modules['torchreid.utils'] = utils
    sys.modules['torchreid.utils.re_ranking'] = utils.re_ranking_module
    loader = types.ModuleType('torchreid.dataset_loader')
    loader.generate_graph = pose.generate_graph
    root.dataset_loader = loader
    sys.modules['torchreid.dataset_loader'] = loader
    return root
