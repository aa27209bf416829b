"""engine.test with torch.distributed initialised (one process per GPU in production; gloo + CPU stand-ins here):
every rank feeds the loader of its own gallery shard, the ranking is the gallery-sharded one, every rank returns the
single-process result.  The per-rank device kernels are replaced by the CpuOps stand-ins of test_sharded_gloo.py."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agrl.pytorch_b200 import engine, sharded
from oracle import distance as odist
from oracle import rank as orank
from test_engine_host import TinyModel, loaders
from test_sharded_gloo import CpuOps, _free_port

K = 20


def _shard(loader_items, lo, hi, batch=5):
    """re-batch rows [lo, hi) of a loader (a list of (imgs, pids, camids, adj) batches)"""
    imgs = torch.cat([b[0] for b in loader_items])[lo:hi]
    pids = torch.cat([b[1] for b in loader_items])[lo:hi]
    cams = torch.cat([b[2] for b in loader_items])[lo:hi]
    adj = torch.cat([b[3] for b in loader_items])[lo:hi]
    return [(imgs[o:o + batch], pids[o:o + batch], cams[o:o + batch], adj[o:o + batch]) for o in range(0, hi - lo, batch)]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ql, gl, labels = loaders(nq=12, ng=41, seed=3)
        lo, hi = sharded.shard_bounds(41, world)[rank]
        model = TinyModel()
        r1, mAP = engine.test(model, ql, _shard(gl, lo, hi), 'avg', True, max_rank=K, verbose=(rank == 0),
                              sharded_ops=CpuOps())
        block = engine.test(model, ql, _shard(gl, lo, hi), 'avg', True, max_rank=K, return_distmat=True, verbose=False,
                            sharded_ops=CpuOps())
        refused = False
        try:
            engine.test(model, ql, _shard(gl, lo, hi), 'avg', True, max_rank=K, re_rank=True, verbose=False)
        except NotImplementedError:
            refused = True
        np.savez(os.path.join(out_dir, 'r%d.npz' % rank), r1=r1, mAP=mAP, block=block, refused=refused, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_engine_gallery_sharded_matches_single_process(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # single-process reference: the reference's sequence on the whole gallery
    ql, gl, (qp, qc, gp, gc) = loaders(nq=12, ng=41, seed=3)
    model = TinyModel()
    with torch.no_grad():
        qf = torch.cat([model(i, a) for i, _, _, a in ql])
        gf = torch.cat([model(i, a) for i, _, _, a in gl])
    # distance per shard, as the ranks compute it (a CPU matmul over a column block rounds differently from the full product)
    d = np.concatenate([odist.distance_matrix(qf, gf[lo:hi], 'euclidean').numpy()
                        for lo, hi in sharded.shard_bounds(41, world)], axis=1)
    cmc, mAP = orank.mars_port(d, qp, gp, qc, gc, K)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), 'r%d.npz' % r))
        assert float(z['r1']) == cmc[0] and float(z['mAP']) == mAP
        assert bool(z['refused'])
        # (worker processes batch / thread their CPU matmuls differently: the block is compared to fp32 accuracy)
        ref_block = d[:, int(z['lo']):int(z['hi'])]
        assert z['block'].shape == ref_block.shape and np.allclose(z['block'], ref_block, rtol=1e-5, atol=1e-2)
