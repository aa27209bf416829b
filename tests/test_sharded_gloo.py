"""world_size-2 (and 3) gloo runs of the gallery-sharded evaluation plumbing on CPU.

The collectives, the global-index bookkeeping and the [part][query][K] merge layout of
agrl.pytorch_b200.sharded are exercised with CPU stand-ins for the per-rank CUDA kernels (test-only
``ops``, written against the same key / class-byte contract as csrc/rank.cu); the result must be
bit-identical to the oracle evaluated on the concatenated gallery."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agrl.pytorch_b200 import sharded
from agrl.pytorch_b200 import synthetic as synth
from oracle import distance as odist
from oracle import rank as orank

KEY_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)


def mono_keys(d):
    """numpy twin of rank_key() in csrc/common.cuh: order-preserving uint32 of the distance."""
    d = np.where(np.isnan(d), np.float32(np.nan), d + np.float32(0.0)).astype(np.float32)
    u = d.view(np.uint32)
    k = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000))
    return np.where(np.isnan(d), np.uint32(0xFFFFFFFF), k).astype(np.uint64)


class CpuOps(object):
    """Stand-ins for the CUDA kernels, same outputs (keys as int64 bit patterns, class bytes, good counts)."""

    def distance(self, qf, gf, metric):
        return odist.distance_matrix(qf, gf, metric)

    def partial(self, d, qp, gp, qc, gc, K, offset):
        d, qp, gp, qc, gc = d.numpy(), qp.numpy(), gp.numpy(), qc.numpy(), gc.numpy()
        nq, ng = d.shape
        keys = np.full((nq, K), KEY_MAX, np.uint64)
        cls = np.zeros((nq, K), np.uint8)
        ngood = np.zeros(nq, np.int32)
        for q in range(nq):
            k = (mono_keys(d[q]) << np.uint64(32)) | np.arange(ng, dtype=np.uint64)
            top = np.sort(k)[:K]
            idx = (top & np.uint64(0xFFFFFFFF)).astype(np.int64)
            good = (gp[idx] == qp[q]) & (gc[idx] != qc[q])
            junk = (gp[idx] == -1) | ((gp[idx] == qp[q]) & (gc[idx] == qc[q]))
            keys[q, :len(top)] = top + np.uint64(offset)
            cls[q, :len(top)] = good.astype(np.uint8) | (junk.astype(np.uint8) << 1)
            ngood[q] = np.sum((gp == qp[q]) & (gc != qc[q]))
        return (torch.from_numpy(keys.view(np.int64)), torch.from_numpy(cls), torch.from_numpy(ngood),
                torch.zeros(1, dtype=torch.int32))

    def merge(self, keys_all, cls_all, ngood, K, status):
        keys = keys_all.numpy().view(np.uint64)
        cls, ngood = cls_all.numpy(), ngood.numpy()
        parts, nq, _ = keys.shape
        ap = np.zeros(nq)
        cmc = np.zeros((nq, K))
        for q in range(nq):
            kk, cc = keys[:, q, :].reshape(-1), cls[:, q, :].reshape(-1)
            order = np.argsort(kk, kind='stable')[:K]
            order = order[kk[order] != KEY_MAX]
            ap[q], cmc[q] = compute_ap(cc[order], int(ngood[q]), K)
        return np.mean(cmc, axis=0), np.mean(ap)


    # ---- market1501 metric stand-ins (same key / count contracts as csrc/rank.cu) ---------------------
    def market_count(self, qp, gp, qc, gc):
        qp, gp = qp.numpy(), gp.numpy()
        m = max(int(np.sum(gp == p)) for p in qp)
        return torch.tensor([m], dtype=torch.int32), torch.zeros(1, dtype=torch.int32)

    def market_gather(self, d, qp, gp, qc, gc, offset, cap):
        d, qp, gp, qc, gc = d.numpy(), qp.numpy(), gp.numpy(), qc.numpy(), gc.numpy()
        nq, ng = d.shape
        keys = np.full((nq, cap), KEY_MAX, np.uint64)
        counts = np.zeros((2, nq), np.int32)
        for q in range(nq):
            idx = np.nonzero(gp == qp[q])[0]
            junk = (gc[idx] == qc[q])
            k = (mono_keys(d[q, idx]) << np.uint64(32)) | ((idx.astype(np.uint64) + np.uint64(offset)) << np.uint64(1)) \
                | junk.astype(np.uint64)
            keys[q, :len(idx)] = k
            counts[0, q], counts[1, q] = np.sum(~junk), np.sum(junk)
        return torch.from_numpy(keys.view(np.int64)), torch.from_numpy(counts), torch.zeros(1, dtype=torch.int32)

    def market_bin(self, d, offset, keys_all):
        d = d.numpy()
        keys = keys_all.numpy().view(np.uint64)
        parts, nq, cap = keys.shape
        n2 = 2
        while n2 < parts * cap:
            n2 *= 2
        cnt = np.zeros((nq, n2), np.int32)
        srt = np.full((nq, n2), KEY_MAX, np.uint64)
        ng = d.shape[1]
        for q in range(nq):
            lst = np.sort(keys[:, q, :].reshape(-1))
            srt[q, :len(lst)] = lst
            real = lst[lst != KEY_MAX] & ~np.uint64(1)
            ek = (mono_keys(d[q]) << np.uint64(32)) | ((np.arange(ng, dtype=np.uint64) + np.uint64(offset)) << np.uint64(1))
            if len(real):
                t = np.searchsorted(real, ek, side='right')          # number of list items <= element
                t = t[ek < real[-1]]
                np.add.at(cnt[q], t, 1)
        return torch.from_numpy(cnt), torch.from_numpy(srt.view(np.int64))

    def market_finalize(self, cnt, srt, counts, ng_total, parts, cap, max_rank, status):
        cnt, srt, counts = cnt.numpy(), srt.numpy().view(np.uint64), counts.numpy()
        nq = cnt.shape[0]
        R = min(max_rank, ng_total)
        all_cmc, all_ap, nvalid, scratch = np.zeros((nq, R), np.float32), np.zeros(nq, np.float32), 0, np.zeros(max(ng_total, R), np.float32)
        for q in range(nq):
            npos, njunk = int(counts[0, q]), int(counts[1, q])
            if npos == 0:
                continue
            m = npos + njunk
            full = np.cumsum(cnt[q, :m])
            junk = (srt[q, :m] & np.uint64(1)).astype(bool)
            kept_rank = full - np.concatenate([[0], np.cumsum(junk)[:-1]])
            pos_ranks = kept_rank[~junk]
            kept = ng_total - njunk
            row = np.zeros(kept, np.float32)
            row[pos_ranks[0]:] = 1
            scratch[:kept] = row
            all_cmc[q] = scratch[:R]                                   # stale tail semantics (rank_cy.pyx:177)
            acc = np.float32(0)
            for c, r in enumerate(pos_ranks):
                acc = np.float32(np.float64(acc) + np.float64(c + 1) / np.float64(r + 1))
            all_ap[q] = acc / np.float32(npos)
            nvalid += 1
        assert nvalid > 0, 'Error: all query identities do not appear in gallery'
        fv = np.float32(nvalid)
        cmc = np.zeros(R, np.float32)
        for q in range(nq):
            cmc += all_cmc[q]
        mAP = np.float32(0)
        for q in range(nq):
            mAP = np.float32(mAP + all_ap[q])
        return (cmc / fv).astype(np.float32), float(np.float32(mAP / fv))


def compute_ap(cls, ngood, K):
    """Compute_AP (rank.py:180-212) on class bytes (bit0 good, bit1 junk)."""
    cmc = np.zeros(K)
    old_recall, old_precision, ap = 0, 1., 0
    inter = j = good_now = njunk = 0
    for n, c in enumerate(cls):
        flag = 0
        if c & 1:
            cmc[n - njunk:] = 1
            flag = 1
            good_now += 1
        if c & 2:
            njunk += 1
            continue
        if flag:
            inter += 1
        recall = inter / ngood
        precision = inter / (j + 1)
        ap += (recall - old_recall) * (old_precision + precision) / 2
        old_recall, old_precision = recall, precision
        j += 1
        if good_now == ngood:
            break
    return ap, cmc


def _worker(rank, world, port, case, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        qp, qc, gp, gc, qf, gf, metric, K = case
        lo, hi = sharded.shard_bounds(len(gp), world)[rank]
        qf_r = qf.clone() if rank == 0 else torch.zeros_like(qf)          # broadcast must repair this
        qp_r = qp if rank == 0 else np.zeros_like(qp)
        qc_r = qc if rank == 0 else np.zeros_like(qc)
        cmc, mAP = sharded.evaluate_mars_sharded(qf_r, gf[lo:hi], qp_r, gp[lo:hi], qc_r, gc[lo:hi],
                                                 metric=metric, max_rank=K, ops=CpuOps())
        counts = [b - a for a, b in sharded.shard_bounds(len(gp), world)]      # known shard sizes: no count exchange
        mcmc, mmAP = sharded.evaluate_market1501_sharded(qf_r, gf[lo:hi], qp_r, gp[lo:hi], qc_r, gc[lo:hi],
                                                         metric=metric, max_rank=K, ops=CpuOps(), gallery_counts=counts)
        np.savez(os.path.join(out_dir, 'r%d.npz' % rank), cmc=cmc, mAP=mAP, mcmc=mcmc, mmAP=mmAP)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world,shape,metric,ties', [(2, (40, 333, 12, 3), 'euclidean', False),
                                                     (2, (24, 101, 5, 2), 'cosine', True),
                                                     (3, (16, 200, 6, 4), 'euclidean', True)])
def test_sharded_mars_matches_unsharded_oracle(tmp_path, world, shape, metric, ties):
    qp, qc, gp, gc = synth.eval_labels(shape, seed=world)
    qf, gf = synth.eval_features(qp, gp, 32, seed=world, clustered=True)
    if ties:                                           # duplicate gallery rows across shard boundaries
        gf[1::2] = gf[0:-1:2][:len(gf[1::2])]
    K = 20
    mp.spawn(_worker, args=(world, _free_port(), (qp, qc, gp, gc, qf, gf, metric, K), str(tmp_path)),
             nprocs=world, join=True)
    d = odist.distance_matrix(qf, gf, metric).numpy()
    ref_cmc, ref_map = orank.mars_port(d, qp, gp, qc, gc, K)
    mref_cmc, mref_map = orank.market1501_port(d, qp, gp, qc, gc, K)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), 'r%d.npz' % r))
        assert np.array_equal(got['cmc'], ref_cmc), r
        assert float(got['mAP']) == float(ref_map), r
        assert np.array_equal(got['mcmc'].view(np.uint32), mref_cmc.view(np.uint32)), r
        assert float(got['mmAP']) == mref_map, r


def test_shard_bounds():
    assert sharded.shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sharded.shard_bounds(9330, 8)[-1][1] == 9330
    assert sharded.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_single_process_path_without_init():
    qp, qc, gp, gc = synth.eval_labels((12, 90, 4, 2), seed=9)
    qf, gf = synth.eval_features(qp, gp, 16, seed=9, clustered=True)
    cmc, mAP = sharded.evaluate_mars_sharded(qf, gf, qp, gp, qc, gc, max_rank=10, ops=CpuOps())
    ref = orank.mars_port(odist.distance_matrix(qf, gf, 'euclidean').numpy(), qp, gp, qc, gc, 10)
    assert np.array_equal(cmc, ref[0]) and float(mAP) == float(ref[1])


def _or_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        st = torch.tensor([[1, 4, 0][rank % 3] | (8 if rank == world - 1 else 0)], dtype=torch.int32)
        sharded.or_across_ranks(st)
        np.save(os.path.join(out_dir, 'or%d.npy' % rank), st.numpy())
    finally:
        dist.destroy_process_group()


def test_status_words_are_or_reduced_across_ranks(tmp_path):
    """different flags on different ranks must all survive (a MAX reduction would keep only the largest word)"""
    world = 3
    mp.spawn(_or_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(np.load(os.path.join(str(tmp_path), 'or%d.npy' % r))[0]) == (1 | 4 | 8)
