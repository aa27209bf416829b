"""Drop-in for the reference's only native component, the Cython module
torchreid/metrics/rank_cylib/rank_cy.pyx: same name, same entry point, same return types --
computed by the sm_100a kernels of libagrl_b200 (csrc/rank.cu) instead of a CPU loop.

    evaluate_cy(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, use_metric_cuhk03=False)
        -> (numpy.float32[min(max_rank, num_g)], float)                    rank_cy.pyx:24-32, :241

Rows are ranked by (distance, gallery index): the result equals the reference's with
numpy.argsort(kind='stable'); the reference's default argsort leaves exact ties undefined.
"""
import ctypes

import numpy as np

from ... import _lib


def _as(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def _is_cuda_tensor(x):
    return type(x).__module__.startswith('torch') and getattr(x, 'is_cuda', False)


def evaluate_cy(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, use_metric_cuhk03=False,
                return_all_ap=False):
    if use_metric_cuhk03:
        # eval_cuhk03_cy (rank_cy.pyx:35-151) draws np.random samples; it is outside the hot path
        # scope and deliberately not re-implemented.
        raise NotImplementedError('the cuhk03 metric is outside the scope of agrl.pytorch_b200')
    if _is_cuda_tensor(distmat):
        from ..rank import _market1501_device
        return _market1501_device(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, return_all_ap)
    lib = _lib.require_device()
    d = _as(distmat, np.float32)                       # the casts of rank_cy.pyx:25-29
    qp, gp = _as(q_pids, np.int64), _as(g_pids, np.int64)
    qc, gc = _as(q_camids, np.int64), _as(g_camids, np.int64)
    assert d.ndim == 2
    num_q, num_g = d.shape
    assert qp.shape == (num_q,) and qc.shape == (num_q,) and gp.shape == (num_g,) and gc.shape == (num_g,)
    max_rank = int(max_rank)
    if num_g < max_rank:
        print('Note: number of gallery samples is quite small, got {}'.format(num_g))   # rank_cy.pyx:160-162
    cmc = np.zeros(max(max_rank, 1), np.float32)
    mAP = ctypes.c_float(0.0)
    all_ap = np.zeros(max(num_q, 1), np.float32) if return_all_ap else None
    rank_len, num_valid = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib.agrl_rank_market1501_host(
        d.ctypes.data, qp.ctypes.data, gp.ctypes.data, qc.ctypes.data, gc.ctypes.data,
        num_q, num_g, max_rank, cmc.ctypes.data, ctypes.addressof(mAP),
        all_ap.ctypes.data if return_all_ap else None,
        ctypes.addressof(rank_len), ctypes.addressof(num_valid))
    _lib.check(rc)
    out = (cmc[:rank_len.value].copy(), float(mAP.value))
    if return_all_ap:
        return out + (all_ap[:num_q], int(num_valid.value))
    return out
