"""Ranking oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two checkers for CMC/mAP:

* ``market1501_port`` / ``mars_port``: ctypes calls into ``oracle/_build/liboracle_rank.so``
  (oracle/rank_oracle.c, the plain-C restatement of rank_cy.pyx:154-249 / rank.py:160-212).
* ``reference_evaluate_cy``: the reference's OWN Cython evaluator, compiled unmodified from
  /root/reference/torchreid/metrics/rank_cylib/rank_cy.pyx into ``oracle/_ref`` by
  ``make -C oracle ref``; numpy.argsort is forced to kind='stable' around the call because the
  reference's default (unstable) argsort leaves tie order undefined ("Ties").
"""
import contextlib
import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "_build", "liboracle_rank.so")
_lib = None


def build(ref=True):
    """Compile the C restatement (and, when /root/reference is mounted, oracle/_ref)."""
    targets = ["port"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True,
                   stdout=subprocess.DEVNULL)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PORT) or os.path.getmtime(_PORT) < os.path.getmtime(
                os.path.join(_HERE, "rank_oracle.c")):
            build(ref=False)
        lib = ctypes.CDLL(_PORT)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.oracle_eval_market1501.restype = ctypes.c_int
        lib.oracle_eval_market1501.argtypes = [f32p, i64p, i64p, i64p, i64p, ctypes.c_int64,
                                               ctypes.c_int64, ctypes.c_int64, f32p, f32p, f32p,
                                               i64p, i64p]
        lib.oracle_eval_mars.restype = ctypes.c_int
        lib.oracle_eval_mars.argtypes = [f32p, i64p, i64p, i64p, i64p, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_int64, f64p, f64p, f64p]
        lib.oracle_pairwise_sum_f64.restype = ctypes.c_double
        lib.oracle_pairwise_sum_f64.argtypes = [f64p, ctypes.c_int64]
        _lib = lib
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _prep(distmat, q_pids, g_pids, q_camids, g_camids):
    d = np.ascontiguousarray(distmat, dtype=np.float32)
    ids = [np.ascontiguousarray(x, dtype=np.int64) for x in (q_pids, g_pids, q_camids, g_camids)]
    return d, ids


def market1501_port(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, return_ap=False):
    """C restatement of rank_cy.evaluate_cy(..., use_metric_cuhk03=False) with stable ties.

    Returns (np.float32[min(max_rank, num_g)], python float) like rank_cy.pyx:241.
    """
    lib = _load()
    d, (qp, gp, qc, gc) = _prep(distmat, q_pids, g_pids, q_camids, g_camids)
    nq, ng = d.shape
    cmc = np.zeros(max(max_rank, 1), np.float32)
    mAP = ctypes.c_float(0)
    ap = np.zeros(max(nq, 1), np.float32)
    rl = ctypes.c_int64(0)
    nv = ctypes.c_int64(0)
    rc = lib.oracle_eval_market1501(_p(d, ctypes.c_float), _p(qp, ctypes.c_int64),
                                    _p(gp, ctypes.c_int64), _p(qc, ctypes.c_int64),
                                    _p(gc, ctypes.c_int64), nq, ng, max_rank,
                                    _p(cmc, ctypes.c_float), ctypes.byref(mAP),
                                    _p(ap, ctypes.c_float), ctypes.byref(rl), ctypes.byref(nv))
    if rc == 1:
        raise AssertionError('Error: all query identities do not appear in gallery')
    if rc != 0:
        raise MemoryError
    out = (cmc[:rl.value].copy(), float(mAP.value))
    return out + (ap[:nq].copy(), nv.value) if return_ap else out


def mars_port(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, return_ap=False):
    """C restatement of rank.evaluate_mars (rank.py:160-212) with stable ties.

    Returns (np.float64[max_rank], np.float64).
    """
    lib = _load()
    d, (qp, gp, qc, gc) = _prep(distmat, q_pids, g_pids, q_camids, g_camids)
    nq, ng = d.shape
    cmc = np.zeros(max(max_rank, 1), np.float64)
    mAP = ctypes.c_double(0)
    ap = np.zeros(max(nq, 1), np.float64)
    rc = lib.oracle_eval_mars(_p(d, ctypes.c_float), _p(qp, ctypes.c_int64), _p(gp, ctypes.c_int64),
                              _p(qc, ctypes.c_int64), _p(gc, ctypes.c_int64), nq, ng, max_rank,
                              _p(cmc, ctypes.c_double), ctypes.byref(mAP), _p(ap, ctypes.c_double))
    if rc == 2:
        raise ZeroDivisionError('division by zero')
    if rc == 3:
        raise ValueError('could not broadcast input array: num_g < max_rank')
    if rc != 0:
        raise MemoryError
    out = (cmc[:max_rank].copy(), np.float64(mAP.value))
    return out + (ap[:nq].copy(),) if return_ap else out


def pairwise_sum_f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return _load().oracle_pairwise_sum_f64(_p(a, ctypes.c_double), a.size)


# ----------------------------------------------------------------------------------------------
# the reference's own compiled evaluator (oracle/_ref)
# ----------------------------------------------------------------------------------------------
_ref_mod = None


def reference_rank_cy():
    """Import oracle/_ref/rank_cy*.so (the reference's rank_cy.pyx, compiled unmodified)."""
    global _ref_mod
    if _ref_mod is None:
        refdir = os.path.join(_HERE, "_ref")
        cands = [f for f in (os.listdir(refdir) if os.path.isdir(refdir) else [])
                 if f.startswith("rank_cy") and f.endswith(".so")]
        if not cands:
            return None
        spec = importlib.util.spec_from_file_location("rank_cy", os.path.join(refdir, cands[0]))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_mod = mod
    return _ref_mod


@contextlib.contextmanager
def stable_argsort():
    """Force np.argsort(kind='stable') -- reaches the call inside rank_cy (a Python-level lookup)."""
    orig = np.argsort

    def _stable(a, axis=-1, kind=None, order=None, **kw):
        return orig(a, axis=axis, kind='stable', order=order)

    np.argsort = _stable
    try:
        yield
    finally:
        np.argsort = orig


def reference_evaluate_cy(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, stable=True):
    mod = reference_rank_cy()
    if mod is None:
        raise RuntimeError("oracle/_ref/rank_cy is not built (make -C oracle ref)")
    ctx = stable_argsort() if stable else contextlib.nullcontext()
    with ctx, contextlib.redirect_stdout(sys.stderr):
        return mod.evaluate_cy(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, False)
