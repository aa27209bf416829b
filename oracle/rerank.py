"""k-reciprocal re-ranking oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the reference's torchreid/utils/re_ranking.py:30-94 (Zhong et al., CVPR 2017), the optional step
between the distance matrix and the ranking in test() (train_vidreid_xent_htri.py:523-527).  Same numpy float32
arithmetic in the same order, organised by stage with sparse rows instead of the dense N x N work matrices; ties in
the initial ranking are broken by index (numpy.argsort(kind='stable'); the reference's default quicksort leaves them
undefined, so the golden vectors are produced with argsort forced stable).  Pinned by tests/test_oracle_rerank.py
against tests/golden/rerank_*.npz, outputs of the reference's own re_ranking().
"""
import numpy as np


def normalised_dist(q_g, q_q, g_g):
    """re_ranking.py:34-40: D[i, j] = orig[j, i]^2 / max_k orig[k, i]^2 (float32), orig the (N, N) block matrix."""
    orig = np.concatenate([np.concatenate([q_q, q_g], axis=1), np.concatenate([q_g.T, g_g], axis=1)], axis=0)
    orig = np.power(orig, 2).astype(np.float32)
    return np.transpose(1. * orig / np.max(orig, axis=0))


def initial_rank(D, k):
    """first k columns of argsort(D) with ties by index (re_ranking.py:42)"""
    return np.argsort(D, axis=1, kind='stable')[:, :k].astype(np.int32)


def reciprocal_row(rank, i, k):
    """re_ranking.py:50-53 (and :57-60 for the candidates): neighbours f of i within the first k with i in f's first k."""
    fwd = rank[i, :k]
    back = rank[fwd, :k]
    return fwd[np.where(back == i)[0]]


def sparse_weights(D, rank, k1):
    """re_ranking.py:48-67: per row the sorted expansion set and exp(-D) weights normalised to sum 1 (float32)."""
    half = int(np.around(k1 / 2.)) + 1
    rows = []
    for i in range(D.shape[0]):
        recip = reciprocal_row(rank, i, k1 + 1)
        expansion = recip
        for cand in recip:
            cr = reciprocal_row(rank, cand, half)
            if len(np.intersect1d(cr, recip)) > 2. / 3 * len(cr):
                expansion = np.append(expansion, cr)
        idx = np.unique(expansion)
        w = np.exp(-D[i, idx])
        rows.append((idx, (1. * w / np.sum(w)).astype(np.float32)))
    return rows


def query_expansion(rows, rank, k2, N):
    """re_ranking.py:69-74: V_qe[i] = mean of the V rows of i's first k2 neighbours (float32, rows added in rank order)."""
    out = []
    for i in range(N):
        acc = np.zeros(N, np.float32)
        dense = np.zeros((k2, N), np.float32)
        for r, n in enumerate(rank[i, :k2]):
            idx, w = rows[n]
            dense[r, idx] = w
        acc = np.mean(dense[:len(rank[i, :k2])], axis=0)
        idx = np.where(acc != 0)[0]
        out.append((idx, acc[idx]))
    return out


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    q_g, q_q, g_g = (np.asarray(a) for a in (q_g_dist, q_q_dist, g_g_dist))
    nq, N = q_g.shape[0], q_g.shape[0] + q_g.shape[1]
    D = normalised_dist(q_g, q_q, g_g)
    rank = initial_rank(D, max(k1 + 1, k2))
    rows = sparse_weights(D, rank, k1)
    if k2 != 1:
        rows = query_expansion(rows, rank, k2, N)
    # re_ranking.py:76-88: inverted index, sum over shared columns of min(V[i, c], V[j, c]) in ascending c
    V = np.zeros((N, N), np.float32)
    for i, (idx, w) in enumerate(rows):
        V[i, idx] = w
    inv = [np.where(V[:, c] != 0)[0] for c in range(N)]
    jac = np.zeros((nq, N), np.float32)
    for i in range(nq):
        t = np.zeros((1, N), np.float32)
        for c in rows[i][0]:
            t[0, inv[c]] = t[0, inv[c]] + np.minimum(V[i, c], V[inv[c], c])
        jac[i] = 1 - t / (2. - t)
    final = jac * (1 - lambda_value) + D[:nq] * lambda_value                      # :90
    return final[:nq, nq:]
