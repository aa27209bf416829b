"""Seeded synthetic inputs for the hot path (labels, features, pose graphs, layer4 maps, head weights),
shared by tests, __graft_entry__.smoke() and bench.py.  Pure input generation: no hot-path compute.

Shapes follow SURVEY.md section 8(d): MARS 1980 x 9330 (626 ids, 6 cams), DukeV 702 x 2636
(702 ids, 8 cams), iLIDS 150 x 150, PRID 89 x 89 (ids = arange, query cam 0 / gallery cam 1 as
in data_manager/ilidsvid.py:84-87, prid2011.py:129,138).  All generators are numpy/torch-CPU and
deterministic in ``seed``.
"""
import numpy as np
import torch

EVAL_SHAPES = {
    # name: (num_q, num_g, num_ids, num_cams)
    'ilidsvid': (150, 150, 150, 2),
    'prid2011': (89, 89, 89, 2),
    'mars': (1980, 9330, 626, 6),
    'dukev': (702, 2636, 702, 8),
}


def eval_labels(name, seed=0, distractor_frac=0.05):
    """(q_pids, q_camids, g_pids, g_camids) int64.  Every query id has >=1 gallery item under a
    different camera (else evaluate_mars divides by zero, rank.py:203); the gallery also holds
    same-id-same-camera items (junk) and, for MARS, pid == -1 distractors (rank.py:167)."""
    nq, ng, nid, ncam = EVAL_SHAPES[name] if isinstance(name, str) else name
    rng = np.random.RandomState(seed)
    if ncam == 2 and nq == ng == nid:
        ids = np.arange(nid, dtype=np.int64)
        return ids.copy(), np.zeros(nq, np.int64), ids.copy(), np.ones(ng, np.int64)
    q_pids = rng.randint(0, nid, size=nq).astype(np.int64)
    q_cam = rng.randint(0, ncam, size=nq).astype(np.int64)
    g_pids = rng.randint(0, nid, size=ng).astype(np.int64)
    g_cam = rng.randint(0, ncam, size=ng).astype(np.int64)
    if distractor_frac > 0 and ng > 4 * nq:
        g_pids[rng.rand(ng) < distractor_frac] = -1
    # guarantee one cross-camera match and one same-camera junk item per query identity
    slots = rng.permutation(ng)[:2 * nq] if ng >= 2 * nq else None
    for q in range(nq):
        if slots is not None:
            a, b = slots[2 * q], slots[2 * q + 1]
        else:
            a, b = rng.randint(ng), rng.randint(ng)
        g_pids[a] = q_pids[q]
        g_cam[a] = (q_cam[q] + 1 + rng.randint(ncam - 1)) % ncam
        if slots is not None:
            g_pids[b] = q_pids[q]
            g_cam[b] = q_cam[q]
    if slots is None:      # tiny galleries: re-check the guarantee after the overwrites
        for q in range(nq):
            ok = np.any((g_pids == q_pids[q]) & (g_cam != q_cam[q]))
            if not ok:
                j = rng.randint(ng)
                g_pids[j] = q_pids[q]
                g_cam[j] = (q_cam[q] + 1) % ncam
    return q_pids, q_cam, g_pids, g_cam


def eval_features(q_pids, g_pids, d, seed=0, clustered=False, num_ids=None):
    """fp32 features ~N(0,1); ``clustered``: centroid[pid] + 0.5*N(0,1) for a non-trivial mAP."""
    g = torch.Generator().manual_seed(seed)
    nq, ng = len(q_pids), len(g_pids)
    qf = torch.randn(nq, d, generator=g)
    gf = torch.randn(ng, d, generator=g)
    if clustered:
        nid = int(max(q_pids.max(), g_pids.max())) + 2 if num_ids is None else num_ids + 1
        cent = torch.randn(nid, d, generator=g)
        qf = cent[torch.as_tensor(q_pids) + 1] + 0.5 * qf
        gf = cent[torch.as_tensor(g_pids) + 1] + 0.5 * gf
    return qf.contiguous(), gf.contiguous()


def quantised_distmat(nq, ng, seed=0, levels=64):
    """Distance matrix with MANY exact ties (values on a coarse grid) to stress stable tie-breaking."""
    rng = np.random.RandomState(seed)
    return (rng.randint(0, levels, size=(nq, ng)).astype(np.float32) / 8.0)


def pose_keypoints(B, S=8, seed=0, missing=0.1, img_h=(96, 320)):
    """Synthetic OpenPose-style detections for B tracklets of S frames: keypoints (B,S,18,3) float64 [x, y, conf] in
    ORIGINAL image coordinates, heights (B,S) float64 (original image height, dataset_loader.py:313 ``size[1]``),
    valid (B,S) uint8 (0: the pose lookup fails for that frame, :337-338).  An upright skeleton with jitter; some
    keypoints below the 0.1 confidence threshold, some exactly on strip boundaries, some outside the image."""
    rng = np.random.RandomState(seed)
    heights = rng.randint(img_h[0], img_h[1], size=(B, S)).astype(np.float64)
    heights[rng.rand(B, S) < 0.3] = 256.0                                   # MARS boxes are 256 x 128
    # nominal vertical position (fraction of the height) of the 18 COCO keypoints
    nominal = np.array([0.08, 0.18, 0.20, 0.32, 0.44, 0.20, 0.32, 0.44, 0.50, 0.70, 0.92, 0.50, 0.70, 0.92,
                        0.06, 0.06, 0.07, 0.07])
    kp = np.zeros((B, S, 18, 3), np.float64)
    shift = rng.uniform(-0.25, 0.25, size=(B, S, 1))                        # person not centred in the box
    scale = rng.uniform(0.6, 1.3, size=(B, S, 1))
    y = (nominal[None, None] * scale + shift + rng.normal(0, 0.03, size=(B, S, 18))) * heights[..., None]
    kp[..., 0] = rng.uniform(0, 128, size=(B, S, 18))
    kp[..., 1] = y
    kp[..., 2] = rng.uniform(0, 1, size=(B, S, 18))
    kp[..., 2][rng.rand(B, S, 18) < 0.15] = 0.1                             # exactly the threshold: not counted
    on_edge = rng.rand(B, S, 18) < 0.1                                       # exactly on a strip boundary
    edge = rng.randint(0, 5, size=(B, S, 18)) * (heights[..., None] / 4)
    kp[..., 1] = np.where(on_edge, edge, kp[..., 1])
    valid = (rng.rand(B, S) >= missing).astype(np.uint8)
    return kp, heights, valid


def pose_adjacency(B, S=8, P=7, seed=0, mode='pose'):
    """(B, S*P, S*P) fp32 pose graph in the format of dataset_loader.py:345-388: binary, symmetric,
    zero diagonal; every node that contains a given body part (head/body/leg) is connected to every
    other such node across all S frames.  ``mode``: 'pose' | 'zeros' (pose lookup failed, :332-333)
    | 'ones' (--use-pose off, :209-212)."""
    V = S * P
    if mode == 'zeros':
        return torch.zeros(B, V, V)
    if mode == 'ones':
        return torch.ones(B, V, V)
    rng = np.random.RandomState(seed)
    adj = np.zeros((B, V, V), np.float32)
    # pyramid for num_split=4: strips 0-3 (quarters), 4-5 (halves), 6 (whole)
    for b in range(B):
        member = np.zeros((3, V), bool)                  # part -> node membership
        for s in range(S):
            # an upright skeleton: head in strip 0(-1), body in 1-2, legs in 2-3, jittered
            spans = [(0, rng.randint(0, 2)), (1, 1 + rng.randint(0, 2)), (2 + rng.randint(0, 2), 3)]
            for part, (lo, hi) in enumerate(spans):
                if rng.rand() < 0.1:                     # part not detected in this frame
                    continue
                for q in range(lo, hi + 1):
                    member[part, s * P + q] = True
                    member[part, s * P + 4 + q // 2] = True
                    member[part, s * P + 6] = True
        for part in range(3):
            idx = np.nonzero(member[part])[0]
            adj[b][np.ix_(idx, idx)] = 1.0
        np.fill_diagonal(adj[b], 0.0)
    return torch.from_numpy(adj)


def head_weights(C=2048, num_gb=2, seed=0, randomise_bn=True):
    """Random-init head parameters with the reference's initialisers (vmgn.py:125-140 W~N(0,0.01),
    BN gamma=1 beta=0; necks gamma~N(1,0.001) via torchtools.py:51-64).  ``randomise_bn`` draws
    running stats (mean~N(0,0.1), var~U(0.5,1.5)) and affine terms away from identity so BN bugs
    cannot hide."""
    g = torch.Generator().manual_seed(seed)
    w = {}

    def bn(prefix, neck):
        w[prefix + '.weight'] = 1.0 + 0.001 * torch.randn(C, generator=g) if neck else torch.ones(C)
        w[prefix + '.bias'] = torch.zeros(C)
        w[prefix + '.running_mean'] = torch.zeros(C)
        w[prefix + '.running_var'] = torch.ones(C)
        if randomise_bn:
            w[prefix + '.weight'] = 1.0 + 0.2 * torch.randn(C, generator=g)
            w[prefix + '.bias'] = 0.1 * torch.randn(C, generator=g)
            w[prefix + '.running_mean'] = 0.1 * torch.randn(C, generator=g)
            w[prefix + '.running_var'] = 0.5 + torch.rand(C, generator=g)

    bn('global_bottleneck', True)
    bn('att_bottleneck', True)
    for i in range(num_gb):
        w['graph_layers.%d.linear.weight' % i] = 0.01 * torch.randn(C, C, generator=g)
        bn('graph_layers.%d.bn' % i, False)
    return w


def feature_maps(B, S=8, C=2048, h=16, w=8, seed=0, scale=1.0):
    """Post-ReLU-like layer4 maps (B*S, C, h, w): |N(0,1)| * scale, sparse like real ReLU outputs."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(B * S, C, h, w, generator=g).clamp_(min=0) * scale
    x2 = torch.randn(B * S, C, h, w, generator=g).clamp_(min=0) * scale
    return x1, x2
