"""Pose-graph oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain-Python restatement of the reference's pose-guided adjacency builder
(torchreid/dataset_loader.py:218-388: generate_graph :218-343, adj_graph :345-388) for the canonical
configuration num_parts = 3, method = 'same', num_scale = 1 (create_multiscale_graph :391-393 is the
identity there).  Pinned by tests/test_oracle_pose.py against tests/golden/pose_graph.npz, which
tests/golden/make_golden.py produced by calling the reference's own generate_graph.

Per frame and body-part class (head / body / leg) the reference collects the horizontal strips that hold
a confident keypoint, makes that set contiguous, adds the coarser pyramid strips above them, and finally
links every pair of DISTINCT nodes that share a class anywhere in the tracklet.  The result is therefore
fully described by three V-bit membership masks per tracklet (V = seq_len * strips): ``part_masks``.
"""
from bisect import bisect_right

import numpy as np

BODY_IDS = (('head', (0, 1, 14, 15, 16, 17)),        # dataset_loader.py:318-320
            ('body', (2, 3, 4, 5, 6, 7)),
            ('leg', (8, 9, 10, 11, 12, 13)))


def calc_splits(num_split):
    """utils/reidtools.py:13-15"""
    return [n for n in range(num_split, 0, -1) if num_split % n == 0]


def frame_sets(kp, height, num_split=4, threshold=0.1):
    """dataset_loader.py:313-336 for one frame: {part: set of 1-based strip ids}.  ``kp`` (18, 3) [x, y, conf]
    or None when the pose lookup fails (:337-338: the sets stay empty)."""
    sets = {}
    if kp is None:
        return sets
    splits = np.arange(0, height + 1, height / num_split)                       # :313
    for name, ids in BODY_IDS:
        for p in ids:
            if kp[p, 2] > threshold:                                             # :323
                loc = bisect_right(splits, kp[p, 1])                             # :326
                loc = min(num_split, max(1, loc))                                # :327
                sets.setdefault(name, set()).add(loc)
    for name, st in sets.items():                                                # :329-333 contiguous range
        if len(st) > 1:
            st.update(range(min(st), max(st) + 1))
    return sets


def pyramid_extend(sets, num_split=4):
    """dataset_loader.py:356-371: strip id -> ids of the coarser strips that contain it (1-based, levels appended)."""
    k = int(np.log2(num_split))
    out = {}
    for name, st in sets.items():
        new = set(st)
        for sid in st:
            new.update(int(np.ceil(sid / 2 ** i)) + (2 ** (k + 1) - 2 ** (k + 1 - i)) for i in range(1, k + 1))
        out[name] = new
    return out


def part_masks(keypoints, heights, valid, num_split=4, threshold=0.1, pyramid_part=True):
    """(S,18,3), (S,), (S,) -> three python ints: bit v = s * P + (strip id - 1) set when node v holds the part."""
    P = sum(calc_splits(num_split)) if pyramid_part else num_split
    masks = [0, 0, 0]
    for s in range(len(heights)):
        sets = frame_sets(keypoints[s] if valid[s] else None, heights[s], num_split, threshold)
        if pyramid_part:
            sets = pyramid_extend(sets, num_split)
        for c, (name, _) in enumerate(BODY_IDS):
            for sid in sets.get(name, ()):
                masks[c] |= 1 << (s * P + sid - 1)
    return masks


def adjacency_from_masks(masks, V):
    """dataset_loader.py:373-387 (method='same'): adj[a, b] = 1 for every ordered pair of distinct members of a class."""
    adj = np.zeros((V, V), np.float32)
    for m in masks:
        idx = [v for v in range(V) if (m >> v) & 1]
        for a in idx:
            for b in idx:
                if a != b:
                    adj[a, b] = 1.0
    return adj


def generate_graph(keypoints, heights, valid, num_split=4, threshold=0.1, pyramid_part=True):
    P = sum(calc_splits(num_split)) if pyramid_part else num_split
    return adjacency_from_masks(part_masks(keypoints, heights, valid, num_split, threshold, pyramid_part), len(heights) * P)
