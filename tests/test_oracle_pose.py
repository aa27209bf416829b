"""Pin oracle/pose_graph.py to the reference's own generate_graph / adj_graph (dataset_loader.py:218-388):
tests/golden/pose_graph.npz holds synthetic detections and the adjacency the reference built from them."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import pose_graph as pg
from oracle import synth


def load_golden():
    g = np.load(os.path.join(GOLDEN, 'pose_graph.npz'))
    adj = np.unpackbits(g['adj_bits'], axis=-1)[..., :56].astype(np.float32)
    return g['keypoints'], g['heights'], g['valid'], adj


def test_pose_graph_oracle_matches_reference_golden():
    kp, heights, valid, adj = load_golden()
    assert adj.shape == (24, 56, 56)
    for b in range(adj.shape[0]):
        got = pg.generate_graph(kp[b], heights[b], valid[b])
        assert np.array_equal(got, adj[b]), b
    assert adj[1].sum() == 0 and adj[2].sum() == 0           # no confident keypoint / no pose entry
    assert 0.3 < adj.mean() < 0.9


def test_masks_describe_the_graph_completely():
    """binary, symmetric, zero diagonal, and equal to the OR over the three classes of mask x mask"""
    kp, heights, valid, adj = load_golden()
    for b in range(adj.shape[0]):
        m = pg.part_masks(kp[b], heights[b], valid[b])
        assert all(0 <= x < (1 << 56) for x in m)
        a = adj[b]
        assert np.array_equal(a, a.T) and np.trace(a) == 0
        bits = np.array([[(x >> v) & 1 for v in range(56)] for x in m], bool)
        rebuilt = np.zeros((56, 56), bool)
        for c in range(3):
            rebuilt |= np.outer(bits[c], bits[c])
        np.fill_diagonal(rebuilt, False)
        assert np.array_equal(rebuilt.astype(np.float32), a)


def test_pyramid_ids_and_strip_boundaries():
    # dataset_loader.py:364-365 with num_split = 4: quarters 1,2 -> half 5, quarters 3,4 -> half 6, all -> whole 7
    assert pg.pyramid_extend({'head': {1}})['head'] == {1, 5, 7}
    assert pg.pyramid_extend({'leg': {3, 4}})['leg'] == {3, 4, 6, 7}
    kp = np.zeros((18, 3))
    kp[:, 2] = 1.0
    for y, strip in ((0.0, 1), (63.999, 1), (64.0, 2), (128.0, 3), (192.0, 4), (256.0, 4), (300.0, 4), (-5.0, 1)):
        kp[:, 1] = y                                           # bisect_right: a point ON a boundary belongs to the strip below it
        assert pg.frame_sets(kp, 256)['head'] == {strip}, y
    kp[:, 1] = 10.0
    kp[8, 1] = 250.0                                           # a leg keypoint far below: the range is made contiguous
    kp[9:14, 2] = 0.0
    assert pg.frame_sets(kp, 256)['leg'] == {4}
    kp[9, 2] = 1.0
    assert pg.frame_sets(kp, 256)['leg'] == {1, 2, 3, 4}
    assert pg.frame_sets(None, 256) == {}
    kp2, h2, v2 = synth.pose_keypoints(2, 8, seed=0)
    assert kp2.shape == (2, 8, 18, 3) and h2.shape == (2, 8) and v2.dtype == np.uint8
