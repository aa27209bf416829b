"""GPU parity of the pose-graph builder (csrc/pose.cu) and the compact-adjacency head path: bit-exact against the
adjacency the reference's own generate_graph produced (tests/golden/pose_graph.npz) and against the oracle on more
detections; the head fed with masks equals the head fed with the dense matrix bit for bit."""
import numpy as np
import pytest
import torch

from oracle import head as ohead
from oracle import pose_graph as opg
from oracle import synth
from test_oracle_pose import load_golden

pytestmark = pytest.mark.gpu


def test_masks_and_adjacency_match_the_reference_golden():
    from agrl.pytorch_b200 import pose
    kp, heights, valid, adj = load_golden()
    masks = pose.part_masks(kp, heights, valid)
    assert masks.dtype == torch.int64 and tuple(masks.shape) == (kp.shape[0], 3)
    got = pose.expand_adjacency(masks, 8).cpu().numpy()
    assert np.array_equal(got, adj)
    for b in range(kp.shape[0]):
        want = opg.part_masks(kp[b], heights[b], valid[b])
        assert [int(x) & ((1 << 64) - 1) for x in masks[b].cpu().tolist()] == want, b


@pytest.mark.parametrize('S', [1, 5, 8, 9])
def test_masks_vs_oracle_more_detections(S):
    from agrl.pytorch_b200 import pose
    B = 67
    kp, heights, valid = synth.pose_keypoints(B, S, seed=100 + S, missing=0.2)
    heights[0] = 4.0
    heights[1, 0] = 0.0                                      # numpy raises on a zero step -> the frame stays empty
    kp[2, :, :, 1] = np.inf
    kp[3, :, :, 1] = -np.inf
    masks = pose.part_masks(kp, heights, valid).cpu().tolist()
    for b in range(B):
        hts = heights[b].copy()
        v = valid[b].copy()
        v[hts == 0] = 0
        hts[hts == 0] = 1.0
        want = opg.part_masks(kp[b], hts, v)
        assert [int(x) & ((1 << 64) - 1) for x in masks[b]] == want, (S, b)
    none_valid = pose.part_masks(kp, heights, None)          # valid = NULL: every frame counts
    assert tuple(none_valid.shape) == (B, 3)


def test_generate_graph_mirror_signature_and_errors():
    from agrl.pytorch_b200 import pose, _lib
    kp, heights, valid, adj = load_golden()
    b, S = 7, 8
    paths = ['data/mars/bbox_test/%04d/%04dC1T%04dF%03d.jpg' % (b, b, b, s) for s in range(S)]
    poses = {p.split('/')[-1]: kp[b, s] for s, p in enumerate(paths) if valid[b, s]}
    sizes = [(128, int(heights[b, s])) for s in range(S)]
    got = pose.generate_graph([None] * S, im_paths=paths, im_sizes=sizes, poses=poses, num_split=4, num_parts=3,
                              num_scale=1, pyramid_part=True)
    assert not got.is_cuda and got.dtype == torch.float32 and np.array_equal(got.numpy(), adj[b])
    with pytest.raises(ValueError):
        pose.generate_graph([None], im_paths=['somewhere/else.jpg'], im_sizes=[(1, 1)], poses={}, num_split=4,
                            num_parts=3, num_scale=1, pyramid_part=True)
    with pytest.raises(NotImplementedError):
        pose.generate_graph([None] * S, paths, sizes, poses, 4, 2, 1, True)
    with pytest.raises(_lib.AgrlError):
        pose.generate_graph([None] * S, paths, sizes, poses, 8, 3, 1, True)
    assert pose.pose_key('data/prid2011/prid_2011/multi_shot/cam_a/person_0115/0006.png') == 'cam_a-person_0115-0006.png'
    assert pose.pose_key('data/ilids-vid/i-LIDS-VID/sequences/cam1/person238/cam1_person238_02519.png') == 'cam1_person238_02519.png'


@pytest.mark.parametrize('split', [1, 2])
def test_head_with_masks_equals_head_with_dense_adjacency(split):
    from agrl.pytorch_b200 import pose, _lib
    from test_gpu_head import make_model, rel_err, TOL
    S, B = 8, 9
    kp, heights, valid = synth.pose_keypoints(B, S, seed=5)
    valid[4] = 0                                             # a tracklet without any pose: zero graph rows
    masks = pose.part_masks(kp, heights, valid)
    adj = pose.expand_adjacency(masks, S)
    x1, x2 = synth.feature_maps(B, S, 2048, 16, 8, seed=6, scale=2.0)
    wts = synth.head_weights(2048, 2, seed=7, randomise_bn=True)
    model = make_model(wts, split=split)
    for lowrank in (True, False):
        model.head_lowrank = lowrank
        with torch.no_grad():
            dense = model.head(x1.cuda(), x2.cuda(), adj, S)
            compact = model.head(x1.cuda(), x2.cuda(), masks, S)
        assert torch.equal(dense, compact), lowrank
    ref = ohead.head_forward(x1, x2, adj.cpu(), wts, dtype=torch.float64)
    emax, enrm = rel_err(compact.cpu(), ref)
    assert emax < TOL and enrm < TOL
