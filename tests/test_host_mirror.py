"""Host-side mirror of the reference interface (no GPU needed): factory behaviour, parameter names
and shapes, argument checks that happen before any device work."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

KW = dict(num_classes=625, loss={'xent', 'htri'}, last_stride=1, num_split=4, num_gb=2, num_scale=1,
          pyramid_part=True, use_pose=True, learn_graph=True, pretrained=False)


def test_state_dict_matches_the_reference_checkpoint_layout():
    from agrl.pytorch_b200 import models
    torch.manual_seed(0)
    m = models.init_model('vmgn', **KW)
    want = json.load(open(os.path.join(GOLDEN, 'vmgn_state_dict_keys.json')))
    got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert got == want and len(got) == 402
    # same seed -> same random init as the reference constructor (same RNG consumption order)
    g = np.load(os.path.join(GOLDEN, 'full_model.npz'))
    assert abs(float(m.state_dict()['conv1.weight'].double().sum()) - float(g['conv1_checksum'])) < 1e-9


def test_factory_behaviour(tmp_path):
    from agrl.pytorch_b200 import models
    assert 'vmgn' in models.get_names()
    with pytest.raises(KeyError, match='Unknown model'):
        models.init_model('resnet50tp', **KW)
    kw = dict(KW, num_gb=1, save_dir=str(tmp_path))
    m = models.init_model('vmgn', **kw)                     # copies the model source next to the logs
    assert (tmp_path / 'vmgn.py').exists()
    assert m.total_split_list == [4, 2, 1] and m.total_split == 7 and m.feature_dim == 2048
    assert not m.global_bottleneck.bias.requires_grad and not m.att_bottleneck.bias.requires_grad
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(1, 8, 3, 32, 16), torch.zeros(1, 56, 56))


def test_distance_argument_checks_precede_device_work():
    from agrl.pytorch_b200.metrics import compute_distance_matrix as cdm
    a = torch.zeros(3, 4)
    with pytest.raises(AssertionError):
        cdm(a, torch.zeros(4, 5))
    with pytest.raises(AssertionError, match='Expected 2-D tensor'):
        cdm(a[0], a)
    with pytest.raises(AssertionError):
        cdm(a.numpy(), a)
    with pytest.raises(ValueError, match='Unknown distance metric: manhattan'):
        cdm(a, a, 'manhattan')


def test_install_as_torchreid_aliases():
    import sys
    import agrl.pytorch_b200 as pkg
    saved = {k: v for k, v in sys.modules.items() if k == 'torchreid' or k.startswith('torchreid.')}
    for k in saved:
        del sys.modules[k]
    try:
        pkg.install_as_torchreid()
        from torchreid import metrics, models
        from torchreid.metrics.rank_cylib.rank_cy import evaluate_cy
        from torchreid.metrics.rank import IS_CYTHON_AVAI
        assert metrics.evaluate_rank is pkg.metrics.evaluate_rank and callable(evaluate_cy) and IS_CYTHON_AVAI
        assert models.init_model is pkg.models.init_model
        from torchreid.utils.re_ranking import re_ranking                     # train_vidreid_xent_htri.py:26
        from torchreid.dataset_loader import generate_graph
        assert re_ranking is pkg.utils.re_ranking and generate_graph is pkg.pose.generate_graph
    finally:
        for k in [k for k in sys.modules if k == 'torchreid' or k.startswith('torchreid.')]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_pose_key_and_keypoint_packing_need_no_device():
    from agrl.pytorch_b200 import pose
    paths = ['data/mars/bbox_test/0001/0001C1T0001F%03d.jpg' % s for s in range(3)]
    poses = {paths[0].split('/')[-1]: np.ones((18, 3)), paths[2].split('/')[-1]: np.full((18, 3), 2.0)}
    kp, heights, valid = pose.pack_keypoints(paths, [(128, 256), (64, 100), (128, 256)], poses)
    assert kp.shape == (3, 18, 3) and list(valid) == [1, 0, 1] and list(heights) == [256.0, 100.0, 256.0]
    assert kp[2, 5, 1] == 2.0 and kp[1].sum() == 0
    with pytest.raises(ValueError, match='is not acceptable'):
        pose.pose_key('elsewhere/img.jpg')
    assert pose.pose_key('data/dukemtmc-vidreid/DukeMTMC-VideoReID/train/0148/0212/0148_C5_F0006_X89499.jpg') == \
        '0148-0212-0148_C5_F0006_X89499.jpg'


def test_rerank_workspace_query_and_limits():
    from agrl.pytorch_b200 import _lib
    lib = _lib.load()
    assert lib.agrl_rerank_workspace_bytes(1980, 9330, 20, 6) > 4 * 11310 * 11310
    assert lib.agrl_rerank_workspace_bytes(1980, 9330, 64, 6) == 0            # k1 beyond one warp of neighbours
    assert lib.agrl_rerank_workspace_bytes(0, 10, 20, 6) == 0
