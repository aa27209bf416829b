"""Tiny head call for `compute-sanitizer --tool racecheck` (shared-memory hazards of the TMA-staged kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from agrl.pytorch_b200 import _lib, models, pose, synthetic as synth

_lib.require_device()
S, B = 8, 2
wts = synth.head_weights(2048, 2, seed=1, randomise_bn=True)
x1, x2 = synth.feature_maps(B, S, 2048, 16, 8, seed=2)
kp, hts, valid = synth.pose_keypoints(B, S, seed=3)
adj = pose.expand_adjacency(pose.part_masks(kp, hts, valid), S)
m = models.init_model('vmgn', num_classes=8, loss={'xent', 'htri'}, last_stride=1, num_split=4, num_gb=2, num_scale=1,
                      pyramid_part=True, use_pose=True, learn_graph=True, pretrained=False)
sd = m.state_dict()
for k, v in wts.items():
    sd[k].copy_(v)
m = m.cuda().eval()
with torch.no_grad():
    for lowrank in (True, False):
        m.head_lowrank = lowrank
        out = m.head(x1.cuda(), x2.cuda(), adj, S)
torch.cuda.synchronize()
print('ok', float(out.double().sum()))
