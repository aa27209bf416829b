"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck):
compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from agrl.pytorch_b200 import _lib, metrics, models, pose, synthetic as synth
from agrl.pytorch_b200.utils.re_ranking import re_ranking_dev

_lib.require_device()
S, B = 8, 3
wts = synth.head_weights(2048, 2, seed=1, randomise_bn=True)
x1, x2 = synth.feature_maps(B, S, 2048, 16, 8, seed=2)
kp, hts, valid = synth.pose_keypoints(B, S, seed=3)
masks = pose.part_masks(kp, hts, valid)
adj = pose.expand_adjacency(masks, S)
outs = []
for split in (2, 1, 3, 4):
    m = models.init_model('vmgn', num_classes=8, loss={'xent', 'htri'}, last_stride=1, num_split=4, num_gb=2, num_scale=1,
                          pyramid_part=True, use_pose=True, learn_graph=True, pretrained=False, head_split=split)
    sd = m.state_dict()
    for k, v in wts.items():
        sd[k].copy_(v)
    m = m.cuda().eval()
    for lowrank in (True, False):
        for tma in (True, False):
            m.head_lowrank, m.pool_tma, m.gemm_pair = lowrank, tma, tma
            with torch.no_grad():
                outs.append(m.head(x1.cuda(), x2.cuda(), adj, S))
                outs.append(m.head(x1.cuda(), x2.cuda(), masks, S))
    m.head_lowrank, m.pool_tma = True, True
    with torch.no_grad():
        outs.append(m.head(x1.cuda().contiguous(memory_format=torch.channels_last),
                           x2.cuda().contiguous(memory_format=torch.channels_last), adj, S))
    outs.append(models.pool_clips(outs[-1], 3, 'avg'))
qp, qc, gp, gc = synth.eval_labels((20, 150, 6, 3), seed=4)
qf, gf = synth.eval_features(qp, gp, 300, seed=5, clustered=True)
for metric in ('euclidean', 'cosine'):
    qg = metrics.compute_distance_matrix(qf.cuda(), gf.cuda(), metric)
    qq = metrics.compute_distance_matrix(qf.cuda(), qf.cuda(), metric)
    gg = metrics.compute_distance_matrix(gf.cuda(), gf.cuda(), metric)
    rr = re_ranking_dev(qg, qq, gg)
    print(metric, metrics.evaluate_rank(rr, qp, gp, qc, gc, use_metric_mars=True)[1],
          metrics.evaluate_rank(qg, qp, gp, qc, gc, use_metric_market1501=True)[1])
# fused distance -> top-k (EpiTopK, compaction, label hash tables) incl. the overflow flag
from agrl.pytorch_b200 import sharded
from agrl.pytorch_b200.metrics.distance import PreparedOperand
ops = sharded.CudaOps()
g = torch.Generator(device='cuda').manual_seed(9)
q2, g2 = torch.randn(70, 200, generator=g, device='cuda'), torch.randn(3000, 200, generator=g, device='cuda')
lab = [torch.randint(0, 9, (n,), generator=g, device='cuda') for n in (70, 3000, 70, 3000)]
for metric in ('euclidean', 'cosine'):
    k, c, n, st = ops.topk_fused(PreparedOperand(q2, metric), PreparedOperand(g2, metric), lab[0], lab[1], lab[2], lab[3], 50, 5)
    outs.append(k.float())
# CTA-pair tiles (more than 128 queries) and the other operand splits, matrix and fused
q3 = torch.randn(200, 200, generator=g, device='cuda')
lab3 = [torch.randint(0, 9, (n,), generator=g, device='cuda') for n in (200, 3000, 200, 3000)]
for split in (_lib.SPLIT_FP16X2, _lib.SPLIT_BF16X3, _lib.SPLIT_BF16X2):
    outs.append(metrics.compute_distance_matrix(q3, g2, 'euclidean', split=split))
    k, c, n, st = sharded.CudaOps(split=split).topk_fused(PreparedOperand(q3, 'euclidean', split), PreparedOperand(g2, 'euclidean', split),
                                                          lab3[0], lab3[1], lab3[2], lab3[3], 50, 0)
    outs.append(k.float())
torch.cuda.synchronize()
print('ok', float(sum(o.double().sum() for o in outs)))
