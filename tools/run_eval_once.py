"""Run the MARS-shaped evaluation (distance + both rankers) a few times -- target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from agrl.pytorch_b200 import metrics, synthetic as synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
d = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
qp, qc, gp, gc = synth.eval_labels('mars', seed=0)
qf, gf = synth.eval_features(qp, gp, d, seed=0)
qf, gf = qf.cuda(), gf.cuda()
lab = [torch.as_tensor(a).cuda() for a in (qp, gp, qc, gc)]
for _ in range(reps):
    for metric in ('euclidean', 'cosine'):
        dm = metrics.compute_distance_matrix(qf, gf, metric)
    r1 = metrics.evaluate_rank(dm, *lab, use_metric_mars=True)
    r2 = metrics.evaluate_rank(dm, *lab, use_metric_market1501=True)
torch.cuda.synchronize()
print('mars mAP %.5f  market mAP %.5f' % (r1[1], r2[1]))
