"""Where the gallery-sharded MARS evaluation spends its time (run under torchrun): CUDA-event time of every phase of
sharded.evaluate_mars_sharded at the bench's shape (1980 queries, 9330 gallery rows per rank, d = 4096), mean of 20
iterations after a barrier -- i.e. without the rank skew a preceding head pass adds in bench.py."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from agrl.pytorch_b200 import _lib, sharded
from agrl.pytorch_b200.sharded import CudaOps, _broadcast_queries, or_across_ranks

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl')
dev = torch.device('cuda', local)
_lib.require_device()
g = torch.Generator(device=dev).manual_seed(rank)
nq, ng, d, K = 1980, 9330, 4096, 50
qf, gf = torch.randn(nq, d, device=dev, generator=g), torch.randn(ng, d, device=dev, generator=g)
qp = torch.randint(0, 625, (nq,), device=dev, generator=g); qc = torch.randint(0, 6, (nq,), device=dev, generator=g)
gp = torch.randint(0, 625, (ng,), device=dev, generator=g); gc = torch.randint(0, 6, (ng,), device=dev, generator=g)
ops = CudaOps()
names = ['bcast', 'distance', 'partial', 'gather_keys', 'gather_cls', 'reduce_ngood', 'or_status', 'merge']
tot = {n: 0.0 for n in names}
whole = 0.0
reps = 20
for it in range(reps + 3):
    dist.barrier(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    _broadcast_queries(qf, qp, qc, None); ev[1].record()
    dm = ops.distance(qf, gf, 'euclidean'); ev[2].record()
    keys, cls, ngood, status = ops.partial(dm, qp, gp, qc, gc, K, rank * ng); ev[3].record()
    keys_all = torch.empty((world * nq, K), dtype=keys.dtype, device=dev); cls_all = torch.empty((world * nq, K), dtype=cls.dtype, device=dev)
    dist.all_gather_into_tensor(keys_all, keys); ev[4].record()
    dist.all_gather_into_tensor(cls_all, cls); ev[5].record()
    dist.all_reduce(ngood); ev[6].record()
    or_across_ranks(status, None); ev[7].record()
    r = ops.merge(keys_all.view(world, nq, K), cls_all.view(world, nq, K), ngood, K, status); ev[8].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i, n in enumerate(names):
            tot[n] += ev[i].elapsed_time(ev[i + 1])
        whole += ev[0].elapsed_time(ev[8])
if rank == 0:
    print(json.dumps({'world': world, 'ms_total': round(whole / reps, 3), 'phases_ms': {n: round(t / reps, 3) for n, t in tot.items()}}))
dist.destroy_process_group()
