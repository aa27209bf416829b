#!/usr/bin/env python
"""bench.py -- the hot path on synthetic MARS-shaped data, one JSON line ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Job (BASELINE.json metric "MARS-shape tracklets/s (graph head) + query x gallery eval ms"):
one STEP is a whole MARS-shaped test pass minus the stock backbone -- the graph head over
J = 1980 + 9330 = 11310 tracklets (8 frames, 2048 x 16 x 8 layer4 maps each, fed from a resident
pool that is larger than L2 and cycled), then the 1980 x 9330 distance matrix on the 4096-d
features the head just produced, then CMC/mAP (MARS metric, what the reference's test() calls).
`value` = J / step time with the maps resident in HBM; `eval_ms` carries the distance + ranking
part; `e2e` is the same job through the host-buffer API (maps, features, distance matrix and labels
all start in pinned host memory, results end on the host).  At N > 1 every rank runs the head on
its own J tracklets (no collective) and owns a 9330-row gallery shard of a N x 9330 gallery; the
per-query top-k / good counts are merged with NCCL (weak scaling).

`--impl reference` times the reference's CPU implementation of the same path on the host cores
(oracle/: torch-CPU restatement of head and distance, the reference's own compiled rank_cy and the
C restatement of evaluate_mars), on a bounded sample of the job, and prints the same line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

NQ, NG, NIDS, NCAMS = 1980, 9330, 626, 6
S, C, H, W = 8, 2048, 16, 8
BYTES_PER_TRACKLET = 2 * S * C * H * W * 4 + 56 * 56 * 4 + 2 * C * 4        # 16 806 144
FLOPS_PER_TRACKLET = 2 * (2 * 56 * C * C + 2 * 2 * 56 * 56 * C)            # ~0.99 GFLOP
METRIC = 'MARS-shape tracklets/s (graph head) + query×gallery eval ms at 1/2/4/8 B200'


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return dict(hbm_gbs=p['hbm_gbs'], bf16_tflops=p['bf16_tflops'], bf16_sustained=p['bf16_tflops_sustained'],
                    source='measured')
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            try:
                power.append(float(f[3]))
            except ValueError:
                pass
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), power_w=float(np.median(power)) if power else None)
        return out


# ------------------------------------------------------------------------------------------------
# synthetic job
# ------------------------------------------------------------------------------------------------
def make_labels(rank, world):
    from agrl.pytorch_b200 import synthetic as synth
    qp, qc, gp, gc = synth.eval_labels((NQ, NG, NIDS, NCAMS), seed=100 + rank)
    return qp, qc, gp, gc


def make_head_weights(seed=0):
    g = torch.Generator().manual_seed(seed)
    w = {}
    for name in ('global_bottleneck', 'att_bottleneck', 'graph_layers.0.bn', 'graph_layers.1.bn'):
        w[name + '.weight'] = 1.0 + 0.1 * torch.randn(C, generator=g)
        w[name + '.bias'] = 0.1 * torch.randn(C, generator=g)
        w[name + '.running_mean'] = 0.1 * torch.randn(C, generator=g)
        w[name + '.running_var'] = 0.5 + torch.rand(C, generator=g)
    for i in range(2):
        w['graph_layers.%d.linear.weight' % i] = 0.01 * torch.randn(C, C, generator=g)
    return w


def make_model(dev, weights):
    from agrl.pytorch_b200 import models
    m = models.init_model('vmgn', num_classes=625, loss={'xent', 'htri'}, last_stride=1, num_split=4, num_gb=2,
                          num_scale=1, pyramid_part=True, use_pose=True, learn_graph=True, pretrained=False)
    sd = m.state_dict()
    for k, v in weights.items():
        sd[k].copy_(v)
    # only the head's parameters are needed on the device (the backbone is not part of this path)
    for name in ('graph_layers', 'global_bottleneck', 'att_bottleneck'):
        getattr(m, name).to(dev)
    return m.eval()


def make_pool(n, dev, seed, pinned=False):
    """n tracklets of post-ReLU-like layer4 maps + a pose adjacency each; generated on the target."""
    shape = (n * S, C, H, W)
    if pinned:
        g = torch.Generator().manual_seed(seed)
        x1 = torch.empty(shape, pin_memory=True)
        x2 = torch.empty(shape, pin_memory=True)
        blk = 64 * S
        for i in range(0, n * S, blk):             # fill in blocks: keeps the transient small
            x1[i:i + blk] = torch.randn((min(blk, n * S - i), C, H, W), generator=g).clamp_(min=0)
            x2[i:i + blk] = torch.randn((min(blk, n * S - i), C, H, W), generator=g).clamp_(min=0)
    else:
        g = torch.Generator(device=dev).manual_seed(seed)
        x1 = torch.randn(shape, generator=g, device=dev).clamp_(min=0)
        x2 = torch.randn(shape, generator=g, device=dev).clamp_(min=0)
    from agrl.pytorch_b200 import synthetic as synth
    adj = synth.pose_adjacency(n, S, 7, seed=seed)
    adj = adj.pin_memory() if pinned else adj.to(dev)
    return x1, x2, adj


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from agrl.pytorch_b200 import _lib, metrics, sharded

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local), timeout=datetime.timedelta(seconds=180))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    lib = _lib.require_device()
    pk = peaks()

    J, pool_n = NQ + NG, args.pool
    weights = make_head_weights()
    model = make_model(dev, weights)
    x1, x2, adj = make_pool(pool_n, dev, seed=1 + rank)
    qp, qc, gp, gc = make_labels(rank, world)
    lab = [torch.as_tensor(a).to(dev) for a in (qp, gp, qc, gc)]
    feats = torch.empty(J, 2 * C, device=dev)
    chunks = [(o, min(pool_n, J - o)) for o in range(0, J, pool_n)]
    stream = torch.cuda.current_stream(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def head_pass():
        for off, n in chunks:
            model.head(x1[:n * S], x2[:n * S], adj[:n], S, out=feats[off:off + n])

    def eval_pass():
        if world == 1:
            d = metrics.compute_distance_matrix(feats[:NQ], feats[NQ:], args.dist_metric)
            return metrics.evaluate_rank(d, lab[0], lab[1], lab[2], lab[3], use_metric_mars=True)
        return sharded.evaluate_mars_sharded(feats[:NQ], feats[NQ:], lab[0], lab[1], lab[2], lab[3],
                                             metric=args.dist_metric, max_rank=50)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for _ in range(args.warmup):
            head_pass(); result = eval_pass()
        # ---- timed region: exactly K steps, device timed, max over ranks -------------------------
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        launches0 = _lib.launch_count()
        t_head, t_eval = [], []
        e0, e3 = ev(), ev()
        e0.record(stream)
        marks = []
        for _ in range(args.steps):
            a, b, c = ev(), ev(), ev()
            a.record(stream); head_pass(); b.record(stream); result = eval_pass(); c.record(stream)
            marks.append((a, b, c))
        e3.record(stream)
        barrier()
        total_ms = e0.elapsed_time(e3)
        launches = _lib.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        head_ms = float(np.mean([a.elapsed_time(b) for a, b, c in marks]))
        eval_ms = float(np.mean([b.elapsed_time(c) for a, b, c in marks]))
        if world > 1:
            t = torch.tensor([total_ms, head_ms, eval_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms, head_ms, eval_ms = [float(v) for v in t.cpu()]
        ms_per_step = total_ms / args.steps
        value = world * J / (ms_per_step * 1e-3)

        # ---- the same head pass with the opt-in single-plane fp16 graph-layer GEMM (reported beside, not as `value`) --
        fast = None
        if not args.no_fast_mode:
            model.head_split = _lib.SPLIT_FP16X1
            for _ in range(2):
                head_pass()
            barrier()
            f0, f1 = ev(), ev()
            f0.record(stream)
            for _ in range(args.steps):
                head_pass()
            f1.record(stream)
            barrier()
            fast_ms = f0.elapsed_time(f1) / args.steps
            if world > 1:
                t = torch.tensor([fast_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                fast_ms = float(t.cpu())
            with _lib.profile(stream.cuda_stream) as fprof:
                head_pass()
            fast = dict(head_ms=fast_ms, kernels={k: round(t, 4) for k, (n, t) in fprof.totals().items()})
            model.head_split = _lib.SPLIT_BF16X2
            head_pass()                                  # features of the default mode again (for e2e / result parity)
            barrier()

        # ---- the same head pass with the opt-in low-rank first layer (option head_lowrank; reported beside) -----------
        lowrank = None
        if not args.no_lowrank:
            try:
                _lib.set_option('head_lowrank', 1)
                for _ in range(2):
                    head_pass()
                torch.cuda.synchronize(dev)      # (no collective inside the guarded block)
                l0, l1 = ev(), ev()
                l0.record(stream)
                for _ in range(args.steps):
                    head_pass()
                l1.record(stream)
                torch.cuda.synchronize(dev)      # (no collective inside the guarded block)
                low_ms = l0.elapsed_time(l1) / args.steps
                with _lib.profile(stream.cuda_stream) as lprof:
                    head_pass()
                lowrank = dict(head_ms=low_ms, kernels={k: round(t, 4) for k, (n, t) in lprof.totals().items()})
            except Exception as exc:                     # an extra figure must never cost the bench line
                lowrank = {'unavailable': '%s: %s' % (type(exc).__name__, exc)}
            finally:
                _lib.set_option('head_lowrank', 0)
            if world > 1:                                # every rank takes part, whatever happened on it
                t = torch.tensor([lowrank.get('head_ms', -1.0)], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if 'head_ms' in lowrank:
                    lowrank['head_ms'] = float(t.cpu())
            head_pass()                                  # features of the default mode again
            barrier()

        # ---- everything opt-in at once, on calls large enough for the partitioned pipeline (reported beside) ----------
        # fp16 plane + low-rank first layer + spatially partitioned pooling, 1764 tracklets per call (6 sub-batches of 294);
        # the configuration of profiles/r1/lowrank_probe_v8.log ("split=1,lr=1,sub=294,mode=0,psms=64,stages=6")
        tuned = None
        if not args.no_tuned:
            tuned_opts = {'head_lowrank': 1, 'head_sub_batch': 294, 'overlap_mode': 0, 'pool_sms': 64, 'pool_stages': 6}
            saved_opts = {k: _lib.get_option(k) for k in tuned_opts}
            big = None
            try:
                big_n = 2 * pool_n
                big = make_pool(big_n, dev, seed=11 + rank)
                big_chunks = [(o, min(big_n, J - o)) for o in range(0, J, big_n)]

                def tuned_pass():
                    for off, n in big_chunks:
                        model.head(big[0][:n * S], big[1][:n * S], big[2][:n], S, out=feats[off:off + n])
                for k, v in tuned_opts.items():
                    _lib.set_option(k, v)
                model.head_split = _lib.SPLIT_FP16X1
                for _ in range(2):
                    tuned_pass()
                torch.cuda.synchronize(dev)
                t0, t1 = ev(), ev()
                t0.record(stream)
                for _ in range(args.steps):
                    tuned_pass()
                t1.record(stream)
                torch.cuda.synchronize(dev)
                tuned = dict(head_ms=t0.elapsed_time(t1) / args.steps, call_tracklets=big_n, options=dict(tuned_opts))
            except Exception as exc:                     # an extra figure must never cost the bench line
                tuned = {'unavailable': '%s: %s' % (type(exc).__name__, exc)}
            finally:
                for k, v in saved_opts.items():
                    _lib.set_option(k, v)
                model.head_split = _lib.SPLIT_BF16X2
                del big
                torch.cuda.empty_cache()
            if world > 1:                                # every rank takes part, whatever happened on it
                t = torch.tensor([tuned.get('head_ms', -1.0)], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if 'head_ms' in tuned:
                    tuned['head_ms'] = float(t.cpu())
            head_pass()                                  # features of the default mode again
            barrier()

        # ---- kernel timeline of one more step (CUDA events after every kernel, same stream) ------
        # (every rank runs the pass -- eval_pass holds collectives at N > 1 -- rank 0's timeline is reported)
        barrier()
        with _lib.profile(stream.cuda_stream) as prof:
            head_pass(); eval_pass()
        timeline = prof.totals()
        barrier()

        # ---- same-box comparator: the reference's module code in stock PyTorch on this GPU (rank 0) ----
        eager = None
        if not args.no_eager and rank == 0:
            try:
                eager = torch_eager_gpu(args, dev, weights, x1, x2, adj, feats)
            except Exception as exc:                     # a comparator must never cost the bench line
                eager = {'unavailable': '%s: %s' % (type(exc).__name__, exc)}
            torch.cuda.empty_cache()
        barrier()

        # ---- end to end through host buffers -------------------------------------------------
        e2e = None
        if not args.no_e2e:
            e2e = run_e2e(args, model, dev, rank, world, (qp, qc, gp, gc))

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    step_ms = sum(t for _, t in timeline.values())
    kern = {k: dict(launches=n, ms=round(t, 4), share=round(t / step_ms, 4)) for k, (n, t) in sorted(
        timeline.items(), key=lambda kv: -kv[1][1])}
    top = max(timeline.items(), key=lambda kv: kv[1][1])[0]
    n_top, ms_top = timeline[top]
    # roofline of the dominant kernel (algorithmic work per launch / measured launch duration)
    if top == 'pool':
        per_launch = pool_n * BYTES_PER_TRACKLET
        roof = dict(kernel=top, bound='hbm', unit='GB/s', peak=pk['hbm_gbs'],
                    achieved=J * BYTES_PER_TRACKLET / (ms_top * 1e-3) / 1e9)
    elif top.startswith('gemm'):
        passes = 3 if top == 'gemm_graph_layer' else 6
        flops = passes * (2 * 56 * C * C * J * 2 if top == 'gemm_graph_layer' else 2 * NQ * NG * 2 * C)
        roof = dict(kernel=top, bound='tensor', unit='TFLOP/s', peak=pk['bf16_sustained'],
                    achieved=flops / (ms_top * 1e-3) / 1e12, passes=passes)
    else:
        roof = dict(kernel=top, bound='hbm', unit='GB/s', peak=pk['hbm_gbs'],
                    achieved=(J * 56 * C * 4 * 2) / (ms_top * 1e-3) / 1e9)
    roof['frac'] = roof['achieved'] / roof['peak']
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/traffic.json),
    # scaled to this run's units per launch
    roof['traffic'] = None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))[top]
        per_unit = tr['dram_bytes_per_launch'] / tr['units_per_launch']
        roof['traffic'] = per_unit * (pool_n if tr['unit'] == 'tracklet' else 1)
        roof['algorithmic_bytes_per_launch'] = pool_n * BYTES_PER_TRACKLET if top == 'pool' else None
    except Exception:
        pass
    roof['peak_source'] = pk['source']
    roof['launches'] = n_top
    roof['avg_launch_ms'] = ms_top / n_top
    # whole-head roofline as SURVEY 8(d) defines it: algorithmic bytes per tracklet / head time
    head_gbs = J * BYTES_PER_TRACKLET / (head_ms * 1e-3) / 1e9
    gemm_d = timeline.get('gemm_distance', (1, float('nan')))[1]
    line = {
        'metric': METRIC, 'value': value, 'unit': 'tracklets/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'fp32 (bf16x2/bf16x3 split operands on tcgen05, fp32 accumulate)',
        'data': 'synthetic',
        'config': {'workload': 'MARS-shaped test pass: graph head over 11310 tracklets (8 frames, 2048x16x8 maps), '
                               '1980x9330 %s distance on the 4096-d features, MARS-metric CMC/mAP' % args.dist_metric,
                   'tracklets_per_step_per_gpu': J, 'pool_tracklets': pool_n,
                   'head': 'bulk-copy pooling (TMA ring), graph layers on tcgen05 (graph_kernel_tc + bf16x2 split GEMM, 3 products); options %s' % (
                       {k: _lib.get_option(k) for k in ('head_sub_batch', 'pool_tma', 'pool_stages', 'graph_variant', 'gemm_pair')},),
                   'cache': 'input pool %.1f GB per GPU, larger than L2; cycled' % (pool_n * BYTES_PER_TRACKLET / 1e9),
                   'parallelism': 'independent head shards + gallery-sharded eval (NCCL merge)' if world > 1 else 'single GPU'},
        'head_ms': head_ms, 'eval_ms': eval_ms,
        'head_tracklets_per_s_per_gpu': J / (head_ms * 1e-3),
        'head_hbm': {'achieved': head_gbs, 'peak': pk['hbm_gbs'], 'frac': head_gbs / pk['hbm_gbs'], 'unit': 'GB/s',
                     'note': 'algorithmic 16806144 B per tracklet / whole-head time (SURVEY 8d)'},
        'distance': {'ms': gemm_d, 'algorithmic_tflops': 2 * NQ * NG * 2 * C / (gemm_d * 1e-3) / 1e12,
                     'tensor_pipe_frac': 6 * 2 * NQ * NG * 2 * C / (gemm_d * 1e-3) / 1e12 / pk['bf16_sustained']},
        'roofline': roof, 'kernels': kern, 'gpu_launches': int(launches),
        'clocks': clocks, 'result': {'mAP': float(result[1]), 'rank1': float(result[0][0])},
    }
    if fast is not None:
        fgbs = J * BYTES_PER_TRACKLET / (fast['head_ms'] * 1e-3) / 1e9
        line['fast_mode'] = {
            'what': 'opt-in head_split=1: ONE fp16 plane per GEMM operand, pow2-scaled per tracklet / per layer (TF32-class, '
                    '11 significant bits); head error vs the reference 1e-5 norm-relative / 3e-5 max-scaled (bar 1e-4, '
                    'tests/test_gpu_head.py); NOT the configuration `value` is measured in',
            'head_ms': fast['head_ms'], 'head_tracklets_per_s_per_gpu': J / (fast['head_ms'] * 1e-3),
            'head_hbm_frac': fgbs / pk['hbm_gbs'], 'kernels_ms': fast['kernels']}
    if lowrank is not None:
        if 'head_ms' in lowrank:
            lgbs = J * BYTES_PER_TRACKLET / (lowrank['head_ms'] * 1e-3) / 1e9
            lowrank = {'what': 'opt-in head_lowrank=1: the first graph layer runs X.W^T on the 32 quarter-strip rows per tracklet '
                               '(G.X.W^T = (G.T).(Q.W^T)) and applies G.T afterwards; same fp32-accurate arithmetic class as the '
                               'default (tests/test_lowrank_layer1.py, test_gpu_head.py::test_lowrank_first_layer_agrees); NOT the '
                               'configuration `value` is measured in this round',
                       'head_ms': lowrank['head_ms'], 'head_tracklets_per_s_per_gpu': J / (lowrank['head_ms'] * 1e-3),
                       'head_hbm_frac': lgbs / pk['hbm_gbs'], 'kernels_ms': lowrank['kernels']}
        line['lowrank_mode'] = lowrank
    if tuned is not None:
        if 'head_ms' in tuned:
            tgbs = J * BYTES_PER_TRACKLET / (tuned['head_ms'] * 1e-3) / 1e9
            tuned.update(what='every opt-in at once: fp16 plane (TF32-class operand rounding, head error 1e-5 vs the 1e-4 bar) + '
                              'low-rank first layer + spatially partitioned pooling (64 SMs stream the maps beside the graph / GEMM '
                              'kernels of the previous sub-batch) on %d-tracklet calls; NOT the configuration `value` is measured in'
                              % tuned['call_tracklets'],
                         head_tracklets_per_s_per_gpu=J / (tuned['head_ms'] * 1e-3), head_hbm_frac=tgbs / pk['hbm_gbs'])
        line['tuned_fast_mode'] = tuned
    if eager is not None:
        if 'head_ms_per_pass' in eager:
            eager['b200_head_speedup'] = eager['head_ms_per_pass'] / head_ms
            eager['b200_distance_speedup'] = eager['distance_ms'] / gemm_d
        line['torch_eager_gpu'] = eager
    if e2e is not None:
        line['e2e'] = e2e
    if not args.no_cpu_baseline and world >= 1:
        line['cpu_baseline'] = cpu_reference(args, steps=1, warmup=0)['cpu_baseline']
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_sweep(args):
    """Scaled retrieval sweep (BASELINE.json config 5): 10 000 queries x 1 000 000 gallery features
    (d = 2048), gallery rows sharded over the ranks (strong scaling), MARS-metric top-50 merge over NCCL.
    One step = prepare the gallery shard's operand planes, then per 2000-query chunk: distance block on
    tcgen05 + per-shard top-k, then one all-gather / all-reduce and the merge."""
    import datetime
    import torch.distributed as dist
    from agrl.pytorch_b200 import _lib, sharded
    from agrl.pytorch_b200.metrics.distance import PreparedOperand, distance_prepared
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local), timeout=datetime.timedelta(seconds=180))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    _lib.require_device()
    nq, ng_total, d, K, qchunk = args.sweep_queries, args.sweep_gallery, 2048, 50, 2000
    lo, hi = sharded.shard_bounds(ng_total, world)[rank]
    ng = hi - lo
    g = torch.Generator(device=dev).manual_seed(5)
    qf = torch.randn(nq, d, generator=g, device=dev)                       # same on every rank (same seed)
    g2 = torch.Generator(device=dev).manual_seed(50 + rank)
    gf = torch.randn(ng, d, generator=g2, device=dev)
    nid = max(ng_total // 20, 2)
    qp = torch.randint(0, nid, (nq,), generator=g, device=dev)
    qc = torch.randint(0, 6, (nq,), generator=g, device=dev)
    gp = torch.randint(0, nid, (ng,), generator=g2, device=dev)
    gc = torch.randint(0, 6, (ng,), generator=g2, device=dev)
    if rank == 0:                                                          # every query has a cross-camera match
        n0 = min(nq, ng)
        gp[:n0] = qp[:n0]; gc[:n0] = (qc[:n0] + 1) % 6
    ops = sharded.CudaOps()
    keys = torch.empty(nq, K, dtype=torch.int64, device=dev)
    cls = torch.empty(nq, K, dtype=torch.uint8, device=dev)
    ngood = torch.empty(nq, dtype=torch.int32, device=dev)
    dbuf = torch.empty(min(qchunk, nq), ng, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        gop = PreparedOperand(gf, args.dist_metric)
        st = None
        for q0 in range(0, nq, qchunk):
            q1 = min(nq, q0 + qchunk)
            qop = PreparedOperand(qf[q0:q1], args.dist_metric)
            dm = distance_prepared(qop, gop, out=dbuf[:q1 - q0])
            k, c, n, st = ops.partial(dm, qp[q0:q1], gp, qc[q0:q1], gc, K, lo)
            keys[q0:q1], cls[q0:q1], ngood[q0:q1] = k, c, n
        if world > 1:
            ka = torch.empty(world * nq, K, dtype=keys.dtype, device=dev)
            ca = torch.empty(world * nq, K, dtype=cls.dtype, device=dev)
            dist.all_gather_into_tensor(ka, keys); dist.all_gather_into_tensor(ca, cls)
            nall = ngood.clone(); dist.all_reduce(nall); dist.all_reduce(st, op=dist.ReduceOp.MAX)
            return ops.merge(ka.view(world, nq, K), ca.view(world, nq, K), nall, K, st)
        return ops.merge(keys.unsqueeze(0), cls.unsqueeze(0), ngood, K, st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(1, args.warmup)):
        res = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        res = step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    barrier()
    with _lib.profile(stream.cuda_stream) as prof:
        step()
    tl = prof.totals()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.cpu())
        dist.barrier()
    if rank == 0:
        pk = peaks()
        flops = 2.0 * nq * ng_total * d
        gemm_ms = tl.get('gemm_distance', (0, float('nan')))[1]
        print(json.dumps({
            'metric': 'scaled retrieval sweep: %d queries x %d gallery eval ms' % (nq, ng_total), 'value': ms, 'unit': 'ms',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(1, args.warmup), 'ms_per_step': ms,
            'higher_is_better': False, 'scaling': 'strong', 'vs_baseline': None, 'data': 'synthetic',
            'dtype': 'fp32 (bf16x3 split operands on tcgen05, fp32 accumulate)',
            'config': {'workload': 'retrieval sweep, gallery rows sharded over ranks, MARS-metric top-50 merged with NCCL',
                       'queries': nq, 'gallery': ng_total, 'dim': d, 'gallery_rows_per_gpu': ng, 'metric': args.dist_metric},
            'algorithmic_tflops_per_gpu': flops / world / (ms * 1e-3) / 1e12,
            'gemm_ms_rank0': gemm_ms,
            'gemm_tensor_pipe_frac': 6 * flops / world / (gemm_ms * 1e-3) / 1e12 / pk['bf16_sustained'],
            'kernels': {k: dict(launches=n, ms=round(t, 3)) for k, (n, t) in sorted(tl.items(), key=lambda kv: -kv[1][1])},
            'result': {'mAP': float(res[1]), 'rank1': float(res[0][0])}}))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, model, dev, rank, world, labels):
    """Same job through the host-facing API: maps in pinned host memory, H2D chunk by chunk on a copy
    stream (double buffered against the head), features back to the host, then the reference's own
    call sequence compute_distance_matrix(CPU tensors) -> .numpy() -> evaluate_rank(numpy)
    (train_vidreid_xent_htri.py:477-531)."""
    import torch.distributed as dist
    from agrl.pytorch_b200 import metrics
    J, n_host, chunk = NQ + NG, args.e2e_pool, args.e2e_chunk
    hx1, hx2, hadj = make_pool(n_host, dev, seed=7 + rank, pinned=True)
    qp, qc, gp, gc = labels
    bufs = [(torch.empty(chunk * S, C, H, W, device=dev), torch.empty(chunk * S, C, H, W, device=dev),
             torch.empty(chunk, 56, 56, device=dev)) for _ in range(2)]
    feats_host = torch.empty(J, 2 * C, pin_memory=True)
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    chunks = [(o, min(chunk, J - o)) for o in range(0, J, chunk)]
    h2d = J * BYTES_PER_TRACKLET - J * 2 * C * 4 + J * 2 * C * 4 + NQ * NG * 4 + (NQ + NG) * 16    # maps+adj, features, distmat, labels
    d2h = J * 2 * C * 4 + NQ * NG * 4 + 51 * 8

    def one_step():
        ready = [torch.cuda.Event() for _ in chunks]
        freed = [torch.cuda.Event() for _ in chunks]
        for i, (off, n) in enumerate(chunks):
            b1, b2, ba = bufs[i % 2]
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(freed[i - 2])
                src = (off % n_host)
                if src + n > n_host:
                    src = 0
                b1[:n * S].copy_(hx1[src * S:(src + n) * S], non_blocking=True)
                b2[:n * S].copy_(hx2[src * S:(src + n) * S], non_blocking=True)
                ba[:n].copy_(hadj[src:src + n], non_blocking=True)
                ready[i].record(copy_stream)
            main.wait_event(ready[i])
            f = model.head(b1[:n * S], b2[:n * S], ba[:n], S)
            feats_host[off:off + n].copy_(f, non_blocking=True)
            freed[i].record(main)
        main.synchronize()
        d = metrics.compute_distance_matrix(feats_host[:NQ], feats_host[NQ:], args.dist_metric)   # CPU tensors
        return metrics.evaluate_rank(d.numpy(), qp, gp, qc, gc, use_metric_mars=True)

    steps = max(1, min(args.steps, args.e2e_steps))
    one_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = one_step()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.cpu())
    del hx1, hx2, bufs
    return {'value': world * J / dt, 'unit': 'tracklets/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
            'ms_per_step': dt * 1e3, 'steps': steps, 'h2d_gb_per_s': h2d / dt / 1e9,
            'note': 'layer4 maps start in pinned host memory (16.8 MB/tracklet over PCIe: the step is bound by the host link, see h2d_gb_per_s); eval via CPU-tensor / numpy API'}


# ------------------------------------------------------------------------------------------------
# same-box GPU comparator: the reference's head / distance lines as stock PyTorch modules
# on the same B200 (library kernels: cuDNN/ATen pooling, cuBLAS fp32 GEMMs, elementwise ATen ops)
# ------------------------------------------------------------------------------------------------
def build_eager(dev, weights):
    """(head, distance) callables: the reference's module code for this path, written with the stock torch.nn modules
    it is built from (vmgn.py:104-123,142-172,237-268,270-278,296-321; distance.py:59-73), on device `dev`."""
    import torch.nn as nn
    import torch.nn.functional as F
    pool3d = nn.AdaptiveAvgPool3d(1)
    part_pools = [nn.AdaptiveAvgPool2d((k, 1)) for k in (4, 2, 1)]

    def bn(prefix):
        m = nn.BatchNorm1d(C).to(dev).eval()
        for k in ('weight', 'bias', 'running_mean', 'running_var'):
            getattr(m, k).data.copy_(weights[prefix + '.' + k])
        return m
    g_neck, a_neck = bn('global_bottleneck'), bn('att_bottleneck')
    layers = []
    for i in range(2):
        lin = nn.Linear(C, C, bias=False).to(dev)
        lin.weight.data.copy_(weights['graph_layers.%d.linear.weight' % i])
        layers.append((lin, bn('graph_layers.%d.bn' % i)))
    act = nn.LeakyReLU(0.1)

    def layer(lin, norm, f, a):
        h = lin(f)
        a = F.normalize(a, p=1, dim=2)
        sq = torch.pow(f, 2).sum(dim=2)
        d = sq.unsqueeze(1) + sq.unsqueeze(2)
        d -= 2 * torch.bmm(f, f.transpose(1, 2))
        g = F.normalize(2 / (d.clamp(1e-12).sqrt().exp() + 1), p=1, dim=2)
        hp = torch.bmm((a + g) / 2, h)
        hp = act(norm(hp.view(-1, C)).view(h.shape))
        return 0.9 * f + 0.1 * hp

    def head(m1, m2, a):
        b = a.shape[0]
        g_bn = g_neck(pool3d(m1.view(b, S, C, H, W).transpose(1, 2).contiguous()).view(b, -1))
        f = torch.cat([p(m2).view(b, S, C, k) for p, k in zip(part_pools, (4, 2, 1))], dim=3)
        f = f.transpose(2, 3).contiguous().view(b, S * 7, C)
        for lin, norm in layers:
            f = layer(lin, norm, f, a)
        f = f.view(b, S, 7, C)
        att = F.normalize(f.norm(p=2, dim=3, keepdim=True), p=1, dim=1)
        return torch.cat([g_bn, a_neck(f.mul(att).sum(dim=1).mean(dim=1))], dim=1)

    def distance(q, g):
        d = torch.pow(q, 2).sum(dim=1, keepdim=True).expand(q.shape[0], g.shape[0]) + \
            torch.pow(g, 2).sum(dim=1, keepdim=True).expand(g.shape[0], q.shape[0]).t()
        return d.addmm(q, g.t(), beta=1, alpha=-2)

    return head, distance


def torch_eager_gpu(args, dev, weights, x1, x2, adj, feats):
    """What the reference's own module code costs on this GPU, left in PyTorch's default fp32 mode (TF32 matmuls
    off).  Not a product path and not the oracle: a reported comparator, timed on a bounded sample of the resident
    pool.  The ranking has no GPU form in the reference (numpy / Cython), so it is not part of it."""
    head, distance = build_eager(dev, weights)
    n = min(args.eager_sample, x1.shape[0] // S)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    chunk = min(n, 64)                                     # the transposed copy of 64 tracklets is 0.5 GB
    with torch.no_grad():
        def head_sample():
            for off in range(0, n - chunk + 1, chunk):
                head(x1[off * S:(off + chunk) * S], x2[off * S:(off + chunk) * S], adj[off:off + chunk])
        done = (n // chunk) * chunk
        head_ms = timed(head_sample, 3) / done
        dist_ms = timed(lambda: distance(feats[:NQ], feats[NQ:]), 5)
    J = NQ + NG
    return {'what': 'the reference head / distance lines as stock torch.nn modules on this GPU, default fp32 (TF32 off), '
                    'maps resident; ranking excluded (numpy / Cython only in the reference)',
            'sample': 'head: %d of %d tracklets in batches of %d, extrapolated; distance %dx%dx%d in full' % (done, J, chunk, NQ, NG, 2 * C),
            'head_ms_per_tracklet': head_ms, 'head_tracklets_per_s': 1e3 / head_ms, 'head_ms_per_pass': head_ms * J,
            'distance_ms': dist_ms}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the same path, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference(args, steps, warmup):
    from oracle import head as ohead, distance as odist, rank as orank
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    J, n_head = NQ + NG, args.cpu_head_sample
    weights = make_head_weights()
    g = torch.Generator().manual_seed(3)
    x1 = torch.randn(n_head * S, C, H, W, generator=g).clamp_(min=0)
    x2 = torch.randn(n_head * S, C, H, W, generator=g).clamp_(min=0)
    from agrl.pytorch_b200 import synthetic as synth
    adj = synth.pose_adjacency(n_head, S, 7, seed=3)
    qp, qc, gp, gc = make_labels(0, 1)
    feats = torch.randn(J, 2 * C, generator=g)
    have_ref_cy = orank.reference_rank_cy() is not None

    def one():
        t0 = time.perf_counter()
        with torch.no_grad():
            ohead.head_forward(x1, x2, adj, weights)
        t1 = time.perf_counter()
        d = odist.distance_matrix(feats[:NQ], feats[NQ:], args.dist_metric)
        t2 = time.perf_counter()
        orank.mars_port(d.numpy(), qp, gp, qc, gc, 50)
        t3 = time.perf_counter()
        cy = None
        if have_ref_cy:
            orank.reference_evaluate_cy(d.numpy(), qp, gp, qc, gc, 50, stable=False)
            cy = time.perf_counter() - t3
        return (t1 - t0) / n_head, t2 - t1, t3 - t2, cy

    for _ in range(warmup):
        one()
    runs = [one() for _ in range(max(1, steps))]
    per_tracklet = min(r[0] for r in runs)
    t_dist = min(r[1] for r in runs)
    t_rank = min(r[2] for r in runs)
    t_cy = min(r[3] for r in runs) if have_ref_cy else None
    job_s = J * per_tracklet + t_dist + t_rank
    value = J / job_s
    base = {'value': value, 'unit': 'tracklets/s', 'cores': cores, 'kind': 'port',
            'sample': 'head: %d of %d tracklets timed with the torch-CPU restatement (%.2f ms/tracklet, extrapolated); '
                      'distance 1980x9330x4096 (%.0f ms) and MARS-metric ranking (C restatement, %.0f ms) in full'
                      % (n_head, J, per_tracklet * 1e3, t_dist * 1e3, t_rank * 1e3),
            'head_ms_per_tracklet': per_tracklet * 1e3, 'distance_ms': t_dist * 1e3, 'rank_mars_ms': t_rank * 1e3,
            'rank_cy_reference_ms': None if t_cy is None else t_cy * 1e3}
    return {'cpu_baseline': base, 'job_s': job_s, 'eval_ms': (t_dist + t_rank) * 1e3}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_reference(args, steps=args.steps, warmup=args.warmup)
    base = r['cpu_baseline']
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'tracklets/s',
            'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': r['job_s'] * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': 'MARS-shaped test pass on the host CPU: graph head (sampled, extrapolated to 11310 '
                                   'tracklets) + 1980x9330 %s distance + MARS-metric CMC/mAP' % args.dist_metric},
            'eval_ms': r['eval_ms'], 'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': 'tracklets/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--pool', type=int, default=882, help='tracklets per head call = resident input pool (14.8 GB at 882)')
    ap.add_argument('--dist-metric', default='euclidean', choices=['euclidean', 'cosine'])
    ap.add_argument('--e2e-pool', type=int, default=128)
    ap.add_argument('--e2e-chunk', type=int, default=64)
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--cpu-head-sample', type=int, default=32)
    ap.add_argument('--workload', default='mars', choices=['mars', 'sweep'])
    ap.add_argument('--sweep-queries', type=int, default=10000)
    ap.add_argument('--sweep-gallery', type=int, default=1000000)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-tuned', action='store_true', help='skip the extra head pass with every opt-in switched on')
    ap.add_argument('--no-lowrank', action='store_true', help='skip the extra head pass with the low-rank first layer')
    ap.add_argument('--no-eager', action='store_true', help='skip the stock-PyTorch-on-this-GPU comparator (SURVEY 8d)')
    ap.add_argument('--eager-sample', type=int, default=256, help='tracklets of the pool the comparator head is timed on')
    ap.add_argument('--no-fast-mode', action='store_true', help='skip the extra head pass with the fp16 single-plane GEMM')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.workload == 'sweep':
        run_sweep(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
