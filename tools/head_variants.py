"""Head-pipeline variants on one resident input pool (one process, options switched at run time).
usage: python tools/head_variants.py POOL "k=v,k=v" ...   (keys: tma stages hint split lr pair nhwc call gb)
`gb` = number of graph layers of the model (0: pooling + attention only -> the pooling kernels run practically alone).
`call` = tracklets per agrl_head_forward_dev call (default POOL).  Prints head ms per 11310 tracklets.
HV_LIB=path: load that build of the library (tools/build_alt.sh) -- A/B comparisons on ONE box, boxes differ by 3 %."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from agrl.pytorch_b200 import _lib
if os.environ.get('HV_LIB'):                      # diagnostic A/B build (tools/build_alt.sh) instead of the product library
    _lib.LIB_PATH = os.path.abspath(os.environ['HV_LIB'])

KEYS = {'tma': 'pool_tma', 'stages': 'pool_stages', 'hint': 'pool_l2_hint', 'lr': 'head_lowrank', 'pair': 'gemm_pair'}      # module attributes
DEFAULTS = {'tma': 1, 'stages': 0, 'hint': 1, 'lr': 1, 'pair': 1}


def main():
    pool_n = int(sys.argv[1])
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    _lib.require_device()
    model = bench.make_model(dev, bench.make_head_weights())
    models_by_gb = {2: model}

    def model_for(gb):
        if gb not in models_by_gb:
            from agrl.pytorch_b200 import models
            m = models.init_model('vmgn', num_classes=625, loss={'xent', 'htri'}, last_stride=1, num_split=4, num_gb=gb,
                                  num_scale=1, pyramid_part=True, use_pose=True, learn_graph=True, pretrained=False)
            sd = m.state_dict()
            for k, v in bench.make_head_weights().items():
                if k in sd:
                    sd[k].copy_(v)
            for name in ('graph_layers', 'global_bottleneck', 'att_bottleneck'):
                getattr(m, name).to(dev)
            models_by_gb[gb] = m.eval()
        return models_by_gb[gb]
    x1, x2, adj = bench.make_pool(pool_n, dev, seed=1)
    J, S = int(os.environ.get('HV_TRACKLETS', bench.NQ + bench.NG)), bench.S
    feats = torch.empty(J, 2 * bench.C, device=dev)
    stream = torch.cuda.current_stream(dev)
    defaults = dict(DEFAULTS)
    for spec in sys.argv[2:]:
        cfg = dict(defaults)
        cfg['call'] = pool_n
        for kv in spec.split(','):
            if kv:
                k, v = kv.split('=')
                cfg[k] = int(v)
        model = model_for(cfg.get('gb', 2))
        for k, name in KEYS.items():
            setattr(model, name, cfg[k] if k == 'stages' else bool(cfg[k]))
        model.head_split = cfg.get('split', 2)
        call = min(cfg['call'], pool_n)
        chunks = [(o, min(call, J - o)) for o in range(0, J, call)]

        if cfg.get('nhwc') and 'cl' not in globals():
            globals()['cl'] = (x1.contiguous(memory_format=torch.channels_last), x2.contiguous(memory_format=torch.channels_last))
        m1, m2 = globals()['cl'] if cfg.get('nhwc') else (x1, x2)

        def head_pass():
            for off, n in chunks:
                model.head(m1[:n * S], m2[:n * S], adj[:n], S, out=feats[off:off + n])
        try:
            with torch.no_grad():
                for _ in range(2):
                    head_pass()
                torch.cuda.synchronize()
                # HV_REPS timed passes (default 3); HV_CLOCKS=1 samples nvidia-smi clocks / power / throttle reasons meanwhile
                reps = int(os.environ.get('HV_REPS', '3'))
                sampler = bench.ClockSampler(0) if os.environ.get('HV_CLOCKS') else None
                if sampler:
                    sampler.start()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
                for i in range(reps):
                    e[i].record(stream); head_pass()
                e[reps].record(stream)
                torch.cuda.synchronize()
                clocks = sampler.stop() if sampler else None
                ms = [e[i].elapsed_time(e[i + 1]) for i in range(reps)]
                with _lib.profile(stream.cuda_stream) as prof:
                    head_pass()
                tot = {k: round(t, 3) for k, (n, t) in prof.totals().items()}
                energy = None
                if os.environ.get('HV_ENERGY'):                  # NVML energy counter over >= HV_ENERGY seconds of passes
                    cx = type('Cx', (), {'local': 0, 'dev': dev})()
                    energy = bench.energy_of(cx, head_pass, min_seconds=float(os.environ['HV_ENERGY']))
            best = min(ms)
            print(json.dumps({'spec': spec, 'head_ms': round(best, 3), 'all': [round(m, 3) for m in ms[:6]], 'clocks': clocks,
                              'ktracklets_s': round(J / best, 1),
                              'hbm_frac': round(J * bench.BYTES_PER_TRACKLET / (best * 1e-3) / 1e9 / 6545.9, 4),
                              'checksum': float(feats.double().sum()), 'kernels': tot, 'energy': energy}), flush=True)
        except Exception as ex:  # noqa
            print(json.dumps({'spec': spec, 'error': repr(ex)}), flush=True)


if __name__ == '__main__':
    main()
