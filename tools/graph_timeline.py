"""clock64 timeline of graph_kernel_tc's phases for a few tracklets of one 882-tracklet call (full graph layers only).
Needs a diagnostic build of the library (the product build has no instrumentation):
    AGRL_NVCC_EXTRA=-DAGRL_TIMELINE python -m agrl.pytorch_b200.build --force
    python tools/graph_timeline.py
    python -m agrl.pytorch_b200.build --force            # back to the product library
Prints, per sampled CTA, the cycles of the Gram / graph-build / message-passing phases and per-block time stamps
(tile landed, operand buffer free, converted, MMAs committed, accumulator ready, accumulator drained)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from agrl.pytorch_b200 import _lib
dev = torch.device('cuda', 0); torch.cuda.set_device(dev)
lib = _lib.require_device()
raw = ctypes.CDLL(_lib.LIB_PATH)
model = bench.make_model(dev, bench.make_head_weights())
model.head_lowrank = False          # both graph layers as full layers (the low-rank first layer has no Y phase)
n = 882
x1, x2, adj = bench.make_pool(n, dev, seed=1)
feats = torch.empty(n, 4096, device=dev)
buf = torch.zeros(16 * 256, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3):
        model.head(x1, x2, adj, 8, out=feats)
    torch.cuda.synchronize()
    raw.agrl_timeline_set.argtypes = [ctypes.c_void_p]
    assert raw.agrl_timeline_set(buf.data_ptr()) == 0
    model.head(x1, x2, adj, 8, out=feats)
    torch.cuda.synchronize()
    raw.agrl_timeline_set(None)
t = buf.cpu().numpy().reshape(16, 256)
for c in range(10):
    r = t[c]
    if r[0] == 0: continue
    t0 = r[0]
    print('CTA %d: gram %d  build %d  Y %d  total %d' % (c * 97 + 5, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[3] - r[0]))
    conv_start = r[48:64] - t0; conv_done = r[32:48] - t0; issue = r[16:32] - t0; acc_ready = r[64:80] - t0; ld_done = r[80:96] - t0
    print('  conv_start', (conv_start - (r[2]-t0)).tolist())
    print('  conv_done ', (conv_done - (r[2]-t0)).tolist())
    print('  committed ', (issue - (r[2]-t0)).tolist())
    print('  acc_ready ', (acc_ready - (r[2]-t0)).tolist())
    print('  ld_done   ', (ld_done - (r[2]-t0)).tolist())
    print('  G landed  ', (r[96:128] - t0).tolist())
    print('  G buffree ', (r[192:224] - t0).tolist())
    print('  G convdone', (r[128:160] - t0).tolist())
    print('  G commit  ', (r[160:192] - t0).tolist())
