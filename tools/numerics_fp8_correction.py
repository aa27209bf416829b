"""CPU numerics study for DESIGN.md section 10 item 2 (no GPU, no product code): how accurate is a graph-layer GEMM
that computes   y.w  ~=  fp16(y).fp16(w)  +  e4m3(r_y).e4m3(w)  +  e4m3(y).e4m3(r_w)      (r = x - fp16(x))
i.e. ONE fp16 tensor-core product plus TWO 8-bit products (twice the rate each: two pass-equivalents instead of the
three of the bf16x2 split), with power-of-two pre-scaling per tracklet (y) / per layer (w) as the fp16 mode does?
Compared on the reference goldens (tests/golden/head_*.npz) with: fp32 (reference), bf16x2 (3 products, the default),
fp16x1 (1 product, the opt-in fast mode).  Emulation: operands rounded with torch dtypes, products / sums in fp64.

usage: python tools/numerics_fp8_correction.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import head as ohead                      # noqa: E402  (a tool, not the product: may use the oracle)
from test_oracle_head import regenerate               # noqa: E402
from conftest import GOLDEN, golden_files             # noqa: E402


def pow2_scale(t, target):
    """power of two s with max|t| * s just below `target` (per leading index)"""
    m = t.abs().amax(dim=tuple(range(1, t.dim())), keepdim=True).clamp(min=1e-30)
    return torch.exp2(torch.floor(torch.log2(target / m)))


def rnd(t, dtype):
    return t.to(dtype).to(torch.float64)


def gemm_variants(y, w):
    """y (B, V, C), w (C, C) fp32 -> dict name -> (B, V, C) fp64 products y.w^T under each operand rounding"""
    y64, w64 = y.double(), w.double()
    out = {'fp32 operands (exact products)': y64 @ w64.t()}
    # bf16 x2 planes, 3 products
    y0 = rnd(y, torch.bfloat16); y1 = rnd((y64 - y0).float(), torch.bfloat16)
    w0 = rnd(w, torch.bfloat16); w1 = rnd((w64 - w0).float(), torch.bfloat16)
    out['bf16x2, 3 products (default)'] = y0 @ w0.t() + y0 @ w1.t() + y1 @ w0.t()
    # fp16 x1, pow2 scaled
    sy = pow2_scale(y, 2.0 ** 14).double(); sw = pow2_scale(w.unsqueeze(0), 2.0 ** 14).double()[0]
    yh = rnd((y64 * sy).float(), torch.float16) / sy
    wh = rnd((w64 * sw).float(), torch.float16) / sw
    out['fp16x1, 1 product (fast mode)'] = yh @ wh.t()
    # fp16 main + two e4m3 corrections
    f8 = torch.float8_e4m3fn
    ry, rw = y64 - yh, w64 - wh
    def q8(t, lead):
        s = pow2_scale(t.float() if lead else t.float().unsqueeze(0), 256.0).double()
        s = s if lead else s[0]
        return rnd((t * s).float(), f8) / s
    ry8, y8 = q8(ry, True), q8(y64, True)
    rw8, w8 = q8(rw, False), q8(w64, False)
    out['fp16 + 2x e4m3 corrections (2 pass-equivalents)'] = yh @ wh.t() + ry8 @ w8.t() + y8 @ rw8.t()
    # the same with scales TIED to the fp16 product's (so that all three products can share one TMEM accumulator):
    # residuals x 2^6, full operands x 2^-6 on top of the fp16 scaling (max 2^14 -> 2^8 <= 448, half-ulp 2^2 -> 2^8)
    def q8_tied(t, s, shift):
        return rnd((t * s * 2.0 ** shift).float(), f8) / (s * 2.0 ** shift)
    out['same, scales tied to the fp16 product (one accumulator)'] = (
        yh @ wh.t() + q8_tied(ry, sy, 6) @ q8_tied(w64, sw, -6).t() + q8_tied(y64, sy, -6) @ q8_tied(rw, sw, 6).t())
    return out


def main():
    rows = {}
    for fname in golden_files('head_'):
        g = np.load(os.path.join(GOLDEN, fname))
        x1, x2, adj, wts = regenerate(g)
        B = int(g['B'])
        x = ohead.pool_nodes(x2, B, 8, [4, 2, 1])
        for layer in range(2):
            prefix = 'graph_layers.%d' % layer
            a = ohead._l1_rows(adj)
            gr = (a + ohead._l1_rows(ohead.affinity(x))) / 2
            y = torch.bmm(gr, x)                                    # the GEMM's left operand in this library's order
            w = wts[prefix + '.linear.weight']
            prods = gemm_variants(y, w)
            ref = prods['fp32 operands (exact products)']
            for name, h in prods.items():
                # through the layer's epilogue: what reaches the next stage
                def layer_out(hh):
                    hp = ohead._bn_eval(hh.reshape(-1, hh.shape[-1]), {k: v.double() for k, v in wts.items()}, prefix + '.bn')
                    hp = hp.reshape(hh.shape)
                    hp = torch.where(hp >= 0, hp, hp * ohead.LEAKY)
                    return (1 - ohead.GAMMA) * x.double() + ohead.GAMMA * hp
                e_prod = float((h - ref).norm() / ref.norm())
                o, o_ref = layer_out(h), layer_out(ref)
                e_out = float((o - o_ref).abs().max() / o_ref.abs().max())
                rows.setdefault(name, []).append((e_prod, e_out))
            x = ohead.graph_layer(x, adj, wts, prefix)
    print('%-52s %14s %14s' % ('operand rounding', 'product (norm)', 'layer out (max)'))
    for name, v in rows.items():
        print('%-52s %14.2e %14.2e' % (name, max(a for a, _ in v), max(b for _, b in v)))


if __name__ == '__main__':
    main()
