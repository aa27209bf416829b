"""Graph-head oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

torch-CPU restatement (fp32 by default, fp64 on request) of the eval-mode VMGN head, i.e. what
the reference computes after ``featuremaps()``:

    pyramid part pooling + global pooling      torchreid/models/vmgn.py:299-308
    GraphLayer (x num_gb)                      torchreid/models/vmgn.py:104-123, :142-172
    temporal attention fusion                  torchreid/models/vmgn.py:270-278
    part mean, BN necks, concat                torchreid/models/vmgn.py:317-321

Inputs are the two layer4 feature maps (B*S, C, h, w), the pose adjacency (B, S*P, S*P) and a
weight dict with the reference's state_dict names.  Pinned by tests/test_oracle_head.py against
tests/golden/head_*.npz (outputs of the reference's GSTA.forward lines run on the same maps).
"""
import torch

BN_EPS = 1e-5          # nn.BatchNorm1d default (vmgn.py:94, :238, :264)
L1_EPS = 1e-12         # F.normalize default eps (vmgn.py:157, :162, :276)
LEAKY = 0.1            # nn.LeakyReLU(0.1)          (vmgn.py:95)
GAMMA = 0.1            # GraphLayer gamma           (vmgn.py:74, :172)


def split_list(num_split=4, pyramid_part=True):
    """calc_splits (utils/reidtools.py:13-15): divisors of num_split, descending -> [4, 2, 1]."""
    if not pyramid_part:
        return [num_split]
    return [n for n in range(num_split, 0, -1) if num_split % n == 0]


def _bn_eval(x, w, prefix):
    mean, var = w[prefix + '.running_mean'], w[prefix + '.running_var']
    return (x - mean) / torch.sqrt(var + BN_EPS) * w[prefix + '.weight'] + w[prefix + '.bias']


def pool_nodes(x4_2, B, S, splits):
    """Strip means -> node tensor (B, S*P, C); node index v = s*P + p, strips ordered by pyramid level."""
    BS, C, h, w = x4_2.shape
    parts = []
    for n in splits:
        rows = h // n                                   # adaptive pooling with h % n == 0
        parts.append(x4_2.reshape(B, S, C, n, rows * w).mean(dim=4))      # (B,S,C,n)
    v = torch.cat(parts, dim=3)                         # (B,S,C,P)
    return v.permute(0, 1, 3, 2).reshape(B, S * sum(splits), C)


def affinity(x):
    """2 / (exp(||xi-xj||) + 1) with the Gram-form squared distance and the 1e-12 clamp (vmgn.py:114-120)."""
    sq = (x * x).sum(dim=2)
    d2 = sq.unsqueeze(1) + sq.unsqueeze(2) - 2.0 * torch.bmm(x, x.transpose(1, 2))
    d = d2.clamp(min=1e-12).sqrt()
    return 2.0 / (d.exp() + 1.0)


def _l1_rows(m):
    return m / m.abs().sum(dim=2, keepdim=True).clamp(min=L1_EPS)


def graph_layer(x, adj, w, prefix, use_pose=True, learn_graph=True):
    h = x @ w[prefix + '.linear.weight'].t()
    a = _l1_rows(adj) if use_pose else adj
    if learn_graph:
        g = _l1_rows(affinity(x))
        if use_pose:
            g = (a + g) / 2
    else:
        g = a
    hp = torch.bmm(g, h)
    B, V, C = hp.shape
    hp = _bn_eval(hp.reshape(B * V, C), w, prefix + '.bn').reshape(B, V, C)
    hp = torch.where(hp >= 0, hp, hp * LEAKY)
    return (1 - GAMMA) * x + GAMMA * hp


def attention_fuse(f):
    """(B,S,P,C) -> (B,P,C): weights = L1-normalised (over S) L2 norms (vmgn.py:276-277)."""
    nrm = f.pow(2).sum(dim=3, keepdim=True).sqrt()
    att = nrm / nrm.abs().sum(dim=1, keepdim=True).clamp(min=L1_EPS)
    return (f * att).sum(dim=1)


def head_forward(x4_1, x4_2, adj, w, S=8, num_split=4, pyramid_part=True, num_gb=2,
                 use_pose=True, learn_graph=True, dtype=torch.float32, return_nodes=False):
    """Eval-mode VMGN head: (B*S,C,h,w) x2 + (B,V,V) -> (B, 2C)."""
    x4_1, x4_2, adj = x4_1.to(dtype), x4_2.to(dtype), adj.to(dtype)
    w = {k: v.to(dtype) for k, v in w.items() if v.is_floating_point()}
    BS, C, h, wd = x4_1.shape
    B = BS // S
    splits = split_list(num_split, pyramid_part)
    P = sum(splits)

    g_f = x4_1.reshape(B, S, C, h * wd).permute(0, 2, 1, 3).reshape(B, C, S * h * wd).mean(dim=2)
    g_bn = _bn_eval(g_f, w, 'global_bottleneck')

    f = pool_nodes(x4_2, B, S, splits)
    nodes0 = f
    for i in range(num_gb):
        f = graph_layer(f, adj, w, 'graph_layers.%d' % i, use_pose, learn_graph)
    fused = attention_fuse(f.reshape(B, S, P, C))
    att_bn = _bn_eval(fused.mean(dim=1), w, 'att_bottleneck')
    out = torch.cat([g_bn, att_bn], dim=1)
    if return_nodes:
        return out, nodes0, f
    return out


HEAD_KEYS = (
    ['global_bottleneck.' + k for k in ('weight', 'bias', 'running_mean', 'running_var')] +
    ['att_bottleneck.' + k for k in ('weight', 'bias', 'running_mean', 'running_var')]
)


def head_keys(num_gb=2):
    keys = list(HEAD_KEYS)
    for i in range(num_gb):
        keys.append('graph_layers.%d.linear.weight' % i)
        keys += ['graph_layers.%d.bn.%s' % (i, k) for k in ('weight', 'bias', 'running_mean', 'running_var')]
    return keys
