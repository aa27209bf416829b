"""Distance-matrix oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

torch-CPU restatement of torchreid/metrics/distance.py:
    squared euclidean  ||a||^2 + ||b||^2 - 2 a.b^T, no clamp, no sqrt      (:59-73)
    cosine             1 - a_hat . b_hat^T, x_hat = x / max(||x||_2, 1e-12)  (:76-89)
plus the argument checks of compute_distance_matrix (:39-54).  ``dtype=torch.float64`` gives the
"true" value used to judge which of two fp32 results is closer.
Pinned by tests/test_oracle_distance.py against tests/golden/distance_*.npz.
"""
import torch


def euclidean_squared(a, b):
    na = (a * a).sum(dim=1, keepdim=True)            # (m,1)
    nb = (b * b).sum(dim=1, keepdim=True).t()        # (1,n)
    return (na + nb) - 2.0 * (a @ b.t())


def cosine(a, b, eps=1e-12):
    an = a / a.norm(p=2, dim=1, keepdim=True).clamp(min=eps)
    bn = b / b.norm(p=2, dim=1, keepdim=True).clamp(min=eps)
    return 1.0 - an @ bn.t()


def distance_matrix(a, b, metric='euclidean', dtype=torch.float32):
    assert isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor)
    assert a.dim() == 2, 'Expected 2-D tensor, but got {}-D'.format(a.dim())
    assert b.dim() == 2, 'Expected 2-D tensor, but got {}-D'.format(b.dim())
    assert a.size(1) == b.size(1)
    a, b = a.detach().cpu().to(dtype), b.detach().cpu().to(dtype)
    if metric == 'euclidean':
        return euclidean_squared(a, b)
    if metric == 'cosine':
        return cosine(a, b)
    raise ValueError('Unknown distance metric: {}. '
                     'Please choose either "euclidean" or "cosine"'.format(metric))
