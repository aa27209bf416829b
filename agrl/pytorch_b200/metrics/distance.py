"""Mirror of torchreid/metrics/distance.py: ``compute_distance_matrix`` with the reference's
signature, checks and error behaviour (distance.py:11-56), computed by the tcgen05 GEMM of
libagrl_b200 (csrc/distance.cu).

CPU tensors in -> CPU tensor out (the reference's test() passes CPU tensors,
train_vidreid_xent_htri.py:477,507,520); the work still happens on the GPU, through the
host-buffer entry point.  CUDA tensors in -> CUDA tensor out with no host round trip.
"""
import torch

from .. import _lib

_METRICS = {'euclidean': _lib.METRIC_EUCLIDEAN, 'cosine': _lib.METRIC_COSINE}
_ws_cache = {}


def _workspace(dev, nbytes):
    ws = _ws_cache.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _ws_cache[dev] = ws
    return ws


def compute_distance_matrix(input1, input2, metric='euclidean', split=_lib.SPLIT_BF16X3):
    """A wrapper function for computing distance matrix (distance.py:11).

    Args:
        input1 (torch.Tensor): 2-D feature matrix.
        input2 (torch.Tensor): 2-D feature matrix.
        metric (str, optional): "euclidean" or "cosine".  Default is "euclidean".
    Returns:
        torch.Tensor: distance matrix (squared euclidean, or 1 - cosine similarity).
    """
    # check input (distance.py:39-44)
    assert isinstance(input1, torch.Tensor)
    assert isinstance(input2, torch.Tensor)
    assert input1.dim() == 2, 'Expected 2-D tensor, but got {}-D'.format(input1.dim())
    assert input2.dim() == 2, 'Expected 2-D tensor, but got {}-D'.format(input2.dim())
    assert input1.size(1) == input2.size(1)
    if metric not in _METRICS:
        raise ValueError(
            'Unknown distance metric: {}. '
            'Please choose either "euclidean" or "cosine"'.format(metric)
        )
    lib = _lib.require_device()
    m, n, d = input1.size(0), input2.size(0), input1.size(1)
    if input1.is_cuda != input2.is_cuda:
        raise RuntimeError('Expected all tensors to be on the same device')
    if not input1.is_cuda:
        a = input1.detach().to(torch.float32).contiguous()
        b = input2.detach().to(torch.float32).contiguous()
        out = torch.empty(m, n, dtype=torch.float32)
        if m and n:
            _lib.check(lib.agrl_distance_host(a.data_ptr(), b.data_ptr(), out.data_ptr(), m, n, d,
                                              _METRICS[metric], split))
        return out
    dev = input1.device
    a = input1.detach().to(torch.float32)
    b = input2.detach().to(device=dev, dtype=torch.float32)
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    out = torch.empty(m, n, dtype=torch.float32, device=dev)
    if m == 0 or n == 0:
        return out
    with torch.cuda.device(dev):
        wsb = lib.agrl_distance_workspace_bytes(m, n, d, split)
        ws = _workspace(dev, wsb)
        _lib.check(lib.agrl_distance_dev(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                         out.data_ptr(), out.stride(0), m, n, d, _METRICS[metric], split,
                                         ws.data_ptr(), wsb, torch.cuda.current_stream(dev).cuda_stream))
    return out
