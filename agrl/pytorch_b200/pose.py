"""Pose-guided adjacency on the GPU -- mirror of the graph builder in the reference's loader
(torchreid/dataset_loader.py: generate_graph :218-343, adj_graph :345-388).

The loader-side work shrinks to looking the detections up (``pack_keypoints``: string handling only, no python sets,
no itertools.permutations); ``part_masks`` turns them into three membership masks per tracklet on the device
(24 bytes -- the compact wire format ``VMGN.head`` / ``VMGN.forward`` accept directly as ``adj``), and
``expand_adjacency`` / ``generate_graph`` give the reference's dense matrix when a caller wants it.
"""
import numpy as np
import torch

from . import _lib

__all__ = ['pose_key', 'pack_keypoints', 'part_masks', 'expand_adjacency', 'generate_graph']

PARTS_PER_FRAME = 7                    # calc_splits(4) = [4, 2, 1] (utils/reidtools.py:13-15)


def pose_key(path):
    """Key of an image in the pose dict (dataset_loader.py:246-255)."""
    if 'ilids-vid' in path:
        return path.split('/')[-1]
    if 'prid2011' in path:
        return '-'.join(path.split('/')[-3:])
    if 'mars' in path:
        return path.split('/')[-1]
    if 'duke' in path:
        return '-'.join(path.split('/')[-3:])
    raise ValueError('{} is not acceptable'.format(path))


def pack_keypoints(im_paths, im_sizes, poses):
    """One tracklet: (S, 18, 3) float64 detections, (S,) float64 image heights, (S,) uint8 found-flags.
    A frame whose key is missing from ``poses`` (or whose entry is unusable) is flagged invalid, as the reference's
    bare ``except`` leaves it empty (dataset_loader.py:337-338)."""
    S = len(im_paths)
    kp = np.zeros((S, 18, 3), np.float64)
    heights = np.zeros(S, np.float64)
    valid = np.zeros(S, np.uint8)
    for s, (path, size) in enumerate(zip(im_paths, im_sizes)):
        key = pose_key(path)
        heights[s] = size[1]
        try:
            entry = np.asarray(poses[key], np.float64)
            kp[s] = entry[:18, :3]
            valid[s] = 1
        except Exception:
            pass
    return kp, heights, valid


def part_masks(keypoints, heights, valid=None, threshold=0.1, num_split=4, device=None):
    """(B, S, 18, 3), (B, S)[, (B, S)] -> int64 (B, 3) membership masks on the device (bit s*7 + strip)."""
    lib = _lib.require_device()
    kp = torch.as_tensor(keypoints, dtype=torch.float64)
    if device is None:
        device = kp.device if kp.is_cuda else torch.device('cuda', torch.cuda.current_device())
    kp = kp.to(device).contiguous()
    assert kp.dim() == 4 and tuple(kp.shape[2:]) == (18, 3), 'keypoints must be (B, S, 18, 3)'
    B, S = kp.shape[:2]
    h = torch.as_tensor(heights, dtype=torch.float64).to(device).contiguous()
    assert tuple(h.shape) == (B, S)
    v = None
    if valid is not None:
        v = torch.as_tensor(valid).to(device=device, dtype=torch.uint8).contiguous()
        assert tuple(v.shape) == (B, S)
    masks = torch.empty(B, 3, dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.agrl_pose_part_masks_dev(kp.data_ptr(), h.data_ptr(), v.data_ptr() if v is not None else None,
                                                B, S, num_split, float(threshold), masks.data_ptr(),
                                                torch.cuda.current_stream(device).cuda_stream))
    return masks


def expand_adjacency(masks, seq_len):
    """int64 (B, 3) masks -> the reference's dense (B, V, V) fp32 graph (binary, symmetric, zero diagonal)."""
    lib = _lib.require_device()
    if not masks.is_cuda:
        raise RuntimeError('agrl.pytorch_b200 has no CPU path: move the masks to a B200')
    masks = masks.contiguous()
    B, V = masks.size(0), seq_len * PARTS_PER_FRAME
    adj = torch.empty(B, V, V, dtype=torch.float32, device=masks.device)
    with torch.cuda.device(masks.device):
        _lib.check(lib.agrl_pose_adjacency_dev(masks.data_ptr(), B, V, adj.data_ptr(),
                                               torch.cuda.current_stream(masks.device).cuda_stream))
    return adj


def generate_graph(ims, im_paths, im_sizes, poses, num_split, num_parts, num_scale, pyramid_part, threshold=0.1):
    """Signature of the reference's generate_graph (dataset_loader.py:218-219) for one tracklet; returns the (V, V)
    fp32 adjacency as a CPU tensor like the reference does.  Canonical configuration only."""
    if num_parts != 3:
        raise NotImplementedError                            # adj_graph, dataset_loader.py:346-349
    if num_split != 4 or not pyramid_part or num_scale != 1:
        _lib.check(_lib.E_UNSUPPORTED)
    kp, heights, valid = pack_keypoints(im_paths, im_sizes, poses)
    masks = part_masks(kp[None], heights[None], valid[None], threshold=threshold, num_split=num_split)
    return expand_adjacency(masks, len(im_paths))[0].cpu()
