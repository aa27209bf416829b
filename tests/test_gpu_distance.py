"""GPU parity of the tcgen05 distance matrix (csrc/distance.cu, gemm_sm100.cuh, split.cu) through
compute_distance_matrix: reference golden vectors, oracle (fp32 and fp64) on larger seeded inputs,
ragged shapes, argument checks.  Tolerance (north star): 1e-4 relative; the fp32-accurate splits (fp16 x 2 with
three products -- the default -- and bf16 x 3 with six) are held to fp32-level accuracy (2e-6 of the matrix scale) on
top of that."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_files
from oracle import distance as odist

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4           # the north-star bar
FP32_TOL = 2e-6          # what the fp32-accurate split actually has to deliver


@pytest.fixture(scope='module')
def cdm():
    from agrl.pytorch_b200.metrics import compute_distance_matrix
    return compute_distance_matrix


def _err(got, ref):
    scale = np.abs(ref).max()
    return np.abs(got - ref).max() / scale, np.linalg.norm(got - ref) / np.linalg.norm(ref)


ACCURATE_SPLITS = [5, 3]     # _lib.SPLIT_FP16X2 (default), _lib.SPLIT_BF16X3


@pytest.mark.parametrize('fname', golden_files('distance_'))
@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
@pytest.mark.parametrize('where', ['cpu', 'cuda'])
@pytest.mark.parametrize('split', ACCURATE_SPLITS)
def test_distance_golden(cdm, fname, metric, where, split):
    g = np.load(os.path.join(GOLDEN, fname))
    a, b = torch.from_numpy(g['a']), torch.from_numpy(g['b'])
    if where == 'cuda':
        a, b = a.cuda(), b.cuda()
    out = cdm(a, b, metric, split=split)
    assert out.shape == (a.size(0), b.size(0)) and out.dtype == torch.float32
    assert out.device.type == where
    emax, enrm = _err(out.cpu().numpy(), g[metric])
    assert emax < FP32_TOL and enrm < FP32_TOL, (emax, enrm)


@pytest.mark.parametrize('m,n,d', [(300, 1000, 2048), (129, 257, 4096), (1, 1, 1), (7, 130, 63), (128, 128, 64),
                                   (255, 383, 200), (702, 2636, 4096)])
@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
@pytest.mark.parametrize('split', ACCURATE_SPLITS)
def test_distance_vs_oracle(cdm, m, n, d, metric, split):
    g = torch.Generator().manual_seed(m * 7 + n)
    a, b = torch.randn(m, d, generator=g), torch.randn(n, d, generator=g)
    ref32 = odist.distance_matrix(a, b, metric).numpy()
    ref64 = odist.distance_matrix(a, b, metric, dtype=torch.float64).numpy()
    out = cdm(a.cuda(), b.cuda(), metric, split=split).cpu().numpy()
    emax, enrm = _err(out, ref32)
    assert emax < REL_TOL and enrm < REL_TOL
    # against the exact value we must be about as good as the reference's own fp32 arithmetic
    e_ours = np.abs(out - ref64).max()
    e_ref = np.abs(ref32 - ref64).max()
    assert e_ours <= max(4 * e_ref, FP32_TOL * np.abs(ref64).max()), (e_ours, e_ref)


def test_two_plane_split_meets_the_bar(cdm):
    from agrl.pytorch_b200 import _lib
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(200, 2048, generator=g), torch.randn(300, 2048, generator=g)
    for metric in ('euclidean', 'cosine'):
        ref = odist.distance_matrix(a, b, metric).numpy()
        out = cdm(a.cuda(), b.cuda(), metric, split=_lib.SPLIT_BF16X2).cpu().numpy()
        emax, enrm = _err(out, ref)
        assert emax < REL_TOL and enrm < REL_TOL


def test_clustered_small_distances(cdm):
    """near-duplicate rows: the Gram form cancels, so the absolute error is set by the norms.  The
    tensor core truncates on accumulate (a small systematic bias on all-positive sums), hence the
    bound is stated against |q|^2+|g|^2: 1.5e-6 relative (the accumulator is drained every 256 k and summed in fp32)."""
    g = torch.Generator().manual_seed(9)
    base = torch.randn(64, 1024, generator=g)
    a = base + 1e-3 * torch.randn(64, 1024, generator=g)
    ref64 = odist.distance_matrix(a, base, 'euclidean', dtype=torch.float64).numpy()
    ref32 = odist.distance_matrix(a, base, 'euclidean').numpy()
    out = cdm(a.cuda(), base.cuda(), 'euclidean').cpu().numpy()
    norms = (a.double() ** 2).sum(1).numpy()[:, None] + (base.double() ** 2).sum(1).numpy()[None, :]
    e_ours, e_ref = np.abs(out - ref64) / norms, np.abs(ref32 - ref64) / norms
    print('clustered: ours %.3e  reference fp32 %.3e (relative to the norms)' % (e_ours.max(), e_ref.max()))
    assert e_ours.max() < 1.5e-6


@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
def test_fp16x2_rows_of_any_magnitude(cdm, metric):
    """the default split scales every row by its own power of two: rows 30 orders of magnitude apart, an all-zero row, a
    row with one dominant element and elements far below it, a NaN row (poisons only itself)"""
    g = torch.Generator().manual_seed(11)
    a, b = torch.randn(70, 520, generator=g), torch.randn(150, 520, generator=g)
    a *= torch.logspace(-15, 15, 70)[:, None]
    b *= torch.logspace(-12, 12, 150)[:, None]
    a[3] = 0
    b[5] = 0
    a[7, 1:] *= 1e-9                                      # one element dominates the row
    b[9, 17] *= 1e7
    ref64 = odist.distance_matrix(a, b, metric, dtype=torch.float64).numpy()
    out = cdm(a.cuda(), b.cuda(), metric).cpu().numpy()
    if metric == 'euclidean':
        scale = (a.double() ** 2).sum(1).numpy()[:, None] + (b.double() ** 2).sum(1).numpy()[None, :]
        scale[scale == 0] = 1.0
    else:
        scale = np.ones_like(ref64)
    assert np.isfinite(out).all()
    assert (np.abs(out - ref64) / scale).max() < FP32_TOL
    a[11, 4] = float('nan')
    out2 = cdm(a.cuda(), b.cuda(), metric).cpu().numpy()
    assert np.isnan(out2[11]).all()
    keep = np.arange(70) != 11
    assert np.array_equal(out2[keep], out[keep])


def test_non_contiguous_and_strided_inputs(cdm):
    g = torch.Generator().manual_seed(3)
    big = torch.randn(50, 300, generator=g).cuda()
    a, b = big[:20, 10:210], big[20:, 10:210]            # row stride 300, offset start
    ref = odist.distance_matrix(a.cpu(), b.cpu(), 'euclidean').numpy()
    out = cdm(a, b, 'euclidean').cpu().numpy()
    assert _err(out, ref)[0] < FP32_TOL


def test_argument_checks(cdm):
    a = torch.zeros(3, 4).cuda()
    with pytest.raises(AssertionError):
        cdm(a, torch.zeros(4, 5).cuda())
    with pytest.raises(AssertionError):
        cdm(a[0], a)
    with pytest.raises(AssertionError):
        cdm(a.cpu().numpy(), a)
    with pytest.raises(ValueError, match='Unknown distance metric'):
        cdm(a, a, 'manhattan')
    assert cdm(a[:0], a).shape == (0, 3)


@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
@pytest.mark.parametrize('split', ACCURATE_SPLITS)
def test_distance_is_symmetric_at_the_mars_shape(cdm, metric, split):
    """d(a, b) == d(b, a)^T to fp32 rounding -- a size-independent property, checked at the MARS shape where the two calls
    run different tile grids (not bit for bit: the correction products a0.b1 + a1.b0 swap roles, and the tensor core's
    accumulate truncation depends on the order)"""
    g = torch.Generator(device='cuda').manual_seed(17)
    a = torch.randn(1980, 2048, generator=g, device='cuda')
    b = torch.randn(9330, 2048, generator=g, device='cuda')
    ab = cdm(a, b, metric, split=split)
    ba = cdm(b, a, metric, split=split)
    scale = float((a ** 2).sum(1).max() + (b ** 2).sum(1).max()) if metric == 'euclidean' else 1.0
    assert float((ab - ba.t()).abs().max()) < FP32_TOL * scale
    # and the diagonal of d(a, a) is (numerically) zero: |a|^2 + |a|^2 - 2 a.a cancels to rounding, 1 - cos = 0 to 1e-6
    aa = cdm(a[:300], a[:300], metric, split=split)
    scale = float((a[:300] ** 2).sum(1).max()) * 2 if metric == 'euclidean' else 1.0
    assert float(aa.diagonal().abs().max()) < FP32_TOL * scale
