"""Pin oracle/distance.py to outputs of the reference's compute_distance_matrix (tests/golden)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_files
from oracle import distance as odist


@pytest.mark.parametrize('fname', golden_files('distance_'))
@pytest.mark.parametrize('metric', ['euclidean', 'cosine'])
def test_distance_oracle_matches_reference_golden(fname, metric):
    g = np.load(os.path.join(GOLDEN, fname))
    a, b = torch.from_numpy(g['a']), torch.from_numpy(g['b'])
    got = odist.distance_matrix(a, b, metric).numpy()
    ref = g[metric]
    # same formula, same library: differences are summation-order rounding only
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-6 * scale
    exact = odist.distance_matrix(a, b, metric, dtype=torch.float64).numpy()
    assert np.abs(exact - ref).max() <= 4e-6 * scale


def test_distance_oracle_argument_checks():
    a = torch.zeros(3, 4)
    with pytest.raises(AssertionError):
        odist.distance_matrix(a, torch.zeros(4, 5))
    with pytest.raises(AssertionError):
        odist.distance_matrix(a[0], a)
    with pytest.raises(ValueError):
        odist.distance_matrix(a, a, 'manhattan')
