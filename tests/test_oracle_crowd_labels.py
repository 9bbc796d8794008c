"""Pins oracle/crowd_labels_oracle.py (kNN maps, point density map) against
tests/golden/crowd_labels.npz, produced by the UNMODIFIED reference functions with their scikit-learn ball tree
(oracle/make_golden_labels.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import crowd_labels_oracle as L

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'crowd_labels.npz')
CASES = ('dense', 'sparse', 'bounded', 'single')


@pytest.mark.parametrize('name', CASES)
def test_label_oracle_bit_exact_with_reference(name):
    g = np.load(GOLDEN)
    heads, size = g[f'{name}/heads'], tuple(int(v) for v in g[f'{name}/size'])
    ub = float(g[f'{name}/upper_bound']) or None
    for k in (1, 2, 3, 4, 5):
        assert np.array_equal(L.generate_knn_map(heads, size, k, ub), g[f'{name}/knn{k}']), (name, k)
    density, oob = L.generate_point_density_map(heads, size)
    assert np.array_equal(density, g[f'{name}/density']) and oob == int(g[f'{name}/oob'])


def test_density_label_oracle_bit_exact_with_reference():
    """generate_density_label (MCNN-style geometry-adaptive Gaussians), the four betas the preprocessor writes."""
    g = np.load(GOLDEN)
    heads, size = g['density/heads'], tuple(int(v) for v in g['density/size'])
    for beta in (0.05, 0.1, 0.3, 0.5):
        got = L.generate_density_label(heads, size, beta)
        assert got.dtype == np.float32 and np.array_equal(got, g[f'density/beta{beta}']), beta
    assert float(g['density/beta0.3'].sum()) == pytest.approx(len(heads), rel=1e-5)     # normalised to the head count (:223-224)
