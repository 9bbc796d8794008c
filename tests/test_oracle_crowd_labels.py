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
