"""SURVEY section 8 row f4 on the GPU: srgan_knn_maps / srgan_point_density_map through the C ABI against the golden vectors
of the unmodified reference (scikit-learn ball tree) and the numpy oracle -- bit-exact float64."""
import os

import numpy as np
import pytest
import torch

from oracle import crowd_labels_oracle as L

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'crowd_labels.npz')
CASES = ('dense', 'sparse', 'bounded', 'single')


@pytest.mark.parametrize('name', CASES)
def test_knn_and_density_maps_bit_exact_with_reference_golden(name):
    from srgan_b200 import crowd_labels
    g = np.load(GOLDEN)
    heads, size = g[f'{name}/heads'], tuple(int(v) for v in g[f'{name}/size'])
    ub = float(g[f'{name}/upper_bound']) or None
    knn, iknn = crowd_labels.generate_knn_maps(heads, size, 5, ub)
    for k in (1, 2, 3, 4, 5):
        ref = g[f'{name}/knn{k}']
        assert np.array_equal(knn[k - 1].cpu().numpy(), ref), (name, k)
        assert np.array_equal(iknn[k - 1].cpu().numpy(), L.iknn_map(ref)), (name, k)          # 1 / (knn + 1) as float16
        one = crowd_labels.generate_knn_map(heads, size, k, ub)                               # the reference's signature
        assert np.array_equal(one.cpu().numpy(), ref), (name, k)
    density, oob = crowd_labels.generate_point_density_map(heads, size)
    assert np.array_equal(density.cpu().numpy().astype(np.float64), g[f'{name}/density']) and oob == int(g[f'{name}/oob'])


def test_knn_maps_full_size_properties_and_oracle_rows():
    """ShanghaiTech-sized label (768 x 1024) with 1 500 heads (more than one shared-memory tile): a strip of rows against the
    oracle bit for bit, and size-independent properties over the whole map -- the k-NN means increase with k, the 1-NN map is
    zero exactly at integer head positions, and every value is bounded by the label's diagonal."""
    from srgan_b200 import crowd_labels
    rng = np.random.RandomState(4)
    H, W = 768, 1024
    heads = rng.rand(1500, 2) * np.array([H, W], dtype=np.float64)
    heads[:200] = np.floor(heads[:200])
    knn, iknn = crowd_labels.generate_knn_maps(heads, (H, W), 5)
    strip = L.generate_knn_map(heads, (8, W), 3)                       # rows 0..7 of the full map (positions are absolute)
    assert np.array_equal(knn[2, :8].cpu().numpy(), strip)
    assert bool((knn[1:] >= knn[:-1]).all())
    iy, ix = heads[:200, 0].astype(int), heads[:200, 1].astype(int)
    assert float(knn[0][iy, ix].abs().max()) == 0.0
    assert float(knn.max()) <= float(np.hypot(H, W))
    assert float(iknn[0][iy, ix].min()) == 1.0                          # 1 / (0 + 1)
    density, oob = crowd_labels.generate_point_density_map(heads, (H, W))
    assert float(density.sum()) == 1500.0 and oob == 0


def test_label_generation_argument_errors():
    from srgan_b200 import crowd_labels
    with pytest.raises(ValueError):
        crowd_labels.generate_knn_maps(np.zeros((0, 2)), (8, 8))
    with pytest.raises(RuntimeError):
        crowd_labels.generate_knn_maps(np.zeros((3, 2)), (8, 8), k_max=9)
    density, oob = crowd_labels.generate_point_density_map(np.zeros((0, 2)), (8, 8))
    assert float(density.abs().sum()) == 0.0 and oob == 0
