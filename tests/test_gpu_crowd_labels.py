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


def test_density_label_matches_reference_golden():
    """srgan_density_label vs the reference's generate_density_label for the four betas the preprocessor writes: fp32 within
    2e-6 of the label's peak (device exp / sums differ from numpy's in the last bits), the saved float16 copy equal except for
    rare one-ulp roundings, and the label sums to the head count."""
    from srgan_b200 import crowd_labels
    g = np.load(GOLDEN)
    heads, size = g['density/heads'], tuple(int(v) for v in g['density/size'])
    for beta in (0.05, 0.1, 0.3, 0.5):
        ref = g[f'density/beta{beta}']
        label, f16 = crowd_labels.generate_density_label(heads, size, beta, half=True)
        got = label.cpu().numpy()
        assert np.abs(got - ref).max() <= 2e-6 * ref.max(), (beta, np.abs(got - ref).max(), ref.max())
        assert (got == 0).sum() == (ref == 0).sum()                       # the same support: clipping and skipped heads
        ref16, got16 = ref.astype(np.float16), f16.cpu().numpy()
        differ = got16 != ref16
        assert differ.mean() < 2e-3, (beta, differ.mean())
        assert np.abs(got16[differ].astype(np.float32) - ref16[differ].astype(np.float32)).max(initial=0) <= 2e-3 * ref.max()
        assert float(label.sum()) == pytest.approx(len(heads), rel=1e-5)


def test_density_label_full_size_vs_oracle_strip():
    """768 x 1024 label, 600 heads incl. some next to every border (the oracle uses the NumPy-1.x integer semantics the
    reference was written for): the whole map against the oracle, and mass conservation."""
    from srgan_b200 import crowd_labels
    rng = np.random.RandomState(9)
    H, W = 768, 1024
    heads = rng.rand(600, 2) * np.array([H, W], dtype=np.float64)
    heads[:8] = [[0.2, 0.3], [1.0, 500.0], [400.0, 0.6], [767.4, 1023.2], [766.0, 3.0], [2.5, 1022.5], [383.5, 511.5], [0.0, 0.0]]
    ref = L.generate_density_label(heads, (H, W), 0.3)
    got = crowd_labels.generate_density_label(heads, (H, W), 0.3).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * ref.max()
    assert (got == 0).sum() == (ref == 0).sum()
    assert float(got.astype(np.float64).sum()) == pytest.approx(600.0, rel=1e-5)
