"""
Generates tests/golden/crowd_labels.npz from the UNMODIFIED reference's label preprocessing functions
(crowd/database_preprocessor.py: generate_knn_map with its scikit-learn ball tree, generate_point_density_map)  --  TEST
INFRASTRUCTURE.  Run in the build container:  python oracle/make_golden_labels.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402

CASES = {            # name: (label size, number of heads, seed, upper bound)
    'dense': ((40, 56), 37, 0, None),
    'sparse': ((33, 21), 3, 1, None),          # fewer heads than k = 4, 5
    'bounded': ((24, 48), 12, 2, 6.5),
    'single': ((16, 16), 1, 3, None),
}


def heads_for(size, n, seed):
    rng = np.random.RandomState(seed)
    heads = rng.rand(n, 2) * np.array(size, dtype=np.float64)
    heads[: n // 3] = np.floor(heads[: n // 3])                 # some annotations on exact pixel centres
    if n > 8:
        heads[-1] = heads[-2]                                    # a duplicated annotation (distance ties)
        heads[-3] = [size[0] + 3.2, size[1] * 0.5]               # outside the label (the archives contain such points)
        heads[-4] = [-0.4, 2.5]                                  # rounds to -0 / 2 (half to even)
        heads[-5] = [-1.7, 5.0]                                  # negative index: wraps like Python's
    return heads


def main():
    ref_harness.install_shims()
    from crowd.database_preprocessor import generate_knn_map, generate_point_density_map
    out = {}
    for name, (size, n, seed, ub) in CASES.items():
        heads = heads_for(size, n, seed)
        out[f'{name}/heads'], out[f'{name}/size'] = heads, np.array(size)
        out[f'{name}/upper_bound'] = np.float64(ub if ub is not None else 0.0)
        for k in (1, 2, 3, 4, 5):
            out[f'{name}/knn{k}'] = generate_knn_map(heads, list(size), number_of_neighbors=k, upper_bound=ub)
        density, oob = generate_point_density_map(heads, size)
        out[f'{name}/density'], out[f'{name}/oob'] = density, np.int64(oob)
    # ---- Gaussian density labels (generate_density_label as generate_labels_for_example calls it, :81-89).  Under NumPy >= 2 the
    # reference's `y - off_center_size` stays uint32 and the function raises for heads within `off` pixels of the top / left
    # border (it was written for the NumPy 1.x promotion rules), so these cases keep every head at least its own kernel
    # half-width away from those two borders; clipping at the bottom / right border and heads beyond them ARE exercised.
    from crowd.database_preprocessor import generate_density_label
    rng = np.random.RandomState(7)
    size = (72, 96)
    heads = np.concatenate([rng.rand(40, 2) * np.array([40.0, 60.0]) + np.array([30.0, 34.0]),       # interior + bottom / right
                            np.array([[71.4, 95.2], [70.0, 50.5], [45.5, 95.0], [73.6, 60.0]])])   # on / just past the far borders
    out['density/heads'], out['density/size'] = heads, np.array(size)
    for beta in (0.05, 0.1, 0.3, 0.5):
        out[f'density/beta{beta}'] = generate_density_label(heads, size, perspective_resizing=True, yx_order=True,
                                                            neighbor_deviation_beta=beta)
    path = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'crowd_labels.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
