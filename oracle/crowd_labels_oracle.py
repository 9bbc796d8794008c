"""
CPU oracle (numpy) of the reference's crowd label preprocessing  --  TEST INFRASTRUCTURE (tests/, smoke(), bench cpu leg only).

  generate_knn_map            crowd/database_preprocessor.py:258-290
  generate_point_density_map  crowd/database_preprocessor.py:246-256
  iknn_map                    crowd/database_preprocessor.py:92-99

The reference queries a scikit-learn ball tree (scikit-learn is an unpinned requirement, requirements.txt; 1.7 is installed
here).  A ball tree is exact, so its published result is restated as a brute-force search: Euclidean distance
sqrt((y - hy)^2 + (x - hx)^2) in float64 from every label position to every head, the k smallest in ascending order,
optional clip, mean over the k columns.  Pinned: oracle/make_golden_labels.py runs the reference functions (with
scikit-learn) on seeded cases and commits tests/golden/crowd_labels.npz; tests/test_oracle_crowd_labels.py compares bit for bit.
"""
import numpy as np


def generate_knn_map(head_positions, label_size, number_of_neighbors=1, upper_bound=None):
    heads = np.asarray(head_positions, dtype=np.float64)
    ys, xs = np.meshgrid(np.arange(label_size[0], dtype=np.float64), np.arange(label_size[1], dtype=np.float64), indexing='ij')
    dy = ys.reshape(-1, 1) - heads[:, 0].reshape(1, -1)
    dx = xs.reshape(-1, 1) - heads[:, 1].reshape(1, -1)
    distances = np.sqrt(dy * dy + dx * dx)
    k = min(number_of_neighbors, len(heads))
    nearest = np.sort(distances, axis=1)[:, :k]
    if upper_bound is not None:
        nearest = np.clip(nearest, a_min=None, a_max=upper_bound)
    return np.ascontiguousarray(nearest).mean(axis=1).reshape(label_size)


def iknn_map(knn_map, epsilon=1):
    return (1 / (knn_map + epsilon)).astype(np.float16)


def generate_point_density_map(head_positions, label_size):
    density_map = np.zeros(label_size)
    out_of_bounds_count = 0
    for y, x in head_positions:
        y, x = int(round(y)), int(round(x))
        if -label_size[0] <= y < label_size[0] and -label_size[1] <= x < label_size[1]:
            density_map[y, x] += 1
        else:
            out_of_bounds_count += 1
    return density_map, out_of_bounds_count
