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


def head_spreads(head_positions, number_of_neighbors=11):
    """crowd/database_preprocessor.py:146-149: mean distance of every head to its min(11, n) nearest heads, ITSELF included
    (distance 0), the MCNN-style geometry-adaptive kernel width."""
    heads = np.asarray(head_positions, dtype=np.float64)
    d = heads[:, None, :] - heads[None, :, :]
    distances = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
    k = min(number_of_neighbors, len(heads))
    return np.ascontiguousarray(np.sort(distances, axis=1)[:, :k]).mean(axis=1)


def make_gaussian(standard_deviation):
    """crowd/database_preprocessor.py:228-243 for a scalar standard deviation."""
    off = int(standard_deviation * 2)
    line = np.linspace(-off, off, off * 2 + 1)
    x, y = np.meshgrid(line, line)
    return np.exp(-((x ** 2) / (2.0 * standard_deviation ** 2) + (y ** 2) / (2.0 * standard_deviation ** 2)))


def generate_density_label(head_positions, label_size, neighbor_deviation_beta=0.15):
    """crowd/database_preprocessor.py:113-225 as generate_labels_for_example calls it (:87-88): perspective=None,
    perspective_resizing=True, yx_order=True, no body, force_full_image_count_normalize=True.  Positions are rounded to
    uint32 like the reference (np.rint(..).astype(np.uint32)); the window arithmetic is in Python ints, which is what the
    reference's expressions evaluate to under the NumPy 1.x promotion rules it was written for (under NumPy >= 2 `y - off` stays
    uint32, wraps for heads within `off` pixels of the top / left border, and the reference raises a broadcasting error)."""
    spreads = head_spreads(head_positions)
    label = np.zeros(shape=label_size, dtype=np.float32)
    head_count = 0
    for head_index, head_position in enumerate(head_positions):
        y, x = (int(v) for v in np.rint(head_position).astype(np.uint32))
        gaussian = make_gaussian(spreads[head_index] * neighbor_deviation_beta)
        gaussian = gaussian / gaussian.sum()
        head_count += 1
        off = int((gaussian.shape[0] - 1) / 2)
        y0, y1 = max(off - y, 0), max(y + off + 1 - label_size[0], 0)
        x0, x1 = max(off - x, 0), max(x + off + 1 - label_size[1], 0)
        if gaussian.shape[0] <= max(y0, y1) or gaussian.shape[1] <= max(x0, x1):
            continue                                         # 'Offset out of head gaussian bounds. Skipping person.' (:189-191)
        person = np.zeros_like(label)
        person[y - off + y0:y + off + 1 - y1, x - off + x0:x + off + 1 - x1] += gaussian[y0:gaussian.shape[0] - y1,
                                                                                         x0:gaussian.shape[1] - x1]
        label += person
    return head_count * (label / label.sum())
