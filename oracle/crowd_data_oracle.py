"""
CPU oracle (numpy) of the reference's crowd input transforms and sliding-window inference  --  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(sr-gan_b200/crowd_data.py -> libsrgan_b200.so) never does.  It restates, function by function, what the reference
computes on the host for SURVEY section 8 rows f1 (input pipeline) and f2 (validation / full-image inference):

  extract_patch            crowd/data.py:377-424 (ExtractPatch.get_patch_for_position, allow_padded=True) + :426-452 (pad_example)
  random_horizontal_flip   crowd/data.py:97-112
  normalize_image          crowd/data.py:120-128
  to_chw_float32           crowd/data.py:46-63
  transformed_position     crowd/shanghai_tech_data.py:80-98 (flat index -> file, y, x)
  transformed_item         crowd/shanghai_tech_data.py:73-104 (__getitem__, index already drawn)
  sliding_positions        crowd/data.py:525-539 (ImageSlidingWindowDataset.__init__; sorted instead of list(set(..)) order)
  predict_full_example     crowd/srgan.py:332-395 (scipy.misc.imresize to the SAME size = identity: patch 224, label 224)
  evaluation_sums          crowd/srgan.py:149-191 (the float64 reductions behind ME / MAE / MSE, kNN MAE / MSE)
  image_label_item         age/data.py:52-60, driving/data.py:44-51 (+ utility.to_normalized_range, utility.py:129-132)

Parity pinned: oracle/make_golden_data.py runs the unmodified reference classes (crowd/data.py, crowd/shanghai_tech_data.py,
crowd/srgan.py) on seeded synthetic examples and commits inputs + outputs as tests/golden/crowd_data.npz;
tests/test_oracle_golden.py checks this module against them bit for bit.
"""
from __future__ import annotations

import numpy as np


def extract_patch(image, label, map_, y, x, patch):
    """Window of `patch` x `patch` pixels centred at (y, x); where it leaves the example, the example is padded first:
    constant 0 for image, label and map (crowd/data.py:442-452; 'edge' padding applies to the perspective only)."""
    half = int(patch // 2)

    def pad(y_pad=(0, 0), x_pad=(0, 0)):
        nonlocal image, label, map_
        image = np.pad(image, (y_pad, x_pad, (0, 0)), 'constant')
        if label is not None:
            label = np.pad(label, (y_pad, x_pad), 'constant')
        if map_ is not None:
            map_ = np.pad(map_, (y_pad, x_pad), 'constant')
    if y - half < 0:
        pad(y_pad=(half - y, 0))
        y += half - y
    if y + half > image.shape[0]:
        pad(y_pad=(0, y + half - image.shape[0]))
    if x - half < 0:
        pad(x_pad=(half - x, 0))
        x += half - x
    if x + half > image.shape[1]:
        pad(x_pad=(0, x + half - image.shape[1]))
    rows, cols = slice(y - half, y + half), slice(x - half, x + half)
    return (image[rows, cols, :], None if label is None else label[rows, cols], None if map_ is None else map_[rows, cols])


def random_horizontal_flip(image, label, map_, flip):
    if flip:
        image = np.flip(image, axis=1).copy()
        label = None if label is None else np.flip(label, axis=1).copy()
        map_ = None if map_ is None else np.flip(map_, axis=1).copy()
    return image, label, map_


def normalize_image(image):
    return (image.astype(np.float32) / (255 / 2)) - 1


def to_chw_float32(image):
    return np.ascontiguousarray(image.transpose((2, 0, 1)), dtype=np.float32)


def position_counts(shapes, patch):
    """Per image: number of valid centre positions along y and x (crowd/shanghai_tech_data.py:63-66; an image smaller than
    the patch has none: Python's empty range)."""
    half = int(patch // 2)
    return [(len(range(half, h - half + 1)), len(range(half, w - half + 1))) for h, w in shapes]


def start_indexes(shapes, patch):
    starts, length = [], 0
    for ny, nx in position_counts(shapes, patch):
        starts.append(length)
        length += ny * nx
    return starts, length


def transformed_position(shapes, patch, index_):
    """Flat position index -> (file index, y, x), crowd/shanghai_tech_data.py:80-98."""
    starts, _ = start_indexes(shapes, patch)
    half = int(patch // 2)
    f = int(np.searchsorted(starts, index_, side='right') - 1)
    h, w = shapes[f]
    ys, xs = range(half, h - half + 1), range(half, w - half + 1)
    yi, xi = np.unravel_index(index_ - starts[f], [len(ys), len(xs)])
    return f, ys[yi], xs[xi]


def transformed_item(examples, patch, index_, flip):
    """ShanghaiTechTransformedDataset.__getitem__ with its two random draws (index_, flip) given."""
    shapes = [e[0].shape[:2] for e in examples]
    f, y, x = transformed_position(shapes, patch, index_)
    image, label, map_ = extract_patch(*examples[f], y, x, patch)
    image, label, map_ = random_horizontal_flip(image, label, map_, flip)
    return to_chw_float32(normalize_image(image)), label.astype(np.float32), map_.astype(np.float32)


def sliding_positions(extent, patch, step):
    """Window centres along one axis (crowd/data.py:530-537): every `step` pixels from patch/2, plus the last full window;
    an image smaller than the patch gets the single (padded) window centred at extent - patch/2."""
    half = int(patch // 2)
    positions = list(range(half, extent - half + 1, step))
    if extent - half > 0:
        positions = sorted(set(positions + [extent - half]))
    return positions


def sliding_item(image, patch, y, x):
    """ImageSlidingWindowDataset.__getitem__ (crowd/data.py:541-557): the normalised CHW patch of an unlabeled example."""
    p, _, _ = extract_patch(image, None, None, y, x, patch)
    return to_chw_float32(normalize_image(p))


def predict_full_example(image, network, patch, step, batch_size):
    """crowd/srgan.py:332-395.  `network(images [n,3,patch,patch] float32) -> (labels [n,patch,patch], counts [n], maps)`.
    Returns (full count (float32 numpy scalar), full label [H,W] float32)."""
    H, W = image.shape[:2]
    ys, xs = sliding_positions(H, patch, step), sliding_positions(W, patch, step)
    sum_density = np.zeros((H, W), dtype=np.float32)
    sum_count = np.zeros((H, W), dtype=np.float32)
    hits = np.zeros((H, W), dtype=np.int32)
    half = patch // 2
    order = [(y, x) for y in ys for x in xs]
    for start in range(0, len(order), batch_size):
        chunk = order[start:start + batch_size]
        images = np.stack([sliding_item(image, patch, y, x) for y, x in chunk])
        labels, counts, _ = network(images)
        for k, (y, x) in enumerate(chunk):
            label = np.asarray(labels[k], dtype=np.float32)
            count_array = np.full(label.shape, np.float32(counts[k]) / label.size)
            y0, y1 = max(half - y, 0), max(y + half - H, 0)
            x0, x1 = max(half - x, 0), max(x + half - W, 0)
            rows = slice(y - half + y0, y + half - y1)
            cols = slice(x - half + x0, x + half - x1)
            sum_density[rows, cols] += label[y0:label.shape[0] - y1, x0:label.shape[1] - x1]
            sum_count[rows, cols] += count_array[y0:label.shape[0] - y1, x0:label.shape[1] - x1]
            hits[rows, cols] += 1
    hits[hits == 0] = 1
    full_label = sum_density / hits.astype(np.float32)
    full_count = np.sum(sum_count / hits.astype(np.float32))
    return full_count, full_label


def evaluation_sums(predicted_counts, densities, predicted_maps, maps):
    """The scalars evaluation_epoch writes (crowd/srgan.py:178-187); every array is float64 there (concatenated onto
    np.array([])).  predicted_maps [n,3,H,W] vs maps [n,H,W] (expanded to [n,1,H,W], :177)."""
    predicted_counts = np.asarray(predicted_counts, dtype=np.float64)
    densities = np.asarray(densities, dtype=np.float64)
    predicted_maps = np.asarray(predicted_maps, dtype=np.float64)
    maps = np.expand_dims(np.asarray(maps, dtype=np.float64), axis=1)
    true_counts = densities.sum(1).sum(1)
    return {'ME': (predicted_counts - true_counts).mean(),
            'MAE': np.abs(predicted_counts - true_counts).mean(),
            'kNN MAE': np.abs(predicted_maps - maps).mean(),
            'MSE': (np.abs(predicted_counts - true_counts) ** 2).mean(),
            'kNN MSE': (np.abs(predicted_maps - maps) ** 2).mean()}


def image_label_item(image, label, hwc=True):
    """AgeDataset.__getitem__ (age/data.py:52-60: imageio HWC uint8 -> transpose((2, 0, 1)) -> float32 -> to_normalized_range,
    utility.py:129-132) / SteeringAngleDataset.__getitem__ (driving/data.py:44-51: the stored array is already CHW)."""
    if hwc:
        image = image.transpose((2, 0, 1))
    return (image.astype(np.float32) / np.float32(127.5)) - np.float32(1), np.float32(label)
