"""Pins oracle/crowd_data_oracle.py (crowd input transforms, sliding-window inference,
evaluation sums) against tests/golden/crowd_data.npz, produced by the UNMODIFIED reference classes
(oracle/make_golden_data.py).  CPU only."""
import os
import random

import numpy as np
import pytest

from oracle import crowd_data_oracle as C
from oracle.make_golden_data import fake_network

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'crowd_data.npz')


@pytest.fixture(scope='module')
def g():
    return np.load(GOLDEN)


def examples(g):
    return [(g[f'image{i}'], g[f'label{i}'], g[f'map{i}']) for i in range(int(g['n_images']))]


def test_patch_transform_chain_bit_exact(g):
    ex, patch = examples(g), int(g['patch'])
    for k, (i, y, x, flip) in enumerate(g['f1_pos']):
        image, label, map_ = C.extract_patch(*ex[i], int(y), int(x), patch)
        image, label, map_ = C.random_horizontal_flip(image, label, map_, bool(flip))
        image = C.to_chw_float32(C.normalize_image(image))
        assert np.array_equal(image, g['f1_images'][k]), (k, i, y, x, flip)
        assert np.array_equal(label, g['f1_labels'][k]) and np.array_equal(map_, g['f1_maps'][k]), (k, i, y, x, flip)


def test_transformed_dataset_items_bit_exact_with_the_reference_draws(g):
    ex, patch = examples(g), int(g['patch'])
    store = [ex[i] for i in g['f1b_store']]
    _, length = C.start_indexes([e[0].shape[:2] for e in store], patch)
    assert length == int(g['f1b_length'])
    random.seed(int(g['f1b_seed']))
    for k in range(len(g['f1b_images'])):
        index_ = random.randrange(length)                  # crowd/shanghai_tech_data.py:80
        flip = random.choice([True, False])                # crowd/data.py:104
        image, label, map_ = C.transformed_item(store, patch, index_, flip)
        assert np.array_equal(image, g['f1b_images'][k]), k
        assert np.array_equal(label, g['f1b_labels'][k]) and np.array_equal(map_, g['f1b_maps'][k]), k


def test_sliding_window_positions_and_full_example_prediction(g):
    ex, patch, step = examples(g), int(g['patch']), int(g['step'])
    for i, (image, _, _) in enumerate(ex):
        H, W = image.shape[:2]
        assert C.sliding_positions(H, patch, step) == list(g[f'f2_ys{i}'])
        assert C.sliding_positions(W, patch, step) == list(g[f'f2_xs{i}'])
        count, label = C.predict_full_example(image, lambda im: [t.numpy() for t in fake_network(im)], patch, step, 7)
        # the reference walks its windows in list(set(..)) order, this oracle in sorted order: same sums up to fp32 rounding
        np.testing.assert_allclose(label, g[f'f2_label{i}'], rtol=2e-6, atol=1e-8)
        assert float(count) == pytest.approx(float(g[f'f2_count{i}']), rel=2e-6)


def test_evaluation_sums(g):
    labels, counts, maps = [t.numpy() for t in fake_network(g['f1_images'][:15])]
    out = C.evaluation_sums(counts, g['f1_labels'][:15], maps[:5], g['f1_maps'][:5])      # maps: first batch only (:170-175)
    for tag, v in out.items():
        assert v == pytest.approx(float(g['f2b_' + tag.replace(' ', '_')]), rel=1e-12), tag


def test_age_and_driving_items_bit_exact(g):
    """image_label_item vs the reference's AgeDataset / SteeringAngleDataset sample path (incl. every byte value)."""
    for k, image in enumerate(g['f1c_hwc']):
        a, _ = C.image_label_item(image, g['f1c_labels'][k], hwc=True)
        assert np.array_equal(a, g['f1c_age_images'][k]), k
        d, v = C.image_label_item(np.ascontiguousarray(image.transpose((2, 0, 1))), g['f1c_labels'][k], hwc=False)
        assert np.array_equal(d, g['f1c_driving_images'][k]) and v == g['f1c_driving_angles'][k], k


def test_product_host_logic_matches_oracle_without_a_gpu(g):
    """The host side of the input pipeline (no kernels): TransformedDataset's flat-index arithmetic and random draws, and the
    SlidingWindow positions, against the oracle / the reference golden -- on a stand-in store that only carries the image shapes."""
    from srgan_b200 import crowd_data
    ex, patch, step = examples(g), int(g['patch']), int(g['step'])
    store_ids = [int(i) for i in g['f1b_store']]
    shapes = [ex[i][0].shape[:2] for i in store_ids]
    fake_store = type('ShapesOnly', (), {'shapes': shapes})()
    ds = crowd_data.TransformedDataset(fake_store, patch, patch)
    starts, length = C.start_indexes(shapes, patch)
    assert ds.start_indexes == starts and len(ds) == length == int(g['f1b_length'])
    for index_ in list(range(0, length, 37)) + [length - 1] + starts:
        assert ds.position(index_) == C.transformed_position(shapes, patch, index_), index_
    # the draws: random.randrange then random.choice per sample, like the reference's __getitem__ + RandomHorizontalFlip
    random.seed(int(g['f1b_seed']))
    table = ds.draw(24)
    random.seed(int(g['f1b_seed']))
    for b in range(24):
        f, y, x = C.transformed_position(shapes, patch, random.randrange(length))
        assert tuple(table[b]) == (f, y, x, int(random.choice([True, False]))), b
    for i, (image, _, _) in enumerate(ex):
        sw = crowd_data.SlidingWindow(*image.shape[:2], patch, step)
        assert sw.y_positions == list(g[f'f2_ys{i}']) and sw.x_positions == list(g[f'f2_xs{i}'])
        assert sw.table(3).shape == (sw.length, 4) and int(sw.table(3)[0, 0]) == 3
    with pytest.raises(NotImplementedError):
        crowd_data.TransformedDataset(fake_store, patch, patch // 2)
    with pytest.raises(RuntimeError):
        import torch
        if torch.cuda.is_available():
            raise RuntimeError('GPU present: the refusal below is the CPU-only behaviour')
        crowd_data.CrowdStore([ex[0]])


def world_expo_frames(g):
    """The frames of every synthetic camera, flattened camera after camera; the label doubles as the map
    (crowd/world_expo_data.py:146)."""
    frames = []
    for c in range(int(g['f1d_cameras'])):
        for image, label in zip(g[f'f1d_images{c}'], g[f'f1d_labels{c}']):
            frames.append((image, label, label))
    return frames


def test_world_expo_items_are_the_same_flat_index_over_flattened_frames(g):
    """WorldExpoTransformedDataset decomposes its flat index into (camera, frame, position); frames of a camera are equally
    sized, so that is the per-example decomposition of the ShanghaiTech dataset over the frames laid out camera after camera:
    the same oracle (and the same CrowdStore + TransformedDataset on the device) serves it."""
    frames, patch = world_expo_frames(g), int(g['patch'])
    _, length = C.start_indexes([f[0].shape[:2] for f in frames], patch)
    assert length == int(g['f1d_length'])
    random.seed(int(g['f1d_seed']))
    for k in range(len(g['f1d_out_images'])):
        index_ = random.randrange(length)
        flip = random.choice([True, False])
        image, label, map_ = C.transformed_item(frames, patch, index_, flip)
        assert np.array_equal(image, g['f1d_out_images'][k]), k
        assert np.array_equal(label, g['f1d_out_labels'][k]) and np.array_equal(map_, g['f1d_out_maps'][k]), k
