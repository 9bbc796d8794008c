"""SURVEY section 8 rows f1 / f2 on the GPU: the device-side crowd input pipeline (srgan_crowd_extract_patches), the
sliding-window merge (srgan_sliding_window_merge) and the evaluation sums (srgan_crowd_eval_sums), through the C ABI,
against the golden vectors of the unmodified reference (tests/golden/crowd_data.npz) and the numpy oracle."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import crowd_data_oracle as C
from oracle import srgan_oracle as O
from oracle.make_golden_data import fake_network

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'crowd_data.npz')


@pytest.fixture(scope='module')
def g():
    return np.load(GOLDEN)


def examples(g):
    return [(g[f'image{i}'], g[f'label{i}'], g[f'map{i}']) for i in range(int(g['n_images']))]


def cpu_network(images):
    """The golden file's stand-in network, evaluated on the host so that its outputs carry the golden bits."""
    return tuple(t.cuda() for t in fake_network(images.cpu()))


def test_extract_patches_bit_exact_with_reference_golden(g):
    from srgan_b200 import crowd_data
    store = crowd_data.CrowdStore(examples(g))
    images, labels, maps = store.extract(g['f1_pos'], int(g['patch']))
    assert torch.equal(images.cpu(), torch.tensor(g['f1_images']))
    assert torch.equal(labels.cpu(), torch.tensor(g['f1_labels']))
    assert torch.equal(maps.cpu(), torch.tensor(g['f1_maps']))
    # the sliding-window form: image only
    only, none_l, none_m = store.extract(g['f1_pos'], int(g['patch']), with_labels=False)
    assert none_l is None and none_m is None and torch.equal(only, images)


def test_transformed_dataset_reproduces_the_reference_dataset_draw_for_draw(g):
    """Same `random` state as the reference's worker-less DataLoader -> the same 24 samples, bit for bit."""
    from srgan_b200 import crowd_data
    ex = examples(g)
    store = crowd_data.CrowdStore([ex[i] for i in g['f1b_store']])
    ds = crowd_data.TransformedDataset(store, int(g['patch']), int(g['patch']))
    assert len(ds) == int(g['f1b_length'])
    random.seed(int(g['f1b_seed']))
    images, labels, maps = ds.batch(len(g['f1b_images']))
    assert torch.equal(images.cpu(), torch.tensor(g['f1b_images']))
    assert torch.equal(labels.cpu(), torch.tensor(g['f1b_labels']))
    assert torch.equal(maps.cpu(), torch.tensor(g['f1b_maps']))
    # a private generator leaves the global one alone and is reproducible
    a = crowd_data.TransformedDataset(store, int(g['patch']), int(g['patch']), rng=random.Random(3)).draw(16)
    b = crowd_data.TransformedDataset(store, int(g['patch']), int(g['patch']), rng=random.Random(3)).draw(16)
    assert np.array_equal(a, b) and a[:, 3].min() == 0 and a[:, 3].max() == 1
    n = sum(1 for _ in zip(range(3), ds.loader(7)))
    assert n == 3
    with pytest.raises(NotImplementedError):
        crowd_data.TransformedDataset(store, 32, 16)


def test_every_byte_value_normalises_like_numpy():
    from srgan_b200 import crowd_data
    image = np.arange(256, dtype=np.uint8).repeat(3 * 4).reshape(32, 32, 3)
    store = crowd_data.CrowdStore([(image, None, None)])
    out, _, _ = store.extract(np.array([[0, 16, 16, 0]], dtype=np.int32), 32)
    assert np.array_equal(out[0].cpu().numpy(), C.to_chw_float32(C.normalize_image(image)))


def test_full_size_batch_matches_oracle_bit_exact():
    """BASELINE's crowd shapes: 64 patches of 224 x 224 from 768 x 1024 / 480 x 640 / 200 x 300 images (the last one padded)."""
    from srgan_b200 import crowd_data
    rng = np.random.RandomState(0)
    ex = []
    for h, w in ((768, 1024), (480, 640), (200, 300)):
        ex.append((rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8), rng.rand(h, w).astype(np.float32),
                   rng.rand(h, w).astype(np.float32)))
    store = crowd_data.CrowdStore(ex)
    pos = np.array([(i % 3, rng.randint(ex[i % 3][0].shape[0]), rng.randint(ex[i % 3][0].shape[1]), rng.randint(2))
                    for i in range(64)], dtype=np.int32)
    images, labels, maps = store.extract(pos, 224)
    for b, (i, y, x, flip) in enumerate(pos):
        im, lb, mp = C.extract_patch(*ex[i], int(y), int(x), 224)
        im, lb, mp = C.random_horizontal_flip(im, lb, mp, bool(flip))
        assert np.array_equal(images[b].cpu().numpy(), C.to_chw_float32(C.normalize_image(im))), b
        assert np.array_equal(labels[b].cpu().numpy(), lb) and np.array_equal(maps[b].cpu().numpy(), mp), b
    with pytest.raises(IndexError):
        store.extract(np.array([[3, 0, 0, 0]], dtype=np.int32), 224)


def test_sliding_window_merge_vs_golden_and_oracle(g):
    from srgan_b200 import crowd_data
    ex, patch, step = examples(g), int(g['patch']), int(g['step'])
    store = crowd_data.CrowdStore(ex)
    for i, (image, _, _) in enumerate(ex):
        sw = crowd_data.SlidingWindow(*image.shape[:2], patch, step)
        assert sw.y_positions == list(g[f'f2_ys{i}']) and sw.x_positions == list(g[f'f2_xs{i}'])
        count, label = crowd_data.predict_full_example(store, i, cpu_network, patch, step, batch_size=7)
        o_count, o_label = C.predict_full_example(image, lambda im: [t.numpy() for t in fake_network(im)], patch, step, 7)
        assert np.array_equal(label.cpu().numpy(), o_label), i          # same per-pixel order of fp32 additions: same bits
        assert count.item() == pytest.approx(float(o_count), rel=2e-6)  # np.sum's pairwise fp32 vs a float64 sum
        np.testing.assert_allclose(label.cpu().numpy(), g[f'f2_label{i}'], rtol=2e-6, atol=1e-8)
        assert count.item() == pytest.approx(float(g[f'f2_count{i}']), rel=2e-6)


def test_sliding_window_constant_network_property_at_full_size():
    """Size-independent property at BASELINE's sizes (768 x 1024 image, patch 224, step 128): a network that returns the same
    label c1 and count c2 for every patch merges to label == c1 everywhere and count == c2 * H * W / 224^2; zeros labels
    (KnnDenseNetCat's, passed as None) stay zero."""
    from srgan_b200 import crowd_data
    image = np.random.RandomState(1).randint(0, 256, size=(768, 1024, 3)).astype(np.uint8)
    store = crowd_data.CrowdStore([(image, None, None)])

    def constant(images):
        n = images.shape[0]
        return torch.full((n, 224, 224), 0.375, device='cuda'), torch.full((n,), 7.0, device='cuda'), None
    count, label = crowd_data.predict_full_example(store, 0, constant)
    assert torch.equal(label, torch.full_like(label, 0.375))
    assert count.item() == pytest.approx(7.0 * 768 * 1024 / 224 ** 2, rel=1e-6)
    count, label = crowd_data.predict_full_example(store, 0, lambda im: (None, torch.full((im.shape[0],), 7.0, device='cuda'), None))
    assert float(label.abs().max()) == 0.0 and count.item() == pytest.approx(7.0 * 768 * 1024 / 224 ** 2, rel=1e-6)


def test_sliver_image_has_no_windows_like_the_reference():
    """A side of at most patch / 2 pixels: ImageSlidingWindowDataset yields no window (crowd/data.py:530-537) and the reference
    returns zero sums; so does the device path (no kernel launch with an empty window list)."""
    from srgan_b200 import crowd_data
    image = np.full((3, 50, 3), 7, dtype=np.uint8)
    store = crowd_data.CrowdStore([(image, None, None)])
    assert crowd_data.SlidingWindow(3, 50, 32, 12).length == 0 and C.sliding_positions(3, 32, 12) == []
    count, label = crowd_data.predict_full_example(store, 0, lambda im: (None, torch.ones(im.shape[0], device='cuda'), None), 32, 12)
    o_count, o_label = C.predict_full_example(image, lambda im: (np.zeros((im.shape[0], 32, 32), np.float32), np.ones(im.shape[0]), None),
                                              32, 12, 4)
    assert count.item() == float(o_count) == 0.0 and label.shape == (3, 50) and float(label.abs().max()) == 0.0
    assert float(np.abs(o_label).max()) == 0.0


def test_evaluation_epoch_matches_reference_golden(g):
    from srgan_b200 import crowd_data
    t = lambda k: torch.tensor(g[k][:15]).cuda()
    images, labels, maps = t('f1_images'), t('f1_labels'), t('f1_maps')
    batches = [(images[s:s + 5], labels[s:s + 5], maps[s:s + 5]) for s in range(0, 15, 5)]

    class Writer:
        scalars = {}

        def add_scalar(self, tag, value):
            self.scalars[tag] = value
    w = Writer()
    out = crowd_data.evaluation_epoch(cpu_network, batches, 5, summary_writer=w, comparison_value=2.0)
    for tag in ('ME', 'MAE', 'kNN MAE', 'MSE', 'kNN MSE'):
        ref = float(g['f2b_' + tag.replace(' ', '_')])
        assert out[tag] == pytest.approx(ref, rel=1e-10), tag
        assert w.scalars[f'Validation/{tag}'] == out[tag]
    assert out['Ratio MAE GAN DNN'] == pytest.approx(out['MAE'] / 2.0)


def test_predict_full_example_through_the_crowd_discriminator():
    """End to end: resident image -> sliding windows -> KnnDenseNetCat forward on the kernels (StepRunner.predict_crowd) ->
    merge, against the oracle's forward of the same (reduced) network through the oracle's predict_full_example."""
    from srgan_b200 import crowd_data
    from tests.gpu_common import runner_from_state
    small = dict(block_config=(2, 2, 2, 2), growth_rate=8, num_init_features=16, bn_size=2, label_patch_size=64)
    st = O.init_crowd(seed=1, image_size=64, z_dim=16, g_conv_dim=8, scale=2.0, **small)
    cfg = O.StepConfig(method='srgan', batch_size=4, map_multiplier=1e-3)
    r = runner_from_state(st, cfg, 'fp32')
    image = np.random.RandomState(2).randint(0, 256, size=(150, 200, 3)).astype(np.uint8)
    store = crowd_data.CrowdStore([(image, None, None)])
    count, label = crowd_data.predict_full_example(store, 0, r.predict_crowd, 64, 40, batch_size=4)

    def oracle_network(images):
        (c, m), _, _ = O.d_forward(st.d_spec, st.D, torch.tensor(images))
        return np.zeros((images.shape[0], 64, 64), np.float32), c.detach().numpy(), m.detach().numpy()
    o_count, o_label = C.predict_full_example(image, oracle_network, 64, 40, 4)
    assert float(label.abs().max()) == 0.0 and float(np.abs(o_label).max()) == 0.0
    assert count.item() == pytest.approx(float(o_count), rel=1e-4)
    # predict_crowd's maps are KnnDenseNetCat's [B, 3, L, L]
    x = torch.tensor(np.stack([C.sliding_item(image, 64, 32, 32), C.sliding_item(image, 64, 100, 150)]))
    _, c, m = r.predict_crowd(x.cuda())
    (c_ref, m_ref), _, _ = O.d_forward(st.d_spec, st.D, x)
    assert m.shape == (2, 3, 64, 64)
    assert (m.cpu() - m_ref).abs().max().item() < 1e-4 * max(m_ref.abs().max().item(), 1.0)
    assert (c.cpu() - c_ref).abs().max().item() < 1e-4 * max(c_ref.abs().max().item(), 1.0)


def test_age_and_driving_batches_bit_exact_with_reference_golden(g):
    """srgan_image_batch: AgeDataset (HWC store) and SteeringAngleDataset (CHW store) samples gathered by index on the device."""
    from srgan_b200 import crowd_data
    hwc, labels = g['f1c_hwc'], g['f1c_labels']
    index = [4, 0, 2, 2, 1, 3]
    age = crowd_data.ImageLabelStore(hwc, labels, hwc=True)
    images, y = age.batch(index)
    assert torch.equal(images.cpu(), torch.tensor(g['f1c_age_images'][index])) and torch.equal(y.cpu(), torch.tensor(labels[index]))
    driving = crowd_data.ImageLabelStore(np.ascontiguousarray(hwc.transpose(0, 3, 1, 2)), labels, hwc=False)
    images, y = driving.batch(index)
    assert torch.equal(images.cpu(), torch.tensor(g['f1c_driving_images'][index]))
    assert torch.equal(y.cpu(), torch.tensor(g['f1c_driving_angles'][index]))
    # the loader is DataLoader(dataset, batch_size, shuffle=True)'s index stream (torch's own samplers) under the same seed
    from torch.utils.data import DataLoader, TensorDataset
    torch.manual_seed(5)
    want = [b[0] for b in DataLoader(TensorDataset(torch.arange(5)), batch_size=2, shuffle=True)]
    torch.manual_seed(5)
    got = list(age.loader(2, shuffle=True))
    assert len(got) == len(want) == 3
    for (im, yy), idx in zip(got, want):
        assert torch.equal(im.cpu(), torch.tensor(g['f1c_age_images'][idx.numpy()])) and torch.equal(yy.cpu(), torch.tensor(labels[idx.numpy()]))
    with pytest.raises(IndexError):
        age.batch([5])


def test_age_full_size_batch_vs_oracle():
    """BASELINE's age shapes: 100 samples of 3 x 128 x 128 from a 512-image resident store."""
    from srgan_b200 import crowd_data
    rng = np.random.RandomState(3)
    images = rng.randint(0, 256, size=(512, 128, 128, 3)).astype(np.uint8)
    labels = (rng.rand(512) * 85 + 10).astype(np.float32)
    store = crowd_data.ImageLabelStore(images, labels, hwc=True)
    index = rng.randint(0, 512, size=100)
    out, y = store.batch(index)
    for b in (0, 17, 99):
        want, lab = C.image_label_item(images[index[b]], labels[index[b]])
        assert np.array_equal(out[b].cpu().numpy(), want) and float(y[b]) == float(lab)
    assert torch.equal(y.cpu(), torch.tensor(labels[index]))


def test_world_expo_dataset_draw_for_draw(g):
    """WorldExpoTransformedDataset (cameras of equally sized frames, label = map) through the same CrowdStore + TransformedDataset:
    the reference's 20 samples under the same `random` seed, bit for bit."""
    from srgan_b200 import crowd_data
    from tests.test_oracle_crowd_data import world_expo_frames
    store = crowd_data.CrowdStore(world_expo_frames(g))
    ds = crowd_data.TransformedDataset(store, int(g['patch']), int(g['patch']))
    assert len(ds) == int(g['f1d_length'])
    random.seed(int(g['f1d_seed']))
    images, labels, maps = ds.batch(len(g['f1d_out_images']))
    assert torch.equal(images.cpu(), torch.tensor(g['f1d_out_images']))
    assert torch.equal(labels.cpu(), torch.tensor(g['f1d_out_labels'])) and torch.equal(maps.cpu(), torch.tensor(g['f1d_out_maps']))
