"""
Generates tests/golden/crowd_data.npz by running the UNMODIFIED reference's crowd data classes and inference helpers
(crowd/data.py, crowd/shanghai_tech_data.py, crowd/srgan.py) on small seeded synthetic examples  --  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference or the staged baseline/_ref):  python oracle/make_golden_data.py
The fixture pins oracle/crowd_data_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path
(tests/test_gpu_crowd_data.py).  Two stand-ins, both outside the arithmetic under test:
  * scipy.misc.imresize was removed from SciPy; predict_full_example calls it with the patch's own size (label 224 = patch 224
    in BASELINE's crowd configuration; here patch 32 = label 32), where it is the identity for mode 'F' -- the stand-in asserts
    the sizes match and returns its input;
  * ShanghaiTechTransformedDataset.__init__ lists a dataset directory of the (undownloadable) ShanghaiTech archive; the
    object is created without __init__ and given the attributes __init__ would have computed, over a temporary directory of
    synthetic .npy files in the reference's layout (images/ labels/ knn_maps/).
"""
from __future__ import annotations

import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness  # noqa: E402

PATCH = 32            # small stand-in for 224 (the arithmetic is size-independent; tests cover 224 through properties)
STEP = 12
SHAPES = [(50, 70), (32, 32), (20, 45), (64, 33), (41, 30)]       # incl. images smaller than the patch (padded windows)


def synthetic_examples(seed=5):
    rng = np.random.RandomState(seed)
    out = []
    for h, w in SHAPES:
        image = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        label = (rng.rand(h, w) < 0.02).astype(np.float32) * rng.rand(h, w).astype(np.float32)
        map_ = (1.0 / (1.0 + 50.0 * rng.rand(h, w))).astype(np.float32)
        out.append((image, label, map_))
    return out


def fake_network(images):
    """A deterministic stand-in for D: per-patch density label, count and three maps computed from the pixels."""
    images = torch.as_tensor(images)
    labels = images.mean(1) * 0.01 + 0.02
    counts = images.sum((1, 2, 3)) * 1e-3 + 1.5
    maps = torch.stack([images[:, 0] * 0.5, images[:, 1] * 0.25 + 0.1, images.abs().mean(1)], dim=1)
    return labels, counts, maps


def main():
    ref_harness.install_shims()
    import scipy.misc

    def imresize_same_size(array, size, mode=None):
        assert tuple(size) == tuple(array.shape) and mode == 'F', (size, array.shape, mode)
        return array
    scipy.misc.imresize = imresize_same_size
    from crowd import data as rdata
    from crowd.shanghai_tech_data import ShanghaiTechTransformedDataset
    from crowd.srgan import CrowdExperiment
    from torch.utils.data import Dataset

    examples = synthetic_examples()
    out = {'patch': np.int64(PATCH), 'step': np.int64(STEP), 'n_images': np.int64(len(examples))}
    for i, (image, label, map_) in enumerate(examples):
        out[f'image{i}'], out[f'label{i}'], out[f'map{i}'] = image, label, map_

    # ---- f1a: ExtractPatchForPosition(allow_padded=True) -> [flip] -> normalise -> tensors, explicit positions
    extract = rdata.ExtractPatchForPosition(PATCH, PATCH, allow_padded=True)
    rng = np.random.RandomState(11)
    seeds = {}
    for k in range(64):                                                # a seed for each outcome of random.choice([True, False])
        random.seed(k)
        seeds.setdefault(int(random.choice([True, False])), k)
    pos, imgs, labs, maps = [], [], [], []
    for i, (image, label, map_) in enumerate(examples):
        h, w = image.shape[:2]
        cands = [(0, 0), (h - 1, w - 1), (h // 2, w // 2), (PATCH // 2, PATCH // 2), (h - PATCH // 2, w - PATCH // 2)]
        cands += [(int(rng.randint(h)), int(rng.randint(w))) for _ in range(3)]
        for k, (y, x) in enumerate(cands):
            flip = (k + i) % 2
            ex = extract(rdata.CrowdExample(image=image, label=label, map_=map_), y, x)
            random.seed(seeds[flip])                                   # RandomHorizontalFlip with its draw known
            ex = rdata.RandomHorizontalFlip()(ex)
            ex = rdata.NumpyArraysToTorchTensors()(rdata.NegativeOneToOneNormalizeImage()(ex))
            pos.append((i, y, x, flip))
            imgs.append(ex.image.numpy()), labs.append(ex.label.numpy()), maps.append(ex.map.numpy())
    out['f1_pos'] = np.array(pos, dtype=np.int32)
    out['f1_images'], out['f1_labels'], out['f1_maps'] = np.stack(imgs), np.stack(labs), np.stack(maps)

    # ---- f1b: ShanghaiTechTransformedDataset.__getitem__ with its own random draws (random.seed(21))
    with tempfile.TemporaryDirectory() as tmp:
        for sub in ('images', 'labels', 'knn_maps'):
            os.makedirs(os.path.join(tmp, sub))
        names = []
        big_ids = [i for i, e in enumerate(examples) if e[0].shape[0] >= PATCH and e[0].shape[1] >= PATCH]
        big = [examples[i] for i in big_ids]
        for i, (image, label, map_) in enumerate(big):
            name = f'IMG_{i}.npy'
            names.append(name)
            np.save(os.path.join(tmp, 'images', name), image)
            np.save(os.path.join(tmp, 'labels', name), label)
            np.save(os.path.join(tmp, 'knn_maps', name), map_)
        ds = object.__new__(ShanghaiTechTransformedDataset)
        ds.dataset_directory, ds.file_names = tmp, names
        ds.image_patch_size = ds.label_patch_size = PATCH
        ds.middle_transform, ds.map_directory_name = rdata.RandomHorizontalFlip(), 'knn_maps'
        half, ds.length, ds.start_indexes = PATCH // 2, 0, []
        for image, _, _ in big:                                        # crowd/shanghai_tech_data.py:60-67
            ds.start_indexes.append(ds.length)
            ds.length += len(range(half, image.shape[0] - half + 1)) * len(range(half, image.shape[1] - half + 1))
        random.seed(21)
        items = [ds[k] for k in range(24)]
        out['f1b_store'] = np.array(big_ids, dtype=np.int64)
        out['f1b_seed'], out['f1b_length'] = np.int64(21), np.int64(ds.length)
        out['f1b_images'] = np.stack([t[0].numpy() for t in items])
        out['f1b_labels'] = np.stack([t[1].numpy() for t in items])
        out['f1b_maps'] = np.stack([t[2].numpy() for t in items])

    # ---- f2a: ImageSlidingWindowDataset positions + predict_full_example with the stand-in network
    exp = object.__new__(CrowdExperiment)
    exp.settings = type('S', (), dict(image_patch_size=PATCH, test_sliding_window_size=STEP, batch_size=7, pin_memory=False,
                                      number_of_data_workers=0))()
    for i, (image, label, map_) in enumerate(examples):
        sw = rdata.ImageSlidingWindowDataset(rdata.CrowdExample(image=image), PATCH, STEP)
        out[f'f2_ys{i}'] = np.array(sorted(sw.y_positions), dtype=np.int32)
        out[f'f2_xs{i}'] = np.array(sorted(sw.x_positions), dtype=np.int32)
        count, full_label = exp.predict_full_example(rdata.CrowdExample(image=image, label=label), fake_network)
        out[f'f2_count{i}'], out[f'f2_label{i}'] = np.float32(count), full_label.astype(np.float32)

    # ---- f2b: evaluation_epoch's scalars (ME, MAE, kNN MAE, MSE, kNN MSE) over 3 batches of 5 patches
    class Patches(Dataset):
        def __len__(self):
            return 15

        def __getitem__(self, k):
            return (torch.as_tensor(out['f1_images'][k]), torch.as_tensor(out['f1_labels'][k]),
                    torch.as_tensor(out['f1_maps'][k]))
    exp.settings.batch_size = 5
    writer = ref_harness._RecordingWriter()
    mae = exp.evaluation_epoch(exp.settings, fake_network, Patches(), writer, 'Validation', shuffle=False)
    for tag in ('ME', 'MAE', 'kNN MAE', 'MSE', 'kNN MSE'):
        out['f2b_' + tag.replace(' ', '_')] = np.float64(writer.scalars[f'Validation/{tag}'][-1][1])
    assert abs(mae - out['f2b_MAE']) < 1e-12

    # ---- f1c: age / driving samples.  SteeringAngleDataset.__getitem__ (driving/data.py:44-51) runs as is over CHW .npy files;
    # AgeDataset.__getitem__ (age/data.py:52-60) needs imageio to decode a JPEG, so its remaining lines are run here on the
    # decoded array with the reference's own utility.to_normalized_range
    from driving.data import SteeringAngleDataset
    from utility import to_normalized_range
    rng = np.random.RandomState(31)
    hwc = rng.randint(0, 256, size=(5, 16, 16, 3)).astype(np.uint8)
    hwc[0] = np.arange(256, dtype=np.uint8).repeat(3).reshape(16, 16, 3)        # every byte value
    values = (rng.rand(5) * 90 - 45).astype(np.float32)
    out['f1c_hwc'], out['f1c_labels'] = hwc, values
    age_items = []
    for image in hwc:
        t = torch.tensor(image.transpose((2, 0, 1)).astype(np.float32))
        age_items.append(to_normalized_range(t).numpy())
    out['f1c_age_images'] = np.stack(age_items)
    with tempfile.TemporaryDirectory() as tmp:
        names = []
        for k, image in enumerate(hwc):
            names.append(f'{k}.jpg')
            np.save(os.path.join(tmp, f'{k}.npy'), np.ascontiguousarray(image.transpose((2, 0, 1))))
        ds = object.__new__(SteeringAngleDataset)
        ds.dataset_path, ds.image_names, ds.angles, ds.length = tmp, np.array(names), values, len(names)
        items = [ds[k] for k in range(len(names))]
        out['f1c_driving_images'] = np.stack([t[0].numpy() for t in items])
        out['f1c_driving_angles'] = np.stack([t[1].numpy() for t in items])

    # ---- f1d: WorldExpoTransformedDataset.__getitem__ (crowd/world_expo_data.py:124-162): cameras hold stacks of equally sized
    # frames, the flat index decomposes into (camera, frame, position) and the label doubles as the map (:146).  The object is
    # built without __init__ (which reads the undownloadable archive) from two synthetic cameras.
    from crowd.world_expo_data import WorldExpoTransformedDataset, CameraData
    rng = np.random.RandomState(41)
    cameras = []
    for c, (frames, h, w) in enumerate(((3, 40, 48), (2, 36, 64))):
        images = rng.randint(0, 256, size=(frames, h, w, 3)).astype(np.uint8)
        labels = rng.rand(frames, h, w).astype(np.float32)
        out[f'f1d_images{c}'], out[f'f1d_labels{c}'] = images, labels
        cameras.append(CameraData(images=images, labels=labels, roi=None, perspective=None))
    ds = object.__new__(WorldExpoTransformedDataset)
    ds.camera_data_list, ds.image_patch_size, ds.label_patch_size = cameras, PATCH, PATCH
    ds.middle_transform = rdata.RandomHorizontalFlip()
    half, ds.length, ds.start_indexes = PATCH // 2, 0, []
    for cam in cameras:                                                # crowd/world_expo_data.py:113-118
        per_image = len(range(half, cam.images.shape[1] - half + 1)) * len(range(half, cam.images.shape[2] - half + 1))
        ds.start_indexes.append(ds.length)
        ds.length += cam.images.shape[0] * per_image
    random.seed(23)
    items = [ds[k] for k in range(20)]
    out['f1d_seed'], out['f1d_length'], out['f1d_cameras'] = np.int64(23), np.int64(ds.length), np.int64(len(cameras))
    out['f1d_out_images'] = np.stack([t[0].numpy() for t in items])
    out['f1d_out_labels'] = np.stack([t[1].numpy() for t in items])
    out['f1d_out_maps'] = np.stack([t[2].numpy() for t in items])

    path = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'crowd_data.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), 'bytes;', len(pos), 'patches,', len(examples), 'full examples')


if __name__ == '__main__':
    main()
