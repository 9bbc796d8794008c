"""
Device-side crowd input pipeline, sliding-window inference and evaluation sums.

The reference feeds `training_loop` (srgan.py:107-117) from 4-worker DataLoaders whose `__getitem__`
(crowd/shanghai_tech_data.py:73-104) memory-maps a full image, crops / pads / flips / normalises it in numpy, and copies
every batch host -> device.  Here the split's full examples stay RESIDENT in HBM (`CrowdStore`; 300 ShanghaiTech part-A
images are ~2.5 GB of the 180 GB) and a batch is one launch of `srgan_crowd_extract_patches` driven by a [B,4] int32
position table -- the only bytes that cross PCIe per step.  The random draws are the reference's (`random.randrange` for the
position, `random.choice` for the flip, in that order per sample), so with the same `random` state and a worker-less
reference DataLoader the batches are bit-identical (tests/test_gpu_crowd_data.py).

Names follow the reference: `TransformedDataset` ~ ShanghaiTechTransformedDataset / UcfQnrfTransformedDataset,
`SlidingWindow` ~ ImageSlidingWindowDataset (crowd/data.py:521-560), `predict_full_example` / `evaluation_epoch` ~
CrowdExperiment's methods (crowd/srgan.py:149-191,332-395).  Everything computes through libsrgan_b200.so; there is no CPU
path (a tensor that is not on the CUDA device raises).
"""
from __future__ import annotations

import os
import random as _random

import numpy as np
import torch

from .ops_cuda import load_library


def _ck(lib, rc, name):
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {lib.srgan_last_error().decode()}')


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class CrowdStore:
    """Full examples of one dataset split, resident on the device: images uint8 HWC, labels / maps fp32 HW, concatenated
    image after image, plus the per-image table (pixel offset, height, width) the kernels index with."""

    def __init__(self, examples, device='cuda:0'):
        if not torch.cuda.is_available():
            raise RuntimeError('CrowdStore needs a CUDA device; the input pipeline has no CPU path')
        self.lib = load_library()
        self.device = torch.device(device)
        examples = list(examples)
        if not examples:
            raise ValueError('empty dataset split')
        has_label = examples[0][1] is not None
        has_map = len(examples[0]) > 2 and examples[0][2] is not None
        self.shapes, offsets, total = [], [], 0
        for e in examples:
            image = e[0]
            if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
                raise TypeError('images must be uint8 [H, W, 3] arrays (the preprocessed .npy files of the reference)')
            for other in e[1:3]:
                if other is not None and tuple(other.shape) != tuple(image.shape[:2]):
                    raise ValueError(f'label / map shape {other.shape} does not match the image {image.shape[:2]}')
            self.shapes.append((int(image.shape[0]), int(image.shape[1])))
            offsets.append(total)
            total += image.shape[0] * image.shape[1]
        self.pixels = total

        def upload(index, dtype, tail):
            host = torch.empty((total,) + tail, dtype=dtype, pin_memory=True)
            view = host.numpy()
            for e, off, (h, w) in zip(examples, offsets, self.shapes):
                view[off:off + h * w] = np.asarray(e[index]).reshape((h * w,) + tail)
            return host.to(self.device, non_blocking=True)
        self.images = upload(0, torch.uint8, (3,))
        self.labels = upload(1, torch.float32, ()) if has_label else None
        self.maps = upload(2, torch.float32, ()) if has_map else None
        self.pixel_offset = torch.tensor(offsets, dtype=torch.int64).to(self.device)
        self.heights = torch.tensor([h for h, _ in self.shapes], dtype=torch.int32).to(self.device)
        self.widths = torch.tensor([w for _, w in self.shapes], dtype=torch.int32).to(self.device)
        torch.cuda.current_stream(self.device).synchronize()          # the pinned staging buffers die with this frame

    @classmethod
    def from_directory(cls, dataset_directory, map_directory_name='knn_maps', number_of_examples=None, device='cuda:0'):
        """The reference's on-disk layout (crowd/shanghai_tech_data.py:24-41): <dir>/{images,labels,<maps>}/<name>.npy."""
        names = [n for n in os.listdir(os.path.join(dataset_directory, 'labels')) if n.endswith('.npy')][:number_of_examples]
        store = cls(((np.load(os.path.join(dataset_directory, 'images', n)), np.load(os.path.join(dataset_directory, 'labels', n)),
                      np.load(os.path.join(dataset_directory, map_directory_name, n))) for n in names), device=device)
        store.file_names = names
        return store

    def __len__(self):
        return len(self.shapes)

    def _upload_table(self, table):
        """Host position table -> device through one of two persistent pinned staging buffers (allocating pinned memory per
        batch cost more than the gather itself); a buffer is reused only after the copy that last read it has completed."""
        n = table.shape[0]
        ring = self.__dict__.setdefault('_staging', [None, None])
        self._turn = (getattr(self, '_turn', 0) + 1) % 2
        slot = ring[self._turn]
        if slot is None or slot[0].shape[0] < n:
            slot = ring[self._turn] = (torch.empty(max(n, 256), 4, dtype=torch.int32, pin_memory=True), torch.cuda.Event())
        else:
            slot[1].synchronize()
        slot[0][:n].copy_(table)
        dev = slot[0][:n].to(self.device, non_blocking=True)
        slot[1].record(torch.cuda.current_stream(self.device))
        return dev

    def extract(self, positions, patch, with_labels=True):
        """positions: [B,4] int32 {image index, y, x, flip} (host array or device tensor).  Returns (images [B,3,P,P],
        labels [B,P,P] | None, maps [B,P,P] | None), fp32 on the device: the tensors `training_loop` hands to the step."""
        if not torch.is_tensor(positions):
            positions = torch.as_tensor(np.ascontiguousarray(positions, dtype=np.int32))
        if positions.dtype != torch.int32 or positions.ndim != 2 or positions.shape[1] != 4:
            raise TypeError('positions must be an int32 [B, 4] table of {image, y, x, flip}')
        if not positions.is_cuda:
            bad = (positions[:, 0] < 0) | (positions[:, 0] >= len(self))
            if bool(bad.any()):
                raise IndexError(f'image index out of range for a store of {len(self)} images')
            positions = self._upload_table(positions)
        B = positions.shape[0]
        f32 = torch.float32
        images = torch.empty(B, 3, patch, patch, device=self.device, dtype=f32)
        labels = torch.empty(B, patch, patch, device=self.device, dtype=f32) if with_labels and self.labels is not None else None
        maps = torch.empty(B, patch, patch, device=self.device, dtype=f32) if with_labels and self.maps is not None else None
        p = lambda t: None if t is None else t.data_ptr()
        _ck(self.lib, self.lib.srgan_crowd_extract_patches(
            p(self.images), p(self.labels if labels is not None else None), p(self.maps if maps is not None else None),
            p(self.pixel_offset), p(self.heights), p(self.widths), len(self), p(positions.contiguous()), B, patch, p(images),
            p(labels), p(maps), _stream(self.device)), 'srgan_crowd_extract_patches')
        return images, labels, maps


class TransformedDataset:
    """crowd/shanghai_tech_data.py:48-107 on a CrowdStore: `length` = number of valid patch centres over all images,
    `start_indexes` = first flat index of each image; a sample is a uniformly random centre (the index argument of the
    reference's __getitem__ is ignored there too, :80) and a fair-coin horizontal flip."""

    def __init__(self, store: CrowdStore, image_patch_size=224, label_patch_size=224, flip=True, rng=None):
        if label_patch_size != image_patch_size:
            raise NotImplementedError('label_patch_size != image_patch_size needs scipy.misc.imresize (crowd/data.py:463-483), '
                                      'which SciPy removed; BASELINE\'s crowd configuration uses 224 / 224')
        if image_patch_size % 4:
            raise ValueError('image_patch_size must be a multiple of 4')
        self.store, self.image_patch_size, self.flip = store, image_patch_size, flip
        self.rng = rng if rng is not None else _random                 # the reference draws from the global `random`
        half = image_patch_size // 2
        self.counts = [(len(range(half, h - half + 1)), len(range(half, w - half + 1))) for h, w in store.shapes]
        self.start_indexes, self.length = [], 0
        for ny, nx in self.counts:
            self.start_indexes.append(self.length)
            self.length += ny * nx
        if self.length == 0:
            raise ValueError('no image of the split is as large as the patch')

    def __len__(self):
        return self.length

    def position(self, index_):
        """Flat index -> (image, y, x): searchsorted(side='right') - 1 over start_indexes, then row-major unravel (:81-97)."""
        f = int(np.searchsorted(self.start_indexes, index_, side='right') - 1)
        half = self.image_patch_size // 2
        yi, xi = divmod(index_ - self.start_indexes[f], self.counts[f][1])
        return f, half + yi, half + xi

    def draw(self, batch_size):
        """The reference's random draws for `batch_size` consecutive __getitem__ calls -> [B,4] int32 position table."""
        table = np.empty((batch_size, 4), dtype=np.int32)
        for b in range(batch_size):
            f, y, x = self.position(self.rng.randrange(self.length))
            table[b] = (f, y, x, int(self.rng.choice([True, False])) if self.flip else 0)
        return table

    def batch(self, batch_size):
        return self.store.extract(self.draw(batch_size), self.image_patch_size)

    def loader(self, batch_size):
        """Iterates like DataLoader(dataset, batch_size) (crowd/srgan.py:59-61): ceil(length / batch_size) batches, the last
        one short."""
        for start in range(0, self.length, batch_size):
            yield self.batch(min(batch_size, self.length - start))


class SlidingWindow:
    """crowd/data.py:521-560: window centres every `window_step_size` pixels from patch/2 plus the last full window along
    each axis (an image smaller than the patch gets one padded window); patch index = yi * len(x_positions) + xi."""

    def __init__(self, height, width, image_patch_size=224, window_step_size=128):
        half = image_patch_size // 2

        def axis(extent):
            positions = list(range(half, extent - half + 1, window_step_size))
            if extent - half > 0:
                positions = sorted(set(positions + [extent - half]))
            return positions
        self.y_positions, self.x_positions = axis(height), axis(width)
        self.image_patch_size = image_patch_size
        self.length = len(self.y_positions) * len(self.x_positions)

    def table(self, image_index=0):
        return np.array([(image_index, y, x, 0) for y in self.y_positions for x in self.x_positions], dtype=np.int32)


def predict_full_example(store: CrowdStore, image_index, network, image_patch_size=224, window_step_size=128, batch_size=64):
    """CrowdExperiment.predict_full_example (crowd/srgan.py:332-395) for image `image_index` of `store`.
    `network(images [n,3,P,P] cuda fp32) -> (labels [n,P,P] | None, counts [n], maps)`; None labels = the zeros
    KnnDenseNetCat returns (crowd/models.py:1153).  Returns (count: 0-d float64 device tensor, label [H,W] fp32 device)."""
    lib, dev = store.lib, store.device
    H, W = store.shapes[image_index]
    sw = SlidingWindow(H, W, image_patch_size, window_step_size)
    if sw.length == 0:
        # a side of at most patch/2 pixels: the reference's position lists are empty (crowd/data.py:530-537), no patch is
        # predicted and its sums stay zero (crowd/srgan.py:345-347,391-394)
        return torch.zeros((), device=dev, dtype=torch.float64), torch.zeros(H, W, device=dev, dtype=torch.float32)
    table = torch.as_tensor(sw.table(image_index)).pin_memory().to(dev, non_blocking=True)
    counts = torch.empty(sw.length, device=dev, dtype=torch.float32)
    labels = None
    for start in range(0, sw.length, batch_size):
        images, _, _ = store.extract(table[start:start + batch_size], image_patch_size, with_labels=False)
        out = network(images)
        n = images.shape[0]
        counts[start:start + n] = out[1].reshape(n).to(torch.float32)
        if out[0] is not None:
            if labels is None:
                labels = torch.zeros(sw.length, image_patch_size, image_patch_size, device=dev, dtype=torch.float32)
            labels[start:start + n] = out[0].reshape(n, image_patch_size, image_patch_size)
    ys = torch.tensor(sw.y_positions, dtype=torch.int32).to(dev)
    xs = torch.tensor(sw.x_positions, dtype=torch.int32).to(dev)
    full_label = torch.empty(H, W, device=dev, dtype=torch.float32)
    full_count = torch.empty((), device=dev, dtype=torch.float64)
    ws_bytes = lib.srgan_sliding_window_workspace_bytes()
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    _ck(lib, lib.srgan_sliding_window_merge(None if labels is None else labels.data_ptr(), counts.data_ptr(), ys.data_ptr(),
                                            len(sw.y_positions), xs.data_ptr(), len(sw.x_positions), H, W, image_patch_size,
                                            full_label.data_ptr(), full_count.data_ptr(), ws.data_ptr(), ws_bytes,
                                            _stream(dev)), 'srgan_sliding_window_merge')
    return full_count, full_label


def eval_sums(densities, predicted_maps, maps):
    """srgan_crowd_eval_sums: float64 per-sample sums [3, n] = {sum(density), sum |map_hat - map|, sum (map_hat - map)^2}."""
    ref = densities if densities is not None else maps
    lib, dev, n = load_library(), ref.device, ref.shape[0]
    HW = ref[0].numel()
    for t in (densities, predicted_maps, maps):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            raise TypeError('eval_sums needs contiguous fp32 CUDA tensors')
    out = torch.zeros(3, n, device=dev, dtype=torch.float64)
    ws_bytes = lib.srgan_crowd_eval_workspace_bytes(n)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    nmaps = 0 if predicted_maps is None else predicted_maps.shape[1]
    p = lambda t: None if t is None else t.data_ptr()
    _ck(lib, lib.srgan_crowd_eval_sums(p(densities), p(predicted_maps), nmaps, p(maps), n, HW, out.data_ptr(), ws.data_ptr(),
                                       ws_bytes, _stream(dev)), 'srgan_crowd_eval_sums')
    return out


def evaluation_epoch(network, batches, batch_size, summary_writer=None, summary_name='Validation', comparison_value=None):
    """CrowdExperiment.evaluation_epoch (crowd/srgan.py:149-191) over device batches (images, labels, maps): the count
    errors accumulate over the batches until `index * batch_size >= 100` (:175), the kNN map errors use the FIRST batch only
    (the reference concatenates maps / predicted maps only while they are empty, :167-172).  Returns the scalars it writes
    (and writes them to `summary_writer` under the reference's tags when one is given)."""
    predicted, truth, map_sums, map_elems = [], [], None, 0
    for index, (images, labels, maps) in enumerate(batches):
        _, counts, predicted_maps = network(images)
        pm = predicted_maps.to(torch.float32).contiguous() if index == 0 else None
        sums = eval_sums(labels.contiguous(), pm, maps.contiguous() if index == 0 else None)
        predicted.append(counts.reshape(-1).to(torch.float64))
        truth.append(sums[0])
        if index == 0:
            map_sums, map_elems = sums[1:].sum(1), pm.numel()
        if index * batch_size >= 100:
            break
    err = torch.cat(predicted) - torch.cat(truth)
    out = torch.stack([err.mean(), err.abs().mean(), map_sums[0] / map_elems, (err.abs() ** 2).mean(),
                       map_sums[1] / map_elems]).tolist()              # ONE device -> host read
    scalars = dict(zip(('ME', 'MAE', 'kNN MAE', 'MSE', 'kNN MSE'), out))
    if comparison_value is not None:
        scalars['Ratio MAE GAN DNN'] = scalars['MAE'] / comparison_value
    if summary_writer is not None:
        for tag, v in scalars.items():
            summary_writer.add_scalar(f'{summary_name}/{tag}', v)
    return scalars


class ImageLabelStore:
    """Age / driving datasets resident on the device (age/data.py:21-60 AgeDataset, driving/data.py:22-51
    SteeringAngleDataset): `images` [N,H,W,3] (imageio order, age) or [N,3,H,W] (the driving .npy files) uint8, `labels` [N]
    fp32.  `batch(index)` is one launch of srgan_image_batch: the reference's __getitem__ + default collate + `.to(gpu)` for the
    samples `index`; `loader(batch_size, shuffle)` walks the dataset with torch's own RandomSampler / BatchSampler, i.e. the
    index stream of `DataLoader(dataset, batch_size, shuffle=True)` (age/srgan.py:27-35) under the same torch seed."""

    def __init__(self, images, labels, hwc=True, device='cuda:0'):
        if not torch.cuda.is_available():
            raise RuntimeError('ImageLabelStore needs a CUDA device; the input pipeline has no CPU path')
        self.lib, self.device, self.hwc = load_library(), torch.device(device), bool(hwc)
        images = np.ascontiguousarray(images)
        if images.dtype != np.uint8 or images.ndim != 4 or images.shape[3 if hwc else 1] != 3:
            raise TypeError('images must be uint8 [N,H,W,3] (hwc=True) or [N,3,H,W] (hwc=False)')
        self.n = images.shape[0]
        self.chw = (3,) + (tuple(images.shape[1:3]) if hwc else tuple(images.shape[2:4]))
        labels = np.ascontiguousarray(labels, dtype=np.float32)
        if labels.shape != (self.n,):
            raise ValueError('one label per image')
        self.images = torch.from_numpy(images).pin_memory().to(self.device, non_blocking=True)
        self.labels = torch.from_numpy(labels).to(self.device)
        torch.cuda.current_stream(self.device).synchronize()

    def __len__(self):
        return self.n

    def batch(self, index):
        index = torch.as_tensor(index, dtype=torch.int64)
        if not index.is_cuda:
            if index.numel() and (int(index.min()) < 0 or int(index.max()) >= self.n):
                raise IndexError(f'sample index out of range for {self.n} images')
            index = index.to(self.device)
        B, (C, H, W) = index.numel(), self.chw
        out = torch.empty(B, C, H, W, device=self.device, dtype=torch.float32)
        labels = torch.empty(B, device=self.device, dtype=torch.float32)
        _ck(self.lib, self.lib.srgan_image_batch(self.images.data_ptr(), int(self.hwc), index.data_ptr(), B, C, H, W,
                                                 out.data_ptr(), self.labels.data_ptr(), labels.data_ptr(),
                                                 _stream(self.device)), 'srgan_image_batch')
        return out, labels

    def loader(self, batch_size, shuffle=True, drop_last=False):
        from torch.utils.data import BatchSampler, RandomSampler, SequentialSampler
        # a DataLoader iterator draws its base seed from the default generator before its sampler draws its own
        # (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__): consume that draw so the index stream is the same
        torch.empty((), dtype=torch.int64).random_()
        sampler = RandomSampler(range(self.n)) if shuffle else SequentialSampler(range(self.n))
        for index in BatchSampler(sampler, batch_size, drop_last):
            yield self.batch(index)
