"""Micro-bench of the device-side crowd input pipeline (rows f1 / f2): srgan_crowd_extract_patches at BASELINE's crowd shapes
(64 patches of 224 x 224 per batch from resident 768 x 1024 images) against the measured HBM copy bandwidth, with the
reference's host path (numpy crop / flip / normalise per sample + the host -> device copy of the batch) timed beside it;
then the sliding-window merge of one 768 x 1024 image.   python tools/crowd_data_bench.py [--batch 64] [--images 32]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from srgan_b200 import crowd_data  # noqa: E402
from oracle import crowd_data_oracle as C  # noqa: E402  (the CPU baseline leg only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--images', type=int, default=32)
    ap.add_argument('--iters', type=int, default=50)
    a = ap.parse_args()
    rng = np.random.RandomState(0)
    H, W, P = 768, 1024, 224
    ex = [(rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8), rng.rand(H, W).astype(np.float32),
           rng.rand(H, W).astype(np.float32)) for _ in range(a.images)]
    store = crowd_data.CrowdStore(ex)
    ds = crowd_data.TransformedDataset(store, P, P)
    tables = [torch.as_tensor(ds.draw(a.batch)).cuda() for _ in range(a.iters)]
    for t in tables[:5]:
        store.extract(t, P)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for t in tables:                                 # every batch gathers different windows of a 200 MB store (> L2 with the outputs)
        e0.record()
        store.extract(t, P)
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    algo = a.batch * P * P * (3 + 4 + 4 + 12 + 4 + 4)
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    peak = None
    for k, v in peaks.items():
        if 'hbm' in k.lower() and isinstance(v, (int, float)):
            peak = float(v)
    gbs = algo / ms / 1e6
    print(f'extract_patches B={a.batch} P={P}: {ms * 1e3:.1f} us/batch, {algo / 1e6:.1f} MB algorithmic, {gbs:.0f} GB/s'
          + (f' = {gbs / peak:.2f} of the measured HBM peak ({peak:.0f} GB/s)' if peak else ''))
    # end to end through the host API incl. the draws and the 1 KB table upload
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        ds.batch(a.batch)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / a.iters
    print(f'TransformedDataset.batch({a.batch}) end to end (draws + table upload + launch): {e2e * 1e3:.3f} ms/batch')
    # the reference's host path: numpy transforms per sample on one core (a DataLoader worker), then the batch's H2D copy
    t0 = time.perf_counter()
    n_cpu = 3
    for _ in range(n_cpu):
        tab = ds.draw(a.batch)
        items = []
        for f, y, x, flip in tab:
            im, lb, mp = C.extract_patch(*ex[f], int(y), int(x), P)
            im, lb, mp = C.random_horizontal_flip(im, lb, mp, bool(flip))
            items.append((C.to_chw_float32(C.normalize_image(im)), lb, mp))
        batch = [torch.from_numpy(np.stack([it[k] for it in items])).pin_memory() for k in range(3)]
        dev = [b.cuda(non_blocking=True) for b in batch]
        torch.cuda.synchronize()
    cpu = (time.perf_counter() - t0) / n_cpu
    print(f'reference host path (numpy per sample on 1 core + collate + H2D): {cpu * 1e3:.1f} ms/batch  -> {cpu / e2e:.0f}x')
    # sliding-window merge, 768 x 1024, patch 224, step 128 (settings.py:60-61)
    sw = crowd_data.SlidingWindow(H, W, P, 128)

    def net(images):
        n = images.shape[0]
        return torch.rand(n, P, P, device='cuda'), torch.rand(n, device='cuda'), None
    crowd_data.predict_full_example(store, 0, net)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        crowd_data.predict_full_example(store, 0, net)
    torch.cuda.synchronize()
    print(f'predict_full_example (extract + merge, {sw.length} windows, stand-in network): '
          f'{(time.perf_counter() - t0) / 10 * 1e3:.3f} ms/image')
    # label preprocessing (row f4): every pixel's 1..5-NN distance maps of a 768 x 1024 label with 1 500 heads
    from srgan_b200 import crowd_labels
    heads = rng.rand(1500, 2) * np.array([H, W], dtype=np.float64)
    crowd_labels.generate_knn_maps(heads, (H, W), 5)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        crowd_labels.generate_knn_maps(heads, (H, W), 5)
    e1.record()
    e1.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 5
    evals = H * W * 1500
    print(f'generate_knn_maps k=1..5, {H}x{W}, 1500 heads: {gpu_ms:.3f} ms = {evals / gpu_ms / 1e6:.0f} G distance evaluations/s (fp64)')
    try:                                                              # the reference's call: one ball-tree query per k
        from sklearn.neighbors import NearestNeighbors
        pos = np.stack(np.meshgrid(np.arange(H), np.arange(W), indexing='ij'), -1).reshape(-1, 2)
        t0 = time.perf_counter()
        NearestNeighbors(n_neighbors=5, algorithm='ball_tree').fit(heads).kneighbors(pos)
        cpu_s = time.perf_counter() - t0
        print(f'scikit-learn ball tree, ONE k=5 query of the same label on the host: {cpu_s:.2f} s (the preprocessor runs k=1..5: '
              f'5 queries) -> >= {cpu_s * 1e3 / gpu_ms:.0f}x')
    except ImportError:
        print('scikit-learn not importable here: no host timing')
    # the Gaussian density label (run.py:67 trains on the beta = 0.3 maps), same label and heads
    crowd_labels.generate_density_label(heads, (H, W), 0.3)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        crowd_labels.generate_density_label(heads, (H, W), 0.3)
    e1.record()
    e1.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 5
    from oracle import crowd_labels_oracle as LO                       # the reference's per-head numpy loop, restated
    t0 = time.perf_counter()
    LO.generate_density_label(heads, (H, W), 0.3)
    cpu_s = time.perf_counter() - t0
    print(f'generate_density_label beta=0.3, {H}x{W}, 1500 heads: {gpu_ms:.3f} ms on the device; the per-head numpy loop on the host '
          f'{cpu_s:.2f} s -> {cpu_s * 1e3 / gpu_ms:.0f}x')


if __name__ == '__main__':
    main()
