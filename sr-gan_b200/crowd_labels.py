"""
Crowd label preprocessing on the device (second half): the reference's one-off, per-image label
generation (crowd/database_preprocessor.py:64-99) keeps a scikit-learn ball tree busy for seconds per image
(`generate_knn_map` queries every pixel, :258-290); here it is one brute-force float64 kernel per image.  Function names and
arguments are the reference's; results are device tensors.  No CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from .ops_cuda import load_library


def _heads(head_positions, device):
    h = torch.as_tensor(np.ascontiguousarray(head_positions, dtype=np.float64)).reshape(-1, 2)
    return h.to(device)


def _ck(lib, rc, name):
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {lib.srgan_last_error().decode()}')


def generate_knn_maps(head_positions, label_size, k_max=5, upper_bound=None, epsilon=1.0, device='cuda:0', inverse=True):
    """All of generate_labels_for_example's k-NN maps in one sweep (crowd/database_preprocessor.py:90-99): returns
    (knn [k_max, H, W] float64 = generate_knn_map(.., number_of_neighbors=k) for k = 1..k_max,
     iknn [k_max, H, W] float16 = (1 / (knn + epsilon)).astype(float16), what the `i{k}nn_maps` directories hold)."""
    if not torch.cuda.is_available():
        raise RuntimeError('crowd_labels needs a CUDA device (no CPU path)')
    lib, dev = load_library(), torch.device(device)
    heads = _heads(head_positions, dev)
    if heads.shape[0] == 0:
        raise ValueError('no head positions (sklearn NearestNeighbors.fit raises on an empty array in the reference)')
    H, W = int(label_size[0]), int(label_size[1])
    knn = torch.empty(k_max, H, W, device=dev, dtype=torch.float64)
    iknn = torch.empty(k_max, H, W, device=dev, dtype=torch.float16) if inverse else None
    _ck(lib, lib.srgan_knn_maps(heads.data_ptr(), heads.shape[0], H, W, k_max, float(upper_bound or 0.0), float(epsilon),
                                knn.data_ptr(), None if iknn is None else iknn.data_ptr(),
                                torch.cuda.current_stream(dev).cuda_stream), 'srgan_knn_maps')
    return knn, iknn


def generate_knn_map(head_positions, label_size, number_of_neighbors=1, upper_bound=None, device='cuda:0'):
    """crowd/database_preprocessor.py:258-290."""
    knn, _ = generate_knn_maps(head_positions, label_size, number_of_neighbors, upper_bound, device=device, inverse=False)
    return knn[number_of_neighbors - 1]


def generate_point_density_map(head_positions, label_size, device='cuda:0'):
    """crowd/database_preprocessor.py:246-256: (density map [H, W] fp32 on the device, out-of-bounds count)."""
    if not torch.cuda.is_available():
        raise RuntimeError('crowd_labels needs a CUDA device (no CPU path)')
    lib, dev = load_library(), torch.device(device)
    heads = _heads(head_positions, dev)
    H, W = int(label_size[0]), int(label_size[1])
    density = torch.empty(H, W, device=dev, dtype=torch.float32)
    oob = torch.empty(1, device=dev, dtype=torch.int32)
    _ck(lib, lib.srgan_point_density_map(heads.data_ptr() if heads.numel() else density.data_ptr(), heads.shape[0], H, W,
                                         density.data_ptr(), oob.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
        'srgan_point_density_map')
    return density, int(oob.item())


def generate_density_label(head_positions, label_size, neighbor_deviation_beta=0.15, device='cuda:0', half=False):
    """crowd/database_preprocessor.py:113-225 in the form generate_labels_for_example uses (:87-88: no perspective map, yx order,
    count-normalised): the geometry-adaptive Gaussian density label, [H, W] fp32 on the device (half=True: also the float16
    copy the `density{beta}` directories hold -- the maps run.py:67 trains the crowd application on)."""
    if not torch.cuda.is_available():
        raise RuntimeError('crowd_labels needs a CUDA device (no CPU path)')
    lib, dev = load_library(), torch.device(device)
    heads = _heads(head_positions, dev)
    if heads.shape[0] == 0:
        raise ValueError('no head positions (sklearn NearestNeighbors.fit raises on an empty array in the reference)')
    H, W = int(label_size[0]), int(label_size[1])
    label = torch.empty(H, W, device=dev, dtype=torch.float32)
    f16 = torch.empty(H, W, device=dev, dtype=torch.float16) if half else None
    ws_bytes = lib.srgan_density_label_workspace_bytes(heads.shape[0], H, W)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    _ck(lib, lib.srgan_density_label(heads.data_ptr(), heads.shape[0], H, W, float(neighbor_deviation_beta), label.data_ptr(),
                                     None if f16 is None else f16.data_ptr(), ws.data_ptr(), ws_bytes,
                                     torch.cuda.current_stream(dev).cuda_stream), 'srgan_density_label')
    return (label, f16) if half else label
