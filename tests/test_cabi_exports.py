"""CPU: the C-ABI library builds, loads, and exports every symbol include/srgan_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'srgan_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(srgan_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    sys.path.insert(0, os.path.join(ROOT, 'sr-gan_b200'))
    import build as srgan_build
    lib_path = srgan_build.build()
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/srgan_b200.h but not exported'
    lib.srgan_version.restype = ctypes.c_int
    assert lib.srgan_version() >= 100


def test_python_binding_covers_header():
    from srgan_b200 import ops_cuda
    assert set(ops_cuda._EXPORTS) == set(declared_symbols())


def test_product_path_refuses_cpu():
    """No CPU fallback: the ops object cannot be created without a CUDA device."""
    import pytest
    import torch
    from srgan_b200 import ops_cuda
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        ops_cuda.CudaOps()
    # the widened rows (input pipeline, label preprocessing) have no CPU path either
    import numpy as np
    from srgan_b200 import crowd_data, crowd_labels
    with pytest.raises(RuntimeError):
        crowd_data.CrowdStore([(np.zeros((8, 8, 3), np.uint8), None, None)])
    with pytest.raises(RuntimeError):
        crowd_data.ImageLabelStore(np.zeros((2, 8, 8, 3), np.uint8), np.zeros(2, np.float32))
    for fn in (crowd_labels.generate_knn_maps, crowd_labels.generate_density_label, crowd_labels.generate_point_density_map):
        with pytest.raises(RuntimeError):
            fn(np.ones((3, 2)), (8, 8))
