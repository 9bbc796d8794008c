import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests are the parity tests proper and need the B200: on a machine without a CUDA device (or without the
    built library) they are skipped, never run against anything else."""
    import torch
    lib = os.path.join(ROOT, 'sr-gan_b200', 'libsrgan_b200.so')
    why = None
    if not torch.cuda.is_available():
        why = 'no CUDA device'
    elif not os.path.exists(lib):
        why = 'libsrgan_b200.so is not built'
    if why is None:
        return
    skip = pytest.mark.skip(reason=why)
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
