"""
Stages the UNMODIFIED reference under baseline/_ref/ so that it travels to the GPU box  --  TEST / BENCH INFRASTRUCTURE.

`pip install --no-index --target baseline/_ref /root/reference` is what the bench contract asks for; it fails ("Neither
'setup.py' nor 'pyproject.toml' found": the reference is a script tree), so the hot-path modules are staged file by file
instead: byte-identical copies of the reference's own *.py files, nothing edited, nothing generated.  baseline/_ref/ is
listed in .gitignore (never committed: the repository holds no reference source) but not in .gpurunignore, exactly like
the built .so files.  Users: bench.py --impl reference / cpu_baseline / gpu_baseline (the reference's own step, timed),
tests/test_gpu_mixin_reference.py (B200StepMixin composed with the reference's Experiment classes on the GPU).

Run here (build container): python oracle/stage_reference.py      (also called by __graft_entry__.build()).
"""
from __future__ import annotations

import filecmp
import glob
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('SRGAN_REFERENCE_SRC', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')
# srgan.py imports the application packages' experiments lazily through run.py only; the step itself needs utility /
# settings / srgan / dnn plus the application packages whose Experiment subclasses and models are on the hot path
PATTERNS = ['utility.py', 'settings.py', 'srgan.py', 'dnn.py', 'sgan.py', 'coefficient/*.py', 'age/*.py', 'driving/*.py',
            'crowd/*.py']


def staged_root():
    """Directory holding an importable reference tree: the checkout when it is mounted, else the staged copy, else None."""
    for p in (SRC, DST):
        if os.path.isfile(os.path.join(p, 'srgan.py')):
            return p
    return None


def stage(verbose=False):
    if not os.path.isfile(os.path.join(SRC, 'srgan.py')):
        return DST if os.path.isfile(os.path.join(DST, 'srgan.py')) else None
    n = 0
    for pat in PATTERNS:
        for src in sorted(glob.glob(os.path.join(SRC, pat))):
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
                n += 1
    if verbose:
        print(f'staged {n} changed files under {DST}')
    return DST


if __name__ == '__main__':
    print(stage(verbose=True))
