"""CPU: the reference arm of bench.py (`--impl reference`: the staged reference, else the oracle port, timed on the host cores) prints ONE JSON line with
the keys the driver reads, on the metric / unit / config of the product arm.  The product arm itself needs a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'coefficient',
                          '--steps', '2', '--warmup', '3'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'SR-GAN train steps/sec' and d['unit'] == 'steps/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 2 and d['vs_baseline'] is None
    assert d['value'] > 0 and abs(d['value'] - 1e3 / d['ms_per_step']) / d['value'] < 1e-6
    assert 'coefficient SR-GAN' in d['config']['workload'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['details']['full_batch_run'] is True and d['details']['sample_batch'] == 5000      # coefficient runs un-scaled
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_synthetic_batches_of_every_workload():
    """bench.make_batches: shapes / label kinds of SURVEY 8(d) configs 1-4 (the bench never reads a dataset)."""
    sys.path.insert(0, ROOT)
    import bench
    import torch
    for name in sorted(bench.WORKLOADS):
        x, y, u = bench.make_batches(name, 4, 7)
        assert x.shape == u.shape and x.shape[0] == 4 and x.dtype == torch.float32
        if name == 'crowd':
            density, label_map = y
            assert x.shape[1:] == (3, 224, 224) and density.shape == label_map.shape == (4, 224, 224)
            assert set(density.unique().tolist()) <= {0.0, 1.0} and 0 < label_map.min() and label_map.max() <= 1
        elif name in ('age', 'driving'):
            assert x.shape[1:] == (3, 128, 128) and y.shape == (4,) and x.abs().max() <= 1
        else:
            assert x.shape[1:] == (50,) and y.shape == (4,)
    ya = bench.make_batches('age', 256, 1)[1]
    yd = bench.make_batches('driving', 256, 1)[1]
    assert ya.min() >= 10 and ya.max() <= 95 and yd.min() < 0 < yd.max()
