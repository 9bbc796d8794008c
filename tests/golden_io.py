"""Loads tests/golden/*.npz (made by oracle/make_golden.py from the unmodified reference) for the parity tests."""
import json
import os

import numpy as np
import torch

from oracle import srgan_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SCALARS = ('dnn_loss', 'labeled_loss', 'unlabeled_loss', 'fake_loss', 'gradient_penalty', 'gradient_norm_mean',
           'generator_loss')


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
        self.cfg = json.loads(bytes(self.z['config_json']).decode())
        self.steps = self.cfg['steps']

    def group(self, prefix, dtype=torch.float32):
        n = len(prefix) + 1
        return {k[n:]: torch.tensor(self.z[k]).to(dtype) for k in self.z.files if k.startswith(prefix + '/')}

    def step_inputs(self, i, dtype=torch.float32):
        return tuple(torch.tensor(self.z[f'step{i}/{k}']).to(dtype) for k in ('x', 'y', 'u', 'z', 'alpha', 'z2'))

    def scalars(self, i):
        return {k: float(self.z[f'step{i}/scalars/{k}']) for k in SCALARS}

    def step_config(self) -> O.StepConfig:
        c = O.StepConfig()
        for k, v in self.cfg.items():
            if hasattr(c, k):
                setattr(c, k, tuple(v) if isinstance(v, list) else v)
        return c

    def oracle_state(self, dtype=torch.float32) -> O.OracleState:
        fam = self.cfg['family']
        dggan = self.cfg['method'] == 'dggan'
        if fam == 'coefficient':
            ds, gs = O.ModelSpec('coefficient', dggan=dggan), O.ModelSpec('coefficient')
        else:
            ds, gs = O.ModelSpec('dcgan', leaky=0.05), O.ModelSpec('dcgan', leaky=0.05)
        return O.OracleState(ds, gs, self.group('init/D', dtype), self.group('init/G', dtype),
                             self.group('init/DNN', dtype))


ALL = ('coefficient_srgan', 'coefficient_srgan_altdist', 'coefficient_dggan', 'dcgan_mini', 'dcgan_sgan_mini', 'coefficient_sgan')
