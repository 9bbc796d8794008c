"""
bench.py -- SR-GAN training steps/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--batch B]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     (the oracle port of the reference step on the host CPU cores)

Workload (config.workload): BASELINE configs[1] "age SR-GAN": DCGAN G/D (age/models.py:32-80), synthetic 3x128x128
inputs ~U(-1,1), labels ~U(10,95), per-GPU batch 100, multipliers of run.py:30-35; one step = dnn_training_step +
gan_training_step with generator_training_step_period=1.  N>1 is weak scaling: every rank holds a
100-sample shard of a global batch 100*N, feature sums and gradients are all-reduced (NCCL) so the loss is the
global-batch loss.

`value`   : steps/s with the step's inputs already resident in HBM (CUDA events, max over ranks).
`e2e`     : steps/s through the public API (srgan_b200.Experiment.*_training_step) with HOST (pinned) input buffers:
            H2D copy of x, y, u every step and a D2H read of the step's scalars inside the timed region.
`roofline`: dominant kernel (the layer-2 discriminator conv over the 4B-row batch) timed live with CUDA events inside
            the timed region; algorithmic FLOPs / duration against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port (PyTorch fp32 autograd on the host cores) on a bounded sample, rank 0, N=1 only.
Inputs total 2 x 19.7 MB per step and activations ~1 GB per step: far larger than L2 (126 MB), so no L2 flush is needed
between iterations (stated in config.l2).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'SR-GAN train steps/sec'
AGE = dict(image=128, conv_dim=64, z_dim=256, batch=100, matching=1e2, contrasting=1e1, gp=1e2)
# algorithmic FLOPs per sample per step = 21 F_D + 4 F_G
F_D, F_G = 0.8305e9, 0.8472e9
FLOPS_PER_SAMPLE = 21 * F_D + 4 * F_G


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms',
                                       '100', '-i', str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if val.strip().lower() == 'active':
                    reasons.add(name)
        if sm:
            sm.sort()
            # "under load" = upper half of the samples (idle samples before/after the region read low)
            load = sm[len(sm) // 2:]
            out.update(sm_mhz=load[len(load) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def make_batches(B, seed, image):
    import torch
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, image, image, generator=gen) * 2 - 1
    u = torch.rand(B, 3, image, image, generator=gen) * 2 - 1
    y = torch.rand(B, generator=gen) * 85 + 10
    return x, y, u


def run_reference(args):
    """The reference's algorithm on the host CPU (oracle port; the reference checkout does not travel to the GPU box)."""
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import srgan_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = args.ref_batch
    st = O.init_dcgan(seed=0, image_size=AGE['image'], conv_dim=AGE['conv_dim'], z_dim=AGE['z_dim'])
    cfg = O.StepConfig(batch_size=Bs, matching_loss_multiplier=AGE['matching'],
                       contrasting_loss_multiplier=AGE['contrasting'], gradient_penalty_multiplier=AGE['gp'])
    x, y, u = make_batches(Bs, 1, AGE['image'])
    gen = torch.Generator().manual_seed(2)
    z, alpha, z2 = torch.randn(Bs, 256, generator=gen), torch.rand(Bs, 1, 1, 1, generator=gen), torch.randn(Bs, 256, generator=gen)
    for _ in range(args.warmup):
        O.training_step(st, cfg, x, y, u, z, alpha, z2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.training_step(st, cfg, x, y, u, z, alpha, z2)
    dt = time.perf_counter() - t0
    full_steps = args.steps * Bs / AGE['batch']          # samples processed / samples per full step
    value = full_steps / dt
    sample = f'{args.steps} steps of the age step on {Bs}-sample batches (full step = {AGE["batch"]}); steps/s scaled by {Bs}/{AGE["batch"]}'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 / value, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'age SR-GAN, DCGAN G/D 3x128x128, batch 100 per step (oracle port of the reference step on host CPU)'},
            'cpu_baseline': {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(budget_s=20.0):
    import torch
    from oracle import srgan_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = 20
    st = O.init_dcgan(seed=0, image_size=AGE['image'], conv_dim=AGE['conv_dim'], z_dim=AGE['z_dim'])
    cfg = O.StepConfig(batch_size=Bs, matching_loss_multiplier=AGE['matching'],
                       contrasting_loss_multiplier=AGE['contrasting'], gradient_penalty_multiplier=AGE['gp'])
    x, y, u = make_batches(Bs, 1, AGE['image'])
    gen = torch.Generator().manual_seed(2)
    z, alpha, z2 = torch.randn(Bs, 256, generator=gen), torch.rand(Bs, 1, 1, 1, generator=gen), torch.randn(Bs, 256, generator=gen)
    O.training_step(st, cfg, x, y, u, z, alpha, z2)      # warm-up (thread pools, allocator)
    n, t0 = 0, time.perf_counter()
    while True:
        O.training_step(st, cfg, x, y, u, z, alpha, z2)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 10:
            break
    value = (n * Bs / AGE['batch']) / dt
    return {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n} oracle steps on {Bs}-sample batches in {dt:.1f} s (full step = {AGE["batch"]} samples; steps/s scaled by {Bs}/{AGE["batch"]})'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=AGE['batch'], help='per-GPU batch')
    ap.add_argument('--ref-batch', type=int, default=10)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import srgan_b200
    from srgan_b200.dist import Comm

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        comm = Comm()
    dev = torch.device('cuda', local_rank)
    B = args.batch

    s = srgan_b200.Settings()
    s.batch_size = B
    s.matching_loss_multiplier, s.contrasting_loss_multiplier, s.gradient_penalty_multiplier = AGE['matching'], AGE['contrasting'], AGE['gp']
    s.precision = args.precision
    exp = srgan_b200.Experiment(s, 'age', device=dev, comm=comm, image_size=AGE['image'], conv_dim=AGE['conv_dim'], z_dim=AGE['z_dim'])
    eng = exp.runner.engine
    xh, yh, uh = make_batches(B, 1 + rank, AGE['image'])
    xh, yh, uh = xh.pin_memory(), yh.pin_memory(), uh.pin_memory()
    x, y, u = xh.to(dev), yh.to(dev), uh.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, xx, yy, uu):
        exp.dnn_training_step(xx, yy, i)
        exp.gan_training_step(xx, yy, uu, i)

    # ---------------- resident-input timing
    for i in range(args.warmup):
        step(i, x, y, u)
    barrier()
    launches0 = eng.ops.launches
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i, x, y, u)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    graphed = exp.runner.use_cuda_graph
    launches = eng.ops.launches - launches0
    ms = e0.elapsed_time(e1)
    # dominant-kernel timing: CUDA events around every launch of the D layer-2 forward conv.  The timed region above
    # replays CUDA graphs (no host launches, so no event records inside it); the same kernel on the same buffers is
    # therefore timed during K eager steps run right after it, still inside this process and clock state.
    exp.runner.use_cuda_graph = False
    launches_eager0 = eng.ops.launches
    eng.probe_begin(layer_index=2, rows=4 * B)
    for i in range(args.steps):
        step(args.warmup + args.steps + i, x, y, u)
    probe = eng.probe_end()
    launches_per_step = (eng.ops.launches - launches_eager0) / args.steps
    exp.runner.use_cuda_graph = graphed
    if graphed:
        launches = int(round(launches_per_step * args.steps))     # kernels executed by the replayed graphs
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    global_steps = 1e3 / ms_per_step                       # optimizer steps/s of the whole job
    # whole-job throughput in units of one-GPU steps (a 100-sample dnn+gan step): every global step at N ranks does N
    # of them (weak scaling), so value = N * global steps/s; at N=1 the two coincide
    value = global_steps * world

    # ---------------- end-to-end timing: host buffers in, scalars out, every step
    for i in range(3):
        step(i, xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True), uh.to(dev, non_blocking=True))
    barrier()
    e0.record()
    for i in range(args.steps):
        xx, yy, uu = xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True), uh.to(dev, non_blocking=True)
        step(i, xx, yy, uu)
        sc = exp.runner.scalars()                          # device -> host read of the step's losses
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * 1e3 / (float(t.item()) / args.steps)
    h2d = xh.numel() * 4 + yh.numel() * 4 + uh.numel() * 4
    d2h = eng.scalars.numel() * 4

    if rank == 0:
        pk, pk_kind = peaks()
        l2 = eng.d_net.layers[1]
        flops_launch = 2.0 * l2.geom.macs_per_sample * 4 * B
        roof = {'bound': 'tensor', 'achieved': None, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': None,
                'traffic': None, 'kernel': probe.get('kernel', 'conv_down layer2 over 4B rows'), 'peak_source': f'{pk_kind} (sustained: kernel timed inside a long step)',
                'launches_timed': probe.get('count', 0),
                'timing': 'CUDA events around each launch during K eager steps run right after the timed region (the timed region replays CUDA graphs)' if graphed else 'CUDA events around each launch inside the timed region'}
        if probe.get('count'):
            avg_ms = probe['ms'] / probe['count']
            roof['achieved'] = flops_launch / (avg_ms * 1e-3) / 1e12
            roof['frac'] = roof['achieved'] / roof['peak']
            roof['avg_launch_ms'] = avg_ms
        step_tflops = FLOPS_PER_SAMPLE * B * world / (ms_per_step * 1e-3) / 1e12
        line = {'metric': METRIC, 'value': value, 'unit': 'steps/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': f'age SR-GAN (BASELINE configs[1]): DCGAN G/D, 3x128x128, per-GPU batch {B}, global batch {B * world}, dnn_training_step + gan_training_step, generator period 1',
                           'precision_mode': args.precision, 'parallelism': f'dp{world}', 'cuda_graph': bool(graphed),
                           'l2': 'inputs (39 MB/step) and activations (~1 GB/step) exceed the 126 MB L2; no flush needed',
                           'global_steps_per_s': global_steps,
                           'value_definition': 'n_gpus x global optimizer steps/s = 100-sample step-equivalents per second over the whole job',
                           'step_algorithmic_tflops': step_tflops,
                           'step_frac_of_bf16_sustained_peak': step_tflops / (pk['bf16_tflops_sustained'] * world)},
                'roofline': roof, 'clocks': clocks, 'gpu_launches': int(launches),
                'e2e': {'value': e2e_value, 'unit': 'steps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
                'last_scalars': sc}
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
