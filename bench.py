"""
bench.py -- SR-GAN training steps/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload age|crowd|coefficient] [--precision bf16|fp32] [--batch B]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     (the oracle port of the reference step on the host CPU cores)

Default workload (config.workload): BASELINE configs[1] "age SR-GAN": DCGAN G/D (age/models.py:32-80), synthetic
3x128x128 inputs ~U(-1,1), labels ~U(10,95), per-GPU batch 100, multipliers of run.py:30-35; one step =
dnn_training_step + gan_training_step with generator_training_step_period=1.  `--workload crowd` runs
BASELINE configs[2] (DCGenerator + KnnDenseNetCat/DenseNet-201 at 224x224, per-GPU batch 64, run.py:57-68 multipliers),
`--workload coefficient` BASELINE configs[0] (MLPs, batch 5000).  N>1 is weak scaling: every rank holds a per-GPU-batch
shard of the global batch, feature sums and gradients are all-reduced (NCCL) so the loss is the global-batch loss.

`value`   : steps/s with the step's inputs already resident in HBM (CUDA events, max over ranks), times N.
`e2e`     : steps/s through the public API (srgan_b200.Experiment.*_training_step) with HOST (pinned) input buffers:
            H2D copy of x, y, u for every step (prefetched one step ahead on a copy stream) and a D2H read of the step's
            scalars inside the timed region.
`roofline`: dominant dense kernel (age: the layer-2 discriminator conv over the 4B-row batch) timed live with CUDA events;
            algorithmic FLOPs / duration against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port (PyTorch fp32 autograd on the host cores) on a bounded sample, rank 0, N=1 only.
Inputs and activations of one step are far larger than L2 (126 MB), so no L2 flush is needed between iterations
(stated in config.l2).
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
WORKLOADS = {
    # BASELINE configs[1] -- the N=1 workload of the default run (the metric's single-GPU configuration)
    'age': dict(batch=100, ref_batch=10, flops_per_sample=21 * 0.8305e9 + 4 * 0.8472e9, mult=(1e2, 1e1, 1e2),
                desc='age SR-GAN (BASELINE configs[1]): DCGAN G/D, 3x128x128'),
    # BASELINE configs[3]: the same DCGAN G/D (driving/models.py == age/models.py), steering-angle labels ~ N(0, 30 deg),
    # run.py:36-44 multipliers, gradient penalty on (B = 100)
    'driving': dict(batch=100, ref_batch=10, flops_per_sample=21 * 0.8305e9 + 4 * 0.8472e9, mult=(1e2, 1e1, 1e2),
                    desc='driving SR-GAN (BASELINE configs[3]): DCGAN G/D, 3x128x128, steering-angle labels'),
    # BASELINE configs[2]: DCGenerator + KnnDenseNetCat (DenseNet-201) at 224x224, run.py:57-68 multipliers
    'crowd': dict(batch=64, ref_batch=2, flops_per_sample=21 * 8.732e9 + 4 * 2.595e9, mult=(1e3, 1e2, 1e2),
                  desc='crowd SR-GAN (BASELINE configs[2]): DCGenerator + KnnDenseNetCat (DenseNet-201), 3x224x224'),
    # BASELINE configs[0]: coefficient MLPs, B = 5000 (run.py:50), one persistent kernel per step method
    'coefficient': dict(batch=5000, ref_batch=5000, flops_per_sample=21 * 1420.0 + 4 * 1600.0, mult=(1.0, 1.0, 10.0),
                        desc='coefficient SR-GAN (BASELINE configs[0]): MLP G/D, 50 observations'),
}


# dram__bytes_read.sum + dram__bytes_write.sum of the probed kernel per launch, from one `ncu --set full` capture of exactly that
# launch (profiles/r1_final_ncu_full_summary.txt: D layer-2 fprop over 400 samples, read 210.1 + write 75.5 MB; algorithmic bytes
# 315 MB = input 210 + output 105: part of the output is still in L2 when the kernel ends)
NCU_TRAFFIC_BYTES = {('age', 100): 285.57e6, ('driving', 100): 285.57e6}


def workload_string(name, B, world):
    return (f'{WORKLOADS[name]["desc"]}, per-GPU batch {B}, global batch {B * world}, dnn_training_step + '
            f'gan_training_step, generator period 1')


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


def make_batches(name, B, seed):
    """(x, y, u) host tensors of one per-GPU batch of the workload; y is the crowd (density, map) pair for crowd."""
    import torch
    gen = torch.Generator().manual_seed(seed)
    if name in ('age', 'driving'):
        x = torch.rand(B, 3, 128, 128, generator=gen) * 2 - 1
        u = torch.rand(B, 3, 128, 128, generator=gen) * 2 - 1
        if name == 'driving':                 # steering angle in degrees
            return x, torch.randn(B, generator=gen) * 30, u
        return x, torch.rand(B, generator=gen) * 85 + 10, u
    if name == 'crowd':
        x = torch.rand(B, 3, 224, 224, generator=gen) * 2 - 1
        u = torch.rand(B, 3, 224, 224, generator=gen) * 2 - 1
        density = (torch.rand(B, 224, 224, generator=gen) < 6.5e-4).float()
        return x, (density, 1 / (1 + torch.rand(B, 224, 224, generator=gen) * 50)), u
    x, u = torch.randn(B, 50, generator=gen), torch.randn(B, 50, generator=gen)
    return x, torch.rand(B, generator=gen) * 2 - 1, u


def oracle_setup(name, Bs):
    """Oracle state, config, inputs and noise of a Bs-sample step of the workload (reference arm / cpu_baseline)."""
    import torch
    from oracle import srgan_oracle as O
    m, c, gp = WORKLOADS[name]['mult']
    cfg = O.StepConfig(batch_size=Bs, matching_loss_multiplier=m, contrasting_loss_multiplier=c, gradient_penalty_multiplier=gp,
                       map_multiplier=1e-3)
    gen = torch.Generator().manual_seed(2)
    if name in ('age', 'driving'):
        st = O.init_dcgan(seed=0, image_size=AGE['image'], conv_dim=AGE['conv_dim'], z_dim=AGE['z_dim'])
        zd, ashape = 256, (Bs, 1, 1, 1)
    elif name == 'crowd':
        st = O.init_crowd(seed=0)
        zd, ashape = 256, (Bs, 1, 1, 1)
    else:
        st = O.init_coefficient(seed=0)
        zd, ashape = 10, (Bs, 1)
    x, y, u = make_batches(name, Bs, 1)
    z, alpha, z2 = torch.randn(Bs, zd, generator=gen), torch.rand(*ashape, generator=gen), torch.randn(Bs, zd, generator=gen)
    return O, st, cfg, (x, y, u, z, alpha, z2)


def run_reference(args):
    """The reference's algorithm on the host CPU (oracle port; the reference checkout does not travel to the GPU box)."""
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    name = args.workload
    full = args.batch or WORKLOADS[name]['batch']
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = args.ref_batch or WORKLOADS[name]['ref_batch']
    O, st, cfg, inputs = oracle_setup(name, Bs)
    for _ in range(args.warmup):
        O.training_step(st, cfg, *inputs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.training_step(st, cfg, *inputs)
    dt = time.perf_counter() - t0
    value = (args.steps * Bs / full) / dt                # full-step equivalents per second (per-sample scaling)
    sample = f'{args.steps} oracle steps on {Bs}-sample batches in {dt:.1f} s (full step = {full} samples; steps/s scaled by {Bs}/{full})'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 / value, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_string(name, full, world),
                       'reference_arm': 'oracle port of the reference step (PyTorch fp32 autograd) on the host CPU cores; the reference is a script tree without packaging metadata and does not travel to the GPU box'},
            'cpu_baseline': {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(name, full, budget_s=20.0):
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = {'age': 20, 'driving': 20, 'crowd': 2, 'coefficient': 5000}[name]
    O, st, cfg, inputs = oracle_setup(name, Bs)
    O.training_step(st, cfg, *inputs)                    # warm-up (thread pools, allocator)
    n, t0 = 0, time.perf_counter()
    while True:
        O.training_step(st, cfg, *inputs)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= (200 if name == 'coefficient' else 10):
            break
    value = (n * Bs / full) / dt
    return {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n} oracle steps on {Bs}-sample batches in {dt:.1f} s (full step = {full} samples; steps/s scaled by {Bs}/{full})'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--workload', default='age', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the workload\'s)')
    ap.add_argument('--ref-batch', type=int, default=0)
    ap.add_argument('--micro-batch', type=int, default=0, help='run every step in micro-batches of this many samples (exact)')
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
    name = args.workload
    wl = WORKLOADS[name]
    B = args.batch or wl['batch']

    s = srgan_b200.Settings()
    s.batch_size = B
    s.matching_loss_multiplier, s.contrasting_loss_multiplier, s.gradient_penalty_multiplier = wl['mult']
    s.map_multiplier = 1e-3
    s.precision = args.precision
    s.micro_batch = args.micro_batch
    if name in ('age', 'driving'):
        exp = srgan_b200.Experiment(s, name, device=dev, comm=comm, image_size=AGE['image'], conv_dim=AGE['conv_dim'], z_dim=AGE['z_dim'])
    else:
        exp = srgan_b200.Experiment(s, name, device=dev, comm=comm)
    eng = exp.runner.engine
    xh, yh, uh = make_batches(name, B, 1 + rank)

    def pin(t):
        return tuple(e.pin_memory() for e in t) if isinstance(t, tuple) else t.pin_memory()

    def dev_copy(t):
        return tuple(e.to(dev, non_blocking=True) for e in t) if isinstance(t, tuple) else t.to(dev, non_blocking=True)

    def nbytes(t):
        return sum(e.numel() * 4 for e in t) if isinstance(t, tuple) else t.numel() * 4
    xh, yh, uh = pin(xh), pin(yh), pin(uh)
    x, y, u = dev_copy(xh), dev_copy(yh), dev_copy(uh)

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
    overlap, exp.runner.overlap_dnn = exp.runner.overlap_dnn, False      # the probed kernel is timed alone on its stream
    launches_eager0 = eng.ops.launches
    # the probed kernel: age = D layer-2 conv (64->128 k4 s2) over the 4B-row batch; crowd = the transition-1 1x1 conv
    # (256->128 at 56x56, a [4B*3136 x 256] x [256 x 128] GEMM); coefficient = no dense kernel to probe (one persistent kernel)
    probe_layer = {'age': 'layer2.0', 'driving': 'layer2.0', 'crowd': 'transition_layers.transition1.conv'}.get(name)
    probe_steps = min(args.steps, 20)
    if probe_layer is not None:
        eng.probe_begin(layer_name=probe_layer, rows=4 * B)
    for i in range(probe_steps):
        step(args.warmup + args.steps + i, x, y, u)
    probe = eng.probe_end() if probe_layer is not None else {'count': 0}
    launches_per_step = (eng.ops.launches - launches_eager0) / probe_steps
    exp.runner.use_cuda_graph = graphed
    exp.runner.overlap_dnn = overlap
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

    # ---------------- end-to-end timing: host buffers in, scalars out, every step.  The host->device copy of step i+1 is
    # issued on a copy stream while step i computes (a two-slot prefetch, what a DataLoader with pin_memory does); every
    # step still waits for ITS inputs and ends with the device->host read of its scalars.
    copy_stream = torch.cuda.Stream(dev)
    slots = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])            # the step that last used this slot has finished with it
            slots[k] = (dev_copy(xh), dev_copy(yh), dev_copy(uh))
            ready[k].record(copy_stream)

    def e2e_loop(n):
        main = torch.cuda.current_stream(dev)
        for k in range(2):
            consumed[k].record(main)
        prefetch(0)
        for i in range(n):
            k = i % 2
            if i + 1 < n:
                prefetch(i + 1)
            main.wait_event(ready[k])
            xx, yy, uu = slots[k]
            step(i, xx, yy, uu)
            consumed[k].record(main)
            sc_ = exp.runner.scalars()                     # device -> host read of the step's losses
        return sc_
    e2e_loop(3)
    barrier()
    e0.record()
    sc = e2e_loop(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * 1e3 / (float(t.item()) / args.steps)
    h2d = nbytes(xh) + nbytes(yh) + nbytes(uh)
    d2h = eng.scalars.numel() * 4

    if rank == 0:
        pk, pk_kind = peaks()
        flops_launch = 2.0 * probe.get('macs_per_sample', 0) * 4 * B
        if name == 'coefficient':
            # one persistent kernel per step method: latency-bound by design; reported against HBM with its algorithmic bytes
            by = B * (50 + 50 + 50 + 10 + 10 + 1 + 1 + 1) * 4.0 + 3 * 2 * 2400 * 4 * 4.0
            ach = by / (ms_per_step * 1e-3) / 1e9
            roof_coef = {'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': ach / pk['hbm_gbs'],
                         'traffic': None, 'kernel': 'coef_step_kernel (two cooperative launches per step: dnn, gan)',
                         'peak_source': pk_kind, 'launches_timed': 2 * args.steps,
                         'note': 'launch/latency-bound tiny MLPs: the figure of merit is us/step, not bandwidth'}
        roof = {'bound': 'tensor', 'achieved': None, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': None,
                'traffic': None, 'kernel': probe.get('kernel', 'conv_down layer2 over 4B rows'), 'peak_source': f'{pk_kind} (sustained: kernel timed inside a long step)',
                'launches_timed': probe.get('count', 0),
                'timing': 'CUDA events around each launch during K eager steps run right after the timed region (the timed region replays CUDA graphs)' if graphed else 'CUDA events around each launch inside the timed region'}
        roof['traffic'] = NCU_TRAFFIC_BYTES.get((name, B))
        if roof['traffic'] is not None:
            roof['traffic_source'] = 'profiles/r1_final_ncu_full_summary.txt (one ncu --set full capture of this launch)'
        if probe.get('count'):
            avg_ms = probe['ms'] / probe['count']
            roof['achieved'] = flops_launch / (avg_ms * 1e-3) / 1e12
            roof['frac'] = roof['achieved'] / roof['peak']
            roof['avg_launch_ms'] = avg_ms
            # which roofline binds this launch: arithmetic intensity (algorithmic FLOP / algorithmic byte: input, output
            # and weights moved once) against the ridge of the two measured peaks.  The age layer-2 conv (AI ~ 680) is
            # tensor-bound; the crowd trunk's 1x1 GEMMs (N = 128: AI ~ 170 and less) are HBM-bound.
            by = probe.get('bytes_per_launch', 0)
            if by and args.precision == 'bf16':
                ai, ridge = flops_launch / by, pk['bf16_tflops_sustained'] * 1e12 / (pk['hbm_gbs'] * 1e9)
                roof['arithmetic_intensity'], roof['ridge'] = ai, ridge
                roof['tensor_tflops'] = roof['achieved']
                if ai < ridge:
                    gbs = by / (avg_ms * 1e-3) / 1e9
                    roof.update({'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': gbs / pk['hbm_gbs'],
                                 'peak_source': f'{pk_kind} (HBM copy bandwidth)', 'algorithmic_bytes_per_launch': by})
        step_tflops = wl['flops_per_sample'] * B * world / (ms_per_step * 1e-3) / 1e12
        line = {'metric': METRIC, 'value': value, 'unit': 'steps/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': workload_string(name, B, world),
                           'precision_mode': args.precision, 'parallelism': f'dp{world}', 'cuda_graph': bool(graphed), 'dnn_gan_overlap': bool(exp.runner.overlap_dnn), 'micro_batch': args.micro_batch,
                           'l2': ('inputs and activations of one step (age: 39 MB + ~1 GB, crowd: ~0.7 GB per sample) exceed the 126 MB L2; no flush needed'
                                  if name != 'coefficient' else 'working set (2 MB) is L2-resident by design: the step is launch/latency-bound, not bandwidth-bound'),
                           'global_steps_per_s': global_steps,
                           'value_definition': 'n_gpus x global optimizer steps/s = per-GPU-batch step-equivalents per second over the whole job',
                           'step_algorithmic_tflops': step_tflops,
                           'step_frac_of_bf16_sustained_peak': step_tflops / (pk['bf16_tflops_sustained'] * world)},
                'roofline': roof_coef if name == 'coefficient' else roof, 'clocks': clocks, 'gpu_launches': int(launches),
                'e2e': {'value': e2e_value, 'unit': 'steps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
                'last_scalars': sc}
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(name, B)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
