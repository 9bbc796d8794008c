"""
bench.py -- SR-GAN training steps/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload crowd|age|driving|coefficient] [--precision bf16|fp32] [--batch B]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     (the UNMODIFIED reference step on the host CPU cores)

Default workload (config.workload) = the configuration BASELINE.json's metric is quoted on: configs[2] "crowd SR-GAN":
DCGenerator + KnnDenseNetCat (DenseNet-201) at 224x224 (crowd/srgan.py:92-96, crowd/models.py:127-147, 1049-1166), synthetic
inputs of SURVEY 8d config 3, run.py:57-68 multipliers, per-GPU batch 64; one step = dnn_training_step +
gan_training_step with generator_training_step_period = 1.  The same line carries `secondary`: the age SR-GAN workload
(configs[1], DCGAN G/D at 128x128, batch 100) measured the same way in the same process.  `--workload age|driving|
coefficient` make one of the other configs the primary.  N>1 is weak scaling: every rank holds a per-GPU-batch shard of
the global batch, feature sums and gradients are all-reduced (NCCL) so the loss is the global-batch loss.

`value`   : steps/s with the step's inputs already resident in HBM (CUDA events, max over ranks), times N.
`e2e`     : steps/s through the public API (srgan_b200.Experiment.*_training_step) with HOST (pinned) input buffers:
            H2D copy of x, y, u for every step (prefetched one step ahead on a copy stream) and a D2H read of the step's
            scalars inside the timed region.
`roofline`: the dominant kernel class of the step timed live with CUDA events around each of its launches (crowd: the
            DenseNet trunk's 1x1 data-gradient GEMMs, HBM-bound; age: the layer-2 discriminator conv, tensor-bound);
            algorithmic bytes or FLOPs per launch / average launch duration against MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own step on the box's host cores on a bounded sample, rank 0, N=1 only ("reference" = the
            unmodified reference staged under baseline/_ref by oracle/stage_reference.py; "port" = the oracle restatement
            when the staged tree is absent).
`gpu_baseline`: the same reference step as stock PyTorch (cuDNN / cuBLAS) ON cuda:0, TF32 off and on -- the "existing
            Blackwell kernels" bar of SURVEY section 0 / 8d.  N=1 only, after the product arm has released its memory.
Inputs and activations of one step are far larger than L2 (126 MB), so no L2 flush is needed between iterations
(stated in details.l2).
"""
from __future__ import annotations

import argparse
import gc
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
    # BASELINE configs[2]: DCGenerator + KnnDenseNetCat (DenseNet-201) at 224x224, run.py:57-68 multipliers -- the
    # configuration the metric ("at 1/2/4/8 B200") and the target sentence are quoted on
    'crowd': dict(batch=64, ref_batch=2, gpu_ref_batch=32, flops_per_sample=21 * 8.732e9 + 4 * 2.595e9, mult=(1e3, 1e2, 1e2),
                  desc='crowd SR-GAN (BASELINE configs[2]): DCGenerator + KnnDenseNetCat (DenseNet-201), 3x224x224'),
    # BASELINE configs[1]
    'age': dict(batch=100, ref_batch=10, gpu_ref_batch=100, flops_per_sample=21 * 0.8305e9 + 4 * 0.8472e9, mult=(1e2, 1e1, 1e2),
                desc='age SR-GAN (BASELINE configs[1]): DCGAN G/D, 3x128x128'),
    # BASELINE configs[3]: the same DCGAN G/D (driving/models.py == age/models.py), steering-angle labels ~ N(0, 30 deg),
    # run.py:36-44 multipliers, gradient penalty on (B = 100)
    'driving': dict(batch=100, ref_batch=10, gpu_ref_batch=100, flops_per_sample=21 * 0.8305e9 + 4 * 0.8472e9, mult=(1e2, 1e1, 1e2),
                    desc='driving SR-GAN (BASELINE configs[3]): DCGAN G/D, 3x128x128, steering-angle labels'),
    # BASELINE configs[0]: coefficient MLPs, B = 5000 (run.py:50), one persistent kernel per step method
    'coefficient': dict(batch=5000, ref_batch=5000, gpu_ref_batch=5000, flops_per_sample=21 * 1420.0 + 4 * 1600.0, mult=(1.0, 1.0, 10.0),
                        desc='coefficient SR-GAN (BASELINE configs[0]): MLP G/D, 50 observations'),
}

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the probed kernel class, from `ncu --set full` captures of
# exactly those launches (profiles/): filled in per round by the builder, None = not captured for this shape
NCU_TRAFFIC = {
    ('age', 100): (285.57e6, 'profiles/r1_final_ncu_full_summary.txt (D layer-2 fprop over 400 samples: read 210.1 + write 75.5 MB)'),
    ('driving', 100): (285.57e6, 'profiles/r1_final_ncu_full_summary.txt (D layer-2 fprop over 400 samples)'),
}


def workload_string(name, B, world):
    return (f'{WORKLOADS[name]["desc"]}, per-GPU batch {B}, global batch {B * world}, dnn_training_step + '
            f'gan_training_step, generator period 1')


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
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
    """Oracle state, config, inputs and noise of a Bs-sample step of the workload (the "port" CPU arm)."""
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


# ------------------------------------------------------------------------------------------------ the reference's own step
class ReferenceStep:
    """One dnn_training_step + gan_training_step of the reference on `device` at batch Bs: the UNMODIFIED reference
    (srgan.py:259-320 through its own Experiment subclass and nn.Modules, its own noise draws and torch.optim.Adam) when
    the staged tree is present, else the oracle restatement ("port")."""

    def __init__(self, name, Bs, device='cpu'):
        import torch
        from oracle import ref_harness
        self.name, self.Bs, self.device = name, Bs, torch.device(device)
        self.kind = 'reference' if ref_harness.reference_available() else 'port'
        self.i = 0

        def put(t):
            return tuple(e.to(self.device) for e in t) if isinstance(t, tuple) else t.to(self.device)
        if self.kind == 'reference':
            m, c, gp = WORKLOADS[name]['mult']
            kw = dict(batch_size=Bs, matching_loss_multiplier=m, contrasting_loss_multiplier=c, gradient_penalty_multiplier=gp,
                      map_multiplier=1e-3, summary_step_period=1000000)
            self.exp = ref_harness.workload_experiment(name, kw, device=self.device)
            x, y, u = make_batches(name, Bs, 1)
            self.inputs = (put(x), put(y), put(u))
        else:
            self.O, st, self.cfg, inputs = oracle_setup(name, Bs)
            for net in ('D', 'G', 'DNN'):
                setattr(st, net, {k: v.to(self.device) for k, v in getattr(st, net).items()})
            self.st = st
            self.inputs = tuple(put(t) for t in inputs)

    def step(self):
        if self.kind == 'reference':
            x, y, u = self.inputs
            self.exp.dnn_training_step(x, y, self.i + 1)          # step 0 would be a summary step (extra .item() syncs)
            self.exp.gan_training_step(x, y, u, self.i + 1)
        else:
            self.O.training_step(self.st, self.cfg, *self.inputs)
        self.i += 1

    def describe(self):
        if self.kind == 'reference':
            return 'the unmodified reference step (baseline/_ref: srgan.py dnn_training_step + gan_training_step, torch autograd)'
        return 'oracle port of the reference step (PyTorch fp32 autograd); the staged reference tree baseline/_ref is absent'


def _timed_cpu(ref, n):
    t0 = time.perf_counter()
    for _ in range(n):
        ref.step()
    return time.perf_counter() - t0


def run_reference(args):
    """The reference's own CPU implementation of the step on the box's host cores, all threads.  Every timed step is the
    full step on a bounded sample of the per-GPU batch: the sample size is chosen from a probe step so that the
    K + W steps end within ~4 minutes (the full batch when that fits: age B=100 on a 16-core host does)."""
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    name = args.workload
    full = args.batch or WORKLOADS[name]['batch']
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = args.ref_batch
    if not Bs:
        b0 = min(full, WORKLOADS[name]['ref_batch'])
        probe = ReferenceStep(name, b0)
        probe.step()                                     # thread pools, allocator
        t1 = _timed_cpu(probe, 1)
        del probe
        gc.collect()
        per_step_budget = args.ref_budget_s / (args.steps + args.warmup)
        Bs = int(max(b0, min(full, b0 * per_step_budget / max(t1, 1e-6))))
    ref = ReferenceStep(name, Bs)
    for _ in range(args.warmup):
        ref.step()
    dt = _timed_cpu(ref, args.steps)
    ms_sample_step = dt / args.steps * 1e3
    value = (args.steps * Bs / full) / dt                # full-step equivalents per second (per-sample scaling)
    exact = Bs == full
    sample = (f'{args.steps} steps of {ref.describe()} on {Bs}-sample batches in {dt:.1f} s'
              + ('' if exact else f' (full step = {full} samples; steps/s scaled by {Bs}/{full}, the CPU cost is linear in the batch)'))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup,
            # wall time of one timed step as executed (a Bs-sample step); ms of a full-batch step = 1e3 / value
            'ms_per_step': ms_sample_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_string(name, full, world)},
            'details': {'reference_arm': ref.describe(), 'sample_batch': Bs, 'full_batch': full, 'full_batch_run': exact,
                        'ms_per_full_step': 1e3 / value},
            'cpu_baseline': {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': ref.kind, 'sample': sample},
            'e2e': {'value': value, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(name, full, budget_s=25.0):
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = {'age': 100, 'driving': 100, 'crowd': 4, 'coefficient': 5000}[name]
    Bs = min(Bs, full)
    ref = ReferenceStep(name, Bs)
    ref.step()                                           # warm-up (thread pools, allocator)
    n, t0 = 0, time.perf_counter()
    while True:
        ref.step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= (200 if name == 'coefficient' else 10):
            break
    value = (n * Bs / full) / dt
    return {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': ref.kind,
            'sample': f'{n} steps of {ref.describe()} on {Bs}-sample batches in {dt:.1f} s'
                      + ('' if Bs == full else f' (full step = {full} samples; steps/s scaled by {Bs}/{full})')}


def gpu_baseline(name, full, dev, budget_s=15.0):
    """Stock PyTorch on the same B200: the reference step with its modules on cuda:0 (cuDNN / cuBLAS), TF32 off and on."""
    import torch
    Bs = min(full, WORKLOADS[name]['gpu_ref_batch'])
    out = {'unit': 'steps/s', 'sample_batch': Bs, 'full_batch': full}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for mode, tf32 in (('fp32_tf32_off', False), ('fp32_tf32_on', True)):
            try:
                torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cudnn.benchmark = True            # run.py:21
                ref = ReferenceStep(name, Bs, dev)
                out['kind'] = ref.kind
                for _ in range(3):
                    ref.step()
                torch.cuda.synchronize(dev)
                n, t0 = 0, time.perf_counter()
                while n < 20 and time.perf_counter() - t0 < budget_s / 2:
                    ref.step()
                    n += 1
                torch.cuda.synchronize(dev)
                dt = (time.perf_counter() - t0) / n
                out[mode] = {'value': (Bs / full) / dt, 'ms_per_sample_step': dt * 1e3, 'steps_timed': n}
                del ref
            except Exception as e:                               # e.g. out of memory at this batch: reported, not hidden
                out[mode] = {'error': f'{type(e).__name__}: {str(e)[:160]}'}
            gc.collect()
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
    out['note'] = ('the reference step as stock PyTorch on cuda:0, wall clock around synchronised steps'
                   + ('' if Bs == full else f'; steps/s scaled by {Bs}/{full}'))
    return out


# ------------------------------------------------------------------------------------------------ the product arm
def probe_selector(name, eng, B):
    """(select, label, bound) of the dominant kernel class of the workload."""
    if name in ('age', 'driving'):
        return (lambda role, st, l, n: role == 'forward' and st is eng.D and l.name == 'layer2.0' and n == 4 * B,
                f'D layer2.0 forward conv (64->128 k4 s2) over {4 * B} samples (umma_conv_persistent_kernel)')
    if name == 'crowd':
        return (lambda role, st, l, n: role in ('dgrad', 'dgrad_bn') and l.name.endswith('.conv1'),
                'DenseNet trunk 1x1 data-gradient GEMMs with the BatchNorm + ReLU backward of norm1 in the epilogue '
                '(bn_dgrad_kernel<false>, csrc/bn_gemm.cu; every dense layer / transition conv1, all passes): '
                '[pixels x 128] x [128 x C], algorithmic bytes rows x (128 + 3 C) x 2')
    return None, None


def run_workload(name, args, dev, comm, world, rank, local_rank, steps, warmup, with_cpu_baseline, with_gpu_baseline):
    import torch
    import torch.distributed as dist
    import srgan_b200

    wl = WORKLOADS[name]
    B = (args.batch if name == args.workload else 0) or wl['batch']
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
        return sum(e.numel() * e.element_size() for e in t) if isinstance(t, tuple) else t.numel() * t.element_size()
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
    for i in range(warmup):
        step(i, x, y, u)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i, x, y, u)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    graphed = exp.runner.use_cuda_graph
    ms = e0.elapsed_time(e1)
    scalars_resident = exp.runner.scalars()
    # dominant-kernel timing: CUDA events around every launch of the probed kernel class.  The timed region above
    # replays CUDA graphs (no host launches, so no event records inside it); the same kernels on the same buffers are
    # therefore timed during eager steps run right after it, still inside this process and clock state, with the side
    # streams off so that each probed kernel runs alone on the device.
    exp.runner.use_cuda_graph = False
    overlap, exp.runner.overlap_dnn = exp.runner.overlap_dnn, False
    side = (eng.wgrad_side_stream, eng.branch_streams)
    eng.wgrad_side_stream = eng.branch_streams = False
    launches_eager0 = eng.ops.launches
    select, label = probe_selector(name, eng, B)
    probe_steps = min(steps, 20 if name != 'crowd' else 3)
    if select is not None:
        eng.probe_begin(select, label)
    for i in range(probe_steps):
        step(warmup + steps + i, x, y, u)
    probe = eng.probe_end() if select is not None else {'count': 0}
    launches_per_step = (eng.ops.launches - launches_eager0) / probe_steps
    exp.runner.use_cuda_graph = graphed
    exp.runner.overlap_dnn = overlap
    eng.wgrad_side_stream, eng.branch_streams = side
    launches = int(round(launches_per_step * steps))         # kernels executed inside the timed region (graph replays)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / steps
    global_steps = 1e3 / ms_per_step                       # optimizer steps/s of the whole job
    # whole-job throughput in units of one-GPU steps (a per-GPU-batch dnn+gan step): every global step at N ranks does N
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
        sc_ = None
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
    sc = e2e_loop(steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * 1e3 / (float(t.item()) / steps)
    h2d = nbytes(xh) + nbytes(yh) + nbytes(uh)
    d2h = eng.scalars.numel() * 4
    overlap_flag = bool(exp.runner.overlap_dnn)

    # ---------------- crowd: the same loop fed by the device-side input pipeline (sr-gan_b200/crowd_data.py): the
    # full images / density labels / kNN maps stay resident in HBM and every step gathers its labeled and unlabeled patches
    # with srgan_crowd_extract_patches from freshly drawn positions; only the two [B,4] position tables cross PCIe
    pipeline = None
    if name == 'crowd':
        import random
        import numpy as np
        from srgan_b200 import crowd_data
        rs = np.random.RandomState(11 + rank)
        shape = (768, 1024)                                # ShanghaiTech part A images are up to 768 x 1024
        full = [(rs.randint(0, 256, size=shape + (3,)).astype(np.uint8),
                 (rs.rand(*shape) < 6.5e-4).astype(np.float32), (1.0 / (1.0 + 50.0 * rs.rand(*shape))).astype(np.float32))
                for _ in range(16)]
        store = crowd_data.CrowdStore(full, device=dev)
        dataset = crowd_data.TransformedDataset(store, 224, 224, rng=random.Random(rank))

        def next_batches():
            # train_dataset_loader's and unlabeled_dataset_loader's batches (crowd/srgan.py:59-68): host draws, then one gather each
            return dataset.batch(B), dataset.batch(B)

        def pipe_loop(n):
            sc_ = None
            nxt = next_batches()
            for i in range(n):
                (xi, li, mi), (ui, _, _) = nxt
                step(i, xi, (li, mi), ui)
                if i + 1 < n:
                    nxt = next_batches()                   # drawn while the step runs, gathered right behind it (a loader's prefetch)
                sc_ = exp.runner.scalars()
            return sc_
        pipe_loop(3)
        barrier()
        e0.record()
        pipe_loop(steps)
        e1.record()
        barrier()
        tp = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        pipeline = {'value': world * 1e3 / (float(tp.item()) / steps), 'unit': 'steps/s',
                    'h2d_bytes_per_step': 2 * B * 16, 'd2h_bytes_per_step': int(d2h),
                    'resident_store_bytes': int(store.pixels * 11),
                    'what': 'e2e with the device-side input pipeline: random 224x224 patches (+ flip, normalise) of 16 resident '
                            '768x1024 examples gathered on the GPU every step for the labeled and the unlabeled batch'}
        del store, dataset, full

    # release the engine's buffers before the baselines / the secondary workload allocate theirs
    del exp, eng, x, y, u, slots
    gc.collect()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    pk, pk_kind = peaks()
    if name == 'coefficient':
        # one persistent kernel per step method: latency-bound by design; reported against HBM with its algorithmic bytes
        by = B * (50 + 50 + 50 + 10 + 10 + 1 + 1 + 1) * 4.0 + 3 * 2 * 2400 * 4 * 4.0
        ach = by / (ms_per_step * 1e-3) / 1e9
        roof = {'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': ach / pk['hbm_gbs'],
                'traffic': None, 'kernel': 'coef_step_kernel (two cooperative launches per step: dnn, gan)',
                'peak_source': pk_kind, 'launches_timed': 2 * steps,
                'note': 'launch/latency-bound tiny MLPs: the figure of merit is us/step, not bandwidth'}
    else:
        roof = {'bound': 'tensor', 'achieved': None, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': None,
                'traffic': None, 'kernel': probe.get('kernel', label), 'peak_source': f'{pk_kind} (sustained: kernel timed inside a long step)',
                'launches_timed': probe.get('count', 0),
                'timing': 'CUDA events around each launch during eager steps run right after the timed region (the timed region replays CUDA graphs)'}
        tr = NCU_TRAFFIC.get((name, B))
        if tr is not None:
            roof['traffic'], roof['traffic_source'] = tr
        elif name == 'crowd':
            # `traffic` is per launch like `achieved`, and the probed class mixes 1 212 launches of many shapes: no single ncu
            # capture is its per-launch average.  The capture that exists is of the largest shape class:
            roof['traffic_reference'] = ('ncu --set full of bn_dgrad_kernel<false> at rows 50176 x C 1024: dram read + write 295.6 MB '
                                         'for 321.1 MB algorithmic in 70.9 us (profiles/r2_final_ncu_full_summary.txt): nothing re-read')
        if probe.get('count'):
            avg_ms = probe['ms'] / probe['count']
            fl, by = probe['flops_per_launch'], probe['bytes_per_launch']
            roof['achieved'] = fl / (avg_ms * 1e-3) / 1e12
            roof['frac'] = roof['achieved'] / roof['peak']
            roof['avg_launch_ms'] = avg_ms
            # which roofline binds this kernel class: arithmetic intensity (algorithmic FLOP / algorithmic byte: operands
            # and result moved once) against the ridge of the two measured peaks.  The age layer-2 conv (AI ~ 340) is
            # tensor-bound; the crowd trunk's 1x1 GEMMs (K or N = 128: AI < 130) are HBM-bound.
            if by and args.precision == 'bf16':
                ai, ridge = fl / by, pk['bf16_tflops_sustained'] * 1e12 / (pk['hbm_gbs'] * 1e9)
                roof['arithmetic_intensity'], roof['ridge'] = ai, ridge
                roof['tensor_tflops'] = roof['achieved']
                if ai < ridge:
                    gbs = by / (avg_ms * 1e-3) / 1e9
                    roof.update({'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': gbs / pk['hbm_gbs'],
                                 'peak_source': f'{pk_kind} (HBM copy bandwidth)', 'algorithmic_bytes_per_launch': by})
    step_tflops = wl['flops_per_sample'] * B * world / (ms_per_step * 1e-3) / 1e12
    line = {'metric': METRIC, 'value': value, 'unit': 'steps/s', 'n_gpus': world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': workload_string(name, B, world)},
            'details': {'precision_mode': args.precision, 'parallelism': f'dp{world}', 'cuda_graph': bool(graphed),
                        'dnn_gan_overlap': overlap_flag, 'micro_batch': args.micro_batch,
                        'l2': ('inputs and activations of one step (age: 39 MB + ~1 GB, crowd: ~0.7 GB per sample) exceed the 126 MB L2; no flush needed'
                               if name != 'coefficient' else 'working set (2 MB) is L2-resident by design: the step is launch/latency-bound, not bandwidth-bound'),
                        'global_steps_per_s': global_steps,
                        'value_definition': 'n_gpus x global optimizer steps/s = per-GPU-batch step-equivalents per second over the whole job',
                        'samples_per_s': global_steps * B * world,
                        'step_algorithmic_tflops': step_tflops,
                        'step_frac_of_bf16_sustained_peak': step_tflops / (pk['bf16_tflops_sustained'] * world)},
            'roofline': roof, 'clocks': clocks, 'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': 'steps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            # the losses of the last timed step (global-batch values: equal across N at equal global batch, tests/test_gpu_dist.py)
            'last_scalars': sc, 'last_scalars_resident': scalars_resident}
    if pipeline is not None:
        line['e2e_device_pipeline'] = pipeline
    if world == 1 and with_gpu_baseline:
        line['gpu_baseline'] = gpu_baseline(name, B, dev)
    if world == 1 and with_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(name, B)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--workload', default='crowd', choices=sorted(WORKLOADS))
    ap.add_argument('--secondary', default=None, help='second workload reported under `secondary` (default: age when the primary is crowd; "none" = off)')
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the workload\'s)')
    ap.add_argument('--ref-batch', type=int, default=0, help='reference arm: samples per timed step (default: from the time budget)')
    ap.add_argument('--ref-budget-s', type=float, default=240.0, help='reference arm: wall-clock budget of the K + W steps')
    ap.add_argument('--micro-batch', type=int, default=0, help='run every step in micro-batches of this many samples (exact)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
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
    line = run_workload(args.workload, args, dev, comm, world, rank, local_rank, args.steps, args.warmup,
                        not args.no_cpu_baseline, not args.no_gpu_baseline)
    sec = args.secondary if args.secondary is not None else ('age' if args.workload == 'crowd' else 'none')
    if sec != 'none' and sec != args.workload:
        # the secondary workload is small next to the primary (age: 4 ms steps): more steps for a stable number
        s2 = run_workload(sec, args, dev, comm, world, rank, local_rank, max(args.steps, 100), max(args.warmup, 10),
                          not args.no_cpu_baseline, not args.no_gpu_baseline)
        if rank == 0:
            line['secondary'] = {k: s2[k] for k in ('value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'dtype',
                                                    'config', 'details', 'roofline', 'gpu_launches', 'e2e', 'last_scalars',
                                                    'cpu_baseline', 'gpu_baseline') if k in s2}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
