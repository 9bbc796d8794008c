"""The "existing Blackwell kernels" bar: the oracle's PyTorch step (autograd + double backward, cuDNN / cuBLAS)
run ON the B200 for the bench workloads -- fp32 with TF32 off, fp32 with TF32 on, and bf16 autocast -- next to this
repository's step.  Test infrastructure (it executes oracle/, so it lives under tests/).  usage: python tests/baselines/torch_gpu_baseline.py [age|crowd] [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else 'age'
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name]['batch']
O, st, cfg, inputs = bench.oracle_setup(name, B)
dev = torch.device('cuda')
def cu(t):
    return tuple(e.to(dev) for e in t) if isinstance(t, tuple) else t.to(dev)
for net in ('D', 'G', 'DNN'):
    setattr(st, net, {k: v.to(dev) for k, v in getattr(st, net).items()})
inputs = tuple(cu(t) for t in inputs)
for mode in ('fp32 (TF32 off)', 'fp32 (TF32 on)', 'bf16 autocast'):
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = (mode != 'fp32 (TF32 off)')
    torch.backends.cudnn.benchmark = True
    s2 = st.clone()
    ctx = torch.autocast('cuda', dtype=torch.bfloat16) if mode.startswith('bf16') else torch.autocast('cuda', enabled=False)
    try:
        with ctx:
            for _ in range(3):
                O.training_step(s2, cfg, *inputs)
            torch.cuda.synchronize()
            n, t0 = 0, time.perf_counter()
            while n < 20 and time.perf_counter() - t0 < 20:
                O.training_step(s2, cfg, *inputs)
                n += 1
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
        print(f'{name} B={B} PyTorch-on-B200 {mode}: {dt * 1e3:.1f} ms/step = {1 / dt:.2f} steps/s ({B / dt:.0f} samples/s)', flush=True)
    except Exception as e:
        print(f'{name} B={B} PyTorch-on-B200 {mode}: failed: {type(e).__name__}: {str(e)[:200]}', flush=True)
