"""Coefficient config (BASELINE configs[0], B=5000, run.py:46-54): us/step of the B200 path with and without CUDA graphs
(the CPU figure next to it comes from `bench.py --workload coefficient`, cpu_baseline)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srgan_b200

B = 5000
method = sys.argv[1] if len(sys.argv) > 1 else 'srgan'
gen = torch.Generator().manual_seed(0)
x, u = torch.randn(B, 50, generator=gen), torch.randn(B, 50, generator=gen)
y = torch.rand(B, generator=gen) * 2 - 1
for precision in ('fp32', 'bf16'):
    for graph in (True, False):
        s = srgan_b200.Settings()
        s.batch_size, s.gradient_penalty_multiplier, s.precision, s.use_cuda_graph = B, 10.0, precision, graph
        exp = srgan_b200.Experiment(s, 'coefficient', method)
        xc, yc, uc = x.cuda(), y.cuda(), u.cuda()
        for i in range(5):
            exp.dnn_training_step(xc, yc, i); exp.gan_training_step(xc, yc, uc, i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        N = 200
        t0 = time.perf_counter()
        e0.record()
        for i in range(N):
            exp.dnn_training_step(xc, yc, i); exp.gan_training_step(xc, yc, uc, i)
        e1.record()
        torch.cuda.synchronize()
        print(f'coefficient {method} B={B} {precision} graph={graph}: {e0.elapsed_time(e1) / N * 1e3:.1f} us/step device, '
              f'{(time.perf_counter() - t0) / N * 1e6:.1f} us/step wall, scalars {exp.runner.scalars()}', flush=True)
