"""Crowd SR-GAN step (BASELINE configs[2]: DCGenerator + KnnDenseNetCat/DenseNet-201 at 224x224) on one GPU: ms/step,
launches/step, peak memory.  usage: python tools/crowd_bench.py [B] [precision] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srgan_b200
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
s = srgan_b200.Settings()
s.batch_size, s.precision = B, precision
s.matching_loss_multiplier, s.contrasting_loss_multiplier, s.gradient_penalty_multiplier, s.map_multiplier = 1e3, 1e2, 1e2, 1e-3
s.use_cuda_graph = os.environ.get('SRGAN_NO_GRAPH', '0') != '1'
exp = srgan_b200.Experiment(s, 'crowd')
x, y, u = bench.make_batches('crowd', B, 3)
x, u, y = x.cuda(), u.cuda(), tuple(t.cuda() for t in y)
t0 = time.perf_counter()
for i in range(3):
    exp.dnn_training_step(x, y, i); exp.gan_training_step(x, y, u, i)
torch.cuda.synchronize()
print(f'warm-up (3 steps incl. graph capture): {time.perf_counter() - t0:.2f} s', flush=True)
l0 = exp.runner.engine.ops.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get('SRGAN_OPTIME', '0') == '1':
    # per-op device time (CUDA events around every C-ABI call of one eager step), grouped by op and shape
    import collections
    ops = exp.runner.engine.ops
    exp.runner.use_cuda_graph = False
    recs = []
    names = ['conv_down', 'conv_up', 'conv_wgrad', 'affine', 'affine_bwd', 'affine_grad', 'copy2d', 'maxpool', 'maxpool_bwd',
             'avgpool', 'avgpool_bwd', 'colsum', 'adam', 'seed_rows', 'rowdot', 'crowd_loss', 'crowd_map_grad', 'im2col', 'col2im',
             'nchw_to_nhwc', 'interpolate', 'gradnorm_penalty', 'feature_norm_seed', 'gp_feature_seed', 'distance', 'repack']
    def wrap(name, fn):
        def w(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(*a, **k); e1.record()
            key = name
            if name.startswith('conv_'):
                g, n = a[4], a[3]
                key = f'{name} n={n} {g.Hs}x{g.Ws}x{g.Ca}<-{g.Hl}x{g.Wl}x{g.Cb} k{g.R}s{g.stride} tensor={ops.lib.srgan_last_path_tensor()}'
            recs.append((key, name, e0, e1))
            return r
        return w
    for nme in names:
        setattr(ops, nme, wrap(nme, getattr(ops, nme)))
    exp.dnn_training_step(x, y, 9); exp.gan_training_step(x, y, u, 9)
    torch.cuda.synchronize()
    by_key, by_name = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(lambda: [0, 0.0])
    for key, name, e0, e1 in recs:
        t = e0.elapsed_time(e1)
        by_key[key][0] += 1; by_key[key][1] += t; by_name[name][0] += 1; by_name[name][1] += t
    tot = sum(v[1] for v in by_name.values())
    print(f'op-timed eager step: {tot:.1f} ms in {len(recs)} calls')
    for k, (c, t) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
        print(f'  {t:9.2f} ms {100 * t / tot:5.1f}% {c:5d}  {k}')
    print('top shapes:')
    for k, (c, t) in sorted(by_key.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f'  {t:9.2f} ms {100 * t / tot:5.1f}% {c:5d}  {k}')
    sys.exit(0)
prof = os.environ.get('SRGAN_PROFILE', '0') == '1'       # ncu --profile-from-start off: only the timed steps are captured
if prof:
    torch.cuda.profiler.start()
e0.record()
for i in range(steps):
    exp.dnn_training_step(x, y, 3 + i); exp.gan_training_step(x, y, u, 3 + i)
e1.record()
torch.cuda.synchronize()
if prof:
    torch.cuda.profiler.stop()
print(f'launches/step {(exp.runner.engine.ops.launches - l0) / steps:.0f}')
ms = e0.elapsed_time(e1) / steps
flops = 193.8e9 * B
print(f'crowd srgan B={B} {precision} graph={exp.runner.use_cuda_graph}: {ms:.2f} ms/step, {1e3 / ms:.2f} steps/s, '
      f'{B * 1e3 / ms:.1f} samples/s, {flops / ms / 1e9:.1f} algorithmic TFLOP/s, '
      f'peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, scalars {exp.runner.scalars()}', flush=True)
