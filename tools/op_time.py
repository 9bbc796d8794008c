"""Per-op device time of one eager training step (CUDA events around every C-ABI call), grouped by op and by shape.
usage: python tools/op_time.py [age|crowd|coefficient] [B] [precision]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srgan_b200
import bench

name = sys.argv[1] if len(sys.argv) > 1 else 'age'
wl = bench.WORKLOADS[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl['batch']
s = srgan_b200.Settings()
s.batch_size, s.precision, s.map_multiplier = B, (sys.argv[3] if len(sys.argv) > 3 else 'bf16'), 1e-3
s.matching_loss_multiplier, s.contrasting_loss_multiplier, s.gradient_penalty_multiplier = wl['mult']
s.use_cuda_graph = False
s.use_persistent_kernel = False
kw = dict(image_size=128, conv_dim=64, z_dim=256) if name == 'age' else {}
exp = srgan_b200.Experiment(s, name, **kw)
x, y, u = bench.make_batches(name, B, 1)
cu = lambda t: tuple(e.cuda() for e in t) if isinstance(t, tuple) else t.cuda()
x, y, u = cu(x), cu(y), cu(u)
for i in range(3):
    exp.dnn_training_step(x, y, i); exp.gan_training_step(x, y, u, i)
torch.cuda.synchronize()
ops = exp.runner.engine.ops
recs = []
skip = {'begin', 'launches', 'pointer_table', 'coefficient_workspace_bytes'}
def wrap(nm, fn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        key = nm
        if nm.startswith('conv_'):
            g, n = a[4], a[3]
            key = f'{nm} n={n} {g.Hs}x{g.Ws}x{g.Ca}<-{g.Hl}x{g.Wl}x{g.Cb} k{g.R}s{g.stride} tensor={ops.lib.srgan_last_path_tensor()}'
        elif nm in ('colsum',):
            key = f'{nm} rows={a[1]} cols={a[2]}'
        elif nm == 'adam':
            key = f'{nm} dims={tuple(a[4])}'
        recs.append((key, nm, e0, e1))
        return r
    return w
for nm in dir(ops):
    if nm.startswith('_') or nm in skip or not callable(getattr(ops, nm)):
        continue
    setattr(ops, nm, wrap(nm, getattr(ops, nm)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
exp.dnn_training_step(x, y, 9); exp.gan_training_step(x, y, u, 9)
e1.record()
torch.cuda.synchronize()
by_key, by_name = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(lambda: [0, 0.0])
for key, nm, a, b in recs:
    t = a.elapsed_time(b)
    by_key[key][0] += 1; by_key[key][1] += t; by_name[nm][0] += 1; by_name[nm][1] += t
tot = sum(v[1] for v in by_name.values())
print(f'{name} B={B} {s.precision}: op-timed eager step {tot:.2f} ms in {len(recs)} calls (wall of the eager step on the device: {e0.elapsed_time(e1):.2f} ms)')
for k, (c, t) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
    print(f'  {t:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {k}')
cat = collections.defaultdict(lambda: [0, 0.0])
for key, nm, a, b in recs:
    m = __import__('re').match(r'(conv_\w+) n=(\d+) (\d+)x(\d+)x(\d+)<-(\d+)x(\d+)x(\d+) k(\d+)s(\d+) tensor=(\d)', key)
    if m:
        op, n, hs, ws, ca, hl, wl, cb, k, st, tc = m.groups()
        c = f'{op} k{k}s{st} ' + ('gemm(1x1)' if hl == '1' and k == '1' else f'{hs}x{ws}') + f' tensor={tc}'
        cat[c][0] += 1; cat[c][1] += a.elapsed_time(b)
print('contractions by class:')
for k, (c, t) in sorted(cat.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f'  {t:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {k}')
print('top shapes:')
for k, (c, t) in sorted(by_key.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'  {t:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {k}')
