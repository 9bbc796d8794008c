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
s.overlap_dnn_step = False           # per-op events need one stream
kw = dict(image_size=128, conv_dim=64, z_dim=256) if name in ('age', 'driving') else {}
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
def algo_bytes(nm, a):
    """Algorithmic HBM bytes of one call (every operand read or written once), 0 = not modelled."""
    try:
        if nm == 'affine':           # (x, x_pitch, x_c0, y, y_pitch, rows, C, ..., href, mode, ...)
            e = a[0].element_size()
            return a[5] * a[6] * e * (3 if a[13] == 1 else 2)
        if nm == 'affine_bwd':       # (dy, dy_pitch, dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, accumulate)
            return a[5] * a[6] * a[0].element_size() * (3 if a[10] else 2)
        if nm == 'affine_grad':      # (dy, dy_pitch, x, x_pitch, x_c0, rows, C, ...)
            return a[5] * a[6] * a[0].element_size() * 2
        if nm == 'affine_bwd_grad':  # (dy, dy_pitch, x, dx, x_pitch, x_c0, rows, C, ..., accumulate)
            return a[6] * a[7] * a[0].element_size() * (4 if a[14] else 3)
        if nm == 'copy2d':           # (src, sp, s0, dst, dp, d0, rows, C, accumulate)
            return a[6] * a[7] * a[0].element_size() * (3 if a[8] else 2)
        if nm in ('conv_down', 'conv_up'):   # (src, W, out, n, g, bias, bias_mod, href, epi, act, slope)
            g, n, e = a[4], a[3], a[0].element_size()
            small, large = n * g.Hs * g.Ws * g.Ca, n * g.Hl * g.Wl * g.Cb
            href = (small if nm == 'conv_down' else large) if (a[7] is not None and a[9] != 0) else 0
            return (small + large + href) * e + a[1].numel() * a[1].element_size()
        if nm == 'conv_wgrad':       # (S, L, dW, n, g)
            g, n, e = a[4], a[3], a[0].element_size()
            return (n * g.Hs * g.Ws * g.Ca + n * g.Hl * g.Wl * g.Cb) * e + 2 * 4 * g.Ca * g.Cb * g.R * g.S
    except Exception:
        pass
    return 0


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
        elif nm in ('affine', 'affine_bwd', 'affine_grad'):
            key = f'{nm} rows={a[5]} C={a[6]}'
        elif nm == 'affine_bwd_grad':
            key = f'{nm} rows={a[6]} C={a[7]}'
        elif nm == 'adam':
            key = f'{nm} dims={tuple(a[4])}'
        recs.append((key, nm, e0, e1, algo_bytes(nm, a)))
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
by_key, by_name = collections.defaultdict(lambda: [0, 0.0, 0]), collections.defaultdict(lambda: [0, 0.0, 0])
for key, nm, a, b, nb in recs:
    t = a.elapsed_time(b)
    by_key[key][0] += 1; by_key[key][1] += t; by_key[key][2] += nb
    by_name[nm][0] += 1; by_name[nm][1] += t; by_name[nm][2] += nb
tot = sum(v[1] for v in by_name.values())
print(f'{name} B={B} {s.precision}: op-timed eager step {tot:.2f} ms in {len(recs)} calls (wall of the eager step on the device: {e0.elapsed_time(e1):.2f} ms)')
gbs = lambda nb, t: f'{nb / 1e9:8.2f} GB {nb / t / 1e6:6.0f} GB/s' if nb else ' ' * 23       # algorithmic bytes / device time
for k, (c, t, nb) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
    print(f'  {t:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {gbs(nb, t)}  {k}')
cat = collections.defaultdict(lambda: [0, 0.0, 0])
for key, nm, a, b, nb in recs:
    m = __import__('re').match(r'(conv_\w+) n=(\d+) (\d+)x(\d+)x(\d+)<-(\d+)x(\d+)x(\d+) k(\d+)s(\d+) tensor=(\d)', key)
    if m:
        op, n, hs, ws, ca, hl, wl, cb, k, st, tc = m.groups()
        c = f'{op} k{k}s{st} ' + ('gemm(1x1)' if hl == '1' and k == '1' else f'{hs}x{ws}') + f' tensor={tc}'
        cat[c][0] += 1; cat[c][1] += a.elapsed_time(b); cat[c][2] += nb
print('contractions by class:')
for k, (c, t, nb) in sorted(cat.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f'  {t:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {gbs(nb, t)}  {k}')
print('top shapes:')
for k, (c, t, nb) in sorted(by_key.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'  {t:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {gbs(nb, t)}  {k}')
# streaming ops by launch size: where the bandwidth is lost
print('streaming ops by algorithmic bytes per launch:')
bins = collections.defaultdict(lambda: [0, 0.0, 0])
for key, nm, a, b, nb in recs:
    if nb and not nm.startswith('conv_'):
        lg = max(0, int(nb).bit_length() - 20)            # 2^lg MB
        e = bins[(nm, lg)]
        e[0] += 1; e[1] += a.elapsed_time(b); e[2] += nb
for (nm, lg), (c, t, nb) in sorted(bins.items()):
    print(f'  {nm:16s} {2 ** lg // 2:5d}-{2 ** lg:<5d} MB {c:5d} calls {t:8.3f} ms  {gbs(nb, t)}')
