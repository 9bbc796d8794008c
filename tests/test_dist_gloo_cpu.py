"""Host-logic test of the data-parallel path with world_size 2 on CPU (gloo): every rank holds half the batch,
feature sums and gradients are all-reduced (sr-gan_b200/dist.py), and both ranks must end with the parameters and the
losses of the single-process full-batch oracle step.  Uses the TEST-ONLY torch emulation of the op set."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import srgan_oracle as O
    from srgan_b200 import nets, engine
    from srgan_b200.dist import Comm, shard
    from tests.golden_io import Golden, SCALARS
    from tests.torch_ops import TorchOps
    from tests.test_engine_schedule_cpu import build_engine, read_scalars
    g = Golden(name)
    dt = torch.float64
    st, cfg = g.oracle_state(dt), g.step_config()
    comm = Comm()
    eng = build_engine(st, dt, g.cfg.get('image_size'), g.cfg.get('conv_dim'), g.cfg.get('z_dim'), comm=comm)
    errs = []
    for i in range(g.steps):
        full = g.step_inputs(i, dt)
        ref = O.training_step(st, cfg, *full, step=i)            # single-process, GLOBAL batch
        x, y, u, z, alpha, z2 = (shard(t, rank, world) for t in full)
        eng.dnn_step(x, y, cfg, O.dnn_lr(cfg, i), cfg.weight_decay)
        eng.gan_step(x, y, u, z, alpha, z2, cfg)
        sc = eng.scalars.clone()
        comm.all_reduce_sum_partial(sc, engine.partial_scalar_slots(cfg.method))
        eng_sc = dict(zip(('dnn_loss', 'labeled_loss', 'unlabeled_loss', 'fake_loss', 'gradient_penalty',
                           'gradient_norm_mean', 'generator_loss'), sc.tolist()))
        for k in SCALARS:
            errs.append(abs(eng_sc[k] - ref[k]) / max(abs(ref[k]), 1e-9))
    perr = 0.0
    for params, mine in ((st.D, eng.D), (st.G, eng.G), (st.DNN, eng.DNN)):
        for k, v in params.items():
            perr = max(perr, (mine.params[k] - v).abs().max().item() / max(v.abs().max().item(), 1e-30))
    q.put((rank, max(errs), perr, comm.calls))
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['coefficient_srgan', 'coefficient_dggan', 'dcgan_mini', 'dcgan_sgan_mini'])
def test_two_rank_step_equals_global_batch_step(name):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, serr, perr, calls in res:
        assert serr < 1e-9, (name, rank, 'scalars', serr)
        assert perr < 1e-9, (name, rank, 'params', perr)
        assert calls > 0
