"""GPU: the data-parallel CUDA path with world_size 2 on ONE device (both ranks on cuda:0, gloo transport, which reduces
CUDA tensors through the host): piecewise CUDA-graph capture around the collectives, the deferred (side-stream) gradient
all-reduce + Adam groups, and global-batch normalisation must reproduce the single-process oracle step on the global
batch -- in eager launches, at capture, and on graph replay (steps 0, 1, 2...)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, family, q, backend='gloo'):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if backend == 'nccl':                                # one GPU per rank, NCCL over NVLink (the bench's transport)
        torch.cuda.set_device(rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    else:
        torch.cuda.set_device(0)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import srgan_oracle as O
    from srgan_b200.dist import Comm, shard
    from tests.gpu_common import runner_from_state, to_cuda
    from tests.golden_io import SCALARS
    comm = Comm()
    gen = torch.Generator().manual_seed(5)
    if family == 'dcgan':
        st = O.init_dcgan(seed=2, image_size=32, conv_dim=16, z_dim=16, scale=3.0)
        cfg = O.StepConfig(batch_size=8, matching_loss_multiplier=1e2, contrasting_loss_multiplier=1e1, gradient_penalty_multiplier=1e2)
        B = 8

        def batch():
            x, u = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1, torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1
            return (x, torch.rand(B, generator=gen) * 85 + 10, u, torch.randn(B, 16, generator=gen),
                    torch.rand(B, 1, 1, 1, generator=gen), torch.randn(B, 16, generator=gen))
    else:
        kw = dict(block_config=(2, 2, 2, 2), growth_rate=8, num_init_features=16, bn_size=2, label_patch_size=64)
        st = O.init_crowd(seed=1, image_size=64, z_dim=16, g_conv_dim=8, scale=2.0, **kw)
        cfg = O.StepConfig(batch_size=4, matching_loss_multiplier=1e3, contrasting_loss_multiplier=1e2,
                           gradient_penalty_multiplier=1e2, map_multiplier=1e-3)
        B = 4
        seeds = iter(range(20, 40))

        def batch():
            return O.synthetic_crowd_batch(B, next(seeds), image=64, label=64, z_dim=16)
    r = runner_from_state(st, cfg, 'fp32', comm=comm)
    assert r.use_cuda_graph
    worst = 0.0
    for i in range(4):                                   # eager, capture, replay, replay
        full = batch()
        ref = O.training_step(st, cfg, *full, step=i)
        mine = tuple((tuple(shard(e, rank, world) for e in t) if isinstance(t, tuple) else shard(t, rank, world)) for t in full)
        x, y, u, z, alpha, z2 = to_cuda(*mine)
        r.dnn_step(x, y)
        r.gan_step(x, y, u, i, noise=(z, alpha, z2))
        got = r.scalars()
        for k in SCALARS:
            worst = max(worst, abs(got[k] - ref[k]) / max(abs(ref[k]), 1e-3 * abs(ref['labeled_loss'])))
    perr = 0.0
    for net, params in (('D', st.D), ('G', st.G), ('DNN', st.DNN)):
        sd = r.modules[net].state_dict()
        for k, v in params.items():
            if O.is_buffer_key(k):
                continue
            perr = max(perr, (sd[k].cpu() - v).abs().max().item() / max(v.abs().max().item(), 1e-30))
    q.put((rank, worst, perr, comm.calls))
    dist.destroy_process_group()


def _run_ranks(world, family, backend):
    import queue
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000) + (7 if backend == 'nccl' else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, family, q, backend)) for r in range(world)]
    for p in procs:
        p.start()
    res, waited = [], 0
    while len(res) < world and waited < 900:             # a worker that died must fail the test, not hang it
        try:
            res.append(q.get(timeout=5))
        except queue.Empty:
            waited += 5
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=120)
        if p.is_alive():
            p.kill()
    assert len(res) == world and all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for rank, serr, perr, calls in res:
        assert serr < 5e-4, (family, rank, 'scalars', serr)          # fp32 mode, 4 steps of drift
        assert perr < 2e-3, (family, rank, 'params', perr)
        assert calls > 0


@pytest.mark.parametrize('family', ['dcgan', 'crowd'])
def test_two_ranks_on_one_gpu_match_global_batch_oracle(family):
    _run_ranks(2, family, 'gloo')


@pytest.mark.parametrize('family', ['dcgan', 'crowd'])
def test_nccl_ranks_match_global_batch_oracle(family):
    """One rank per GPU over NCCL, as bench.py runs N > 1: with the global batch fixed (8 / 4 samples), every world size
    the box offers (2, 4, 8 -- the largest that divides the batch) must reproduce the losses and parameters of the
    single-process global-batch step, i.e. the step's scalars are equal across N = 1/2/4/8 at equal global batch."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least two GPUs (gpurun --gpus N)')
    batch = 8 if family == 'dcgan' else 4
    for world in (8, 4, 2):
        if world <= n and batch % world == 0:
            _run_ranks(world, family, 'nccl')
