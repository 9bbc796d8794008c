"""CPU: checkpoint.AsyncWriter / shard / merge_shards keep the reference's save_models format
(srgan.py:88-97): a dict of state dicts that torch.load + load_state_dict accept."""
import os

import torch

from srgan_b200 import checkpoint


def _model(seed=0):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 2))
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
    for _ in range(3):
        opt.zero_grad()
        net(torch.randn(4, 5)).square().mean().backward()
        opt.step()
    return net, opt, {'D': net.state_dict(), 'd_optimizer': opt.state_dict(), 'step': 3}


def _same(a, b):
    if torch.is_tensor(a):
        return torch.is_tensor(b) and a.dtype == b.dtype and torch.equal(a, b)
    if isinstance(a, dict):
        return isinstance(b, dict) and list(a.keys()) == list(b.keys()) and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return type(a) == type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


def test_async_writer_writes_the_reference_format(tmp_path):
    net, opt, model = _model()
    path = os.path.join(tmp_path, 'model_3.pth')
    w = checkpoint.AsyncWriter()
    w.save(model, path)
    with torch.no_grad():                       # training goes on: the snapshot must not see this
        for p in net.parameters():
            p.add_(1.0)
    w.wait()
    assert os.listdir(tmp_path) == ['model_3.pth']
    loaded = torch.load(path, weights_only=False)
    _, _, again = _model()
    assert _same(loaded, again)
    net2, opt2, _ = _model(seed=1)
    net2.load_state_dict(loaded['D'])
    opt2.load_state_dict(loaded['d_optimizer'])


def test_shards_merge_back_to_the_full_checkpoint(tmp_path):
    _, _, model = _model()
    base = os.path.join(tmp_path, 'model_3.pth')
    W = 3
    sizes = []
    for r in range(W):
        part = checkpoint.shard(model, r, W)
        torch.save(part, checkpoint.shard_path(base, r, W))
        sizes.append(sum(1 for _ in _tensors(part['model'])))
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == sum(1 for _ in _tensors(model))
    merged = checkpoint.load(base)
    assert _same(merged, model)
    try:
        checkpoint.merge_shards([checkpoint.shard(model, 0, 2)])
    except ValueError:
        pass
    else:
        raise AssertionError('a missing shard must raise')


def _tensors(obj):
    if torch.is_tensor(obj):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _tensors(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _tensors(v)
