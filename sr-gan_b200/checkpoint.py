"""
Asynchronous and sharded checkpoint writing for the reference's `save_models` (srgan.py:88-97).

The reference serialises six state dicts (three networks + three Adam states, ~0.8 GB for the crowd models) with one
blocking `torch.save` of CUDA tensors: the training loop stops for the device -> host copy AND the pickling / disk write.
Here:
  * `snapshot`   copies every tensor of the checkpoint into pinned host memory on a side stream (the loop only waits for
                 the step that produced the parameters, not for the copy);
  * `AsyncWriter` pickles and writes the snapshot on a background thread once the copy's event has completed, to a
                 temporary name (`.model_{step}.tmp`) renamed on completion, so a reader never sees a partial `model_{step}.pth`;
  * `shard` / `merge_shards` split the checkpoint's tensors round-robin over the ranks of a data-parallel job (every rank
                 holds identical parameters, so each writes 1/W of the bytes: `model_{step}.shard{r}of{W}.pth`) and put
                 them back together into exactly the dict `load_models` (srgan.py:182-199) expects.
The file format stays `torch.save` of plain dicts: a checkpoint written here loads in the unmodified reference.
"""
from __future__ import annotations

import os
import threading

import torch


def _map_tensors(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return type(obj)((k, _map_tensors(v, fn)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_tensors(v, fn) for v in obj)
    return obj


def snapshot(model: dict, stream=None):
    """Copies every tensor of `model` to the host (pinned memory for CUDA tensors, asynchronously on `stream`).  Returns
    (host copy, event to wait for before reading it | None)."""
    cuda = [False]

    def to_host(t):
        if not t.is_cuda:
            return t.detach().clone()
        cuda[0] = True
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t.detach(), non_blocking=True)
        return host
    devices = []
    _map_tensors(model, lambda t: devices.append(t.device) if t.is_cuda else None)
    if not devices:
        return _map_tensors(model, to_host), None
    dev = devices[0]                                           # one process drives one GPU: the checkpoint lives on it
    with torch.cuda.device(dev):
        stream = stream or torch.cuda.Stream(dev)
        stream.wait_stream(torch.cuda.current_stream(dev))     # the parameters as of the steps enqueued so far
        with torch.cuda.stream(stream):
            host = _map_tensors(model, to_host)
            event = torch.cuda.Event()
            event.record(stream)
    return host, (event if cuda[0] else None)


class AsyncWriter:
    """One background writer per experiment: `save(model, path)` returns as soon as the device -> host copies are enqueued;
    `wait()` blocks until every pending file is complete (called before the next save, and by the mix-in at the end of
    training so that the reference's contract -- the file exists when `train()` returns -- holds)."""

    def __init__(self):
        self._threads = []
        self._stream = None
        self.errors = []

    def save(self, model: dict, path: str):
        self.wait()
        host, event = snapshot(model, self._stream)

        def write():
            try:
                if event is not None:
                    event.synchronize()
                # not `model_*.pth*`: load_models' regex (srgan.py:227) must never pick up a file that is still being written
                tmp = os.path.join(os.path.dirname(path), '.' + os.path.splitext(os.path.basename(path))[0] + '.tmp')
                torch.save(host, tmp)
                os.replace(tmp, path)
            except Exception as e:                          # surfaced by wait()
                self.errors.append(e)
        t = threading.Thread(target=write, name='srgan-checkpoint', daemon=False)
        t.start()
        self._threads.append(t)

    def wait(self):
        for t in self._threads:
            t.join()
        self._threads = []
        if self.errors:
            raise self.errors.pop(0)


def _flatten(obj, prefix, out):
    if torch.is_tensor(obj):
        out.append(prefix)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            _flatten(v, prefix + (k,), out)
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            _flatten(v, prefix + (i,), out)


def shard(model: dict, rank: int, world_size: int) -> dict:
    """The part of `model` rank `rank` writes: tensor leaves are dealt round-robin over the ranks in traversal order (the
    other ranks' tensors are dropped from the copy); non-tensor leaves (step counters, param_groups) travel with rank 0."""
    paths = []
    _flatten(model, (), paths)
    mine = {p for i, p in enumerate(paths) if i % world_size == rank}

    def take(obj, prefix):
        if torch.is_tensor(obj):
            return obj if prefix in mine else None
        if isinstance(obj, dict):
            return {k: take(v, prefix + (k,)) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(take(v, prefix + (i,)) for i, v in enumerate(obj))
        return obj if rank == 0 else None
    return {'shard': rank, 'world_size': world_size, 'model': take(model, ())}


def merge_shards(shards) -> dict:
    """Inverse of `shard`: the full checkpoint dict from all W shard dicts (any order)."""
    shards = sorted(shards, key=lambda s: s['shard'])
    W = shards[0]['world_size']
    if [s['shard'] for s in shards] != list(range(W)):
        raise ValueError(f'need shards 0..{W - 1}, got {[s["shard"] for s in shards]}')

    def join(parts):
        first = parts[0]
        if isinstance(first, dict):
            return {k: join([p[k] for p in parts]) for k in first}
        if isinstance(first, (list, tuple)):
            return type(first)(join([p[i] for p in parts]) for i in range(len(first)))
        for p in parts:
            if p is not None:
                return p
        return None
    return join([s['model'] for s in shards])


def shard_path(path: str, rank: int, world_size: int) -> str:
    base, ext = os.path.splitext(path)
    return f'{base}.shard{rank}of{world_size}{ext}'


def load(path: str, map_location='cpu') -> dict:
    """`torch.load` of a checkpoint written whole or in shards (`path` is the un-sharded name)."""
    if os.path.exists(path):
        return torch.load(path, map_location=map_location, weights_only=False)
    base, ext = os.path.splitext(path)
    directory = os.path.dirname(path) or '.'
    names = [n for n in os.listdir(directory) if n.startswith(os.path.basename(base) + '.shard') and n.endswith(ext)]
    if not names:
        raise FileNotFoundError(path)
    return merge_shards([torch.load(os.path.join(directory, n), map_location=map_location, weights_only=False) for n in names])
