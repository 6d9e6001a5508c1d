"""Minimal stand-in for the three Ray calls the reference's Trainer uses (`Actor.remote(...)`, `handle.method.remote(...)`,
`ray.get(futures, timeout=...)`; distributed_trainer.py:191-200, :307, :326-342, :374) so that the trainer loop in
trainer.py reads like the reference's and runs in ONE process when Ray is not installed (build image, GPU box).

Semantics kept: every actor executes its method calls serially on its own thread (Ray's default max_concurrency = 1,
SURVEY.md 8b), different actors run concurrently, exceptions surface at get().  Each actor thread binds its CUDA device
once (CUDA's current device is per thread; the library launches on the current device)."""
from __future__ import annotations

import concurrent.futures as cf


class _Method:
    def __init__(self, handle, name):
        self._h, self._name = handle, name

    def remote(self, *args, **kwargs):
        return self._h._pool.submit(self._h._call, self._name, args, kwargs)


class ActorHandle:
    def __init__(self, factory, device=None, own_stream=False):
        """factory() builds the actor object INSIDE the actor's thread (so CUDA state belongs to that thread).
        own_stream: give the thread its own CUDA stream (torch's current stream is per thread) — needed when several
        actors share one GPU, else their kernels would serialise on the default stream (and a learner spinning in the
        P2P flag barrier would block the peer it is waiting for)."""
        self._pool = cf.ThreadPoolExecutor(max_workers=1)
        self._device = device
        self._own_stream = own_stream
        self._obj = None
        self._pool.submit(self._init, factory).result()

    def _init(self, factory):
        if self._device is not None:
            import torch
            torch.cuda.set_device(self._device)
            if self._own_stream:
                torch.cuda.set_stream(torch.cuda.Stream(self._device))
        self._obj = factory()

    def _call(self, name, args, kwargs):
        return getattr(self._obj, name)(*args, **kwargs)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return _Method(self, name)

    def shutdown(self):
        self._pool.shutdown(wait=True)


def get(futures, timeout=None):
    """ray.get: a single future or a list of futures -> result(s)."""
    if isinstance(futures, (list, tuple)):
        return [f.result(timeout=timeout) for f in futures]
    return futures.result(timeout=timeout)
