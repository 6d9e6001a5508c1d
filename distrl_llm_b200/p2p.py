"""Multi-learner gradient exchange: peer-mapped flat buffers + the fused one-shot reduce/Adam kernel.

Replaces the reference's exchange (distributed_trainer.py:325-342 driver hop; distributed_actor.py:289-293
D2H export, :311-328 CPU merge + H2D, :331-333 step on learner 0 only).  Each learner allocates its flat
LoRA parameter / gradient buffers and a flag array with b200rl_p2p_alloc (cudaMalloc + cudaIpcGetMemHandle),
the 64-byte IPC handles are exchanged ONCE (torch.distributed under torchrun; two Ray RPCs in the reference's
process model), and every peer maps them with cudaIpcOpenMemHandle.  One step then is:
    barrier kernel (all gradients complete)  ->  reduce_adam kernel: each learner reads its 1/N slice of
    every peer's gradient over NVLink, averages, applies Adam(W) and stores the updated slice into every
    peer's parameter buffer  ->  barrier kernel (all stores landed)  ->  local bf16 operand refresh.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from ._capi import check, lib, stream
from .policy import Policy, tensor_from_ptr

_NAMES = ("params", "grads", "flags")


def lora_numel(cfg, max_batch, P, T) -> int:
    ccfg = _capi.ModelConfig(cfg.vocab, cfg.hidden, cfg.inter, cfg.n_layers, cfg.n_q_heads, cfg.n_kv_heads,
                             cfg.head_dim, cfg.lora_r, cfg.lora_scale, cfg.rms_eps, cfg.rope_theta,
                             max_batch * (P + T), max_batch, P + T, max_batch * T)
    n = int(_capi.lib().b200rl_model_lora_numel(C.byref(ccfg)))
    if n <= 0:
        raise RuntimeError("libb200rl: " + _capi.lib().b200rl_last_error().decode())
    return n


class P2PGroup:
    """Two-phase setup: alloc_local() -> handles (send them to every peer) ; open_peers(all handles)."""

    def __init__(self, rank, world, device):
        _capi.load_library()
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.epoch = 0
        self.local = {}    # name -> local device pointer
        self.peers = {}    # name -> [pointer on every rank]
        self._opened = []
        self.numel = 0

    # ---- phase 1 ---------------------------------------------------------------------------------
    def alloc_local(self, numel):
        """Allocate the IPC-exportable buffers; returns ({name: 64-byte handle}, kwargs for Policy(...))."""
        self.numel = numel
        handles = {}
        for name, nbytes in (("params", numel * 4), ("grads", numel * 4), ("flags", 64 * 4)):
            p = C.c_void_p()
            h = (C.c_ubyte * 64)()
            check(lib().b200rl_p2p_alloc(nbytes, C.byref(p), h), "p2p_alloc")
            self.local[name] = p.value
            handles[name] = bytes(h)
        kw = {"lora_flat": tensor_from_ptr(self.local["params"], numel, torch.float32, self.device),
              "lora_grad": tensor_from_ptr(self.local["grads"], numel, torch.float32, self.device)}
        return handles, kw

    # ---- phase 2 ---------------------------------------------------------------------------------
    def open_peers(self, all_handles):
        """all_handles[r] = the dict returned by rank r's alloc_local()."""
        assert len(all_handles) == self.world
        for name in _NAMES:
            ptrs = []
            for r, hd in enumerate(all_handles):
                if r == self.rank:
                    ptrs.append(self.local[name])
                else:
                    q = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(hd[name])
                    check(lib().b200rl_p2p_open(buf, C.byref(q)), "p2p_open")
                    self._opened.append(q.value)
                    ptrs.append(q.value)
            self.peers[name] = ptrs

    @classmethod
    def from_torch_distributed(cls, cfg, max_batch, P, T, device):
        """torchrun path (bench.py): exchange the handles with all_gather_object."""
        import torch.distributed as dist
        g = cls(dist.get_rank(), dist.get_world_size(), device)
        handles, kw = g.alloc_local(lora_numel(cfg, max_batch, P, T))
        allh = [None] * g.world
        dist.all_gather_object(allh, handles)
        g.open_peers(allh)
        return g, kw

    def attach(self, policy: Policy):
        assert policy.lora_flat.data_ptr() == self.local["params"]
        self.barrier()  # every learner finished initialising its parameters

    def _ptr_array(self, name):
        return (C.c_void_p * self.world)(*self.peers[name])

    def barrier(self):
        self.epoch += 1
        check(lib().b200rl_p2p_barrier(C.cast(self._ptr_array("flags"), C.c_void_p), self.world, self.rank,
                                       self.epoch, stream()), "p2p_barrier")

    def reduce_adam_step(self, policy: Policy, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        policy.opt_step += 1
        self.barrier()   # all learners' gradients are complete
        check(lib().b200rl_lora_reduce_adamw(
            policy.lora_flat.data_ptr(), policy.adam_m.data_ptr(), policy.adam_v.data_ptr(),
            C.cast(self._ptr_array("grads"), C.c_void_p), C.cast(self._ptr_array("params"), C.c_void_p),
            self.world, self.rank, self.numel, policy.opt_step, lr, betas[0], betas[1], eps, weight_decay, 0,
            stream()), "lora_reduce_adamw")
        self.barrier()   # every peer has read my gradients and written its slice of my parameters
        policy.lora_grad.zero_()
        policy.sync_lora()

    def close(self):
        for p in self._opened:
            lib().b200rl_p2p_close(p)
        self._opened = []


def owned_slice(n, world, rank):
    """[lo, hi) of the flat buffer that `rank` reduces and updates — mirrors b200rl_lora_reduce_adamw
    (csrc/optim.cu): contiguous, 4-element aligned, remainder absorbed by the last ranks' clamping."""
    n4 = n // 4
    per = (n4 + world - 1) // world
    lo = 4 * min(per * rank, n4)
    hi = 4 * min(per * (rank + 1), n4)
    return lo, hi
