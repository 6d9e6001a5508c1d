"""Multi-learner gradient exchange: peer-mapped flat buffers + the fused one-shot reduce/Adam kernel.

Replaces the reference's exchange (distributed_trainer.py:325-342 driver hop; distributed_actor.py:289-293
D2H export, :311-328 CPU merge + H2D, :331-333 step on learner 0 only).  Each learner allocates its flat
LoRA parameter / gradient buffers and a flag array with b200rl_p2p_alloc (cudaMalloc + cudaIpcGetMemHandle),
the 64-byte IPC handles are exchanged ONCE (torch.distributed here; Ray RPC in the reference's process model,
see INTEGRATION.md), and every peer maps them with cudaIpcOpenMemHandle.  One step then is:
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


class P2PGroup:
    def __init__(self, rank, world, device, exchange=None):
        """exchange(obj) -> list of every rank's obj (default: torch.distributed.all_gather_object)."""
        _capi.load_library()
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.exchange = exchange or self._dist_exchange
        self.epoch = 0
        self.local = {}    # name -> local device pointer
        self.peers = {}    # name -> [pointer on every rank]
        self._opened = []

    @staticmethod
    def _dist_exchange(obj):
        import torch.distributed as dist
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out

    def _alloc_shared(self, name, nbytes):
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        check(lib().b200rl_p2p_alloc(nbytes, C.byref(p), h), "p2p_alloc")
        self.local[name] = p.value
        handles = self.exchange(bytes(h))
        ptrs = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                ptrs.append(p.value)
            else:
                q = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(hb)
                check(lib().b200rl_p2p_open(buf, C.byref(q)), "p2p_open")
                self._opened.append(q.value)
                ptrs.append(q.value)
        self.peers[name] = ptrs
        return p.value

    def alloc_lora_buffers(self, cfg, max_batch, P, T):
        """Allocate IPC-exportable flat parameter / gradient buffers; returns kwargs for Policy(...)."""
        probe = Policy.__new__(Policy)  # only to size the buffer
        ccfg = _capi.ModelConfig(cfg.vocab, cfg.hidden, cfg.inter, cfg.n_layers, cfg.n_q_heads, cfg.n_kv_heads,
                                 cfg.head_dim, cfg.lora_r, cfg.lora_scale, cfg.rms_eps, cfg.rope_theta,
                                 max_batch * (P + T), max_batch, P + T, max_batch * T)
        n = int(lib().b200rl_model_lora_numel(C.byref(ccfg)))
        self.numel = n
        pp = self._alloc_shared("params", n * 4)
        gp = self._alloc_shared("grads", n * 4)
        self._alloc_shared("flags", 64 * 4)
        return {"lora_flat": tensor_from_ptr(pp, n, torch.float32, self.device),
                "lora_grad": tensor_from_ptr(gp, n, torch.float32, self.device)}

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
