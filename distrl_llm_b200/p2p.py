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
        self.last_timing = None   # CUDA events of the last reduce_adam_step(timing=True)
        self.step_rendezvous = None   # host-side rendezvous before each exchange (same-process learners, see wire_same_process)

    # ---- phase 1 ---------------------------------------------------------------------------------
    def alloc_local(self, numel):
        """Allocate the IPC-exportable buffers; returns ({name: 64-byte handle}, kwargs for Policy(...))."""
        self.numel = numel
        handles = {}
        with torch.cuda.device(self.device):   # cudaMalloc on THIS learner's device
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

    @staticmethod
    def wire_same_process(groups, threaded=False):
        """Several 'learners' living in ONE process (the thread-per-actor harness of actors.py; tests: fake multi-GPU on a
        single GPU, one stream per rank): peer pointers are the other groups' local pointers, no IPC handle involved.
        threaded=True (one host thread per learner): every exchange starts with a host-side rendezvous of the learner
        threads.  Learners of one process share a CUDA context, and CUDA loads a kernel's code at its first launch, which
        synchronises with the device: a learner still launching new kernels while a peer's flag barrier is already
        spinning would dead-lock the two.  After the rendezvous every learner has ENQUEUED its whole step."""
        if threaded:
            import threading
            bar = threading.Barrier(len(groups))
            for g in groups:
                g.step_rendezvous = bar.wait
        for g in groups:
            assert g.world == len(groups)
            for name in _NAMES:
                g.peers[name] = [h.local[name] for h in groups]
            with torch.cuda.device(g.device):   # plain pointers across devices need explicit peer access
                for h in groups:
                    if h.device != g.device:
                        check(lib().b200rl_p2p_enable_peer_access(h.device.index), "p2p_enable_peer_access")

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

    def attach(self, policy: Policy, host_rendezvous=None):
        """Called once every learner has built its policy on the buffers of alloc_local().  Learners reach this point at
        very different times (weight loading, NF4 quantisation), so rendezvous on the HOST first when a host-side barrier
        is available (`host_rendezvous` callable; default: torch.distributed's barrier when a process group exists; under
        Ray the driver's two-phase RPC sequence of INTEGRATION.md is that rendezvous) and only then on the GPU flags."""
        assert policy.lora_flat.data_ptr() == self.local["params"]
        if host_rendezvous is None:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() == self.world:
                    host_rendezvous = dist.barrier
            except Exception:  # pragma: no cover
                host_rendezvous = None
        if host_rendezvous is not None:
            # this learner's stream only: a DEVICE-wide synchronize would also wait for a peer learner of the same process
            # and device whose flag barrier is already spinning — and that barrier waits for the one launched below
            torch.cuda.current_stream(self.device).synchronize()
            host_rendezvous()
        self.barrier()  # every learner finished initialising its parameters

    def _ptr_array(self, name):
        return (C.c_void_p * self.world)(*self.peers[name])

    def barrier(self, timeout_s=0.0):
        """GPU-side flag barrier on this stream.  A peer that does not arrive within timeout_s (0 = B200RL_P2P_TIMEOUT_S,
        default 600 s) is reported by check() instead of trapping the CUDA context."""
        self.epoch += 1
        check(lib().b200rl_p2p_barrier_timeout(C.cast(self._ptr_array("flags"), C.c_void_p), self.world, self.rank,
                                               self.epoch, float(timeout_s), stream()), "p2p_barrier")

    @staticmethod
    def reset_status():
        """Clear the process-wide 'a barrier gave up' word without raising (tests)."""
        lib().b200rl_p2p_status(1)

    def check(self, reset=False):
        """Raise if a barrier of this process timed out (call after a synchronisation; the loss .item() of a step is one)."""
        check(lib().b200rl_p2p_status(1 if reset else 0), "p2p_status")

    def reduce_adam_step(self, policy: Policy, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, timing=False):
        """barrier -> fused P2P reduce + Adam(W) + write-back -> barrier -> zero_grad + bf16 operand refresh.
        timing=True records CUDA events around the three phases on this stream (read them with exchange_ms())."""
        policy.opt_step += 1
        if self.step_rendezvous is not None:
            self.step_rendezvous()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timing else None
        if ev:
            ev[0].record()
        self.barrier()   # all learners' gradients are complete
        if ev:
            ev[1].record()
        check(lib().b200rl_lora_reduce_adamw(
            policy.lora_flat.data_ptr(), policy.adam_m.data_ptr(), policy.adam_v.data_ptr(),
            C.cast(self._ptr_array("grads"), C.c_void_p), C.cast(self._ptr_array("params"), C.c_void_p),
            self.world, self.rank, self.numel, policy.opt_step, lr, betas[0], betas[1], eps, weight_decay, 0,
            stream()), "lora_reduce_adamw")
        self.barrier()   # every peer has read my gradients and written its slice of my parameters
        if ev:
            ev[2].record()
        policy.lora_grad.zero_()
        policy.sync_lora()
        if ev:
            ev[3].record()
            self.last_timing = ev

    def exchange_ms(self):
        """(wait for the slowest learner, reduce + Adam + write-back + closing barrier, zero_grad + operand refresh) in ms
        for the last reduce_adam_step(timing=True); synchronises on the last event."""
        ev = self.last_timing
        if ev is None:
            return None
        ev[3].synchronize()
        return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])

    def nvlink_bytes_per_step(self):
        """Bytes this learner moves over NVLink per step: loads of its slice from world-1 peers + stores to world-1 peers."""
        lo, hi = owned_slice(self.numel, self.world, self.rank)
        return 2 * (self.world - 1) * (hi - lo) * 4

    def close(self, free_local=True):
        """Unmap the peers' buffers and (free_local) release this learner's own allocations.  The policy built on them
        must not be used afterwards."""
        for p in self._opened:
            lib().b200rl_p2p_close(p)
        self._opened = []
        if free_local:
            for name, p in list(self.local.items()):
                lib().b200rl_p2p_free(p)
            self.local = {}
        self.peers = {}


def owned_slice(n, world, rank):
    """[lo, hi) of the flat buffer that `rank` reduces and updates — mirrors b200rl_lora_reduce_adamw
    (csrc/optim.cu): contiguous, 4-element aligned, remainder absorbed by the last ranks' clamping."""
    n4 = n // 4
    per = (n4 + world - 1) // world
    lo = 4 * min(per * rank, n4)
    hi = 4 * min(per * (rank + 1), n4)
    return lo, hi
