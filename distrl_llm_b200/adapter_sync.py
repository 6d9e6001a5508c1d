"""Learner -> generator LoRA hand-off WITHOUT the filesystem.

Reference: after every trainer step learner 0 writes the adapter to `lora_save_path` (`save_lora`,
distributed_actor.py:84-86, called from distributed_trainer.py:346) and every generator re-reads it from disk on every
`generate` (`load_lora`, distributed_actor.py:150).  Here the adapter is ONE flat fp32 device buffer (policy.lora_flat,
161 MB at rank 16), so the hand-off is a single device-to-device copy:

  AdapterPublisher (learner 0)   owns an IPC-exportable mirror of the flat buffer plus a version word.  publish()
                                 = one D2D copy inside the learner's stream (0.1 ms) and a version bump (seqlock: odd while
                                 the copy is in flight, even when complete).
  AdapterSubscriber (generator)  maps the mirror (CUDA IPC handle across processes; plain pointer + peer access inside one
                                 process) and pull()s it into its own flat buffer: NVLink copy GPU -> GPU, 161 MB at
                                 ~700 GB/s = 0.25 ms, retried if the version changed under the copy.  as_peft_tensors()
                                 views the pulled buffer under the PEFT names vLLM's in-memory LoRA loading expects
                                 (`load_lora(..., load_tensors=True)` in unsloth_zoo [3P]) — no safetensors file involved.

The file-based path (`save_adapter()` writing a PEFT directory) remains available for generators that are not on this
node (config["adapter_sync"] = "file").
"""
from __future__ import annotations

import ctypes as C

import torch

from ._capi import check, lib
from .policy import tensor_from_ptr


class AdapterPublisher:
    def __init__(self, policy):
        self.policy = policy
        self.device = policy.device
        self.numel = int(policy.lora_numel)
        self._ptrs, self._handles = {}, {}
        with torch.cuda.device(self.device):
            for name, nbytes in (("mirror", self.numel * 4), ("version", 256)):
                p = C.c_void_p()
                h = (C.c_ubyte * 64)()
                check(lib().b200rl_p2p_alloc(nbytes, C.byref(p), h), "p2p_alloc")
                self._ptrs[name], self._handles[name] = p.value, bytes(h)
        self.mirror = tensor_from_ptr(self._ptrs["mirror"], self.numel, torch.float32, self.device)
        self.version_word = tensor_from_ptr(self._ptrs["version"], 1, torch.int64, self.device)
        self.version = 0          # even = complete

    def publish(self):
        """Enqueue (on the current stream, i.e. after the optimizer step that produced the parameters):
        version -> odd, mirror <- lora_flat, version -> even.  Returns the new version."""
        self.version += 2
        self.version_word.fill_(self.version - 1)
        self.mirror.copy_(self.policy.lora_flat)
        self.version_word.fill_(self.version)
        return self.version

    def describe(self):
        """What a subscriber needs: picklable, safe to send through Ray / a queue."""
        return {"numel": self.numel, "device_index": self.device.index, "handles": dict(self._handles),
                "pointers": dict(self._ptrs)}

    def close(self):
        for p in self._ptrs.values():
            lib().b200rl_p2p_free(p)
        self._ptrs = {}


class AdapterSubscriber:
    def __init__(self, desc, device, same_process=False):
        self.device = torch.device(device)
        self.numel = desc["numel"]
        self._opened = []
        src_dev = torch.device("cuda", desc["device_index"])
        with torch.cuda.device(self.device):
            if same_process:
                ptrs = desc["pointers"]
                if src_dev != self.device:
                    check(lib().b200rl_p2p_enable_peer_access(src_dev.index), "p2p_enable_peer_access")
            else:
                ptrs = {}
                for name, h in desc["handles"].items():
                    q = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(h)
                    check(lib().b200rl_p2p_open(buf, C.byref(q)), "p2p_open")
                    ptrs[name] = q.value
                    self._opened.append(q.value)
        # the mapped tensors are addressed from THIS device (peer mapping)
        self._src = tensor_from_ptr(ptrs["mirror"], self.numel, torch.float32, self.device)
        self._ver = tensor_from_ptr(ptrs["version"], 1, torch.int64, self.device)
        self.flat = torch.zeros(self.numel, device=self.device, dtype=torch.float32)
        self.version = 0

    def pull(self, max_tries=8):
        """Copy the publisher's current adapter into self.flat; returns its version (0 = nothing published yet)."""
        for _ in range(max_tries):
            v1 = int(self._ver.item())
            if v1 % 2 == 1:          # a publish is in flight (the .item() above already waited on this stream): look again
                continue
            if v1 == self.version:
                return v1            # already have it
            self.flat.copy_(self._src)
            v2 = int(self._ver.item())
            if v1 == v2:
                self.version = v1
                return v1
        raise RuntimeError("adapter pull: the publisher kept writing; no consistent snapshot")

    def as_peft_tensors(self, policy_like):
        """{PEFT name: view} of the pulled buffer (policy_like: anything with named_views(), e.g. a Policy of the same
        architecture or policy.LoraLayout)."""
        return policy_like.named_views(self.flat)

    def close(self):
        for p in self._opened:
            lib().b200rl_p2p_close(p)
        self._opened = []
