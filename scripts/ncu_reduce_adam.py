"""reduce_adam_kernel<world> alone, for ncu: world buffer sets of the Qwen2.5-7B rank-16 LoRA size
(40 370 176 fp32 parameters) on ONE GPU, each "rank" launches the fused reduce + Adam + write-back over its owned slice on
the same stream — no barrier kernels (they spin until all ranks arrived, which cannot happen while ncu serialises
kernels).  Peer reads / writes stay inside this GPU's HBM, so the capture shows the kernel's own efficiency (bytes per
launch against the HBM roofline), not NVLink; the in-step NVLink figures come from bench.py --gpus N (`exchange`)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distrl_llm_b200._capi import check, lib, stream  # noqa: E402
from distrl_llm_b200.p2p import P2PGroup  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 40_370_176
dev = torch.device("cuda:0")
groups, bufs = [], []
for r in range(world):
    g = P2PGroup(r, world, dev)
    _, kw = g.alloc_local(n)
    kw["lora_flat"].normal_(0, 0.1)
    kw["lora_grad"].normal_(0, 1.0)
    groups.append(g)
    bufs.append((kw["lora_flat"], kw["lora_grad"], torch.zeros(n, device=dev), torch.zeros(n, device=dev)))
P2PGroup.wire_same_process(groups)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    if it == 2:
        e0.record()
    for r, g in enumerate(groups):
        flat, grad, m, v = bufs[r]
        check(lib().b200rl_lora_reduce_adamw(flat.data_ptr(), m.data_ptr(), v.data_ptr(),
                                             C.cast(g._ptr_array("grads"), C.c_void_p), C.cast(g._ptr_array("params"), C.c_void_p),
                                             world, r, n, it + 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, stream()), "lora_reduce_adamw")
    if it == 2:
        e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / world
slice_n = n / world
# per launch: reads world gradient slices + own param/m/v slice, writes m, v and world parameter slices
bytes_per_launch = 4 * slice_n * (world + 3 + 2 + world)
print(f"world {world}: {ms * 1e3:.1f} us per reduce_adam launch, {bytes_per_launch / 1e6:.1f} MB per launch -> {bytes_per_launch / ms / 1e6:.0f} GB/s "
      f"(local HBM; measured copy peak 6570 GB/s)")
