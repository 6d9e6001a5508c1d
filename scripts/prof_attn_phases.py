"""Where does the tcgen05 attention forward spend its cycles?  Runs the packed BASELINE config-2 pass shape (2 groups x 8
completions of 512 tokens behind 350-token prompts, 28 q / 4 kv heads) with the kernel's per-phase clock64 counters on
(b200rl_attn_set_prof) and prints cycles per key block for the softmax warpgroups and the MMA issuer."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distrl_llm_b200 import _capi, packing  # noqa: E402

NAMES = ["s_full wait", "S tmem ld", "mask+max+exchange", "o_full wait (absorb)", "absorb ld+fma", "exp + P store", "fence+arrive", "loop total",
         "prologue", "epilogue", "MMA: k_full wait", "MMA: s_empty wait", "MMA: p_full wait", "MMA: v_full wait", "MMA: o_empty wait", "MMA total",
         "key blocks", "CTAs"]
dev = torch.device("cuda:0")
nq, nkv, hd, P, T, G, n = 28, 4, 128, 350, 512, 2, 8
B = G * n
ids = np.zeros((B, P + T), np.int32)
rng = np.random.default_rng(0)
for g in range(G):
    pr = rng.integers(1, 1000, size=P)
    for j in range(n):
        ids[g * n + j, :P] = pr
        ids[g * n + j, P:] = rng.integers(1, 1000, size=T)
am = np.ones_like(ids)
host = packing.pack_microbatch(ids, am, P, T)
pk = packing.PackedDevice(host, dev)
rows = host.rows
qkv = (torch.randn(rows, (nq + 2 * nkv) * hd, device=dev) * 0.5).to(torch.bfloat16)
out = torch.empty(rows, nq * hd, device=dev, dtype=torch.bfloat16)
lse = torch.empty(nq, rows, device=dev, dtype=torch.float32)
lib = _capi.lib()
c = pk.c


def fwd():
    _capi.check(lib.b200rl_attn_seg_fwd(qkv.data_ptr(), c.key_mask, out.data_ptr(), lse.data_ptr(), rows, nq, nkv, hd ** -0.5,
                                        c.qblocks, c.n_qblocks, _capi.stream()), "attn_seg_fwd")


for _ in range(3):
    fwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fwd()
e1.record()
torch.cuda.synchronize()
print(f"rows {rows}, q-blocks {c.n_qblocks}, fwd {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch (counters off)")
dout = (torch.randn(rows, nq * hd, device=dev) * 0.5).to(torch.bfloat16)
delta = torch.empty_like(lse)
dqkv = torch.zeros_like(qkv)
kvpart = torch.empty(host.part_rows, 2 * nkv * hd, device=dev, dtype=torch.float32)


def bwd():
    _capi.check(lib.b200rl_attn_seg_bwd(qkv.data_ptr(), c.key_mask, out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                        delta.data_ptr(), dqkv.data_ptr(), kvpart.data_ptr(), rows, nq, nkv, hd ** -0.5,
                                        c.qblocks, c.n_qblocks, c.kblocks, c.n_kblocks, c.red_start, c.red_list,
                                        _capi.stream()), "attn_seg_bwd")


for _ in range(3):
    bwd()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    bwd()
e1.record()
torch.cuda.synchronize()
print(f"k-blocks {c.n_kblocks}, bwd (delta + dQ + dK/dV + reduce) {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch (counters off)")

BWD_NAMES = ["S/dP full wait", "S,dP tmem ld", "mask ballot / stats stage", "dS buffer free wait", "-", "exp + dS math + store", "fence+arrive",
             "loop total", "prologue", "epilogue", "MMA1: operand full wait", "MMA1: S/dP tmem free wait", "MMA2: dS full wait", "MMA2 total", "-", "MMA1 total",
             "blocks", "CTAs"]
for which, title, names, run in ((0, "forward", NAMES, fwd), (1, "dQ kernel", BWD_NAMES, bwd), (2, "dK/dV kernel", BWD_NAMES, bwd)):
    prof = torch.zeros(18, device=dev, dtype=torch.int64)
    lib.b200rl_attn_set_prof(prof.data_ptr(), which)
    run()
    torch.cuda.synchronize()
    lib.b200rl_attn_set_prof(None, 0)
    v = prof.cpu().numpy().astype(np.float64)
    blocks, ctas = v[16], v[17]
    print(f"== {title}: CTAs {int(ctas)}, inner blocks {int(blocks)} ({blocks / ctas:.2f} per CTA)")
    for i in range(10):
        if names[i] == "-":
            continue
        per = v[i] / (2 * ctas) if i >= 7 else v[i] / (2 * blocks)      # two reporting threads (one per softmax warpgroup)
        print(f"  softmax  {names[i]:26s} {per:9.0f} cycles per {'CTA' if i >= 7 else 'block'}")
    for i in range(10, 16):
        if names[i] == "-":
            continue
        per_cta = i == 15 or (which > 0 and i == 13)
        per = v[i] / ctas if per_cta else v[i] / blocks
        print(f"  issuer   {names[i]:26s} {per:9.0f} cycles per {'CTA' if per_cta else 'block'}")
