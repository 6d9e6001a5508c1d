"""Attention at the BASELINE config-2 shape (B=8, L=862, 28/4 heads, hd 128): timing + a target for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distrl_llm_b200 import ops, _capi
dev = torch.device("cuda:0")
B, L, nq, nkv, hd = 8, 862, 28, 4, 128
qkv = (torch.randn(B * L, (nq + 2 * nkv) * hd, device=dev) * 0.5).to(torch.bfloat16)
mask = torch.ones(B, L, dtype=torch.int32, device=dev)
dout = (torch.randn(B * L, nq * hd, device=dev) * 0.1).to(torch.bfloat16)
for mode in (1, 0):
    _capi.lib().b200rl_attn_set_tc(mode)
    for _ in range(3):
        out, lse = ops.attn_fwd(qkv, mask, B, L, nq, nkv, hd)
        dq = ops.attn_bwd(qkv, mask, out, dout, lse, B, L, nq, nkv, hd)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(10):
        out, lse = ops.attn_fwd(qkv, mask, B, L, nq, nkv, hd)
    e[1].record()
    for _ in range(10):
        dq = ops.attn_bwd(qkv, mask, out, dout, lse, B, L, nq, nkv, hd)
    e[2].record()
    torch.cuda.synchronize()
    f = 2.0 * B * nq * L * L * hd
    tf, tb = e[0].elapsed_time(e[1]) / 10, e[1].elapsed_time(e[2]) / 10
    print(f"mode {'tcgen05' if mode else 'mma.sync'}: fwd {tf*1e3:.0f} us ({f/tf/1e9:.0f} TFLOP/s causal-algorithmic), bwd {tb*1e3:.0f} us ({2*f/tb/1e9:.0f} TFLOP/s)", flush=True)
