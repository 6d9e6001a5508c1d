"""Rasterisation sweep of the K-long CTA-pair GEMMs (run under ncu: scripts/gpu_runs/r2_run13_raster.sh).
For each hot shape whose DRAM traffic is far above the algorithmic bytes (profiles/r2_gemm_dram_traffic.json) launch the
GEMM once per rasterisation group size; ncu reads dram bytes / L2 hit rate / duration per launch, in this order."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distrl_llm_b200 import _capi, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _capi.lib()
SHAPES = [("lm_head dX", 8192, 3584, 152064, True), ("gate|up dX", 8892, 3584, 37888, True), ("down fwd", 8892, 3584, 18944, False)]
GMS = [int(x) for x in os.environ.get("GMS", "0,1,2,4,8,11,16,35").split(",")]
WIDE = [int(x) for x in os.environ.get("WIDE", "2,0").split(",")]
g = torch.Generator(device=dev).manual_seed(0)
order = []
for name, M, N, K, b_mn in SHAPES:
    a = (torch.randn(M, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    b = (torch.randn((K, N) if b_mn else (N, K), device=dev, generator=g) * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for wide in WIDE:
        lib.b200rl_gemm_set_wide(wide)
        for gm in GMS:
            lib.b200rl_gemm_set_raster(gm)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ops.gemm(a, b, out=out, b_mn=b_mn)
            e0.record()
            ops.gemm(a, b, out=out, b_mn=b_mn)
            ops.gemm(a, b, out=out, b_mn=b_mn)
            e1.record()
            torch.cuda.synchronize()
            order.append(f"{name} M={M} N={N} K={K} wide={wide} gm={gm} ms={e0.elapsed_time(e1) / 2:.3f} TF={2.0 * M * N * K / (e0.elapsed_time(e1) / 2) / 1e9:.0f}")
    del a, b, out
lib.b200rl_gemm_set_raster(0)
lib.b200rl_gemm_set_wide(2)
print("\n".join(order))
