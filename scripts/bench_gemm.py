"""GEMM micro-benchmark on the Qwen2.5-7B shapes of BASELINE config 2 (M = 8 x 862 tokens).
CUDA-event timing, 3 warm-ups, inputs larger than L2 are rotated between iterations."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distrl_llm_b200 import ops

dev = torch.device("cuda:0")
M = 8 * 862
shapes = [("qkv", M, 4608, 3584), ("o", M, 3584, 3584), ("gate_up", M, 37888, 3584), ("down", M, 3584, 18944),
          ("lm_head", 4096, 152064, 3584), ("dX_gu", M, 3584, 37888)]
res = []
for name, m, n, k in shapes:
    a = [(torch.randn(m, k, device=dev) * 0.1).to(torch.bfloat16) for _ in range(2)]
    b = [(torch.randn(n, k, device=dev) * 0.1).to(torch.bfloat16) for _ in range(2)]
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    for bn in (0, 256):
        for _ in range(3):
            ops.gemm(a[0], b[0], out=out, force_bn=bn)
        torch.cuda.synchronize()
        iters = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            ops.gemm(a[i & 1], b[i & 1], out=out, force_bn=bn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = 2.0 * m * n * k / ms / 1e9
        # cuBLAS reference point (library, not product)
        res.append(dict(shape=name, M=m, N=n, K=k, bn=bn, ms=round(ms, 4), tflops=round(tf, 1)))
        print(res[-1], flush=True)
    for _ in range(3):
        torch.matmul(a[0], b[0].T, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        torch.matmul(a[i & 1], b[i & 1].T, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(dict(shape=name, impl="cublas", ms=round(ms, 4), tflops=round(2.0 * m * n * k / ms / 1e9, 1)), flush=True)
    del a, b, out
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)
