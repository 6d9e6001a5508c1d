"""GEMM micro-benchmark on the Qwen2.5-7B shapes of BASELINE config 2, packed layout (M = 350 + 8 x 512 rows) and
classic layout (M = 8 x 862), with the rank-64 LoRA K-extension, for every pair-tile width.
CUDA-event timing, 3 warm-ups, two operand sets rotated between iterations."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distrl_llm_b200 import ops

dev = torch.device("cuda:0")
res = []
H, I, QKV, QD = 3584, 18944, 4608, 3584
for M in (350 + 8 * 512, 8 * 862):
    shapes = [("qkv", M, QKV, H, False), ("o", M, H, QD, False), ("gate_up", M, 2 * I, H, False), ("down", M, H, I, False),
              ("dX_down", M, I, H, True), ("dX_gu", M, H, 2 * I, True), ("dX_o", M, QD, H, True), ("dX_qkv", M, H, QKV, True)]
    for name, m, n, k, b_mn in shapes:
        a = [(torch.randn(m, k, device=dev) * 0.1).to(torch.bfloat16) for _ in range(2)]
        b = [(torch.randn((k, n) if b_mn else (n, k), device=dev) * 0.1).to(torch.bfloat16) for _ in range(2)]
        a2 = (torch.randn(m, 64, device=dev) * 0.1).to(torch.bfloat16)
        b2 = (torch.randn((64, n) if b_mn else (n, 64), device=dev) * 0.1).to(torch.bfloat16)
        out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        for bn in (0, 256, 224, 192):
            for _ in range(3):
                ops.gemm(a[0], b[0], a2, b2, out=out, force_bn=bn, b_mn=b_mn)
            torch.cuda.synchronize()
            iters = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                ops.gemm(a[i & 1], b[i & 1], a2, b2, out=out, force_bn=bn, b_mn=b_mn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            tf = 2.0 * m * n * (k + 64) / ms / 1e9
            res.append(dict(shape=name, M=m, N=n, K=k, bn=bn, us=round(ms * 1e3, 1), tflops=round(tf, 1)))
            print(res[-1], flush=True)
        del a, b, out
# fused SwiGLU epilogues vs GEMM + row kernel
from distrl_llm_b200 import _capi
lib, st = _capi.lib(), _capi.stream()
M = 350 + 8 * 512
h = (torch.randn(M, H, device=dev) * 0.1).to(torch.bfloat16)
u = (torch.randn(M, 64, device=dev) * 0.1).to(torch.bfloat16)
Wgu = (torch.randn(2 * I, H, device=dev) * 0.05).to(torch.bfloat16)
Bgu = (torch.randn(2 * I, 64, device=dev) * 0.05).to(torch.bfloat16)
Wd = (torch.randn(H, I, device=dev) * 0.05).to(torch.bfloat16)
Ad = (torch.randn(64, I, device=dev) * 0.05).to(torch.bfloat16)
gu = torch.empty(M, 2 * I, device=dev, dtype=torch.bfloat16)
act = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
dact = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
dgu = torch.empty(M, 2 * I, device=dev, dtype=torch.bfloat16)
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
def fwd_unfused():
    ops.gemm(h, Wgu, u, Bgu, out=gu)
    lib.b200rl_swiglu_fwd(gu.data_ptr(), act.data_ptr(), M, I, st)
def fwd_fused():
    lib.b200rl_gemm_swiglu(1, h.data_ptr(), H, Wgu.data_ptr(), H, H, u.data_ptr(), 64, Bgu.data_ptr(), 64, 64, gu.data_ptr(), 2 * I, act.data_ptr(), I, M, I, st)
def bwd_unfused():
    ops.gemm(h, Wd, u, Ad, out=dact, b_mn=True)
    lib.b200rl_swiglu_bwd(gu.data_ptr(), dact.data_ptr(), dgu.data_ptr(), M, I, st)
def bwd_fused():
    lib.b200rl_gemm_swiglu(2, h.data_ptr(), H, Wd.data_ptr(), I, H, u.data_ptr(), 64, Ad.data_ptr(), I, 64, dgu.data_ptr(), 2 * I, gu.data_ptr(), 2 * I, M, I, st)
for name, fn in (("gate_up gemm only", lambda: ops.gemm(h, Wgu, u, Bgu, out=gu)), ("gate_up gemm+swiglu", fwd_unfused), ("gate_up fused", fwd_fused),
                 ("dX_down gemm only", lambda: ops.gemm(h, Wd, u, Ad, out=dact, b_mn=True)), ("dX_down gemm+swiglu_bwd", bwd_unfused), ("dX_down fused", bwd_fused)):
    print("FUSE", os.environ.get("B200RL_GEMM_FUSE_EW", "8"), name, round(t(fn), 1), "us", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)
