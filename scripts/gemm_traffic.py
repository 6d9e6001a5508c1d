"""profiles/r2_gemm_dram_traffic.json from the ncu --set full summary of the 10 large GEMMs of one pass
(scripts/gpu_runs/r2_run07_ncu.sh: `-k regex:gemm_pair_kernel -s 108 -c 10` inside one profiled learner step =
forward layer 27 (qkv, o, gate|up, down), lm_head fwd, lm_head dX, backward layer 27 (down dX, gate|up dX, o dX, qkv dX)).
Per-launch average over the 226 large-GEMM launches of a pass = (28 x the layer's 8 GEMMs + the 2 lm_head GEMMs) / 226.
Algorithmic bytes = every operand and the output touched once (bf16), LoRA operands included."""
import json
import sys

M, R, H, I, V, QKV, K2 = 8892, 8192, 3584, 18944, 152064, 4608, 64
NAMES = ["qkv fwd", "o fwd", "gate|up fwd (fused SwiGLU)", "down fwd", "lm_head fwd", "lm_head dX", "down dX (fused SwiGLU bwd)",
         "gate|up dX", "o dX", "qkv dX"]
b = 2


def alg(m, n, k, extra_out=0, extra_in=0):
    return b * (m * k + n * k + m * n) + b * (m * K2 + n * K2 + K2 * k) + extra_out + extra_in


ALG = [alg(M, QKV, H), alg(M, H, H) + b * M * H, alg(M, 2 * I, H, extra_out=b * M * I), alg(M, H, I) + b * M * H,
       b * (R * H + V * H + R * V), b * (R * V + V * H + R * H),
       alg(M, I, H, extra_out=b * M * I, extra_in=b * M * 2 * I), alg(M, H, 2 * I), alg(M, H, H), alg(M, H, QKV)]

recs = json.load(open(sys.argv[1]))
assert len(recs) >= 10, len(recs)
per = []
for name, a, r in zip(NAMES, ALG, recs[:10]):
    rd = r.get("dram__bytes_read.sum", 0.0)
    wr = r.get("dram__bytes_write.sum", 0.0)
    per.append({"gemm": name, "kernel": r["kernel"][:80], "dram_read_bytes": rd, "dram_write_bytes": wr, "algorithmic_bytes": float(a),
                "duration_ms": r.get("gpu__time_duration.sum", 0.0) / 1e6 if r.get("gpu__time_duration.sum", 0) > 1e3 else r.get("gpu__time_duration.sum", 0.0),
                "tensor_active_pct": r.get("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                "l2_hit_pct": r.get("lts__t_sector_hit_rate.pct")})
layer = [0, 1, 2, 3, 6, 7, 8, 9]
head = [4, 5]
tr = lambda i: per[i]["dram_read_bytes"] + per[i]["dram_write_bytes"]
traffic = (28 * sum(tr(i) for i in layer) + sum(tr(i) for i in head)) / 226
algo = (28 * sum(per[i]["algorithmic_bytes"] for i in layer) + sum(per[i]["algorithmic_bytes"] for i in head)) / 226
out = {"source": sys.argv[1] + " (ncu --set full --clock-control none, one learner step of the default bench.py command: packed rows, "
       "2 reference micro-batches per pass = 8892 rows; GEMMs of the last forward layer, lm_head fwd + dX, first backward layer)",
       "note": "per-launch average over the 226 large-GEMM launches of a pass; algorithmic = each operand and the output touched once",
       "per_launch_traffic_bytes": traffic, "per_launch_algorithmic_bytes": algo, "ratio": traffic / algo, "per_gemm": per}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print("traffic / algorithmic =", round(traffic / algo, 3), " per launch", round(traffic / 1e9, 3), "GB vs", round(algo / 1e9, 3), "GB")
for p in per:
    print(f"{p['gemm']:32s} {(p['dram_read_bytes'] + p['dram_write_bytes']) / 1e9:7.3f} GB  alg {p['algorithmic_bytes'] / 1e9:7.3f}  x{(p['dram_read_bytes'] + p['dram_write_bytes']) / p['algorithmic_bytes']:5.2f}  {p['duration_ms']:.3f} ms  tensor {p['tensor_active_pct']}")
