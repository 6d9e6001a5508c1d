#!/bin/bash
# round 2, run 5: wide tiles with the early-release epilogue: kernel tests + A/B of the wide modes on one box
mkdir -p gpurun_out
L=gpurun_out/r2_run05.log
: > $L
echo "== gemm kernel tests" >> $L
timeout 1200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm" >> $L 2>&1
echo "exit $?" >> $L
echo "== model tests (trainer / p2p / cfg2 shapes)" >> $L
timeout 1500 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_p2p.py tests/test_gpu_cfg2_shapes.py -q -m gpu >> $L 2>&1
echo "exit $?" >> $L
for mode in 0 2 1 0; do
  echo "== bench B200RL_GEMM_WIDE=$mode" >> $L
  B200RL_GEMM_WIDE=$mode timeout 300 python bench.py --steps 4 --warmup 3 --no_cpu_baseline > gpurun_out/r2_run05_bench_wide$mode.json 2>> $L
  echo "exit $?" >> $L
  python - "$mode" >> $L 2>&1 <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_run05_bench_wide{tag}.json").read().strip().splitlines()[-1])
    print("wide", tag, "ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "gemm TF", d["roofline"]["achieved"],
          "frac", d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"], d["profile_ms"])
except Exception as e:
    print(tag, "no result:", e)
PY
done
grep -n "passed\|failed\|rror\|exit\|ms " $L | tail -40
