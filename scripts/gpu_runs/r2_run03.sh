#!/bin/bash
# round 2, run 3: NF4-in-mainloop GEMM + host-rendezvous P2P: tests, then A/B of the three weight modes on one box
mkdir -p gpurun_out
L=gpurun_out/r2_run03.log
: > $L
echo "== gemm kernel tests (nf4 / lora / wide)" >> $L
timeout 1200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "nf4 or lora or wide or swiglu" >> $L 2>&1
echo "exit $?" >> $L
echo "== model-level tests" >> $L
timeout 1500 python -m pytest tests/test_gpu_learner.py tests/test_gpu_trainer.py tests/test_gpu_p2p.py -q -m gpu >> $L 2>&1
echo "exit $?" >> $L
for w in cache inkernel scratch; do
  echo "== bench --weights $w" >> $L
  timeout 400 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --weights $w > gpurun_out/r2_run03_bench_$w.json 2>> $L
  echo "exit $?" >> $L
  python - "$w" >> $L 2>&1 <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_run03_bench_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "gemm TF", d["roofline"]["achieved"],
          "frac", d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"], d["profile_ms"], d["profile_launches"])
except Exception as e:
    print(tag, "no result:", e)
PY
done
grep -n "passed\|failed\|rror\|exit\|ms " $L | tail -40
