#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run28.log
echo "=== learner tests" > $L
timeout 1200 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> $L 2>&1
echo "exit $?" >> $L
for k in 2 1 4 2; do
  echo "=== bench fuse=$k" >> $L
  timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --fuse_microbatches $k > gpurun_out/bench_run28_k$k.json 2>> $L
  python - "$k" <<'PY' >> $L
import json,sys
try:
    d=json.load(open(f"gpurun_out/bench_run28_k{sys.argv[1]}.json")); print("k", sys.argv[1], "ms", round(d["ms_per_step"],1), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "gemm TF", d["roofline"]["achieved"], d["clocks"]["sm_mhz"]); print(d.get("profile_ms"))
except Exception as e: print("k", sys.argv[1], "failed", e)
PY
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> $L
grep -v "^$" $L | tail -40 | cut -c1-260
