#!/bin/bash
mkdir -p gpurun_out
echo "=== attention tc" > gpurun_out/run9.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 120 -k "attention and tcgen05" -x >> gpurun_out/run9.log 2>&1
rc=$?
echo "exit $rc" >> gpurun_out/run9.log
if [ $rc -ne 0 ]; then
  echo "TC ATTENTION FAILED -> mma.sync for the rest" >> gpurun_out/run9.log
  export B200RL_ATTN_TC=0
fi
echo "=== kernels (all)" >> gpurun_out/run9.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 200 -k "not (attention and tcgen05)" >> gpurun_out/run9.log 2>&1
echo "exit $?" >> gpurun_out/run9.log
echo "=== learner" >> gpurun_out/run9.log
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> gpurun_out/run9.log 2>&1
echo "exit $?" >> gpurun_out/run9.log
echo "=== bench" >> gpurun_out/run9.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run9.json 2>> gpurun_out/run9.log
echo "exit $?" >> gpurun_out/run9.log
grep -v "^$" gpurun_out/run9.log | tail -60 | cut -c1-250
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_run9.json")); print("ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), "gemm TF", d["roofline"]["achieved"]); print(d["profile_ms"])
except Exception as e: print(e)
PY
