#!/bin/bash
mkdir -p gpurun_out
echo "=== attention tests" > gpurun_out/run13.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 120 -x >> gpurun_out/run13.log 2>&1
rc=$?
echo "exit $rc" >> gpurun_out/run13.log
echo "=== attention timing" >> gpurun_out/run13.log
timeout 300 python scripts/prof_attn.py >> gpurun_out/run13.log 2>&1
if [ $rc -ne 0 ]; then export B200RL_ATTN_TC=0; echo "TC FAILED -> mma.sync" >> gpurun_out/run13.log; fi
echo "=== learner" >> gpurun_out/run13.log
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> gpurun_out/run13.log 2>&1
echo "exit $?" >> gpurun_out/run13.log
echo "=== bench" >> gpurun_out/run13.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run13.json 2>> gpurun_out/run13.log
echo "exit $?" >> gpurun_out/run13.log
grep -v "^$" gpurun_out/run13.log | grep -v "==PROF==" | tail -40 | cut -c1-250
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_run13.json")); print("ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), "gemm TF", d["roofline"]["achieved"]); print(d["profile_ms"])
except Exception as e: print(e)
PY
