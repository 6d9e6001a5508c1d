#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run16.log
echo "=== kernel tests" > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 120 >> $L 2>&1
echo "exit $?" >> $L
echo "=== learner" >> $L
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> $L 2>&1
echo "exit $?" >> $L
echo "=== bench packed" >> $L
timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run16.json 2>> $L
echo "exit $?" >> $L
echo "=== bench classic" >> $L
timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_share_prompts > gpurun_out/bench_run16_classic.json 2>> $L
echo "exit $?" >> $L
grep -v "^$" $L | grep -v "==PROF==" | tail -60 | cut -c1-300
python - <<'PY'
import json
for f in ("gpurun_out/bench_run16.json", "gpurun_out/bench_run16_classic.json"):
    try:
        d=json.load(open(f)); print(f, "ms", round(d["ms_per_step"],1), "tok/s", round(d["value"]), "e2e", d["e2e"], "gemm TF", d["roofline"]["achieved"]); print(d.get("profile_ms"))
    except Exception as e: print(f, e)
PY
