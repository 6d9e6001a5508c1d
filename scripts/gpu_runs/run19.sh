#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run19.log
echo "=== gemm tests" > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 120 -k gemm >> $L 2>&1
rc=$?
echo "exit $rc" >> $L
if [ $rc -ne 0 ]; then grep -v "^$" $L | tail -60 | cut -c1-250; exit 1; fi
echo "=== gemm bench (tail split on)" >> $L
timeout 600 python scripts/bench_gemm.py 2>&1 | grep "'bn': 0" >> $L
echo "=== learner" >> $L
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> $L 2>&1
echo "exit $?" >> $L
echo "=== bench" >> $L
timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run19.json 2>> $L
echo "exit $?" >> $L
B200RL_GEMM_TAIL_SPLIT=0 timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run19_nosplit.json 2>> $L
grep -v "^$" $L | tail -40 | cut -c1-200
python - <<'PY'
import json
for f in ("gpurun_out/bench_run19.json","gpurun_out/bench_run19_nosplit.json"):
    try:
        d=json.load(open(f)); print(f, "ms", round(d["ms_per_step"],1), "tok/s", round(d["value"]), "e2e", d["e2e"]["ms_per_step"], "gemm TF", d["roofline"]["achieved"]); print(d.get("profile_ms"))
    except Exception as e: print(f, e)
PY
