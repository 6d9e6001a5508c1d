#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run21.log
echo "=== gemm tests" > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 120 -k "swiglu" >> $L 2>&1
echo "exit $?" >> $L
B200RL_GEMM_FUSE_EW=4 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 120 -k "swiglu" >> $L 2>&1
echo "exit $?" >> $L
timeout 600 python scripts/bench_gemm.py 2>&1 | grep "FUSE" >> $L
B200RL_GEMM_FUSE_EW=4 timeout 600 python scripts/bench_gemm.py 2>&1 | grep "FUSE" >> $L
echo "=== bench EW8" >> $L
timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run21.json 2>> $L
echo "=== bench EW4" >> $L
B200RL_GEMM_FUSE_EW=4 timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run21_ew4.json 2>> $L
grep -v "^$" $L | tail -40 | cut -c1-200
python - <<'PY'
import json
for f in ("gpurun_out/bench_run21.json","gpurun_out/bench_run21_ew4.json"):
    try:
        d=json.load(open(f)); print(f, "ms", round(d["ms_per_step"],1), "tok/s", round(d["value"]), "e2e", d["e2e"]["ms_per_step"], "gemm TF", d["roofline"]["achieved"], d["clocks"]); print(d.get("profile_ms"))
    except Exception as e: print(f, e)
PY
