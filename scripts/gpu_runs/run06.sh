#!/bin/bash
mkdir -p gpurun_out
echo "=== kernels" > gpurun_out/run6.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 200 >> gpurun_out/run6.log 2>&1
echo "exit $?" >> gpurun_out/run6.log
echo "=== learner" >> gpurun_out/run6.log
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> gpurun_out/run6.log 2>&1
echo "exit $?" >> gpurun_out/run6.log
echo "=== bench pair" >> gpurun_out/run6.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run6_pair.json 2>> gpurun_out/run6.log
echo "exit $?" >> gpurun_out/run6.log
echo "=== bench single" >> gpurun_out/run6.log
B200RL_GEMM_CTA_PAIR=0 timeout 1500 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run6_single.json 2>> gpurun_out/run6.log
echo "exit $?" >> gpurun_out/run6.log
tail -30 gpurun_out/run6.log | cut -c1-300
python - <<'PY'
import json
for f in ["gpurun_out/bench_run6_pair.json","gpurun_out/bench_run6_single.json"]:
    try:
        d=json.load(open(f)); print(f, "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), "gemm TF", d["roofline"]["achieved"], d["clocks"]); print(d["profile_ms"])
    except Exception as e: print(f, e)
PY
