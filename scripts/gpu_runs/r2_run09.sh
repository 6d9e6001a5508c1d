#!/bin/bash
# attention tweaks: correctness (attention + learner tests), forward phase counters, bench A/B is across calls (compare with r2_run05)
mkdir -p gpurun_out
L=gpurun_out/r2_run09.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" >> $L 2>&1
echo "exit $?" >> $L
timeout 900 python -m pytest tests/test_gpu_learner.py -q -m gpu -x >> $L 2>&1
echo "exit $?" >> $L
python scripts/prof_attn_phases.py > gpurun_out/r2_run09_attn_phases.txt 2>&1
cat gpurun_out/r2_run09_attn_phases.txt >> $L
timeout 300 python bench.py --steps 4 --warmup 3 --no_cpu_baseline > gpurun_out/r2_run09_bench.json 2>> $L
python - >> $L 2>&1 <<'PY'
import json
d = json.loads(open("gpurun_out/r2_run09_bench.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "gemm TF", d["roofline"]["achieved"], "clk", d["clocks"]["sm_mhz"], d["profile_ms"])
PY
grep -v "^$" $L | tail -32
