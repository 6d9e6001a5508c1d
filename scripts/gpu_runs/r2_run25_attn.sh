#!/bin/bash
# round 2, run 25: dQ kernel with Q / dO back in shared memory (TMA), dS in TMEM; forward as in run 23:
# kernel tests, learner / cfg2-shape parity tests, phase counters, bench
mkdir -p gpurun_out
L=gpurun_out/r2_run25.log
: > $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "attn or attention" >> $L 2>&1
echo "exit $?" >> $L
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_cfg2_shapes.py -q -x -m gpu >> $L 2>&1
echo "exit $?" >> $L
timeout 100 python -u scripts/prof_attn_phases.py > gpurun_out/r2_run25_attn_phases.txt 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 --no_cpu_baseline > gpurun_out/r2_run25_bench.json 2>> $L
echo "exit $?" >> $L
tail -12 $L
sed -n 1,20p gpurun_out/r2_run25_attn_phases.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_run25_bench.json').read().strip().splitlines()[-1])
print('bench ms', d['ms_per_step'], 'tok/s', d['value'], 'e2e', d['e2e']['value'], d['profile_ms'], d['clocks'])
PY
