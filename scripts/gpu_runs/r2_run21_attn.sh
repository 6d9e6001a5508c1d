#!/bin/bash
# round 2, run 21: dQ kernel: S / dP stages released when read (dS in its own TMEM columns, Q in TMEM, dO in smem): tests, phase counters, bench
mkdir -p gpurun_out
L=gpurun_out/r2_run21.log
: > $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "attn or attention" >> $L 2>&1
echo "exit $?" >> $L
timeout 100 python -u scripts/prof_attn_phases.py > gpurun_out/r2_run21_attn_phases.txt 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 --no_cpu_baseline > gpurun_out/r2_run21_bench.json 2>> $L
echo "exit $?" >> $L
tail -8 $L
sed -n 1,2p gpurun_out/r2_run21_attn_phases.txt; sed -n 20,52p gpurun_out/r2_run21_attn_phases.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_run21_bench.json').read().strip().splitlines()[-1])
print('bench ms', d['ms_per_step'], 'tok/s', d['value'], 'e2e', d['e2e']['value'], d['profile_ms'], d['clocks'])
PY
