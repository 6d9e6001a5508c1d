#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run37.log
: > $L
timeout 100 python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run37.json 2>> $L
echo "default exit $?" >> $L
python -c "
import json
d=json.loads(open('gpurun_out/bench_run37.json').read().strip().splitlines()[-1]); print('default: ms', round(d['ms_per_step'],1), 'tok/s', round(d['value']), d['config']['passes'][:40])" >> $L 2>&1
timeout 110 python bench.py --steps 1 --warmup 3 --no_cpu_baseline --new_tokens 2048 --seqs 16 > gpurun_out/bench_run37_t2048.json 2>> $L
echo "T=2048 exit $?" >> $L
python -c "
import json
d=json.loads(open('gpurun_out/bench_run37_t2048.json').read().strip().splitlines()[-1]); print('T=2048: ms', round(d['ms_per_step'],1), 'tok/s', round(d['value']), d['config']['passes'][:40])" >> $L 2>&1
tail -8 $L | cut -c1-220
