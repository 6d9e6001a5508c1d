#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run36.log
echo "=== all gpu tests (HEAD)" > $L
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 200 -x >> $L 2>&1
echo "exit $?" >> $L
timeout 200 python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run36.json 2>> $L
echo "exit $?" >> $L
python -c "
import json
d=json.loads(open('gpurun_out/bench_run36.json').read().strip().splitlines()[-1]); print('ms', round(d['ms_per_step'],1), 'tok/s', round(d['value']), 'e2e', round(d['e2e']['value']))" >> $L 2>&1
grep -v "^$" $L | tail -8 | cut -c1-200
