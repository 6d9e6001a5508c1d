#!/bin/bash
# final build on 8 GPUs: bench at N=8 and N=2 exactly as the driver launches it
mkdir -p gpurun_out
L=gpurun_out/run33.log
: > $L
for n in 8 2; do
  echo "=== bench N=$n" >> $L
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_run33_n$n.json 2>> $L
  echo "exit $?" >> $L
  wc -l gpurun_out/bench_run33_n$n.json >> $L
  tail -1 gpurun_out/bench_run33_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms', round(d['ms_per_step'],1), 'tok/s', round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],1), d['clocks'])" >> $L 2>&1
done
grep -v "^$" $L | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -12 | cut -c1-300
