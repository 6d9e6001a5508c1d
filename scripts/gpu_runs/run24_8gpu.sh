#!/bin/bash
# 8-GPU validation: P2P reduce+Adam check at N=8, bench at N=8 and N=4
mkdir -p gpurun_out
L=gpurun_out/run24.log
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
echo "=== p2p check N=8" > $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/p2p_check.py >> $L 2>&1
echo "exit $?" >> $L
for n in 8 4; do
  echo "=== bench N=$n" >> $L
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_run24_n$n.json 2>> $L
  echo "exit $?" >> $L
  tail -1 gpurun_out/bench_run24_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms', round(d['ms_per_step'],1), 'tok/s', round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],1), d['clocks'])" >> $L 2>&1
done
grep -v "^$" $L | grep -v "Warning\|warn" | tail -30 | cut -c1-300
