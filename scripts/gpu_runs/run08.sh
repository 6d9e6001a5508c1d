#!/bin/bash
# 2-GPU: P2P reduce+Adam check, then bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
echo "=== p2p check" > gpurun_out/run8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/p2p_check.py >> gpurun_out/run8.log 2>&1
echo "exit $?" >> gpurun_out/run8.log
echo "=== bench N=2" >> gpurun_out/run8.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_run8_n2.json 2>> gpurun_out/run8.log
echo "exit $?" >> gpurun_out/run8.log
cat gpurun_out/bench_run8_n2.json >> gpurun_out/run8.log
tail -40 gpurun_out/run8.log | cut -c1-700
