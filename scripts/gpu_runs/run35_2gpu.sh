#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run35.log
echo "=== bench N=2 (stdout must be ONE line)" > $L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_run35_n2.json 2>> $L
echo "exit $?" >> $L
wc -l gpurun_out/bench_run35_n2.json >> $L
head -c 150 gpurun_out/bench_run35_n2.json >> $L; echo >> $L
echo "=== reference arm under torchrun N=2" >> $L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_run35_ref_n2.json 2>> $L
echo "exit $?" >> $L
wc -l gpurun_out/bench_run35_ref_n2.json >> $L
head -c 200 gpurun_out/bench_run35_ref_n2.json >> $L; echo >> $L
grep -v "OMP_NUM\|\*\*\*\*\|^$" $L | tail -14 | cut -c1-260
