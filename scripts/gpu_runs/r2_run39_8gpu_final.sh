#!/bin/bash
# round 2, run 39 (8 GPUs): the final build at N=8 launched like the driver does (cfg2, weak scaling), exchange breakdown
mkdir -p gpurun_out
L=gpurun_out/r2_run39.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_run39_bench_cfg2_n8.json 2>> $L
echo "exit $?" >> $L
python - >> $L 2>&1 <<'PY'
import json
d = json.loads(open("gpurun_out/r2_run39_bench_cfg2_n8.json").read().strip().splitlines()[-1])
print("cfg2_n8 ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "clocks", d["clocks"])
print(json.dumps(d["exchange"]))
PY
grep -v "^$" $L | grep -v "^\*\*\*\|OMP_NUM\|NCCL version" | tail -8
