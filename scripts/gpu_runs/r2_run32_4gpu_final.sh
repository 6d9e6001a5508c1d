#!/bin/bash
# round 2, run 32 (4 GPUs): the final build on BASELINE config 3 (256 x 1024 over 4 learners = 64 x 1024 each), config 4
# (1024 x 2048 sampled, top-k 128 of 256 -> 512 sequences over 4 learners) and cfg2 weak scaling at N = 4
mkdir -p gpurun_out
L=gpurun_out/r2_run32.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
for cfg in cfg2 cfg3 cfg4; do
  echo "== $cfg N=4" >> $L
  ST=3; [ $cfg = cfg4 ] && ST=1; [ $cfg = cfg3 ] && ST=2
  timeout 1500 $TR bench.py --gpus 4 --config $cfg --steps $ST --warmup 3 --lean > gpurun_out/r2_run32_bench_${cfg}_n4.json 2>> $L
  echo "exit $?" >> $L
  python - $cfg >> $L 2>&1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_run32_bench_{sys.argv[1]}_n4.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "n4: ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "frac_packed", d["step_roofline"]["frac_of_packed_roofline"])
    print(json.dumps(d["exchange"]))
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
done
grep -v "^$" $L | grep -v "^\*\*\*\|OMP_NUM\|NCCL version" | tail -14
