#!/bin/bash
# round 2, run 36: the driver's own commands on a fresh box: both arms with --steps 20 --warmup 5 (wall time of each)
mkdir -p gpurun_out
L=gpurun_out/r2_run36.log
: > $L
t0=$(date +%s)
timeout 870 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_run36_bench_reference_20_5.json 2>> $L
echo "reference arm exit $? wall $(( $(date +%s) - t0 )) s" >> $L
t0=$(date +%s)
timeout 870 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_run36_bench_n1_20_5.json 2>> $L
echo "b200 arm exit $? wall $(( $(date +%s) - t0 )) s" >> $L
cat $L | tail -5
python - <<'PY'
import json
for f in ("r2_run36_bench_reference_20_5", "r2_run36_bench_n1_20_5"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 1), d["unit"], round(d["value"], 1), "e2e", d.get("e2e", {}).get("value"), "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
