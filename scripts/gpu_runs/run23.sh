#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run23.log
echo "=== gpu tests" > $L
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -x >> $L 2>&1
echo "exit $?" >> $L
for cfg in "" "B200RL_GROUPED_DW=0" ""; do
  tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  echo "=== bench [$cfg]" >> $L
  env $cfg timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run23_$tag.json 2>> $L
  python - "$tag" <<'PY' >> $L
import json,sys
d=json.load(open(f"gpurun_out/bench_run23_{sys.argv[1]}.json")); print(sys.argv[1], "ms", round(d["ms_per_step"],1), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "gemm TF", d["roofline"]["achieved"], d["clocks"]["sm_mhz"]); print(d.get("profile_ms")); print(d.get("profile_launches"))
PY
done
grep -v "^$" $L | tail -40 | cut -c1-260
