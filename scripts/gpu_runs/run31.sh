#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run31.log
echo "=== gemm tests" > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 120 -k gemm >> $L 2>&1
echo "exit $?" >> $L
for cfg in "" "B200RL_GEMM_GM=8" ""; do
  tag=$(echo "$cfg" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  echo "=== bench [$cfg]" >> $L
  env $cfg timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run31_$tag.json 2>> $L
  python - "$tag" <<'PY' >> $L
import json,sys
d=json.load(open(f"gpurun_out/bench_run31_{sys.argv[1]}.json")); print(sys.argv[1], "ms", round(d["ms_per_step"],1), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "gemm TF", d["roofline"]["achieved"], d["clocks"]["sm_mhz"]); print(d.get("profile_ms"))
PY
done
grep -v "^$" $L | tail -12 | cut -c1-260
