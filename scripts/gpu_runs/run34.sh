#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run34.log
echo "=== packed attention + learner tests (ragged rows default)" > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 -k "packed or learner or shared or fused or hf_state or cfg1 or kl_term or train_step" >> $L 2>&1
echo "exit $?" >> $L
for cfg in "" "--ragged"; do
  tag=$(echo "$cfg" | tr -d ' -'); [ -z "$tag" ] && tag=default
  echo "=== bench [$cfg]" >> $L
  timeout 600 python bench.py --steps 3 --warmup 3 --no_cpu_baseline $cfg > gpurun_out/bench_run34_$tag.json 2>> $L
  python - "$tag" <<'PY' >> $L
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench_run34_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print(sys.argv[1], "ms", round(d["ms_per_step"],1), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "e2e tok/s", round(d["e2e"]["value"]), "gemm TF", d["roofline"]["achieved"], d["clocks"]["sm_mhz"]); print(d.get("profile_ms"))
except Exception as e: print(sys.argv[1], "failed", e)
PY
done
grep -v "^$" $L | tail -16 | cut -c1-260
