#!/bin/bash
# round 2, run 2: wide (256x512) pair tiles + in-kernel LoRA intermediates: kernel tests, model tests, A/B bench on one box
mkdir -p gpurun_out
L=gpurun_out/r2_run02.log
: > $L
echo "== gemm kernel tests" >> $L
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" >> $L 2>&1
echo "exit $?" >> $L
echo "== model-level tests" >> $L
timeout 1500 python -m pytest tests/test_gpu_learner.py tests/test_gpu_cfg2_shapes.py tests/test_gpu_trainer.py tests/test_gpu_p2p.py -q -m gpu >> $L 2>&1
echo "exit $?" >> $L
for cfg in "default" "B200RL_GEMM_WIDE=0" "B200RL_GEMM_EXT=0" "B200RL_GEMM_WIDE=0 B200RL_GEMM_EXT=0"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  echo "== bench $cfg" >> $L
  if [ "$cfg" = "default" ]; then
    timeout 300 python bench.py --steps 4 --warmup 3 --no_cpu_baseline > gpurun_out/r2_run02_bench_$tag.json 2>> $L
  else
    env $cfg timeout 300 python bench.py --steps 4 --warmup 3 --no_cpu_baseline > gpurun_out/r2_run02_bench_$tag.json 2>> $L
  fi
  echo "exit $?" >> $L
  python - "$tag" >> $L 2>&1 <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_run02_bench_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "ms", round(d["ms_per_step"], 1), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "gemm TF", d["roofline"]["achieved"],
          "frac", d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"], d["profile_ms"], d["profile_launches"])
except Exception as e:
    print(tag, "no result:", e)
PY
done
grep -n "passed\|failed\|rror\|exit\|ms " $L | tail -40
