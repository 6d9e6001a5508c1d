#!/bin/bash
# round 2, run 42: single-lane roles entered through elect.sync (MMA issuers / TMA producers keep their operands in uniform
# registers: 21 -> ~2 SASS instructions per tcgen05.mma in the GEMM issuer) vs the previous build, A/B on ONE box
mkdir -p gpurun_out
L=gpurun_out/r2_run42.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu >> $L 2>&1
echo "kernel tests exit $?" >> $L
timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_cfg2_shapes.py tests/test_gpu_p2p.py tests/test_gpu_trainer.py -q -x -m gpu >> $L 2>&1
echo "model tests exit $?" >> $L
run() { timeout 300 python bench.py --steps 4 --warmup 3 --no_cpu_baseline --lean 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],1), round(d['value']), d['profile_ms'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"; }
cp distrl_llm_b200/lib/libb200rl.so /tmp/new.so
run new >> $L
cp distrl_llm_b200/lib/libb200rl_prev.so distrl_llm_b200/lib/libb200rl.so
run prev >> $L
cp /tmp/new.so distrl_llm_b200/lib/libb200rl.so
run new >> $L
cp distrl_llm_b200/lib/libb200rl_prev.so distrl_llm_b200/lib/libb200rl.so
run prev >> $L
cp /tmp/new.so distrl_llm_b200/lib/libb200rl.so
python scripts/prof_attn_phases.py 2>&1 | head -2 >> $L
grep -v "^$" $L | tail -14
