#!/bin/bash
# ncu --set full of the last forward layer, lm_head and the first backward layer of one packed learner step
mkdir -p gpurun_out
L=gpurun_out/run25.log
export B200RL_PROFILE_ONE_STEP=1
echo "=== ncu full" > $L
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"gemm_pair_kernel|attn_.*_tc_kernel|dw_grouped" -s 135 -c 14 -f -o gpurun_out/r1_run25_full \
  python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r1_run25_full.ncu-rep --page raw --csv > gpurun_out/r1_run25_full_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r1_run25_full_raw.csv gpurun_out/r1_run25_full_summary.json > gpurun_out/r1_run25_full_summary.txt 2>> $L
ls -la gpurun_out | tail -8 >> $L
grep -v "==PROF==" $L | tail -12
grep -A3 "^###" gpurun_out/r1_run25_full_summary.txt | head -80
