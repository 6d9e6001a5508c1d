#!/bin/bash
# ncu --set full of the large GEMMs with the default pass fusion (k = 2, M = 8892 rows per pass)
mkdir -p gpurun_out
L=gpurun_out/run32.log
export B200RL_PROFILE_ONE_STEP=1
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:"gemm_pair_kernel" -s 108 -c 10 -f -o gpurun_out/r1_run32_gemm \
  python bench.py --steps 1 --warmup 1 --no_cpu_baseline > $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r1_run32_gemm.ncu-rep --page raw --csv > gpurun_out/r1_run32_gemm_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r1_run32_gemm_raw.csv gpurun_out/r1_run32_gemm_summary.json > gpurun_out/r1_run32_gemm_summary.txt 2>> $L
rm -f gpurun_out/r1_run32_gemm_raw.csv gpurun_out/r1_run32_gemm.ncu-rep
grep -v "==PROF==" $L | tail -5
