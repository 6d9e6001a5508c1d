#!/bin/bash
# round 2, run 7: evidence of the final build: ncu launch list of one learner step, ncu --set full of the 10 large GEMMs of
# a pass (DRAM traffic -> profiles/r2_gemm_dram_traffic.json) and of the attention / reduce kernels; torch-GPU comparator
mkdir -p gpurun_out
L=gpurun_out/r2_run07.log
: > $L
export B200RL_PROFILE_ONE_STEP=1
echo "=== launch list" >> $L
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run07_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
python scripts/ncu_summarize.py gpurun_out/r2_run07_launches.csv > gpurun_out/r2_run07_launch_summary.txt 2>&1
python scripts/ncu_summarize.py gpurun_out/r2_run07_launches.csv byname > gpurun_out/r2_run07_launch_summary_byname.txt 2>&1
gzip -f gpurun_out/r2_run07_launches.csv
echo "=== ncu full: the 10 large GEMMs of a pass" >> $L
timeout 1200 ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_pair_kernel" -s 108 -c 10 -f -o gpurun_out/r2_run07_gemm \
  python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r2_run07_gemm.ncu-rep --page raw --csv > gpurun_out/r2_run07_gemm_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r2_run07_gemm_raw.csv gpurun_out/r2_run07_gemm_summary.json > gpurun_out/r2_run07_gemm_summary.txt 2>> $L
python scripts/gemm_traffic.py gpurun_out/r2_run07_gemm_summary.json gpurun_out/r2_gemm_dram_traffic.json >> $L 2>&1
echo "=== ncu full: attention + reduce_adam + logprob + grouped dW" >> $L
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"attn_.*_tc_kernel|dw_grouped_kernel|logprob_kernel|reduce_adam" -s 30 -c 8 -f -o gpurun_out/r2_run07_other \
  python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r2_run07_other.ncu-rep --page raw --csv > gpurun_out/r2_run07_other_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r2_run07_other_raw.csv gpurun_out/r2_run07_other_summary.json > gpurun_out/r2_run07_other_summary.txt 2>> $L
rm -f gpurun_out/r2_run07_gemm_raw.csv gpurun_out/r2_run07_other_raw.csv gpurun_out/r2_run07_other.ncu-rep
unset B200RL_PROFILE_ONE_STEP
echo "=== torch-GPU comparator (oracle port, dense bf16 autocast, reference layout, 28 layers)" >> $L
timeout 900 python bench.py --impl torch_gpu --steps 2 --warmup 1 > gpurun_out/r2_run07_bench_torch_gpu.json 2>> $L
echo "exit $?" >> $L
tail -c 900 gpurun_out/r2_run07_bench_torch_gpu.json >> $L
grep -v "==PROF==" $L | grep -v "^$" | tail -30
head -25 gpurun_out/r2_run07_launch_summary_byname.txt
