#!/bin/bash
# round 2, run 26: evidence of the final build: ncu launch list of one learner step, ncu --set full of the 10 large GEMMs of
# a pass (DRAM traffic -> profiles/r2_gemm_dram_traffic.json), of the three tcgen05 attention kernels and grouped dW / logprob,
# and of reduce_adam_kernel at world 8 (the single-GPU fake world of tests/test_gpu_p2p.py: 8 buffer sets, 8 streams)
mkdir -p gpurun_out
L=gpurun_out/r2_run26.log
: > $L
export B200RL_PROFILE_ONE_STEP=1
echo "=== launch list" >> $L
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run26_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
python scripts/ncu_summarize.py gpurun_out/r2_run26_launches.csv > gpurun_out/r2_run26_launch_summary.txt 2>&1
python scripts/ncu_summarize.py gpurun_out/r2_run26_launches.csv byname > gpurun_out/r2_run26_launch_summary_byname.txt 2>&1
gzip -f gpurun_out/r2_run26_launches.csv
echo "=== ncu full: the 10 large GEMMs of a pass" >> $L
timeout 1200 ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_pair_kernel" -s 108 -c 10 -f -o gpurun_out/r2_run26_gemm \
  python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r2_run26_gemm.ncu-rep --page raw --csv > gpurun_out/r2_run26_gemm_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r2_run26_gemm_raw.csv gpurun_out/r2_run26_gemm_summary.json > gpurun_out/r2_run26_gemm_summary.txt 2>> $L
python scripts/gemm_traffic.py gpurun_out/r2_run26_gemm_summary.json gpurun_out/r2_gemm_dram_traffic.json >> $L 2>&1
echo "=== ncu full: attention + logprob + grouped dW" >> $L
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"attn_.*_tc_kernel|dw_grouped_kernel|logprob_kernel" -s 30 -c 8 -f -o gpurun_out/r2_run26_other \
  python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r2_run26_other.ncu-rep --page raw --csv > gpurun_out/r2_run26_other_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r2_run26_other_raw.csv gpurun_out/r2_run26_other_summary.json > gpurun_out/r2_run26_other_summary.txt 2>> $L
unset B200RL_PROFILE_ONE_STEP
echo "=== ncu full: reduce_adam_kernel<8> + barrier, fake world 8 on one GPU" >> $L
timeout 600 ncu --set full --clock-control none -k regex:"reduce_adam_kernel|p2p_barrier_kernel" -c 24 -f -o gpurun_out/r2_run26_reduce \
  python -m pytest tests/test_gpu_p2p.py -q -m gpu -k "fake_world and 8" >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r2_run26_reduce.ncu-rep --page raw --csv > gpurun_out/r2_run26_reduce_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r2_run26_reduce_raw.csv gpurun_out/r2_run26_reduce_summary.json > gpurun_out/r2_run26_reduce_summary.txt 2>> $L
rm -f gpurun_out/r2_run26_gemm_raw.csv gpurun_out/r2_run26_other_raw.csv gpurun_out/r2_run26_reduce_raw.csv gpurun_out/r2_run26_gemm.ncu-rep
grep -v "==PROF==" $L | grep -v "^$" | tail -30
head -28 gpurun_out/r2_run26_launch_summary_byname.txt
head -20 gpurun_out/r2_run26_reduce_summary.txt
