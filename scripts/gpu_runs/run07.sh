#!/bin/bash
# ncu evidence: launch list of one learner step + full capture of the top kernels
mkdir -p gpurun_out
export B200RL_PROFILE_ONE_STEP=1
echo "=== ncu launch list" > gpurun_out/run7.log
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline >> gpurun_out/run7.log 2>&1
echo "exit $?" >> gpurun_out/run7.log
echo "=== ncu full (gemm pair + attention bwd)" >> gpurun_out/run7.log
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_pair_kernel|attn_bwd_dkv|attn_bwd_dq|attn_fwd" -c 8 -o gpurun_out/prof_r1_top python bench.py --steps 1 --warmup 3 --no_cpu_baseline >> gpurun_out/run7.log 2>&1
echo "exit $?" >> gpurun_out/run7.log
ls -la gpurun_out/ >> gpurun_out/run7.log
tail -15 gpurun_out/run7.log
