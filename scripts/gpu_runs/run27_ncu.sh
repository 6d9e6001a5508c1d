#!/bin/bash
# final-build evidence: (1) launch list of one packed learner step, (2) ncu --set full of one instance of every
# remaining hot kernel (row kernels, skinny GEMM, reductions, log-prob) + the big GEMMs with the new rasterisation
mkdir -p gpurun_out
L=gpurun_out/run27.log
export B200RL_PROFILE_ONE_STEP=1
echo "=== gpu tests (final build)" > $L
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 >> $L 2>&1
echo "exit $?" >> $L
echo "=== launch list" >> $L
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_run27_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
python scripts/ncu_summarize.py gpurun_out/r1_run27_launches.csv > gpurun_out/r1_run27_launch_summary.txt 2>&1
python scripts/ncu_summarize.py gpurun_out/r1_run27_launches.csv byname > gpurun_out/r1_run27_launch_summary_byname.txt 2>&1
gzip -f gpurun_out/r1_run27_launches.csv
echo "=== ncu full: small kernels" >> $L
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:"rmsnorm|rope_kernel|logprob_kernel|grad_accum|reduce_slabs|gemm_kernel|kv_reduce|attn_delta|embed|gather_rows|scatter_add|loss_" -s 1500 -c 40 -f -o gpurun_out/r1_run27_small \
  python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r1_run27_small.ncu-rep --page raw --csv > gpurun_out/r1_run27_small_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r1_run27_small_raw.csv gpurun_out/r1_run27_small_summary.json > gpurun_out/r1_run27_small_summary.txt 2>> $L
echo "=== ncu full: big GEMMs, new rasterisation" >> $L
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:"gemm_pair_kernel" -s 108 -c 10 -f -o gpurun_out/r1_run27_gemm \
  python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r1_run27_gemm.ncu-rep --page raw --csv > gpurun_out/r1_run27_gemm_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r1_run27_gemm_raw.csv gpurun_out/r1_run27_gemm_summary.json > gpurun_out/r1_run27_gemm_summary.txt 2>> $L
rm -f gpurun_out/r1_run27_small.ncu-rep gpurun_out/r1_run27_small_raw.csv gpurun_out/r1_run27_gemm_raw.csv
grep -v "==PROF==" $L | grep -v "^$" | tail -14
head -30 gpurun_out/r1_run27_launch_summary_byname.txt
