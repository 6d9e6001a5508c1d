export B200RL_PROFILE_ONE_STEP=1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"gemm_pair_kernel" -s 114 -c 1 -f -o gpurun_out/r2_run40_down_dx python bench.py --steps 1 --warmup 1 --no_cpu_baseline > /dev/null 2>&1
echo "exit $?"
ls -la gpurun_out/r2_run40_down_dx.ncu-rep
