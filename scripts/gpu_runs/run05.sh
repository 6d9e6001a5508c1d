#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests (pair + single)" > gpurun_out/run5.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 200 -k "gemm" >> gpurun_out/run5.log 2>&1
rc=$?
echo "exit $rc" >> gpurun_out/run5.log
if [ $rc -ne 0 ]; then
  echo "PAIR KERNEL FAILED -> falling back to single-CTA for the rest" >> gpurun_out/run5.log
  export B200RL_GEMM_CTA_PAIR=0
fi
echo "=== gemm bench" >> gpurun_out/run5.log
timeout 600 python scripts/bench_gemm.py >> gpurun_out/run5.log 2>&1
echo "=== learner" >> gpurun_out/run5.log
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> gpurun_out/run5.log 2>&1
echo "exit $?" >> gpurun_out/run5.log
echo "=== bench full" >> gpurun_out/run5.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run5.json 2>> gpurun_out/run5.log
echo "exit $?" >> gpurun_out/run5.log
cat gpurun_out/bench_run5.json >> gpurun_out/run5.log
grep -v "^{'shape'.*'bn': \(128\|192\)" gpurun_out/run5.log | tail -60
