#!/bin/bash
mkdir -p gpurun_out
echo "=== attention" > gpurun_out/run2.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 200 -k "attention" >> gpurun_out/run2.log 2>&1
echo "exit $?" >> gpurun_out/run2.log
echo "=== learner" >> gpurun_out/run2.log
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> gpurun_out/run2.log 2>&1
echo "exit $?" >> gpurun_out/run2.log
echo "=== gemm bench" >> gpurun_out/run2.log
timeout 600 python scripts/bench_gemm.py >> gpurun_out/run2.log 2>&1
echo "exit $?" >> gpurun_out/run2.log
tail -150 gpurun_out/run2.log
