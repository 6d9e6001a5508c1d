#!/bin/bash
mkdir -p gpurun_out
echo "=== kernels" > gpurun_out/run4.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 200 -x >> gpurun_out/run4.log 2>&1
echo "exit $?" >> gpurun_out/run4.log
echo "=== learner" >> gpurun_out/run4.log
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> gpurun_out/run4.log 2>&1
echo "exit $?" >> gpurun_out/run4.log
echo "=== bench full" >> gpurun_out/run4.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run4.json 2>> gpurun_out/run4.log
echo "exit $?" >> gpurun_out/run4.log
cat gpurun_out/bench_run4.json >> gpurun_out/run4.log
tail -40 gpurun_out/run4.log
