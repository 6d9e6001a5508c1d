#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run18.log
echo "=== gemm tests" > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 120 -k gemm >> $L 2>&1
echo "exit $?" >> $L
echo "=== gemm bench" >> $L
timeout 600 python scripts/bench_gemm.py >> $L 2>&1
echo "exit $?" >> $L
grep -v "^$" $L | tail -90 | cut -c1-200
