#!/bin/bash
# first bring-up run: kernel unit tests, each family in its own process (a trap poisons the context)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/env.txt 2>&1; nproc >> gpurun_out/env.txt; free -g >> gpurun_out/env.txt
nvidia-smi topo -m >> gpurun_out/env.txt 2>&1
for k in "test_gemm_tn" "test_gemm_epilogues" "test_gemm_dw" "not gemm"; do
  echo "=== -k $k" >> gpurun_out/run1.log
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 200 -k "$k" >> gpurun_out/run1.log 2>&1
  echo "exit $?" >> gpurun_out/run1.log
done
tail -100 gpurun_out/run1.log
