#!/bin/bash
mkdir -p gpurun_out
echo "=== learner" > gpurun_out/run3.log
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> gpurun_out/run3.log 2>&1
echo "exit $?" >> gpurun_out/run3.log
echo "=== smoke" >> gpurun_out/run3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/run3.log 2>&1
echo "exit $?" >> gpurun_out/run3.log
echo "=== bench 4 layers quick" >> gpurun_out/run3.log
timeout 600 python bench.py --layers 4 --steps 2 --warmup 3 --no_cpu_baseline >> gpurun_out/run3.log 2>&1
echo "exit $?" >> gpurun_out/run3.log
echo "=== bench full" >> gpurun_out/run3.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_n1.json 2>> gpurun_out/run3.log
echo "exit $?" >> gpurun_out/run3.log
cat gpurun_out/bench_r1_n1.json >> gpurun_out/run3.log
tail -60 gpurun_out/run3.log
