#!/bin/bash
# round 2, run 1: new parity tests (cfg-2 shapes, hot GEMM shapes, fake-world P2P, IPC), full GPU suite, smoke, bench N=1,
# reference-arm sample timing
mkdir -p gpurun_out
L=gpurun_out/r2_run01.log
: > $L
nvidia-smi --query-gpu=name,memory.total --format=csv >> $L 2>&1
nproc >> $L
echo "== new tests" >> $L
timeout 900 python -m pytest tests/test_gpu_cfg2_shapes.py tests/test_gpu_p2p.py -x -q -m gpu >> $L 2>&1
echo "exit $?" >> $L
echo "== hot gemm shapes" >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tail_split" >> $L 2>&1
echo "exit $?" >> $L
echo "== full suite" >> $L
timeout 1500 python -m pytest tests -q -m gpu >> $L 2>&1
echo "exit $?" >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "exit $?" >> $L
echo "== bench N=1" >> $L
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_run01_bench_n1.json 2>> $L
echo "exit $?" >> $L
echo "== reference arm (3 samples)" >> $L
timeout 900 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r2_run01_bench_ref.json 2>> $L
echo "exit $?" >> $L
tail -c 1500 gpurun_out/r2_run01_bench_n1.json
tail -c 1200 gpurun_out/r2_run01_bench_ref.json
grep -n "passed\|failed\|error\|exit\|smoke" $L | tail -30
