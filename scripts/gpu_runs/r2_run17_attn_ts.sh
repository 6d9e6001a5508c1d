#!/bin/bash
# round 2, run 17: attention backward with TMEM A operands (dQ: Q, dO, dS; dK/dV: P^T, dS^T): tests, phase counters,
# ncu source-level capture of the dQ and dK/dV kernels (stall reasons per SASS line)
mkdir -p gpurun_out
L=gpurun_out/r2_run17.log
: > $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "attn or attention" >> $L 2>&1
echo "exit $?" >> $L
timeout 100 python -u scripts/prof_attn_phases.py > gpurun_out/r2_run17_attn_phases.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_bwd_d(q|kv)_tc_kernel" -s 8 -c 2 -f -o gpurun_out/r2_run17_attn_bwd python scripts/prof_attn_phases.py >> $L 2>&1
echo "exit $?" >> $L
grep -v "==PROF==" $L | tail -12
sed -n 1,2p gpurun_out/r2_run17_attn_phases.txt; sed -n 20,52p gpurun_out/r2_run17_attn_phases.txt
