#!/bin/bash
# round 2, run 29: reduce_adam_kernel<8> / <4> / <2> alone (no barrier kernels) timed and under ncu --set full
mkdir -p gpurun_out
L=gpurun_out/r2_run29.log
: > $L
for w in 2 4 8; do timeout 120 python scripts/ncu_reduce_adam.py $w >> $L 2>&1; done
timeout 300 ncu --set full --clock-control none -k regex:"reduce_adam_kernel" -s 16 -c 8 -f -o gpurun_out/r2_run29_reduce python scripts/ncu_reduce_adam.py 8 >> $L 2>&1
echo "exit $?" >> $L
ncu -i gpurun_out/r2_run29_reduce.ncu-rep --page raw --csv > gpurun_out/r2_run29_reduce_raw.csv 2>> $L
python scripts/ncu_raw_summary.py gpurun_out/r2_run29_reduce_raw.csv gpurun_out/r2_run29_reduce_summary.json > gpurun_out/r2_run29_reduce_summary.txt 2>> $L
rm -f gpurun_out/r2_run29_reduce_raw.csv
grep -v "==PROF==" $L | grep -v "^$" | tail
head -24 gpurun_out/r2_run29_reduce_summary.txt
