#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_attn.py > gpurun_out/run11.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_.*_tc_kernel" -s 9 -c 3 -o gpurun_out/prof_r1_attn_tc python scripts/prof_attn.py >> gpurun_out/run11.log 2>&1
tail -20 gpurun_out/run11.log
