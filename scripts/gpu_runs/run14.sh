#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_d.*_tc_kernel" -s 6 -c 2 -o gpurun_out/prof_r1_attn_bwd_tc python scripts/prof_attn.py > gpurun_out/run14.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attn_" -c 40 --csv --log-file gpurun_out/attn_launches.csv python scripts/prof_attn.py >> gpurun_out/run14.log 2>&1
grep -v "==PROF==" gpurun_out/run14.log | tail -8
