#!/bin/bash
# round 2, run 13: (a) attention backward with two MMA-issuing threads: tests + phase profile + bench A/B is implicit vs run10
# (b) rasterisation sweep of the K-long GEMMs, timing without ncu and DRAM bytes with ncu (3 launches per point: take the last)
mkdir -p gpurun_out
L=gpurun_out/r2_run13.log
: > $L
echo "== attention tests" >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "attn or attention" >> $L 2>&1
echo "exit $?" >> $L
echo "== learner tests (incl. compact scored rows)" >> $L
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_cfg2_shapes.py -q -x -m gpu >> $L 2>&1
echo "exit $?" >> $L
timeout 120 python scripts/prof_attn_phases.py > gpurun_out/r2_run13_attn_phases.txt 2>&1
echo "== raster sweep (timing, no profiler)" >> $L
timeout 600 python scripts/gemm_raster_sweep.py > gpurun_out/r2_run13_raster_timing.txt 2>> $L
echo "exit $?" >> $L
echo "== raster sweep under ncu" >> $L
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:gemm_pair_kernel --csv --log-file gpurun_out/r2_run13_raster_ncu.csv python scripts/gemm_raster_sweep.py > gpurun_out/r2_run13_raster_order.txt 2>> $L
echo "exit $?" >> $L
echo "== bench" >> $L
timeout 600 python bench.py --steps 4 --warmup 3 --no_cpu_baseline > gpurun_out/r2_run13_bench.json 2>> $L
echo "exit $?" >> $L
cat $L | tail -20
cat gpurun_out/r2_run13_attn_phases.txt | grep -v "^  " 
cat gpurun_out/r2_run13_raster_timing.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_run13_bench.json').read().strip().splitlines()[-1])
print('bench ms', d['ms_per_step'], 'tok/s', d['value'], 'e2e', d['e2e']['value'], {k: v for k, v in d.get('profile_ms', {}).items()})
PY
