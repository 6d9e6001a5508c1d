#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run17.log
echo "=== learner" > $L
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q -p no:cacheprovider --timeout 300 >> $L 2>&1
echo "exit $?" >> $L
echo "=== bench packed + weight cache" >> $L
timeout 900 python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/bench_run17.json 2>> $L
echo "exit $?" >> $L
echo "=== ncu launch list, one step" >> $L
B200RL_PROFILE_ONE_STEP=1 timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_run17_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "exit $?" >> $L
python scripts/ncu_summarize.py gpurun_out/r1_run17_launches.csv > gpurun_out/r1_run17_launch_summary.txt 2>&1
python scripts/ncu_summarize.py gpurun_out/r1_run17_launches.csv byname > gpurun_out/r1_run17_launch_summary_byname.txt 2>&1
gzip -f gpurun_out/r1_run17_launches.csv
grep -v "^$" $L | grep -v "==PROF==" | tail -30 | cut -c1-300
head -50 gpurun_out/r1_run17_launch_summary.txt
python - <<'PY'
import json
for f in ("gpurun_out/bench_run17.json",):
    try:
        d=json.load(open(f)); print(f, "ms", round(d["ms_per_step"],1), "tok/s", round(d["value"]), "e2e", d["e2e"]["ms_per_step"], "gemm TF", d["roofline"]["achieved"]); print(d.get("profile_ms"))
    except Exception as e: print(f, e)
PY
