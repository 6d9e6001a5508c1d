#!/bin/bash
# round 2, run 43: final build (elected issuers): wide-tile modes A/B, then validation like the driver (pytest -m gpu, smoke,
# default bench with cpu_baseline) and the ncu launch list of one step
mkdir -p gpurun_out
L=gpurun_out/r2_run43_final_validation.log
: > $L
run() { B200RL_GEMM_WIDE=$1 timeout 300 python bench.py --steps 4 --warmup 3 --no_cpu_baseline --lean 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wide mode $1:', round(d['ms_per_step'],1), round(d['value']), d['profile_ms']['gemm'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"; }
echo "== wide-tile modes (2 = default)" >> $L
run 2 >> $L; run 1 >> $L; run 0 >> $L; run 2 >> $L
echo "== pytest -m gpu" >> $L
timeout 2400 python -m pytest tests -q -m gpu -x >> $L 2>&1
echo "exit $?" >> $L
echo "== smoke()" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
echo "exit $?" >> $L
echo "== python bench.py" >> $L
timeout 1200 python bench.py > gpurun_out/r2_run43_bench_n1_default.json 2>> $L
echo "exit $?" >> $L
export B200RL_PROFILE_ONE_STEP=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_run43_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline >> $L 2>&1
echo "ncu exit $?" >> $L
unset B200RL_PROFILE_ONE_STEP
python scripts/ncu_summarize.py gpurun_out/r2_run43_launches.csv byname > gpurun_out/r2_run43_launch_summary_byname.txt 2>&1
gzip -f gpurun_out/r2_run43_launches.csv
grep -v "^$" $L | grep -v "==PROF==" | tail -22
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_run43_bench_n1_default.json").read().strip().splitlines()[-1])
print("default bench: ms", round(d["ms_per_step"], 1), round(d["value"], 1), "e2e", d["e2e"]["value"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"], d["profile_ms"])
PY
head -12 gpurun_out/r2_run43_launch_summary_byname.txt
