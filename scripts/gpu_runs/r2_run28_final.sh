#!/bin/bash
# round 2, run 28: validation of the final build exactly like the driver: pytest -m gpu, smoke(), default bench.py (with
# cpu_baseline), bench.py --impl reference (bounded samples)
mkdir -p gpurun_out
L=gpurun_out/r2_run28_final_validation.log
: > $L
echo "== pytest -m gpu" >> $L
timeout 2400 python -m pytest tests -q -m gpu -x >> $L 2>&1
echo "exit $?" >> $L
echo "== smoke()" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
echo "exit $?" >> $L
echo "== python bench.py" >> $L
timeout 1200 python bench.py > gpurun_out/r2_run28_bench_n1_default.json 2>> $L
echo "exit $?" >> $L
echo "== python bench.py --ragged" >> $L
timeout 600 python bench.py --ragged --steps 4 --warmup 3 --no_cpu_baseline > gpurun_out/r2_run28_bench_n1_ragged.json 2>> $L
echo "exit $?" >> $L
echo "== python bench.py --impl reference --steps 3 --warmup 0" >> $L
timeout 1500 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r2_run28_bench_reference_arm.json 2>> $L
echo "exit $?" >> $L
grep -v "^$" $L | tail -25
python - <<'PY'
import json
for f in ("r2_run28_bench_n1_default", "r2_run28_bench_n1_ragged", "r2_run28_bench_reference_arm"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 1), d["unit"], round(d["value"], 1), "e2e", d.get("e2e", {}).get("value"), "roofline", d.get("roofline"), "cpu", d.get("cpu_baseline"), "clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(f, "ERR", e)
PY
