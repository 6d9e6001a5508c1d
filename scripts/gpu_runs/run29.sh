#!/bin/bash
# driver-facing commands, defaults only
mkdir -p gpurun_out
L=gpurun_out/run29.log
echo "=== smoke" > $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" >> $L 2>&1
echo "exit $?" >> $L
echo "=== bench default" >> $L
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_run29_default.json 2> gpurun_out/bench_run29_default.err
echo "exit $?" >> $L
echo "wall $(( $(date +%s) - T0 )) s" >> $L
tail -3 gpurun_out/bench_run29_default.err | cut -c1-300 >> $L
echo "=== bench --impl reference" >> $L
T0=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_run29_reference.json 2> gpurun_out/bench_run29_reference.err
echo "exit $?" >> $L
echo "wall $(( $(date +%s) - T0 )) s" >> $L
cat gpurun_out/bench_run29_reference.json | cut -c1-900 >> $L
python - <<'PY' >> $L
import json
d=json.loads(open("gpurun_out/bench_run29_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in d if k not in ("config","roofline","profile_ms","profile_launches")})
print(d["roofline"]); print(d["profile_ms"])
PY
grep -v "^$" $L | tail -30 | cut -c1-1200
