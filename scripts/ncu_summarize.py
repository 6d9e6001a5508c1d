"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name -> launches, total, mean."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    key = f"{name} grid={r.get('Grid Size','')} block={r.get('Block Size','')}"
    if len(sys.argv) > 2 and sys.argv[2] == "byname":
        key = name
    agg[key][0] += 1
    agg[key][1] += v_us
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={n:5d}  mean {t/n:8.1f} us  {k[:150]}")
