"""Summarise `ncu -i X.ncu-rep --page raw --csv` into the per-kernel metrics quoted in DESIGN.md / bench.py."""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "sm__inst_executed_pipe_uniform.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
out = []
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for r in data:
    name = r[col["Kernel Name"]]
    print(f"### {name[:110]}  grid={r[col.get('Grid Size', 0)]}")
    rec = {"kernel": name}
    for k in hdr:
        if any(k.endswith(x) or k == x for x in KEEP):
            v, u = r[col[k]], units[col[k]]
            print(f"  {k:95s} {v:>16s} {u}")
            try:
                f = float(v.replace(",", ""))
                if k.endswith("dram__bytes_read.sum") or k.endswith("dram__bytes_write.sum"):
                    f *= UNIT.get(u, 1)
                rec[k.split("TriageCompute.")[-1]] = f
            except ValueError:
                pass
    out.append(rec)
    print()
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
