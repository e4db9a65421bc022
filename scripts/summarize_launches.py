"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share for the LAST
`--launches-per-step` launches (one bench step).  Usage: python scripts/summarize_launches.py launches.csv 640 > profiles/x.md"""
import csv
import re
import sys
from collections import OrderedDict

path, per_step = sys.argv[1], int(sys.argv[2])
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        rows.append((r["Kernel Name"], v * scale))
skip = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows) - per_step
sel = rows[skip:skip + per_step]
agg = OrderedDict()
for name, us in sel:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void (lam::)?", "", short)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"launches in window: {len(sel)} (rows {skip}..{skip + len(sel)} of {len(rows)}), total device time {tot / 1e3:.2f} ms (ncu-serialised, cold cache)\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {us / 1e3:.3f} | {100 * us / tot:.1f}% | {us / c:.1f} |")
