"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name -> markdown table."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v))
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v in rows)
print(f"launches: {len(rows)}, total device time {tot / 1e3:.2f} ms (cold-cache, serialised under ncu: compare SHARES)\n")
print("| kernel | launches | total us | share | avg us |")
print("|---|---:|---:|---:|---:|")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {v:.0f} | {100 * v / tot:.1f}% | {v / c:.1f} |")
