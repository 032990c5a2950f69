"""Summarise the LAST training step of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv` of
`bench.py --workload train`): per-kernel totals and the slowest individual launches."""
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches_train.csv"
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    rows.append((int(row["ID"]), row["Kernel Name"][:52], row["Grid Size"], row["Block Size"],
                 float(row["Metric Value"].replace(",", "")) / 1000))
if not per_step:   # a step ends with adamw_kernel
    ends = [i for i, r in enumerate(rows) if "adamw_kernel" in r[1]]
    per_step = ends[-1] - ends[-2]
    last = rows[ends[-2] + 1:ends[-1] + 1]
else:
    last = rows[-per_step:]
print(f"launches in the last step: {len(last)}, total {sum(x[4] for x in last):.0f} us")
agg = {}
for x in last:
    a = agg.setdefault(x[1], [0.0, 0])
    a[0] += x[4]
    a[1] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{v[0]:9.0f} us {v[1]:4d}  {k}")
for x in sorted(last, key=lambda x: -x[4])[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
    print(x)
