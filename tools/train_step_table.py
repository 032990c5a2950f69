"""Per-kernel totals of the LAST complete training step in an ncu launch list with time + DRAM bytes
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`): time share and achieved DRAM GB/s.
    python tools/train_step_table.py gpurun_out/launches_train_full.csv"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
cur = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = cur.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ysp::", "").replace("ysp::", ""), "bytes": 0.0, "us": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
    else:
        d["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
rows = list(cur.values())
ends = [i for i, d in enumerate(rows) if "adamw_kernel" in d["name"]]
step = rows[ends[-2] + 1:ends[-1] + 1] if len(ends) >= 2 else rows
tot = sum(d["us"] for d in step)
print(f"launches in the last step: {len(step)}, kernel time {tot / 1e3:.2f} ms (cold-cache, serialised under ncu: compare SHARES), "
      f"DRAM traffic {sum(d['bytes'] for d in step) / 1e9:.1f} GB\n")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in step:
    a = agg[d["name"]]
    a[0] += 1; a[1] += d["us"]; a[2] += d["bytes"]
print("| kernel | launches | us | share | DRAM GB/s |\n|---|---:|---:|---:|---:|")
for k, (c, us, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:60]}` | {c} | {us:.0f} | {100 * us / tot:.1f}% | {b / us / 1e3 if us else 0:.0f} |")
