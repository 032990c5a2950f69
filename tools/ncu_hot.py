"""Source lines of one kernel sorted by warp-stall SAMPLES (where the time goes), from an `ncu --set full --import-source on`
report:  python tools/ncu_hot.py report.ncu-rep "<substring of the kernel name>" [top]"""
import csv
import subprocess
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


path, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        secs.append(cur)
    elif r[0] == "Function Name":
        cur["fn"] = r[1]
    elif r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for s in secs:
    if pat not in s.get("fn", ""):
        continue
    h = s["hdr"]
    isamp, ii = h.index("# Samples"), h.index("Instructions Executed")
    rows = [r for r in s["rows"] if r[0] != ""]
    tot = sum(num(r[isamp]) for r in rows) or 1
    rows.sort(key=lambda r: -num(r[isamp]))
    print(f"## {s['file'].split('/')[-1]} ({tot} samples)")
    for r in rows[:top]:
        st = {k[6:]: num(r[h.index(k)]) for k in h if k.startswith("stall_")}
        tp = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
        print(f"{r[0]:>5} samp {100 * num(r[isamp]) / tot:5.1f}%  instr {num(r[ii]):>11,}  [{tp}]  {r[1].strip()[:90]}")
