"""Per-source-line instruction share and warp-stall samples of one kernel from an `ncu --set full --import-source on`
report (read here without a GPU):  python tools/ncu_lines.py report.ncu-rep "<substring of the kernel name>" [top]"""
import csv
import subprocess
import sys

STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_not_selected",
          "stall_selected", "stall_sleep", "stall_membar", "stall_lg", "stall_branch_resolving", "stall_no_inst",
          "stall_dispatch", "stall_tex", "stall_misc", "stall_drain"]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
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
    tot = smp = 0
    agg = dict.fromkeys(STALLS, 0)
    lines = []
    for s in secs:
        if pat not in s.get("fn", ""):
            continue
        h = s["hdr"]
        ii, isamp = h.index("Instructions Executed"), h.index("# Samples")
        for r in s["rows"]:
            if r[0] == "":          # SASS rows under a source line
                continue
            n, m = num(r[ii]), num(r[isamp])
            tot += n
            smp += m
            st = {k: num(r[h.index(k)]) for k in STALLS}
            for k in STALLS:
                agg[k] += st[k]
            lines.append((n, m, s["file"].split("/")[-1], r[0], r[1].strip()[:90], max(st.items(), key=lambda kv: kv[1])[0]))
    print(f"kernel `{pat}`: {tot:,} warp instructions, {smp:,} stall samples\n")
    print("stall share of samples: " + ", ".join(f"{k[6:]} {100 * v / smp:.1f} %" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > smp))
    print("\n| line | instr % | samples % | top stall | source |\n|---|---:|---:|---|---|")
    for n, m, f, l, src, st in sorted(lines, reverse=True)[:top]:
        print(f"| {f}:{l} | {100 * n / tot:.1f} | {100 * m / smp:.1f} | {st[6:]} | `{src.replace('|', '/')}` |")


if __name__ == "__main__":
    main()
