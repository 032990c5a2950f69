"""Summarise `ncu --set full` reports (.ncu-rep, read here without a GPU) into a markdown table of the metrics that
the roofline discussion uses."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "warp-instr")]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"### {path}\n")
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|" + "---:|" * len(KEYS))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].replace("ysp::", "")[:60]
        cells = []
        for k, _ in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            else:
                cells.append("-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")
    print()
