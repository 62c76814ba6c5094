#!/usr/bin/env python
"""One markdown row per kernel of an `ncu --set full` report: duration, DRAM bytes, throughputs, occupancy and the
two largest stall reasons.  usage: tools/ncu_summary.py rep.ncu-rep [...]"""
import csv
import re
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def to_unit(v, unit, want):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return float("nan")
    if want == "MB":
        return x * scale.get(unit, 1) / 1e6
    if want == "us":
        return x * scale.get(unit, 1)
    return x


print("| kernel | grid×block | regs | µs | DRAM read MB | DRAM write MB | DRAM thr % | SM thr % | issue active % | "
      "warps active % | warp-inst (M) | top stalls (warps per issue) |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, unit = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "")
        g = lambda k, want=None: to_unit(r[ix[k]], unit[ix[k]], want) if k in ix else float("nan")
        st = sorted(((float(r[ix[s]] or 0), s.split("stalled_")[1].split("_per_issue")[0]) for s in stalls), reverse=True)
        top = ", ".join(f"{n} {v:.1f}" for v, n in st[:2] if n not in ("selected",))
        print(f"| `{name}` | {int(g('launch__grid_size'))}×{int(g('launch__block_size'))} | {int(g('launch__registers_per_thread'))} | "
              f"{g('gpu__time_duration.sum', 'us'):.1f} | {g('dram__bytes_read.sum', 'MB'):.1f} | {g('dram__bytes_write.sum', 'MB'):.1f} | "
              f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{g('smsp__inst_executed.sum') / 1e6:.1f} | {top} |")
