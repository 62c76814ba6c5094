#!/usr/bin/env python
"""Summarise an .ncu-rep per CUDA source line: share of executed warp instructions and stall samples.
usage: tools/ncu_lines.py rep.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; funcs = {}; cur = None
for x in rows:
    if len(x) >= 2 and x[0] == "Function Name":
        cur = x[1].split("(")[0]; funcs.setdefault(cur, []); continue
    if len(x) > 7 and x[0] == "Line No":
        hdr = x; continue
    if hdr and cur and len(x) >= 8 and x[0] not in ("", "Line No"):
        try:
            funcs[cur].append((float(x[7]), int(x[6] or 0), x[0], x[1].strip()[:110]))
        except ValueError:
            pass
for fn, lines in funcs.items():
    tot = sum(l[0] for l in lines) or 1; ts = sum(l[1] for l in lines) or 1
    lines.sort(reverse=True)
    print(f"== {fn}: total warp-instructions {tot:.0f}, samples {ts}")
    for v, smp, n, s in lines[:top]:
        print(f"{v / tot * 100:5.1f}% inst  {smp / ts * 100:5.1f}% smp  L{n:>4s}  {s}")
