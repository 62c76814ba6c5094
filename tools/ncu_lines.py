#!/usr/bin/env python
"""Per CUDA source line of an .ncu-rep (compiled with -lineinfo, captured with --import-source on): share of executed
warp instructions and of the stall samples, summed over the SASS of the line and over the captured launches.
usage: tools/ncu_lines.py rep.ncu-rep [topN] [kernel-substring]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
only = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, hdr, agg = None, None, {}
for x in csv.reader(out.splitlines()):
    if len(x) >= 2 and x[0] == "Function Name":
        cur = x[1].split("(")[0][:60]
        continue
    if len(x) > 7 and x[0] == "Line No":
        hdr = x
        continue
    if hdr and cur and len(x) >= 8 and x[0] not in ("", "Line No"):
        try:
            a = agg.setdefault(cur, collections.OrderedDict())
            v = a.setdefault((x[0], x[1].strip()[:110]), [0.0, 0])
            v[0] += float(x[7])
            v[1] += int(x[6] or 0)
        except ValueError:
            pass
for fn, a in agg.items():
    if only not in fn:
        continue
    tot = sum(v[0] for v in a.values()) or 1
    ts = sum(v[1] for v in a.values()) or 1
    print(f"== {fn}: warp-instructions {tot:.0f}, stall samples {ts}")
    for (ln, src), v in sorted(a.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * v[0] / tot:5.1f}% inst {100 * v[1] / ts:5.1f}% smp  L{ln:>4s}  {src}")
