#!/usr/bin/env python
"""Executed warp instructions per SASS opcode of an .ncu-rep captured with --set full (source page).
usage: tools/ncu_opcodes.py rep.ncu-rep [kernel-substring] [topN]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, ix, agg = None, None, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1].split("(")[0][-60:]
        continue
    if r and r[0] == "Address":
        ix = {n: i for i, n in enumerate(r)}
        continue
    if not ix or not cur or len(r) < len(ix):
        continue
    op = r[ix["Source"]].split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    o = o.rstrip(";")
    a = agg.setdefault(cur, [collections.Counter(), collections.Counter()])
    a[0][o] += int(r[ix["Instructions Executed"]] or 0)
    a[1][o] += int(r[ix["# Samples"]] or 0)
for k, (cnt, smp) in agg.items():
    if only not in k:
        continue
    tot = sum(cnt.values()) or 1
    ts = sum(smp.values()) or 1
    print(f"== {k}: {tot} warp instructions, {ts} samples")
    for o, n in cnt.most_common(top):
        print(f"  {o:28s} {n:10d} {100 * n / tot:5.1f} %   samples {100 * smp[o] / ts:5.1f} %")
