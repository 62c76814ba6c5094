#!/usr/bin/env python
"""Markdown table (kernel, launches, mean / min microseconds, grid) from an `ncu --metrics gpu__time_duration.sum --csv`
launch list.  usage: tools/launch_table.py list.csv [name-regex]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("<unnamed>::", "").replace("void ", "")
    if pat and not pat.search(name):
        continue
    agg.setdefault(name, []).append((float(r[14]) / 1000, r[8]))
tot = sum(sum(t for t, _ in v) / len(v) for v in agg.values())
print("| kernel | launches | mean µs | min µs | grid | share of one pass |")
print("|---|---:|---:|---:|---|---:|")
for k, v in agg.items():
    ts = [t for t, _ in v]
    print(f"| `{k}` | {len(ts)} | {sum(ts) / len(ts):.1f} | {min(ts):.1f} | {v[0][1]} | {100 * sum(ts) / len(ts) / tot:.1f} % |")
