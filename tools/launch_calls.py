#!/usr/bin/env python
"""Per-kernel device time of the LAST call of each launch sequence in an `ncu --metrics gpu__time_duration.sum` CSV
(calls are split at k4_init).  usage: tools/launch_calls.py launches.csv [call indices...]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
seq = []
for r in rows[hi + 1:]:
    if len(r) != len(h):
        continue
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "")
    seq.append((name, v / 1000 if r[ix["Metric Unit"]] in ("ns", "nsecond") else v))
ki = [i for i, (n, _) in enumerate(seq) if n.startswith("k4_init")]
calls = [int(a) for a in sys.argv[2:]] or [len(ki) - 1]
for call in calls:
    a = ki[call]; b = ki[call + 1] if call + 1 < len(ki) else len(seq)
    part = [(n, us) for n, us in seq[a:b] if not n.startswith("at::")]
    tot = sum(us for _, us in part)
    print(f"--- call {call}: {len(part)} launches, {tot:.1f} us (serialised, cold caches)")
    for n, us in part:
        print(f"   {n:20s} {us:8.1f} us  {us / tot * 100:5.1f}%")
