"""Print the headline numbers and the per-kernel table of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["unit"], "| launches", d["gpu_launches"])
r = d.get("roofline") or {}
print("roofline:", r.get("kernel"), "| frac", round(r.get("frac") or 0, 4), "|", round(r.get("achieved") or 0, 1), "GB/s")
for k, v in list((d.get("kernel_times") or {}).items())[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"  {k:22s} {1000 * v['ms_per_step']:9.1f} us/step  x{v['launches_per_step']:5.1f}  "
          f"{100 * (v.get('share_of_own_kernel_time') or 0):5.1f} %  frac {v.get('frac', float('nan')):.3f}")
