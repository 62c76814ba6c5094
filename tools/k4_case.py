#!/usr/bin/env python
"""K4 on cfg4 (16 x 2048^2 gray line-scan frames, composite: Otsu + cross morphology + CCL + stats + labels): per-kernel
durations from the library's own events.  K4_FRAMES=n for a smaller case (ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qcgraft import ffi, synth
qc = ffi.Context.current(0)
n = int(os.environ.get("K4_FRAMES", 16))
g = torch.from_numpy(synth.make_linescans(300, n)).cuda()
regs = [dict(img=i, x=0, y=0, w=2048, h=2048, min_area=20) for i in range(n)]
for _ in range(3):
    res = qc.segment(g, regs, morph_iters=1)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
for s, e in ev:
    s.record(); res = qc.segment(g, regs, morph_iters=1); e.record()
torch.cuda.synchronize()
ts = sorted(s.elapsed_time(e) for s, e in ev)
by = n * 2048 * 2048 * 6 + 20 * int(res["ncomp"].sum())
print(f"K4 cfg4 {n} frames: median {ts[len(ts) // 2] * 1000:.1f} us, {by / ts[len(ts) // 2] / 1e6:.0f} GB/s, components {int(res['ncomp'].sum())}")
qc.profile(True)
for _ in range(5):
    qc.segment(g, regs, morph_iters=1)
torch.cuda.synchronize()
acc = {}
for name, ms in qc.profile_collect():
    acc.setdefault(name, []).append(ms * 1000)
qc.profile(False)
print("  " + ", ".join(f"{k} {sum(v) / 5:.1f}" for k, v in acc.items()))
