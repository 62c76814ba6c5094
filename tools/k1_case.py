#!/usr/bin/env python
"""K1 with a real resize: 32 x 1080x1920 -> letterbox 640 (auto: 384x640) bf16; short enough for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qcgraft import ffi
qc = ffi.Context.current(0)
rng = np.random.default_rng(0)
b = [torch.from_numpy(rng.integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).cuda() for _ in range(2)]
plan = qc.make_pre_plan(ffi.QC_PRE_LETTERBOX_AUTO, 1080, 1920, 640, 640)
outs = [torch.empty((32, 3, plan.dst_h, plan.dst_w), dtype=torch.bfloat16, device="cuda") for _ in range(2)]
for i in range(4):
    qc.preprocess(b[i % 2], plan, out=outs[i % 2])
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
for i, (s, e) in enumerate(ev):
    s.record(); qc.preprocess(b[i % 2], plan, out=outs[i % 2]); e.record()
torch.cuda.synchronize()
ts = sorted(s.elapsed_time(e) for s, e in ev)
tapped = 32 * (720 * 1920 * 3 + 3 * plan.dst_h * plan.dst_w * 2)
allb = 32 * (1080 * 1920 * 3 + 3 * plan.dst_h * plan.dst_w * 2)
print(f"K1 1080p letterbox: median {ts[4] * 1000:.1f} us; all source bytes {allb / ts[4] / 1e6:.0f} GB/s; "
      f"source rows actually tapped (2 of 3) {tapped / ts[4] / 1e6:.0f} GB/s")
