#!/usr/bin/env python
"""K3 on 512 ROIs (~80x80 -> 224x224 bf16) of 16 frames: the kernel micro-benchmark case, short enough for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qcgraft import ffi, synth
qc = ffi.Context.current(0)
rng = np.random.default_rng(0)
frames = torch.from_numpy(synth.make_frames(7, 16)).cuda()
K = int(os.environ.get("K3_ROIS", 512))
r = np.zeros((K, 5), np.float32)
r[:, 0] = rng.integers(0, 16, K)
r[:, 1] = rng.integers(0, 500, K); r[:, 2] = rng.integers(0, 500, K)
if os.environ.get("K3_MIX"):
    # the bench step's box sizes: sides (l + r) * stride with l, r in 0.5..7.5 cells, strides 8 / 16 / 32 at 0.6 / 0.3 / 0.1
    st = rng.choice([8, 16, 32], K, p=[0.6, 0.3, 0.1])
    bw = ((rng.integers(1, 8, K) + rng.integers(1, 8, K)) * st).clip(8, 600); bh = ((rng.integers(1, 8, K) + rng.integers(1, 8, K)) * st).clip(8, 600)
    r[:, 1] = (rng.random(K) * (640 - bw)).astype(np.int32); r[:, 2] = (rng.random(K) * (640 - bh)).astype(np.int32)
    r[:, 3] = r[:, 1] + bw; r[:, 4] = r[:, 2] + bh
else:
    r[:, 3] = r[:, 1] + rng.integers(40, 120, K); r[:, 4] = r[:, 2] + rng.integers(40, 120, K)
if os.environ.get("K3_SORT"):          # biggest boxes first (the work items are handed out in table order)
    r = r[np.argsort(-(r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2]), kind="stable")]
rois = torch.from_numpy(r).cuda()
out = torch.empty((K, 3, 224, 224), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    qc.roi_align(frames, rois, out=out)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
for s, e in ev:
    s.record(); qc.roi_align(frames, rois, out=out); e.record()
torch.cuda.synchronize()
by = int(((r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2]) * 3).sum()) + K * 3 * 224 * 224 * 2
ts = sorted(s.elapsed_time(e) for s, e in ev)
print(f"K3 {K} rois: median {ts[len(ts) // 2] * 1000:.1f} us, best {ts[0] * 1000:.1f} us, {by / ts[len(ts) // 2] / 1e6:.0f} GB/s",
      "NOTMA" if os.environ.get("QC_K3_NOTMA") else "TMA")
qc.profile(True)
for _ in range(5):
    qc.roi_align(frames, rois, out=out)
torch.cuda.synchronize()
acc = {}
for name, ms in qc.profile_collect():
    acc.setdefault(name, []).append(ms * 1000)
qc.profile(False)
print("   per kernel (events inside the library):", ", ".join(f"{k} {sorted(v)[len(v) // 2]:.1f} us" for k, v in acc.items()))
