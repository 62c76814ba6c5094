#!/usr/bin/env python
"""K4 in detection-ROI mode (the reference's segment_defects usage): 64 RGB frames, ~1400 boxes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qcgraft import ffi, synth
qc = ffi.Context.current(0)
rng = np.random.default_rng(0)
frames = torch.from_numpy(synth.make_frames(0, 64)).cuda()
K = 1400
w = rng.integers(17, 300, K); h = rng.integers(17, 300, K)
x = (rng.random(K) * (640 - w)).astype(int); y = (rng.random(K) * (640 - h)).astype(int)
regs = dict(img=rng.integers(0, 64, K), x=x, y=y, w=w, h=h, polarity=rng.integers(0, 2, K), thr_scale=1.0, min_area=20,
            max_comp=np.minimum(((w + 1) // 2) * ((h + 1) // 2) + 1, 1024))
for _ in range(3):
    res = qc.segment(frames, regs, morph_iters=1, swap_rb=True, want_mask=False, want_labels=False)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
for s, e in ev:
    s.record(); qc.segment(frames, regs, morph_iters=1, swap_rb=True, want_mask=False, want_labels=False); e.record()
torch.cuda.synchronize()
print("roi-mode K4 ms:", sorted(s.elapsed_time(e) for s, e in ev), "pixels", int(res["params"].total_pixels), "bands", res["params"].total_bands)
