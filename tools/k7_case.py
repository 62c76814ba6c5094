#!/usr/bin/env python
"""K7 (the reference's threshold / adaptive segmentation methods) on ~1400 detection ROIs of 64 RGB frames, stats-only
outputs as the batched pipeline runs them; short enough for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qcgraft import ffi, synth
qc = ffi.Context.current(0)
rng = np.random.default_rng(0)
frames = torch.from_numpy(synth.make_frames(0, 64)).cuda()
K = int(os.environ.get("K7_ROIS", 1400))
w = rng.integers(16, 240, K); h = rng.integers(16, 240, K)
x = (rng.random(K) * (640 - w)).astype(int); y = (rng.random(K) * (640 - h)).astype(int)
img = rng.integers(0, 64, K)
which = os.environ.get("K7_METHOD", "both")
for name, code, iters, marea in (("threshold", ffi.QC_SEG_REF_THRESHOLD, 2, 10), ("adaptive", ffi.QC_SEG_REF_ADAPTIVE, 1, 20)):
    if which not in ("both", name):
        continue
    regs = dict(img=img, x=x, y=y, w=w, h=h, polarity=rng.integers(0, 2, K), thr_scale=0.8, min_area=marea)
    regs["max_comp"] = np.full(K, 2) if name == "threshold" else np.minimum(((w + 1) // 2) * ((h + 1) // 2) + 1, 1024)
    full = os.environ.get("K7_FULL", "1") == "1"      # with K8 contour statistics and K9 Canny confidence, as the pipeline runs it
    call = lambda: qc.segment(frames, regs, morph_iters=iters, swap_rb=True, want_mask=False, want_labels=False,
                              method=code, want_sums=full or name == "threshold", want_contours=full)
    for _ in range(3):
        res = call()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for s, e in ev:
        s.record(); call(); e.record()
    torch.cuda.synchronize()
    print(f"K7 {name}: ms", [round(s.elapsed_time(e), 3) for s, e in ev], "pixels", int(res["params"].total_pixels),
          "bands", res["params"].total_bands, "area sum", int(res["area"].sum()))
