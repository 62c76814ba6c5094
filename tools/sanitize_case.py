#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel once, small sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qcgraft import ffi, synth

qc = ffi.Context.current(0)
rng = np.random.default_rng(0)
# K1 fast + generic
imgs = synth.make_frames(0, 2, 96, 160)
d = torch.from_numpy(imgs).cuda()
qc.preprocess(d, qc.make_pre_plan(ffi.QC_PRE_LETTERBOX, 96, 160, 96, 160))
qc.preprocess(d, qc.make_pre_plan(ffi.QC_PRE_LETTERBOX, 96, 160, 128, 128), out_dtype=torch.float32)
# K2 (fused + nms-only, smem and global sort paths)
lv = [l.cuda() for l in synth.make_head(1, 2, 320, 320, 150)]
qc.yolo_decode_nms(lv, qc.yolo_params(lv, 5))
n = 5000
b = torch.rand(1, n, 4).cuda(); b[..., 2:] += b[..., :2]
qc.nms((b * 300).contiguous(), torch.rand(1, n).cuda(), torch.randint(0, 5, (1, n), dtype=torch.int32).cuda(),
       torch.tensor([n], dtype=torch.int32).cuda())
# K3 (separable + fallback)
rois = torch.tensor([[0, 3, 4, 60, 70], [1, 0, 0, 160, 96], [0, -500, -500, 900, 900], [1, 10, 10, 11, 11]], dtype=torch.float32).cuda()
qc.roi_align(d, rois)
qc.roi_align(d, rois, sampling_ratio=0, out_hw=(7, 7), out_dtype=torch.float32)
# K3b (PIL-exact resize: down-scale with runtime tap groups, up-scale, odd output)
qc.pil_resize(d)
qc.pil_resize(d, out_hw=(201, 333), out_dtype=torch.float32)
qc.pil_resize(torch.from_numpy(synth.make_frames(3, 1, 700, 1300)).cuda(), out_hw=(31, 45))
# K4 (full frame gray, RGB ROIs, odd sizes)
g = torch.from_numpy(np.stack([synth.make_linescan(5, 130, 203), synth.make_linescan(6, 130, 203)])).cuda()
qc.segment(g, [dict(img=0, x=0, y=0, w=203, h=130), dict(img=1, x=7, y=9, w=65, h=33, polarity=0)], morph_iters=2)
qc.segment(d, [dict(img=1, x=1, y=2, w=150, h=90, polarity=0, thr_scale=0.9)], morph_iters=1, swap_rb=True)
big = torch.from_numpy(synth.make_linescan(7, 300, 1100)[None]).cuda()
qc.segment(big, [dict(img=0, x=0, y=0, w=1100, h=300)], morph_iters=1)
# K5
det = torch.rand(2, 8, 6).cuda() * 100
qc.spc_stats(det, torch.tensor([3, 8], dtype=torch.int32).cuda())
torch.cuda.synchronize()
print("sanitize_case done, launches:", qc.launch_count)
