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
# K7 / K8 / K9 (reference segmentation methods: run union-find fill, sweeps fallback on a noise ROI, adaptive + Canny)
gn = torch.from_numpy(rng.integers(0, 256, (1, 300, 310), dtype=np.uint8)).cuda()
for code, iters, marea in ((ffi.QC_SEG_REF_THRESHOLD, 2, 10), (ffi.QC_SEG_REF_ADAPTIVE, 1, 20)):
    regs = [dict(img=0, x=0, y=0, w=203, h=130, polarity=1, min_area=marea), dict(img=1, x=7, y=9, w=65, h=33, polarity=0, min_area=marea),
            dict(img=0, x=100, y=50, w=33, h=70, polarity=1, min_area=marea)]
    if code == ffi.QC_SEG_REF_THRESHOLD:
        for r in regs:
            r["max_comp"] = 2
    qc.segment(g, regs, morph_iters=iters, method=code, want_sums=True, want_contours=True)
    rn = [dict(img=0, x=3, y=2, w=301, h=290, polarity=1, min_area=marea)]
    if code == ffi.QC_SEG_REF_THRESHOLD:
        rn[0]["max_comp"] = 2
    qc.segment(gn, rn, morph_iters=iters, method=code, want_sums=True, want_contours=True)
# K6 (clustering alone + the whole chain)
rows = np.zeros((2, 40), ffi.RULES_ROW_DTYPE)
xy = rng.integers(0, 250, (2, 40, 2)); wh = rng.integers(4, 60, (2, 40, 2))
rows["x1"], rows["y1"], rows["x2"], rows["y2"] = xy[..., 0], xy[..., 1], xy[..., 0] + wh[..., 0], xy[..., 1] + wh[..., 1]
rows["cls"] = rng.integers(0, 3, (2, 40)); rows["severity"] = rng.integers(1, 4, (2, 40)); rows["conf"] = rng.uniform(0.3, 1, (2, 40))
rows["ext"] = -1
qc.rules_apply(rows, np.array([40, 17], np.int32), np.array([[320, 320], [300, 310]], np.int32))
# K10 (4:2:0 with restart intervals, 4:4:4 odd size, gray)
import cv2
streams = [cv2.imencode(".jpg", np.ascontiguousarray(imgs[0][..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_RST_INTERVAL, 3])[1],
           cv2.imencode(".jpg", np.ascontiguousarray(imgs[1][:77, :131, ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])[1],
           cv2.imencode(".jpg", imgs[0][:50, :9, 0])[1]]
qc.jpeg_decode(streams)
torch.cuda.synchronize()
print("sanitize_case done, launches:", qc.launch_count)
