#!/usr/bin/env python
"""cProfile of QualityControlDetector.predict_batch on the bench's cfg2 load (64 frames, YOLOv8s, ~30 injected candidates
per frame): where the host time of the reference-facing call goes."""
import cProfile, pstats, sys, os, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from qcgraft import synth
from qcgraft.detector import QualityControlDetector
n = 64
cfg = {"model": {"confidence_threshold": 0.7, "nms_threshold": 0.5},
       "quality_control": {"defect_classes": ["crack", "scratch", "dent", "discoloration", "contamination"]},
       "qcgraft": {"yolo_scale": "s", "device": 0, "synthetic_weights": True, "max_batch": n}}
d = QualityControlDetector("synthetic", "synthetic", cfg)
d.ensemble_predictor.yolo_detector.model.zero_detect_bias()
imgs = [synth.make_frame(8000 + i) for i in range(n)]
inj = [l.cuda().to(torch.bfloat16) for l in synth.make_head(8100, n, 640, 640, 30)]
for _ in range(2):
    d.predict_batch(imgs, _inject=inj)
torch.cuda.synchronize()
for _ in range(3):
    t = time.time(); r = d.predict_batch(imgs, _inject=inj); print("predict_batch ms", (time.time() - t) * 1000)
print("detections", sum(len(x["detections"]) for x in r))
pr = cProfile.Profile(); pr.enable()
d.predict_batch(imgs, _inject=inj)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40); print(s.getvalue()[:8000])
