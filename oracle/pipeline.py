"""CPU restatement of the reference's per-image inspection chain -- the "port" baseline that
bench.py's ``cpu_baseline`` / ``--impl reference`` legs time on the host cores.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

One ``predict(image)`` call = what ``QualityControlDetector.predict`` does for one frame
(inference/detector.py:72-126), stage by stage, with the libraries the reference calls:
  1. _preprocess_image: BGR2RGB swap                              inference/detector.py:233-234
  2. YOLODetector.predict: swap, [ultralytics: LetterBox, [...,::-1], NCHW, /255, net,
     Detect decode, NMS, scale_boxes], _parse_results            models/yolo_model.py:84-160
  3. ResNetClassifier.predict on the whole frame (PIL Resize 224 + ToTensor + Normalize)
                                                                  models/resent_model.py:114-179
  4. the ROI hand-off: the reference MOCKS it (models/ensemble.py:129-142,181-195: class = YOLO class, confidence x 1.1)
     -- ``classify_rois=False`` does exactly that and is what the reference's CPU path costs; ``classify_rois=True``
     really crops (roi_align, oracle/roi.py fixes the convention) and runs ResNet-50 on every kept box, the work the
     GPU pipeline does (the "same work" baseline)
  5. _merge_detection_classification                              models/ensemble.py:197-218
  6. segment_defects per detection: the reference's per-class methods (oracle/refseg.py), or the K4
     composite (seg_mode='composite')                             inference/segmentation.py:39-298
  7. PostProcessor.process_results (refine / filter / merge / validate / grade / risk)  inference/postprocess.py:32-109,
     through the array-form restatement pinned against the reference's class (``rules=`` object, optional)
``stage_s`` accumulates wall seconds per stage (pre | nets | decode_nms | roi | seg | rules) for bench.py's stage split.
The conv nets are the same random-init torch modules the GPU pipeline runs (float32 on CPU).
The reference runs YOLO and ResNet in a 2-thread pool (models/ensemble.py:88); torch already
uses every core for each net, so they run back to back here.
"""
from __future__ import annotations

import time

import cv2
import numpy as np
import torch
from PIL import Image
from torchvision import transforms

from . import preprocess as opre, yolo_head as oyh, roi as oroi, segment as oseg, refseg as orefseg


class CpuReferencePipeline:
    def __init__(self, yolo, resnet, nc=5, conf=0.7, iou=0.5, imgsz=640, max_det=300, seg_mode="reference",
                 classify_rois=True, rules=None):
        self.seg_mode, self.classify_rois, self.rules = seg_mode, classify_rois, rules
        self.stage_s = {k: 0.0 for k in ("pre", "nets", "decode_nms", "roi", "seg", "rules")}
        self.yolo, self.resnet = yolo.float().eval(), resnet.float().eval()
        self.nc, self.conf, self.iou, self.imgsz, self.max_det = nc, conf, iou, imgsz, max_det
        self.transform = transforms.Compose([                    # models/resent_model.py:59-66
            transforms.Resize((224, 224)), transforms.ToTensor(),
            transforms.Normalize(mean=list(oroi.RESNET_MEAN), std=list(oroi.RESNET_STD))])
        self.ensemble_weights = {"yolo": 0.6, "resnet": 0.4}     # models/ensemble.py:38-41

    @torch.no_grad()
    def predict(self, image: np.ndarray, inject=None, head_override=None, include_segmentation=True,
                morph_iters=1, min_area=20) -> dict:
        t0 = time.time()
        st = self.stage_s
        tick = time.perf_counter
        t = tick()
        processed = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)                      # detector.py:234
        # ---- YOLO -------------------------------------------------------------------------------
        yimg = cv2.cvtColor(processed, cv2.COLOR_BGR2RGB)                       # yolo_model.py:80
        x, g = opre.yolo_preprocess([yimg], self.imgsz, self.imgsz, dtype=torch.float32)
        st["pre"] += tick() - t; t = tick()
        if head_override is not None:
            levels = [l.float() for l in head_override]
        else:
            levels = self.yolo(x)
            if inject is not None:
                levels = [l + j.float() for l, j in zip(levels, inject)]
        st["nets"] += tick() - t; t = tick()
        dets = oyh.yolo_postprocess(levels, self.nc, x.shape[2:], yimg.shape[:2], self.conf, self.iou, self.max_det)
        detections = oyh.parse_results(dets[0].numpy())
        st["decode_nms"] += tick() - t; t = tick()
        # ---- whole-frame ResNet ---------------------------------------------------------------------
        rimg = cv2.cvtColor(processed, cv2.COLOR_BGR2RGB)                       # resent_model.py:123
        tin = self.transform(Image.fromarray(rimg)).unsqueeze(0)
        st["pre"] += tick() - t; t = tick()
        gp = torch.softmax(self.resnet(tin), 1)[0]
        st["nets"] += tick() - t; t = tick()
        gconf, gcls = torch.max(gp, 0)
        global_cls = {"predicted_class": oyh.CLASS_NAMES[int(gcls)], "confidence": float(gconf),
                      "class_probabilities": {oyh.CLASS_NAMES[i]: float(p) for i, p in enumerate(gp)}}
        # ---- ROI hand-off ------------------------------------------------------------------------------
        boxes = [(d["bbox"]["x1"], d["bbox"]["y1"], d["bbox"]["x2"], d["bbox"]["y2"]) for d in detections]
        valid = [i for i, b in enumerate(boxes) if b[2] > b[0] and b[3] > b[1]]
        roi_probs = None
        if valid and self.classify_rois:
            rois = np.array([[0, *boxes[i]] for i in valid], np.float32)
            crops = oroi.roi_align_normalize(rimg[None], rois)
            st["roi"] += tick() - t; t = tick()
            roi_probs = torch.softmax(self.resnet(crops), 1).numpy()
            st["nets"] += tick() - t; t = tick()
        enhanced = []
        for i, d in enumerate(detections):
            m = dict(d)
            yolo_conf = d["confidence"]
            resnet_conf = min(yolo_conf * 1.1, 1.0)                               # ensemble.py:186-195 (mock)
            m.update({"ensemble_confidence": yolo_conf * 0.6 + resnet_conf * 0.4, "yolo_confidence": yolo_conf,
                      "resnet_confidence": resnet_conf, "final_severity": d["severity"]})
            if roi_probs is not None and i in valid:
                m["roi_class_probabilities"] = roi_probs[valid.index(i)]
            enhanced.append(m)
        st["roi"] += tick() - t; t = tick()
        # ---- segmentation ---------------------------------------------------------------------------------
        seg = []
        seg_full = {"segmented_regions": [], "defect_density": 0.0}
        if include_segmentation and detections:
            gray = cv2.cvtColor(processed, cv2.COLOR_RGB2GRAY)                  # segmentation.py:66-69
            tot = 0
            for i in valid:
                x1, y1, x2, y2 = boxes[i]
                roi = np.ascontiguousarray(gray[y1:y2, x1:x2])
                if self.seg_mode == "reference":
                    m, conf, method = orefseg.segment_region(roi, detections[i]["class"])
                    area = int((m > 0).sum())
                    seg.append({"detection_id": i, "area_pixels": area, "method": method, "confidence_score": conf})
                    if self.rules is not None:      # what segment_defects also computes per region (:102-127)
                        cs = orefseg.find_contours(m)
                        per = float(sum(cv2.arcLength(c_, True) for c_ in cs))
                        gm = np.zeros(gray.shape, np.uint8)
                        gm[y1:y2, x1:x2] = m
                        seg_full["segmented_regions"].append({
                            "detection_id": i, "confidence_score": conf, "area_pixels": area,
                            "area_percentage": area / gray.size * 100, "perimeter": per,
                            "compactness": 0.0 if (not cs or area == 0 or per == 0) else min(4 * np.pi * area / per ** 2, 1.0),
                            "contours": cs, "mask": gm})
                        tot += area
                    continue
                dark = detections[i]["class"] in oseg.DARK_CLASSES
                r = oseg.segment_k4(roi, 0 if dark else 1, 1.0, morph_iters, min_area)
                seg.append({"detection_id": i, "area_pixels": int((r["mask"] > 0).sum()), "thr": r["thr"],
                            "ncomp": r["ncomp"]})
            seg_full["defect_density"] = tot / gray.size * 100
        st["seg"] += tick() - t; t = tick()
        out = {"detections": enhanced, "global_classification": global_cls, "segmented_regions": seg,
               "total_inference_time_ms": (time.time() - t0) * 1000}
        if self.rules is not None and enhanced:
            final = self.rules.process(enhanced, seg_full, processed.shape)
            quality = self.rules._assess_quality(final, seg_full)
            out["final"] = {"detections": final, "quality_assessment": quality,
                            "risk_analysis": self.rules._analyze_risks(final, quality)}
        st["rules"] += tick() - t
        return out

    def benchmark(self, images, injects=None, iterations=1):
        """inference/detector.py:331-374: sequential loop, images_per_second = N*iters / sum(t)."""
        times = []
        for _ in range(iterations):
            for k, im in enumerate(images):
                t = time.time()
                self.predict(im, inject=None if injects is None else injects[k])
                times.append(time.time() - t)
        return {"images_per_second": len(times) / sum(times), "times": times}
