"""Seeded inputs of the rule-engine golden cases, shared by the generator (which feeds them to the reference's
PostProcessor) and by tests/test_rules_cpu.py (which feeds them to qcgraft.rules.DetectionRules)."""
import numpy as np

CLASSES = ["crack", "scratch", "dent", "discoloration", "contamination"]
SEVERITIES = ["minor", "major", "critical"]

CASES = {
    "sparse": dict(seed=1, k=6, shape=(480, 640, 3), cluster=0.0, seg_frac=0.5),
    "clustered": dict(seed=2, k=40, shape=(640, 640, 3), cluster=0.8, seg_frac=0.6),
    "dense_150": dict(seed=3, k=150, shape=(1280, 1280, 3), cluster=0.7, seg_frac=0.3),
    "isolated_quirk": dict(seed=4, k=12, shape=(640, 640, 3), cluster=0.0, seg_frac=0.0, one_class=True),
    "low_conf_and_slivers": dict(seed=5, k=30, shape=(600, 800, 3), cluster=0.5, seg_frac=0.4, slivers=True,
                                 thresholds={"confidence_threshold": 0.6, "area_threshold_percent": 2.0}),
    "edge_boxes": dict(seed=6, k=25, shape=(320, 320, 3), cluster=0.6, seg_frac=0.8, edges=True),
    "empty": dict(seed=7, k=0, shape=(640, 640, 3), cluster=0.0, seg_frac=0.0),
    "single": dict(seed=8, k=1, shape=(640, 640, 3), cluster=0.0, seg_frac=1.0),
}


def jsonable(o):
    if isinstance(o, dict):
        return {str(k): jsonable(v) for k, v in o.items() if k not in ("mask",)}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (np.floating, np.integer, np.bool_)):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    return o


def build_case(seed, k, shape, cluster, seg_frac, one_class=False, slivers=False, edges=False, thresholds=None):
    """-> (detections, segmentation_results, image_shape, config) in the reference's dict formats
    (models/yolo_model.py:137-156 + models/ensemble.py:197-218 keys; inference/segmentation.py:118-132 regions)."""
    rng = np.random.default_rng(seed)
    H, W = shape[:2]
    centres = rng.uniform(0.1, 0.9, (max(1, k // 6), 2)) * (W, H)
    dets, regions = [], []
    for i in range(k):
        if rng.random() < cluster:
            c = centres[rng.integers(0, len(centres))] + rng.normal(0, 6, 2)
        else:
            c = rng.uniform(0.02, 0.98, 2) * (W, H)
        w, h = rng.uniform(8, 140), rng.uniform(8, 140)
        if slivers and rng.random() < 0.3:
            w, h = (rng.uniform(1, 4), rng.uniform(60, 200)) if rng.random() < 0.5 else (rng.uniform(100, 300), rng.uniform(2, 9))
        if edges and rng.random() < 0.4:
            c = np.array([rng.choice([1.0, W - 2.0]), rng.choice([1.0, H - 2.0])]) + rng.normal(0, 3, 2)
        x1, y1 = int(max(0, c[0] - w / 2)), int(max(0, c[1] - h / 2))
        x2, y2 = int(min(W, c[0] + w / 2)), int(min(H, c[1] + h / 2))
        x2, y2 = max(x2, x1), max(y2, y1)
        conf = float(np.float32(rng.uniform(0.3, 0.99)))
        cls = CLASSES[2] if one_class else CLASSES[int(rng.integers(0, 5))]
        sev = SEVERITIES[int(rng.integers(0, 3))]
        rconf = min(conf * 1.1, 1.0)
        d = {"id": i, "class": cls, "confidence": conf,
             "bbox": {"x1": x1, "y1": y1, "x2": x2, "y2": y2, "width": x2 - x1, "height": y2 - y1,
                      "center_x": (x1 + x2) / 2, "center_y": (y1 + y2) / 2},
             "area": (x2 - x1) * (y2 - y1), "severity": sev, "ensemble_confidence": conf * 0.6 + rconf * 0.4,
             "yolo_confidence": conf, "resnet_confidence": rconf, "final_severity": sev}
        if i % 7 == 3:
            del d["ensemble_confidence"]                      # falls back to 'confidence'
        dets.append(d)
        if rng.random() < seg_frac and x2 > x1 and y2 > y1:
            mask = np.zeros((H, W), np.uint8)
            nrect = int(rng.integers(0, 4))                   # 0 rectangles -> empty mask -> bbox stays
            for _ in range(nrect):
                ry0 = int(rng.integers(y1, y2)); rx0 = int(rng.integers(x1, x2))
                ry1 = int(rng.integers(ry0, y2)) + 1; rx1 = int(rng.integers(rx0, x2)) + 1
                mask[ry0:ry1, rx0:rx1] = 255
            area = int((mask > 0).sum())
            regions.append({"detection_id": i, "defect_class": cls, "confidence": conf, "bbox": d["bbox"],
                            "mask": mask if rng.random() < 0.9 else None, "contours": [None] * int(rng.integers(0, 4)),
                            "area_pixels": area if rng.random() < 0.8 else area * 5 + 50000,
                            "area_percentage": float(area / (H * W) * 100 * (1 if rng.random() < 0.8 else 40)),
                            "perimeter": float(rng.uniform(0, 400)), "compactness": float(rng.uniform(0, 1)),
                            "confidence_score": float(rng.uniform(0, 1))})
    cfg = {"quality_control": {"thresholds": thresholds}} if thresholds else {}
    return dets, {"segmented_regions": regions}, tuple(shape), cfg
