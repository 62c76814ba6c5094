"""Post-NMS detection rules on packed arrays (host form).

The reference refines, filters, merges and validates the <= 300 detections of an image with Python loops over
dicts (inference/postprocess.py: ``_refine_detections`` :111-184, ``_filter_detections`` :186-231,
``_merge_overlapping_detections`` :233-303 with ``_merge_detection_cluster`` :305-358, ``_validate_detections``
:360-410 with ``_calculate_validation_score`` :412-464).  Here the same rules run on one structure-of-arrays table
per image -- the layout a device kernel takes over next (boxes int64[k,4], confidences f64[k], class / severity
codes, segmentation columns) -- and only the last step turns rows back into the reference's dicts.

Arithmetic notes that matter for parity (pinned against the reference's own class on golden cases,
tests/golden/rules_golden.json.gz):
  * IoU is integer intersection / integer union in float64; pairs are linked when ``1.0 - iou <= 1.0 - 0.3``.
  * The reference clusters with DBSCAN(metric='precomputed', eps=0.7, min_samples=1) on ``1 - iou_matrix`` whose
    DIAGONAL is 1.0 (the IoU matrix is zero there), so a box is not its own neighbour: boxes with no partner are
    noise (label -1) and ALL noise boxes of a class form one "cluster" that is merged when it has >= 2 members;
    the real clusters are the connected components of the link graph, numbered by their first member.  np.unique
    puts the noise group first.
  * A merged row keeps the dict of its most confident member (first maximum), takes the union box, np.mean of the
    confidences, the maximum severity (first maximum) and the sum of the positive areas.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

SEVERITY_ORDER = {"minor": 1, "major": 2, "critical": 3}
OVERLAP_THRESHOLD = 0.3
MASK_PADDING = 5


# ---------------------------------------------------------------------------------------------------
# array kernels (what a device version replaces)
def iou_matrix(box: np.ndarray) -> np.ndarray:
    """box int64[k,4] (x1,y1,x2,y2; width/height = differences) -> float64[k,k], zero diagonal."""
    b = np.asarray(box, np.int64)
    x1 = np.maximum(b[:, None, 0], b[None, :, 0]); y1 = np.maximum(b[:, None, 1], b[None, :, 1])
    x2 = np.minimum(b[:, None, 2], b[None, :, 2]); y2 = np.minimum(b[:, None, 3], b[None, :, 3])
    inter = np.where((x2 > x1) & (y2 > y1), (x2 - x1) * (y2 - y1), 0)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    union = area[:, None] + area[None, :] - inter
    out = np.zeros(inter.shape, np.float64)
    np.divide(inter, union, out=out, where=(inter > 0) & (union > 0))
    np.fill_diagonal(out, 0.0)
    return out


def cluster_labels(iou: np.ndarray) -> np.ndarray:
    """DBSCAN(precomputed, eps = 1 - 0.3, min_samples = 1) labels of ``1 - iou`` with a non-zero diagonal:
    -1 for boxes without a partner, else the index of the connected component in order of its first member."""
    k = iou.shape[0]
    link = (1.0 - iou) <= (1.0 - OVERLAP_THRESHOLD)
    np.fill_diagonal(link, False)
    parent = np.arange(k)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    ii, jj = np.nonzero(np.triu(link, 1))
    for a, b in zip(ii.tolist(), jj.tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    labels = np.full(k, -1, np.int64)
    has_partner = link.any(1)
    nxt = 0
    seen = {}
    for i in range(k):
        if not has_partner[i]:
            continue
        r = find(i)
        if r not in seen:
            seen[r] = nxt
            nxt += 1
        labels[i] = seen[r]
    return labels


def filter_keep(conf: np.ndarray, area_pct: np.ndarray, width: np.ndarray, height: np.ndarray,
                min_conf: float = 0.5, max_area_pct: float = 50.0) -> np.ndarray:
    """inference/postprocess.py:186-231 as one boolean column."""
    w = np.asarray(width, np.float64); h = np.asarray(height, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        aspect = np.where(h > 0, w / np.where(h > 0, h, 1), np.inf)
    keep = ~(np.asarray(conf, np.float64) < min_conf)
    keep &= ~(np.asarray(area_pct, np.float64) > max_area_pct)
    keep &= ~((aspect > 10) | (aspect < 0.1))
    keep &= ~((np.asarray(width) < 5) | (np.asarray(height) < 5))
    return keep


def validation_score(conf, width, height, cx, cy, has_seg, seg_conf, image_shape) -> np.ndarray:
    """inference/postprocess.py:412-464 on columns (same operation order per row)."""
    conf = np.asarray(conf, np.float64)
    H, W = image_shape[0], image_shape[1]
    confidence_score = np.minimum(conf / 0.8, 1.0)
    area_ratio = (np.asarray(width) * np.asarray(height)) / (H * W)
    size_score = np.where((0.001 <= area_ratio) & (area_ratio <= 0.5), 1.0, 0.5)
    aspect = np.asarray(width) / np.maximum(np.asarray(height), 1)
    aspect_score = np.where((0.2 <= aspect) & (aspect <= 5.0), 1.0, 0.5)
    seg_score = np.where(np.asarray(has_seg, bool), np.minimum(np.asarray(seg_conf, np.float64) / 0.5, 1.0), 0.7)
    nx = np.asarray(cx, np.float64) / W
    ny = np.asarray(cy, np.float64) / H
    location_score = np.minimum(np.minimum(np.minimum(nx, 1 - nx) / 0.05, np.minimum(ny, 1 - ny) / 0.05), 1.0)
    return 0.3 * confidence_score + 0.2 * size_score + 0.2 * aspect_score + 0.2 * seg_score + 0.1 * location_score


# ---------------------------------------------------------------------------------------------------
def _bbox_dict(x1, y1, x2, y2) -> Dict:
    return {"x1": x1, "y1": y1, "x2": x2, "y2": y2, "width": x2 - x1, "height": y2 - y1,
            "center_x": (x1 + x2) / 2, "center_y": (y1 + y2) / 2}


def _conf_of(d: Dict):
    return d.get("ensemble_confidence", d.get("confidence", 0.0))


class DetectionRules:
    """The refine -> filter -> merge -> validate chain of the reference's PostProcessor with the same method names
    and results; thresholds from ``config['quality_control']['thresholds']`` (inference/postprocess.py:30)."""

    def __init__(self, config: Optional[Dict] = None):
        self.config = config or {}
        self.quality_thresholds = self.config.get("quality_control", {}).get("thresholds", {})

    # -- :111-184 ------------------------------------------------------------------------------------
    def _refine_detections(self, detections: Sequence[Dict], segmentation_results: Dict, image_shape: Tuple) -> List[Dict]:
        regions = {}
        for r in (segmentation_results or {}).get("segmented_regions", []):
            regions.setdefault(r.get("detection_id"), r)          # first region of an id wins, like the linear search
        out = []
        for i, det in enumerate(detections):
            d = det.copy()
            r = regions.get(i)
            if r:
                d.update({"segmentation_confidence": r.get("confidence_score", 0.0), "area_pixels": r.get("area_pixels", 0),
                          "area_percentage": r.get("area_percentage", 0.0), "perimeter": r.get("perimeter", 0.0),
                          "compactness": r.get("compactness", 0.0), "contour_count": len(r.get("contours", [])),
                          "has_segmentation": True})
                d["bbox"] = self._refine_bbox_from_mask(det["bbox"], r.get("mask"), image_shape)
                d["final_severity"] = self._refine_severity(det, r)
            else:
                bb = det["bbox"]
                d.update({"segmentation_confidence": 0.0, "area_pixels": int(bb["width"] * bb["height"] * 0.6),
                          "area_percentage": 0.0, "perimeter": 0.0, "compactness": 0.0, "contour_count": 0,
                          "has_segmentation": False})
                d["final_severity"] = det.get("severity", "minor")
            bb = d["bbox"]
            d.update({"aspect_ratio": bb["width"] / max(bb["height"], 1), "bbox_area": bb["width"] * bb["height"],
                      "normalized_position": {"x": bb["center_x"] / 1000, "y": bb["center_y"] / 1000}})
            out.append(d)
        return out

    @staticmethod
    def mask_extent(mask) -> Optional[Tuple[int, int, int, int]]:
        """(y_min, y_max, x_min, x_max) of mask > 0, or None -- the four numbers K4's per-component bounding boxes
        already hold on the device."""
        if mask is None:
            return None
        if hasattr(mask, "extent"):              # detector._LazyFullMask: same numbers from the ROI-sized mask
            return mask.extent()
        m = np.asarray(mask) > 0
        rows = np.flatnonzero(m.any(1))
        if rows.size == 0:
            return None
        cols = np.flatnonzero(m.any(0))
        return rows[0], rows[-1], cols[0], cols[-1]

    def _refine_bbox_from_mask(self, original_bbox: Dict, mask, image_shape: Tuple) -> Dict:   # :764-812
        ext = self.mask_extent(mask)
        if ext is None:
            return original_bbox
        y_min, y_max, x_min, x_max = ext
        y_min = max(0, y_min - MASK_PADDING); x_min = max(0, x_min - MASK_PADDING)
        y_max = min(image_shape[0], y_max + MASK_PADDING); x_max = min(image_shape[1], x_max + MASK_PADDING)
        return _bbox_dict(x_min, y_min, x_max, y_max)

    @staticmethod
    def _refine_severity(detection: Dict, seg_region: Dict) -> str:                             # :814-840
        sev = detection.get("severity", "minor")
        if seg_region.get("area_percentage", 0.0) > 5.0:
            if sev == "minor":
                return "major"
            if sev == "major":
                return "critical"
        if seg_region.get("compactness", 0.0) < 0.3 and detection["class"] in ("crack", "scratch") and sev == "minor":
            return "major"
        return sev

    # -- :186-231 ------------------------------------------------------------------------------------
    def _filter_detections(self, detections: Sequence[Dict]) -> List[Dict]:
        if not detections:
            return []
        keep = filter_keep([_conf_of(d) for d in detections], [d.get("area_percentage", 0.0) for d in detections],
                           [d["bbox"]["width"] for d in detections], [d["bbox"]["height"] for d in detections],
                           self.quality_thresholds.get("confidence_threshold", 0.5),
                           self.quality_thresholds.get("area_threshold_percent", 50.0))
        return [d for d, k in zip(detections, keep) if k]

    # -- :233-358 ------------------------------------------------------------------------------------
    def _merge_overlapping_detections(self, detections: List[Dict], labels=None) -> List[Dict]:
        """``labels``: optional per-detection cluster labels computed elsewhere (K6 on the device: -1 isolated,
        >= 0 component of its class); without them the IoU matrix and the clustering run here."""
        if len(detections) <= 1:
            return detections
        groups: Dict[str, List[int]] = {}
        for i, d in enumerate(detections):
            groups.setdefault(d["class"], []).append(i)           # classes in order of first appearance
        merged = []
        for member_idx in groups.values():
            members = [detections[i] for i in member_idx]
            if len(members) == 1:
                merged.extend(members)
                continue
            if labels is not None:
                lab_c = np.asarray(labels)[member_idx]
                for lab in np.unique(lab_c):
                    idx = np.nonzero(lab_c == lab)[0]
                    merged.append(members[idx[0]] if len(idx) == 1 else self._merge_detection_cluster([members[i] for i in idx]))
                continue
            box = np.array([[m["bbox"]["x1"], m["bbox"]["y1"], m["bbox"]["x1"] + m["bbox"]["width"],
                             m["bbox"]["y1"] + m["bbox"]["height"]] for m in members], np.int64)
            # the reference intersects on (x1,y1,x2,y2) and takes areas from width*height: identical while
            # width == x2-x1, which every producer on this path guarantees; keep its exact form otherwise
            xyxy = np.array([[m["bbox"]["x1"], m["bbox"]["y1"], m["bbox"]["x2"], m["bbox"]["y2"]] for m in members], np.int64)
            lab_c = cluster_labels(iou_matrix(xyxy) if np.array_equal(box, xyxy) else self._iou_matrix_general(members))
            for lab in np.unique(lab_c):
                idx = np.nonzero(lab_c == lab)[0]
                merged.append(members[idx[0]] if len(idx) == 1 else self._merge_detection_cluster([members[i] for i in idx]))
        return merged

    @staticmethod
    def _iou_matrix_general(members: List[Dict]) -> np.ndarray:
        k = len(members)
        out = np.zeros((k, k))
        for i in range(k):
            a = members[i]["bbox"]
            for j in range(i + 1, k):
                b = members[j]["bbox"]
                x1, y1 = max(a["x1"], b["x1"]), max(a["y1"], b["y1"])
                x2, y2 = min(a["x2"], b["x2"]), min(a["y2"], b["y2"])
                if x2 <= x1 or y2 <= y1:
                    continue
                inter = (x2 - x1) * (y2 - y1)
                union = a["width"] * a["height"] + b["width"] * b["height"] - inter
                out[i, j] = out[j, i] = inter / union if union > 0 else 0.0
        return out

    @staticmethod
    def _merge_detection_cluster(members: List[Dict]) -> Dict:
        confs = [_conf_of(m) if ("ensemble_confidence" in m or "confidence" in m) else 0 for m in members]
        out = members[int(np.argmax(confs))].copy()               # first maximum, like max(..., key=)
        out["bbox"] = _bbox_dict(min(m["bbox"]["x1"] for m in members), min(m["bbox"]["y1"] for m in members),
                                 max(m["bbox"]["x2"] for m in members), max(m["bbox"]["y2"] for m in members))
        out["ensemble_confidence"] = np.mean(confs)
        out["confidence"] = np.mean(confs)
        sev = [m.get("final_severity", m.get("severity", "minor")) for m in members]
        out["final_severity"] = sev[int(np.argmax([SEVERITY_ORDER.get(s, 0) for s in sev]))]
        areas = [m.get("area_pixels", 0) for m in members if m.get("area_pixels", 0) > 0]
        if areas:
            out["area_pixels"] = sum(areas)
        out["merged_from"] = len(members)
        out["original_confidences"] = confs
        return out

    # -- :360-464 ------------------------------------------------------------------------------------
    def _validate_detections(self, detections: List[Dict], image_shape: Tuple) -> List[Dict]:
        if not detections:
            return []
        height, width = image_shape[:2]
        for d in detections:                                       # clamp in place, like the reference
            bb = d["bbox"]
            bb["x1"] = max(0, min(bb["x1"], width - 1)); bb["y1"] = max(0, min(bb["y1"], height - 1))
            bb["x2"] = max(bb["x1"] + 1, min(bb["x2"], width)); bb["y2"] = max(bb["y1"] + 1, min(bb["y2"], height))
            bb["width"] = bb["x2"] - bb["x1"]; bb["height"] = bb["y2"] - bb["y1"]
            bb["center_x"] = (bb["x1"] + bb["x2"]) / 2; bb["center_y"] = (bb["y1"] + bb["y2"]) / 2
            bbox_area = bb["width"] * bb["height"]
            if d.get("area_pixels", bbox_area) > bbox_area * 2:
                d["area_pixels"] = bbox_area
                d["area_consistency_warning"] = True
        score = validation_score([_conf_of(d) if ("ensemble_confidence" in d or "confidence" in d) else 0 for d in detections],
                                 [d["bbox"]["width"] for d in detections], [d["bbox"]["height"] for d in detections],
                                 [d["bbox"]["center_x"] for d in detections], [d["bbox"]["center_y"] for d in detections],
                                 [d.get("has_segmentation", False) for d in detections],
                                 [d.get("segmentation_confidence", 0) for d in detections], image_shape)
        out = []
        for d, s in zip(detections, score.tolist()):
            d["validation_score"] = s
            if s > 0.3:
                out.append(d)
        return out

    # -- :466-599 -------------------------------------------------------------------------------------
    GRADES = (("F", "FAIL", "high"), ("D", "FAIL", "high"), ("C", "CONDITIONAL", "medium"), ("B", "CONDITIONAL", "low"),
              ("A", "PASS", "low"))
    ACTIONS = {("PASS", "low"): "accept", ("CONDITIONAL", "low"): "accept_with_monitoring",
               ("CONDITIONAL", "medium"): "review_required", ("FAIL", "high"): "reject", ("FAIL", "medium"): "rework_required"}
    PENALTY = {"critical": 30, "major": 15, "minor": 5}

    def _assess_quality(self, detections: Sequence[Dict], segmentation_results: Dict) -> Dict:
        """Severity counts -> grade / pass-fail / risk, quality score and the per-image aggregates the SPC layer reads."""
        counts = {"critical": 0, "major": 0, "minor": 0}
        if not detections:
            return {"quality_grade": "A", "pass_fail_status": "PASS", "risk_level": "low", "total_defects": 0,
                    "severity_breakdown": counts, "total_defect_area_pixels": 0, "defect_density": 0.0,
                    "average_confidence": 1.0, "quality_score": 100.0, "meets_requirements": True,
                    "recommended_action": "accept"}
        for d in detections:
            counts[d.get("final_severity", "minor")] += 1
        t = self.quality_thresholds
        rule = [counts["critical"] > t.get("critical_defect_limit", 0), counts["major"] > t.get("major_defect_limit", 1),
                counts["minor"] > t.get("minor_defect_limit", 3), counts["major"] > 0, True]
        grade, pass_fail, risk = self.GRADES[rule.index(True)]
        avg_conf = np.mean([_conf_of(d) if ("ensemble_confidence" in d or "confidence" in d) else 0 for d in detections])
        base = 100.0
        for sev, c in counts.items():
            base -= c * self.PENALTY[sev]
        score = max(0, base * min(avg_conf / 0.8, 1.0))
        return {"quality_grade": grade, "pass_fail_status": pass_fail, "risk_level": risk, "total_defects": len(detections),
                "severity_breakdown": counts, "total_defect_area_pixels": sum(d.get("area_pixels", 0) for d in detections),
                "defect_density": (segmentation_results or {}).get("defect_density", 0.0), "average_confidence": avg_conf,
                "quality_score": score, "meets_requirements": pass_fail == "PASS",
                "recommended_action": self.ACTIONS.get((pass_fail, risk), "manual_inspection")}

    # -- :601-682 -------------------------------------------------------------------------------------
    @staticmethod
    def _analyze_defect_clustering(detections: Sequence[Dict]) -> Dict:
        """DBSCAN(eps=100, min_samples=2) on the box centres: with two samples per core point every point that has
        another one within 100 px is a core point, so the clusters are the connected components of that graph and
        the rest is noise.  (Squared distances against 100^2, like the KD-tree sklearn picks for 2-D points.)"""
        k = len(detections)
        if k < 3:
            return {"has_clusters": False, "cluster_count": 0}
        p = np.array([[d["bbox"]["center_x"], d["bbox"]["center_y"]] for d in detections], np.float64)
        d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(2)
        near = d2 <= 100.0 * 100.0
        np.fill_diagonal(near, False)
        has_nb = near.any(1)
        parent = np.arange(k)

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a
        for a, b in zip(*np.nonzero(np.triu(near, 1))):
            ra, rb = find(int(a)), find(int(b))
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        n_clusters = len({find(i) for i in range(k) if has_nb[i]})
        return {"has_clusters": n_clusters > 0, "cluster_count": n_clusters, "clustered_defects": np.sum(has_nb),
                "isolated_defects": np.sum(~has_nb)}

    def _analyze_risks(self, detections: Sequence[Dict], quality_assessment: Dict) -> Dict:
        """Risk score from clustering, large (> 1000 px) and critical defects.  The reference leaves its clustering
        variable unset for an empty list and fails there (process_results then returns its error dict); an empty
        list is reported here as no clusters."""
        factors, score = [], 0.0
        cluster = {"has_clusters": False, "cluster_count": 0}
        if len(detections) > 0:
            cluster = self._analyze_defect_clustering(detections)
            if cluster["has_clusters"]:
                factors.append("Clustered defects detected - possible systematic issue")
                score += 0.3
            n_large = sum(1 for d in detections if d.get("area_pixels", 0) > 1000)
            if n_large:
                factors.append(f"{n_large} large defects detected")
                score += 0.2 * n_large
            n_crit = sum(1 for d in detections if d.get("final_severity") == "critical")
            if n_crit:
                factors.append("Critical defects present - immediate attention required")
                score += 0.5 * n_crit
        level = "high" if score >= 1.0 else ("medium" if score >= 0.5 else "low")
        return {"overall_risk_level": level, "risk_score": min(score, 1.0), "risk_factors": factors,
                "defect_clustering": cluster, "requires_immediate_action": quality_assessment["pass_fail_status"] == "FAIL"}

    # -- :64-81 ---------------------------------------------------------------------------------------
    def process(self, raw_detections: Sequence[Dict], segmentation_results: Dict, image_shape: Tuple) -> List[Dict]:
        refined = self._refine_detections(raw_detections, segmentation_results, image_shape)
        return self._validate_detections(self._merge_overlapping_detections(self._filter_detections(refined)), image_shape)


# ---------------------------------------------------------------------------------------------------
class DeviceDetectionRules(DetectionRules):
    """The same chain with every number computed by K6 on the device (``qc_rules_apply``: refine, filter, cluster,
    merge, validate for a whole batch in one launch); the host only packs the detections of the batch into one table
    and turns the returned rows back into the reference's dictionaries.  Results are identical to
    ``DetectionRules.process`` -- and through it to the reference's PostProcessor (tests/test_gpu_rules.py)."""

    SEV_NAMES = {1: "minor", 2: "major", 3: "critical"}

    def __init__(self, config: Optional[Dict] = None, device: int = 0):
        super().__init__(config)
        self.device = device

    def process(self, raw_detections, segmentation_results, image_shape):
        return self.process_many([(raw_detections, segmentation_results, image_shape)])[0]

    def process_many(self, items) -> List[List[Dict]]:
        """items: [(raw_detections, segmentation_results, image_shape)] -> the validated detection list of each."""
        from . import ffi
        n = len(items)
        out: List[Optional[List[Dict]]] = [None] * n
        cap = max([len(it[0]) for it in items] + [1])
        rows = np.zeros((n, cap), ffi.RULES_ROW_DTYPE)
        rows["ext"] = -1
        count = np.zeros(n, np.int32)
        shapes = np.ones((n, 2), np.int32)
        class_ids: Dict[str, int] = {}
        prepared = []
        for a, (raw, seg, shape) in enumerate(items):
            regions = {}
            for r in (seg or {}).get("segmented_regions", []):
                regions.setdefault(r.get("detection_id"), r)
            ok = len(raw) > 0
            exts, tups = [], []
            no_ext = (-1, -1, -1, -1)
            for i, det in enumerate(raw):
                sev = SEVERITY_ORDER.get(det.get("severity", "minor"))
                cid = class_ids.setdefault(det["class"], len(class_ids))
                bb = det["bbox"]
                if sev is None or cid >= 64 or bb["width"] != bb["x2"] - bb["x1"] or bb["height"] != bb["y2"] - bb["y1"]:
                    ok = False                           # a table the kernel does not model: the host form below
                    break
                conf = float(_conf_of(det)) if ("ensemble_confidence" in det or "confidence" in det) else 0.0
                r = regions.get(i)
                ext = None
                if r:
                    ext = self.mask_extent(r.get("mask"))
                    # one tuple per row in RULES_ROW_DTYPE's field order (a single structured assignment per image)
                    tups.append((bb["x1"], bb["y1"], bb["x2"], bb["y2"], cid, sev, 1, 0,
                                 tuple(int(v) for v in ext) if ext is not None else no_ext, r.get("area_pixels", 0), conf,
                                 r.get("confidence_score", 0.0), r.get("area_percentage", 0.0), r.get("compactness", 0.0)))
                else:
                    tups.append((bb["x1"], bb["y1"], bb["x2"], bb["y2"], cid, sev, 0, 0, no_ext, 0, conf, 0.0, 0.0, 0.0))
                exts.append(ext)
            if not ok:
                out[a] = super().process(raw, seg, shape) if raw else []
                prepared.append(None)
                continue
            if tups:
                rows[a, :len(tups)] = np.array(tups, dtype=ffi.RULES_ROW_DTYPE)
            count[a] = len(raw)
            shapes[a] = shape[0], shape[1]
            prepared.append((regions, exts))
        if not any(p is not None for p in prepared):
            return out
        elong = sum(1 << cid for name, cid in class_ids.items() if name in ("crack", "scratch"))
        res = ffi.Context.current(self.device).rules_apply(
            rows, count, shapes, self.quality_thresholds.get("confidence_threshold", 0.5),
            self.quality_thresholds.get("area_threshold_percent", 50.0), elong, OVERLAP_THRESHOLD)
        for a, (raw, seg, shape) in enumerate(items):
            if prepared[a] is None:
                continue
            regions, exts = prepared[a]
            refined = []
            # plain tuples instead of numpy records: (x1, y1, x2, y2, severity, kept, area_pixels)
            ref_l = res["refined"][a, :len(raw)].tolist()
            for i, det in enumerate(raw):                       # the dictionaries of _refine_detections (:111-184)
                d = det.copy()
                t = ref_l[i]
                rf = {"x1": t[0], "y1": t[1], "x2": t[2], "y2": t[3], "severity": t[4], "area_pixels": t[6]}
                r = regions.get(i)
                if r:
                    d.update({"segmentation_confidence": r.get("confidence_score", 0.0), "area_pixels": r.get("area_pixels", 0),
                              "area_percentage": r.get("area_percentage", 0.0), "perimeter": r.get("perimeter", 0.0),
                              "compactness": r.get("compactness", 0.0), "contour_count": len(r.get("contours", [])),
                              "has_segmentation": True})
                    if exts[i] is not None:
                        d["bbox"] = _bbox_dict(int(rf["x1"]), int(rf["y1"]), int(rf["x2"]), int(rf["y2"]))
                else:
                    d.update({"segmentation_confidence": 0.0, "area_pixels": int(rf["area_pixels"]), "area_percentage": 0.0,
                              "perimeter": 0.0, "compactness": 0.0, "contour_count": 0, "has_segmentation": False})
                d["final_severity"] = self.SEV_NAMES[int(rf["severity"])]
                bb = d["bbox"]
                d.update({"aspect_ratio": bb["width"] / max(bb["height"], 1), "bbox_area": bb["width"] * bb["height"],
                          "normalized_position": {"x": bb["center_x"] / 1000, "y": bb["center_y"] / 1000}})
                refined.append(d)
            member = res["member"][a, :len(raw)]
            final = []
            keys = ("base", "merged_from", "x1", "y1", "x2", "y2", "severity", "flags", "area_pixels", "conf", "validation_score")
            for j, t in enumerate(res["rows"][a, :int(res["count"][a])].tolist()):
                o = dict(zip(keys, t))
                d = refined[int(o["base"])]
                if o["merged_from"] > 1:                        # _merge_detection_cluster (:305-358)
                    idx = np.nonzero(member == j)[0]
                    confs = [_conf_of(refined[i]) if ("ensemble_confidence" in refined[i] or "confidence" in refined[i]) else 0
                             for i in idx]
                    d = d.copy()
                    d["bbox"] = {}
                    d["ensemble_confidence"] = np.float64(o["conf"])
                    d["confidence"] = np.float64(o["conf"])
                    d["final_severity"] = self.SEV_NAMES[int(o["severity"])]
                    d["merged_from"] = int(o["merged_from"])
                    d["original_confidences"] = confs
                x1, y1, x2, y2 = int(o["x1"]), int(o["y1"]), int(o["x2"]), int(o["y2"])
                d["bbox"].update(_bbox_dict(x1, y1, x2, y2))    # _validate_detections clamps in place (:376-386)
                if o["merged_from"] > 1 or (o["flags"] & 1):
                    d["area_pixels"] = int(o["area_pixels"])
                if o["flags"] & 1:
                    d["area_consistency_warning"] = True
                d["validation_score"] = float(o["validation_score"])
                if o["flags"] & 2:
                    final.append(d)
            out[a] = final
        return out
