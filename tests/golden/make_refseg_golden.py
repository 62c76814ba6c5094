#!/usr/bin/env python
"""Golden vectors of the REFERENCE's own ``ImageSegmentator.segment_defects`` (inference/segmentation.py:39-146),
produced by importing /root/reference/inference/segmentation.py UNCHANGED on top of oracle/skimage_shim.py
(scikit-image and matplotlib are absent from the image).  Run in the build container only:

    python tests/golden/make_refseg_golden.py

Writes tests/golden/refseg_golden.json (per-region area / method / confidence / perimeter / compactness) and
tests/golden/refseg_masks.npz (the local masks, bit-packed).  The frames are regenerated from seeds at test time
(``refseg_cases.py``), so nothing but numbers is committed.

Classes: crack / scratch -> the reference's 'adaptive'; discoloration / unknown names -> 'threshold'.  dent ->
'watershed' and contamination -> 'region_growing' are out of scope (under the shim the watershed
method has no skimage.segmentation.watershed to call and returns an empty mask, an artefact of this container): their
rows come from calling the reference's ``_threshold_segmentation`` directly -- the method the reference itself falls
back to (:175-183) and what the drop-in runs for those two classes under the name 'threshold_fallback'.
"""
import importlib.util
import json
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))
sys.path.insert(0, ROOT)
sys.path.insert(0, OUT)


def main():
    from oracle import skimage_shim
    skimage_shim.install()
    spec = importlib.util.spec_from_file_location("ref_segmentation", os.path.join(REF, "inference/segmentation.py"))
    seg_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(seg_mod)
    import refseg_cases
    seg = seg_mod.ImageSegmentator({})
    golden, masks = {"frames": []}, {}
    nroi = 0
    for fi, (frame, dets) in enumerate(refseg_cases.cases()):
        growing = [d for d in dets if d["class"] in ("dent", "contamination")]
        dets_ref = [d for d in dets if d["class"] not in ("dent", "contamination")]
        sd = seg.segment_defects(frame, dets_ref)
        assert "error" not in sd, sd
        rows = []
        for r in sd["segmented_regions"]:
            d = dets_ref[r["detection_id"]]
            rows.append({"id": d["id"], "class": d["class"], "bbox": d["bbox"], "area_pixels": int(r["area_pixels"]),
                         "method": r["segmentation_method"], "confidence_score": float(r["confidence_score"]),
                         "perimeter": float(r["perimeter"]), "compactness": float(r["compactness"]),
                         "n_contours": len(r["contours"])})
            masks[f"f{fi}_d{d['id']}"] = np.packbits(r["local_mask"] > 0)
            nroi += 1
        import cv2
        gray = cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY) if frame.ndim == 3 else frame
        for d in growing:
            b = d["bbox"]
            roi = gray[b["y1"]:b["y2"], b["x1"]:b["x2"]]
            if roi.size == 0:
                continue
            m, conf = seg._threshold_segmentation(roi, d["class"])
            cs = seg._find_contours(m)
            area = int(np.sum(m > 0))
            rows.append({"id": d["id"], "class": d["class"], "bbox": b, "area_pixels": area,
                         "method": "threshold_fallback", "confidence_score": float(conf),
                         "perimeter": float(seg._calculate_perimeter(cs)),
                         "compactness": float(seg._calculate_compactness(area, cs)), "n_contours": len(cs)})
            masks[f"f{fi}_d{d['id']}"] = np.packbits(m > 0)
            nroi += 1
        rows.sort(key=lambda r: r["id"])
        golden["frames"].append({"regions": rows, "total_defect_area": int(sd["total_defect_area"]),
                                 "n_reference_regions": len(sd["segmented_regions"])})
    golden["n_rois"] = nroi
    with open(os.path.join(OUT, "refseg_golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "refseg_masks.npz"), **masks)
    print("wrote", nroi, "regions;", {m: sum(r["method"] == m for fr in golden["frames"] for r in fr["regions"])
                                      for m in ("adaptive", "threshold", "threshold_fallback")})


if __name__ == "__main__":
    main()
