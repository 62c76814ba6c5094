#!/usr/bin/env python
"""Generates tests/golden/rules_golden.json.gz by running the REFERENCE's own PostProcessor
(/root/reference/inference/postprocess.py, imported unchanged -- it only needs numpy, cv2 and scikit-learn) on
seeded detection lists: the outputs of _refine_detections, _filter_detections, _merge_overlapping_detections and
_validate_detections, snapshotted stage by stage (the reference mutates its dicts in place).
Run in the build container only; the fixture is committed, the reference never travels.

    python tests/golden/make_rules_golden.py
"""
import importlib.util
import json
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
from rules_cases import build_case, jsonable, CASES  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_postprocess", os.path.join(REF, "inference/postprocess.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    golden = {}
    for name, kw in CASES.items():
        dets, seg, shape, cfg = build_case(**kw)
        pp = mod.PostProcessor(cfg)
        refined = pp._refine_detections(dets, seg, shape)
        snap = {"refined": jsonable(refined)}
        filtered = pp._filter_detections(refined)
        snap["filtered"] = jsonable(filtered)
        merged = pp._merge_overlapping_detections(filtered)
        snap["merged"] = jsonable(merged)
        validated = pp._validate_detections(merged, shape)
        snap["validated"] = jsonable(validated)
        seg_q = dict(seg, defect_density=1.25)
        snap["quality"] = jsonable(pp._assess_quality(validated, seg_q))
        snap["quality_refined"] = jsonable(pp._assess_quality(refined, seg_q))      # more rows: other grades
        if validated:
            snap["risk"] = jsonable(pp._analyze_risks(validated, pp._assess_quality(validated, seg_q)))
        if refined:
            snap["risk_refined"] = jsonable(pp._analyze_risks(refined, pp._assess_quality(refined, seg_q)))
        golden[name] = snap
        print(name, len(dets), "->", len(refined), len(filtered), len(merged), len(validated))
    import gzip
    with gzip.open(os.path.join(OUT, "rules_golden.json.gz"), "wt", compresslevel=9) as f:
        json.dump(golden, f)


if __name__ == "__main__":
    main()
