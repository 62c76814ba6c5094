"""GPU tests of the reference-facing layer: drop-in classes, batched pipeline and the SPC kernel, against
the golden fixtures (reference's own code) and the CPU oracle pipeline."""
import json
import os

import numpy as np
import pytest
import torch

from gpu_common import qc  # noqa: F401
from oracle import yolo_head as oyh, segment as oseg

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
GOLD = json.load(open(os.path.join(G, "reference_golden.json")))
CFG = {"model": {"confidence_threshold": 0.7, "nms_threshold": 0.5},
       "quality_control": {"defect_classes": ["crack", "scratch", "dent", "discoloration", "contamination"]},
       "qcgraft": {"yolo_scale": "n", "device": 0}}


def test_spc_kernel_matches_reference_extract_metrics(qc):
    rows = np.array(GOLD["parse_results"]["rows"], np.float32)
    counts = GOLD["spc"]["counts"]
    det = np.zeros((len(counts), 16, 6), np.float32)
    for i, c in enumerate(counts):
        det[i, :c] = rows[:c]
    v = qc.spc_stats(torch.from_numpy(det).cuda(), torch.tensor(counts, dtype=torch.int32).cuda()).cpu().numpy()
    assert np.allclose(v, np.array(GOLD["spc"]["vector"], np.float64), rtol=1e-12, atol=0)


def test_segmentator_matches_reference_chain():
    """ImageSegmentator.segment_defects on the fixture's binary image == the reference's own
    morphologyEx CLOSE/OPEN + _remove_small_components + contour statistics."""
    from qcgraft.detector import ImageSegmentator
    z = np.load(os.path.join(G, "segmentation_chain.npz"))
    m0 = z["m0"]
    h, w = m0.shape
    seg = ImageSegmentator(CFG)
    det = [{"id": 0, "class": "dent", "confidence": 0.9, "bbox": {"x1": 0, "y1": 0, "x2": w, "y2": h}}]
    out = seg.segment_defects(m0, det)
    assert "error" not in out
    r = out["segmented_regions"][0]
    g = GOLD["segmentation_chain"]
    assert np.array_equal(r["local_mask"], z["m2"])
    assert r["area_pixels"] == g["area_pixels"] and len(r["contours"]) == g["n_contours"]
    assert abs(r["perimeter"] - g["perimeter"]) < 1e-6 and abs(r["compactness"] - g["compactness"]) < 1e-9
    assert np.array_equal(np.asarray(r["mask"]), z["m2"])                 # lazy full-frame mask
    assert out["total_defect_area"] == g["area_pixels"]
    assert seg.segment_defects(m0, [])["segmented_regions"] == []
    # a box partly outside the frame is clipped like numpy slicing; an empty box is skipped
    det2 = [{"id": 0, "class": "crack", "confidence": 0.9, "bbox": {"x1": 100, "y1": 50, "x2": 400, "y2": 300}},
            {"id": 1, "class": "crack", "confidence": 0.9, "bbox": {"x1": 10, "y1": 10, "x2": 10, "y2": 40}}]
    out2 = seg.segment_defects(m0, det2)
    assert [r["detection_id"] for r in out2["segmented_regions"]] == [0]
    ref = oseg.segment_k4(np.ascontiguousarray(m0[50:300, 100:400]), 0, 1.0, 1, 20)
    assert np.array_equal(out2["segmented_regions"][0]["local_mask"], ref["mask"])


def test_yolo_detector_dropin_matches_oracle():
    from qcgraft import synth
    from qcgraft.detector import YOLODetector
    img = synth.make_frame(5, 480, 640)
    levels = synth.make_head(21, 1, 640, 640, 25)
    y = YOLODetector("yolov8n.pt", 0.7, 0.5, config=CFG)
    out = y.predict(img, _head_override=[l.cuda() for l in levels])
    assert "error" not in out and out["image_shape"] == (480, 640)
    dets = oyh.yolo_postprocess(levels, 5, (640, 640), (480, 640))
    ref = oyh.parse_results(dets[0].numpy())
    assert len(out["detections"]) == len(ref) == out["total_detections"]
    for a, b in zip(out["detections"], ref):
        assert a["class"] == b["class"] and a["severity"] == b["severity"] and a["id"] == b["id"]
        assert abs(a["confidence"] - b["confidence"]) <= 1e-3
        for k in ("x1", "y1", "x2", "y2"):
            assert abs(a["bbox"][k] - b["bbox"][k]) <= 1       # int() of coords that agree to 1e-3
    y.update_thresholds(confidence=0.9)
    out9 = y.predict(img, _head_override=[l.cuda() for l in levels])
    assert all(d["confidence"] > 0.9 for d in out9["detections"]) and len(out9["detections"]) < len(ref)
    assert y.get_model_info()["confidence_threshold"] == 0.9


def test_quality_control_detector_end_to_end_and_error_convention():
    from qcgraft import synth
    from qcgraft.detector import QualityControlDetector
    d = QualityControlDetector("y.pt", "r.pt", dict(CFG))
    img = synth.make_frame(6)
    levels = [l.cuda() for l in synth.make_head(22, 1, 640, 640, 12)]
    out = d.predict(img, _head_override=levels)
    assert "error" not in out
    for k in ("detections", "quality_assessment", "risk_analysis", "recommendations", "metadata",
              "total_inference_time_ms", "processing_pipeline", "image_metadata"):
        assert k in out
    assert out["detections"] and all("ensemble_confidence" in x and "final_severity" in x for x in out["detections"])
    x0 = out["detections"][0]
    assert abs(x0["ensemble_confidence"] - (0.6 * x0["yolo_confidence"] + 0.4 * min(x0["yolo_confidence"] * 1.1, 1.0))) < 1e-12
    assert "roi_classification" in x0["classification_details"]
    assert abs(sum(x0["classification_details"]["roi_classification"]["class_probabilities"].values()) - 1) < 1e-3
    assert "segmentation" in x0 and x0["segmentation"]["area_pixels"] >= 0
    assert d.get_performance_stats()["total_predictions"] == 1
    # error convention: never raises
    assert d.predict(None) == {"error": "Invalid image input"}
    assert d.predict(np.zeros((2, 2, 2, 2), np.uint8)) == {"error": "Invalid image input"}
    res = d.predict_batch([img, None])
    assert res[0]["batch_index"] == 0 and "error" in res[1] and "batch_statistics" in res[0]
    bm = d.benchmark([img], iterations=2)          # random-init net: no detections, still a full pass
    assert bm["total_images"] == 2 and bm["throughput"]["images_per_second"] > 0
    ok, buf = __import__("cv2").imencode(".png", img)
    assert "error" not in d.predict(np.frombuffer(buf.tobytes(), np.uint8))     # 1-D encoded upload (app.py:190-194)


@pytest.mark.parametrize("graphs", [False, True])
def test_batched_pipeline_matches_cpu_reference_pipeline(graphs):
    """InspectionPipeline.run vs the oracle's per-image chain on identical inputs: kept boxes, classes,
    ROI tables, segmentation areas/thresholds identical; ROI class probabilities close (bf16 nets)."""
    from qcgraft import nets, synth
    from qcgraft.pipeline import InspectionPipeline
    from oracle.pipeline import CpuReferencePipeline
    n = 3
    frames = synth.make_frames(700, n)
    heads = synth.make_head(701, n, 640, 640, 15)
    pipe = InspectionPipeline("n", 5, device=0, use_graphs=graphs, net_dtype=torch.float32)
    res = pipe.run(torch.from_numpy(frames).cuda(), head_override=[l.cuda() for l in heads], keep_seg_tensors=True)
    y = nets.build_yolo("n", 5, 0)
    cp = CpuReferencePipeline(y, nets.build_resnet50(5, 0))
    k0 = 0
    for i in range(n):
        ref = cp.predict(frames[i], head_override=[l[i:i + 1] for l in heads])
        c = int(res.count[i])
        assert c == len(ref["detections"])
        for j, d in enumerate(ref["detections"]):
            row = res.det[i, j]
            assert int(row[5]) == oyh.CLASS_NAMES.index(d["class"]) and abs(row[4] - d["confidence"]) <= 1e-3
            assert [int(v) for v in row[:4]] == [d["bbox"][k] for k in ("x1", "y1", "x2", "y2")]
        sel = np.nonzero(res.roi_img == i)[0]
        assert len(sel) == len(ref["segmented_regions"])
        for kk, r in zip(sel, ref["segmented_regions"]):
            assert int(res.roi_det[kk]) == r["detection_id"]
            assert int(res.seg_area[kk]) == r["area_pixels"], (i, kk)
            assert int(res.seg_thr[kk]) == r["thr"] and int(res.seg_ncomp[kk]) == r["ncomp"]
            p_ref = ref["detections"][r["detection_id"]]["roi_class_probabilities"]
            assert np.abs(res.roi_probs[kk] - p_ref).max() < 2e-2
        k0 += len(sel)
        # whole-frame classifier: K3b feeds the net the reference's exact PIL/ToTensor/Normalize tensor (bit-exact in
        # tests/test_gpu_k3b_resize.py); the probabilities differ by the GPU conv arithmetic (TF32) only
        gref = np.array([ref["global_classification"]["class_probabilities"][c] for c in oyh.CLASS_NAMES])
        assert np.abs(res.global_probs[i] - gref).max() < 2e-2
    assert res.spc[0] == n and res.spc[1] == res.count.sum()


def test_pipeline_host_frames_and_stats_only_area():
    """The public call with pinned HOST frames (H2D on the pipeline's copy stream, two staging buffers) and without
    pixel outputs (kept area from the stats rows) returns what the resident / full-output form returns."""
    from qcgraft import synth
    from qcgraft.pipeline import InspectionPipeline
    pipe = InspectionPipeline("n", 5, device=0, use_graphs=False, net_dtype=torch.float32)
    batches = [synth.make_frames(710 + 10 * b, 4) for b in range(3)]
    heads = [synth.make_head(720 + b, 4, 640, 640, 25) for b in range(3)]
    full = [pipe.run(torch.from_numpy(f).cuda(), head_override=[l.cuda() for l in h], keep_seg_tensors=True)
            for f, h in zip(batches, heads)]
    pend = [pipe.run_async(torch.from_numpy(f).pin_memory(), head_override=[l.cuda() for l in h])
            for f, h in zip(batches, heads)]                     # three batches in flight over two staging buffers
    for a, p in zip(full, pend):
        b = p.result()
        assert np.array_equal(a.det, b.det) and np.array_equal(a.count, b.count)
        assert np.array_equal(a.seg_area, b.seg_area) and np.array_equal(a.seg_thr, b.seg_thr)
        assert np.array_equal(a.seg_ncomp, b.seg_ncomp) and np.allclose(a.roi_probs, b.roi_probs, atol=1e-6)
        assert np.array_equal(a.spc, b.spc)
