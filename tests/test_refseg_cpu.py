"""CPU pins of oracle/refseg.py (the reference's per-class segmentation methods, inference/segmentation.py:185-298,
449-484): restated forms == library forms bit for bit, and both == the golden vectors recorded from the reference's own
file (tests/golden/make_refseg_golden.py)."""
import json
import os
import sys

import cv2
import numpy as np
import pytest

from oracle import refseg as R, skimage_shim as sk

G = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, G)
import refseg_cases  # noqa: E402

GOLD = json.load(open(os.path.join(G, "refseg_golden.json")))
MASKS = np.load(os.path.join(G, "refseg_masks.npz"))


def _rand_roi(rng, t, lo=1, hi=90):
    h, w = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
    kind = t % 4
    if kind == 0:
        return rng.integers(0, 256, (h, w)).astype(np.uint8)
    if kind == 1:
        return np.clip(rng.normal(128, 20, (h, w)), 0, 255).astype(np.uint8)
    if kind == 2:
        return np.where(rng.random((h, w)) < 0.3, rng.integers(0, 100), rng.integers(150, 256)).astype(np.uint8)
    return np.full((h, w), rng.integers(0, 256), np.uint8)


def test_gaussian_weights_are_scipys():
    from scipy.ndimage._filters import _gaussian_kernel1d
    w = _gaussian_kernel1d(1.0, 0, 4)
    assert tuple(w[4:]) == R.GAUSS_W and np.array_equal(w, w[::-1])
    assert R.U8_TO_FLOAT[255] == 255 * (1.0 / 255) and R.U8_TO_FLOAT[3] == 3 * (1.0 / 255)


def test_blur_histogram_otsu_restated_bit_exact():
    rng = np.random.default_rng(0)
    for t in range(120):
        g = _rand_roi(rng, t)
        a, b = R.gaussian_sigma1(g), R.gaussian_sigma1_restated(g)
        assert np.array_equal(a, b), t
        if a.min() != a.max():
            c1, e1 = np.histogram(a.reshape(-1), 256)
            c2, e2 = R.histogram256_restated(a)
            assert np.array_equal(c1, c2) and np.array_equal(e1, e2), t
        assert sk.threshold_otsu(a) == R.otsu_float_restated(a), t


def test_threshold_method_restated_equals_library():
    rng = np.random.default_rng(1)
    for t in range(60):
        g = _rand_roi(rng, t)
        for cls in ("discoloration", "dent", "contamination", "crack", "something_else"):
            m1, c1, _ = R.threshold_segmentation(g, cls, False)
            m2, c2, _ = R.threshold_segmentation(g, cls, True)
            assert np.array_equal(m1, m2) and c1 == c2, (t, cls)


def test_adaptive_method_restated_equals_cv2():
    """The float32 blur model (fused row / column passes, cv2's three column groups) reproduces cv2's uint8 mean --
    and therefore the mask -- on every pixel of these cases; a mismatch budget of 0 is asserted and the count printed."""
    rng = np.random.default_rng(2)
    tot = bad = 0
    for t in range(80):
        g = _rand_roi(rng, t, 3, 260)
        h, w = g.shape
        big = rng.integers(0, 256, (h + 5, w + 6)).astype(np.uint8)      # the reference passes a strided view (:80)
        big[2:2 + h, 3:3 + w] = g
        roi = big[2:2 + h, 3:3 + w]
        for cls in ("crack", "anything"):
            m1, c1, a1 = R.adaptive_segmentation(roi, cls, False)
            m2, c2, a2 = R.adaptive_segmentation(roi, cls, True)
            tot += h * w
            bad += int((a1["raw"] != a2["raw"]).sum())
            assert np.array_equal(m1, m2), (t, cls)
    print(f"adaptive threshold: {bad} of {tot} pre-morphology pixels differ from cv2")
    assert bad == 0


def test_adaptive_block_size_rule():
    assert R.adaptive_block_size(120, 120) == 13 and R.adaptive_block_size(100, 300) == 11
    assert R.adaptive_block_size(320, 320) == 33 and R.adaptive_block_size(5, 5) == 11
    assert R.adaptive_block_size(640, 130) == 13 and R.adaptive_block_size(140, 999) == 15


@pytest.mark.parametrize("restated", [False, True])
def test_oracle_reproduces_reference_segment_defects(restated):
    """Golden = the reference's own ImageSegmentator.segment_defects (through the skimage shim) on 32 ROIs."""
    n = 0
    for fi, (frame, dets) in enumerate(refseg_cases.cases()):
        got = {r["detection_id"]: r for r in R.segment_defects(frame, dets, restated)}
        gold = GOLD["frames"][fi]["regions"]
        assert sorted(got) == [g["id"] for g in gold]
        for g in gold:
            r = got[g["id"]]
            b = g["bbox"]
            h = min(b["y2"], frame.shape[0]) - b["y1"]
            w = min(b["x2"], frame.shape[1]) - b["x1"]
            gm = np.unpackbits(MASKS[f"f{fi}_d{g['id']}"])[:h * w].reshape(h, w)
            assert np.array_equal(r["mask"] > 0, gm > 0), (fi, g["id"])
            assert r["area_pixels"] == g["area_pixels"] and r["method"] == g["method"], (fi, g["id"])
            assert abs(r["confidence_score"] - g["confidence_score"]) <= 1e-12, (fi, g["id"])
            assert abs(r["perimeter"] - g["perimeter"]) <= 1e-9 * max(1.0, g["perimeter"])
            assert abs(r["compactness"] - g["compactness"]) <= 1e-12
            assert r["n_contours"] == g["n_contours"]
            n += 1
    assert n == GOLD["n_rois"] >= 20


def test_contour_statistics_restated_equal_cv2():
    """oracle/contours.py: perimeter / contour count / contourArea from 2x2 window moves of the hole-filled mask ==
    cv2.findContours(RETR_EXTERNAL) + contourArea + arcLength (inference/segmentation.py:486-525)."""
    from oracle import contours as C
    rng = np.random.default_rng(3)
    for t in range(160):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        kind = t % 4
        if kind == 0:
            m = rng.random((h, w)) < rng.uniform(0.05, 0.9)
        elif kind == 1:
            m = np.zeros((h, w), bool)
            for _ in range(8):
                y, x = rng.integers(0, h), rng.integers(0, w)
                m[y:y + rng.integers(1, 20), x:x + rng.integers(1, 20)] = True
            for _ in range(4):
                y, x = rng.integers(0, h), rng.integers(0, w)
                m[y:y + rng.integers(1, 8), x:x + rng.integers(1, 8)] = False
        elif kind == 2:
            yy, xx = np.mgrid[0:h, 0:w]
            m = (yy + xx) % int(rng.integers(2, 6)) == 0
        else:
            m = cv2.morphologyEx((rng.random((h, w)) < 0.5).astype(np.uint8), cv2.MORPH_OPEN, np.ones((3, 3), np.uint8)) > 0
        a, b = C.contour_stats(m), C.contour_stats_restated(m)
        assert a[1] == b[1] and a[2] == b[2], t
        assert abs(a[0] - b[0]) <= 1e-6 * max(1.0, a[0]), t


def test_canny_restated_equals_cv2():
    """oracle/canny.py == cv2.Canny(roi, 50, 150) bit for bit, on strided ROI views like the reference passes."""
    from oracle import canny as K
    rng = np.random.default_rng(4)
    for t in range(90):
        g = _rand_roi(rng, t, 1, 150)
        h, w = g.shape
        big = rng.integers(0, 256, (h + 5, w + 6)).astype(np.uint8)
        big[2:2 + h, 3:3 + w] = g
        roi = big[2:2 + h, 3:3 + w]
        assert np.array_equal(cv2.Canny(roi, 50, 150), K.canny_restated(np.ascontiguousarray(roi))), t
        m, conf, _ = R.adaptive_segmentation(roi, "crack")
        assert K.adaptive_confidence_restated(np.ascontiguousarray(roi), m) == conf
