"""K7 parity (through the C ABI): the reference's own per-class segmentation methods on device
(inference/segmentation.py:197-298,449-484) vs oracle/refseg.py, and the drop-in ImageSegmentator.segment_defects vs the
golden vectors recorded from the reference's own file (tests/golden/make_refseg_golden.py).

Bars: masks, areas, extents, Otsu bins, labels: bit-exact.  The threshold method's float64 blur / histogram / Otsu are
restated operation by operation, so exact.  The adaptive method follows cv2's fused float32 blur; its mismatch budget is
stated in each test (0 pixels on every case here) and the count is printed.  Confidences are float64 scalars computed
from integer sums: <= 1e-9."""
import json
import os
import sys

import cv2
import numpy as np
import pytest
import torch

from gpu_common import qc  # noqa: F401
from oracle import refseg as R, segment as oseg, contours as ocont
from qcgraft import ffi, synth

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, G)
import refseg_cases  # noqa: E402

CLASSES = ["crack", "scratch", "dent", "discoloration", "contamination", "something_else"]


def _rand_rois(rng, n, H, W, wmax=200, hmax=200):
    regs = []
    for i in range(n):
        w, h = int(rng.integers(1, min(wmax, W) + 1)), int(rng.integers(1, min(hmax, H) + 1))
        regs.append(dict(img=0, x=int(rng.integers(0, W - w + 1)), y=int(rng.integers(0, H - h + 1)), w=w, h=h))
    return regs


def _frame(seed, H, W, gray=True):
    rng = np.random.default_rng(seed)
    f = synth.make_frame(seed, H, W)
    for _ in range(10):           # flat patches: exactly constant ROIs / plateaus in the blurred image
        y, x = int(rng.integers(0, H - 10)), int(rng.integers(0, W - 10))
        f[y:y + int(rng.integers(4, 40)), x:x + int(rng.integers(4, 40))] = int(rng.integers(0, 256))
    return cv2.cvtColor(f, cv2.COLOR_RGB2GRAY) if gray else f


def _check_contours(res, i, mask):
    """K8 (n_axis, n_diag, n_contours) of region i vs cv2.findContours / contourArea / arcLength on the same mask
    (inference/segmentation.py:486-525): contour count exact, perimeter to the float32 rounding of cv2's segments."""
    per, n, _areas = ocont.contour_stats(mask)
    n_axis, n_diag, n_cont, _ = (int(v) for v in res["contours"][i].cpu().numpy())
    assert n_cont == n, f"region {i}: {n_cont} contours vs cv2 {n}"
    got = n_axis + np.sqrt(2.0) * n_diag
    assert abs(got - per) <= 1e-6 * max(1.0, per), f"region {i}: perimeter {got} vs cv2 {per}"


def _run_threshold(qc, img, regs, classes):
    rr = []
    for r, c in zip(regs, classes):
        rr.append(dict(r, polarity=0 if c in R.DARK_CLASSES else 1, thr_scale=R.THRESHOLD_FACTORS.get(c, 1.0),
                       min_area=10, max_comp=2))
    res = qc.segment(torch.from_numpy(img[None]).cuda(), rr, morph_iters=2, method=ffi.QC_SEG_REF_THRESHOLD,
                     want_sums=True, want_contours=True)
    torch.cuda.synchronize()
    return res


def _check_threshold(res, i, img, reg, cls):
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) if img.ndim == 3 else img
    roi = gray[reg["y"]:reg["y"] + reg["h"], reg["x"]:reg["x"] + reg["w"]]
    m, conf, aux = R.threshold_segmentation(roi, cls, restated=True)
    o = int(res["regions"][i].out_offset)
    w, h = reg["w"], reg["h"]
    got = res["mask"][o:o + w * h].view(h, w).cpu().numpy()
    assert np.array_equal(got, m), f"region {i} {cls} {w}x{h}: {int((got != m).sum())} mask pixels differ"
    area = int((m > 0).sum())
    assert int(res["area"][i]) == area
    assert int(res["ncomp"][i]) == (2 if area else 1)
    lab = res["labels"][o:o + w * h].view(h, w).cpu().numpy()
    assert np.array_equal(lab, (m > 0).astype(np.int32))
    st = res["stats"][int(res["regions"][i].stats_offset) + 1].cpu().numpy()
    if area:
        ys, xs = np.nonzero(m)
        assert list(st) == [xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1, area]
    else:
        assert list(st) == [0, 0, 0, 0, 0]
    s = res["sums"][i].cpu().numpy()
    r64 = roi.astype(np.int64)
    assert list(s) == [int(r64[m > 0].sum()), int(r64.sum()), int((r64 * r64).sum()), roi.size]
    from qcgraft.detector import ImageSegmentator
    assert abs(ImageSegmentator._threshold_confidence(s, area) - conf) <= 1e-9
    _check_contours(res, i, m)


def test_threshold_method_random_rois(qc):
    img = _frame(1, 480, 640)
    rng = np.random.default_rng(10)
    regs = _rand_rois(rng, 120, 480, 640)
    classes = [CLASSES[i % len(CLASSES)] for i in range(len(regs))]
    res = _run_threshold(qc, img, regs, classes)
    for i, (r, c) in enumerate(zip(regs, classes)):
        _check_threshold(res, i, img, r, c)


def test_threshold_method_rgb_input_and_shapes(qc):
    """RGB frames (gray computed on device), ragged shapes: 1-pixel strips, widths around the 32-bit word borders,
    a whole frame."""
    img = _frame(2, 300, 333, gray=False)
    regs = [dict(img=0, x=0, y=0, w=333, h=300), dict(img=0, x=5, y=7, w=1, h=1), dict(img=0, x=0, y=0, w=333, h=1),
            dict(img=0, x=332, y=0, w=1, h=300), dict(img=0, x=1, y=1, w=32, h=40), dict(img=0, x=3, y=9, w=33, h=17),
            dict(img=0, x=2, y=2, w=31, h=64), dict(img=0, x=10, y=20, w=64, h=3), dict(img=0, x=100, y=100, w=3, h=3),
            dict(img=0, x=0, y=250, w=97, h=50), dict(img=0, x=200, y=0, w=133, h=130)]
    classes = ["discoloration", "dent", "crack", "contamination", "discoloration", "x", "scratch", "dent",
               "discoloration", "contamination", "discoloration"]
    res = _run_threshold(qc, img, regs, classes)
    for i, (r, c) in enumerate(zip(regs, classes)):
        _check_threshold(res, i, img, r, c)


def test_threshold_method_fill_holes_patterns(qc):
    """Rings, nested rings, spirals and a maze: holes that need long, winding 4-connected background paths."""
    H = W = 160
    img = np.full((H, W), 60, np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    d = np.maximum(np.abs(yy - 80), np.abs(xx - 80))
    for rad in range(10, 75, 9):
        img[(d >= rad) & (d < rad + 3)] = 220                 # nested square rings (every gap but the outside is a hole)
    img[80, 80:] = 60                                         # ... cut open on one side: a spiral-like open background
    img2 = np.full((H, W), 60, np.uint8)
    for k in range(4, 150, 8):
        img2[k:k + 3, 4 + (k // 8 % 2) * 12:W - 16 + (k // 8 % 2) * 12] = 220      # serpentine walls
    img2[:, 0:3] = 220; img2[:, W - 3:] = 220
    img3 = np.full((H, W), 60, np.uint8)
    img3[(xx + yy) % 7 == 0] = 220                            # diagonals: 8-connected foreground, 4-connected background blocked
    for im in (img, img2, img3):
        regs = [dict(img=0, x=0, y=0, w=W, h=H), dict(img=0, x=3, y=5, w=150, h=149), dict(img=0, x=40, y=40, w=80, h=81)]
        res = _run_threshold(qc, im, regs, ["discoloration"] * 3)
        for i, r in enumerate(regs):
            _check_threshold(res, i, im, r, "discoloration")


def test_threshold_method_large_regions(qc):
    """A 640x640 and a 1200x1100 ROI (wider than one warp of words: the multi-chunk flood path)."""
    img = synth.make_linescan(7, 1200, 1100)
    regs = [dict(img=0, x=0, y=0, w=1100, h=1200), dict(img=0, x=100, y=200, w=640, h=640)]
    res = _run_threshold(qc, img, regs, ["discoloration", "x"])
    for i, r in enumerate(regs):
        _check_threshold(res, i, img, r, ["discoloration", "x"][i])


def test_threshold_method_noise_rois_many_background_runs(qc):
    """White-noise ROIs: tens of thousands of background runs with maze-like 4-connected paths.  The small ones go
    through the run union-find of the flood, the 420x430 one has more runs than the CTA's node table holds and takes
    the sweep fallback."""
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (480, 480)).astype(np.uint8)
    regs = [dict(img=0, x=10, y=20, w=420, h=430), dict(img=0, x=0, y=0, w=200, h=190), dict(img=0, x=100, y=50, w=97, h=301),
            dict(img=0, x=300, y=300, w=33, h=31), dict(img=0, x=5, y=400, w=470, h=70)]
    cls = ["discoloration", "x", "dent", "discoloration", "x"]
    res = _run_threshold(qc, img, regs, cls)
    for i, r in enumerate(regs):
        _check_threshold(res, i, img, r, cls[i])


def _run_adaptive(qc, img, regs, classes, max_comp=None):
    rr = [dict(r, polarity=0 if c in R.DARK_CLASSES else 1, min_area=20) for r, c in zip(regs, classes)]
    res = qc.segment(torch.from_numpy(img[None]).cuda(), rr, morph_iters=1, method=ffi.QC_SEG_REF_ADAPTIVE,
                     want_contours=True, want_sums=True)
    torch.cuda.synchronize()
    return res


def _check_adaptive(res, i, img, reg, cls, budget=0):
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) if img.ndim == 3 else img
    roi = gray[reg["y"]:reg["y"] + reg["h"], reg["x"]:reg["x"] + reg["w"]]
    m, conf, aux = R.adaptive_segmentation(roi, cls, restated=False)            # cv2 itself
    o = int(res["regions"][i].out_offset)
    w, h = reg["w"], reg["h"]
    got = res["mask"][o:o + w * h].view(h, w).cpu().numpy()
    bad = int((got != m).sum())
    assert bad <= budget, f"region {i} {cls} {w}x{h} block {aux['block']}: {bad} mask pixels differ from cv2"
    assert int(res["thr"][i]) == aux["block"]
    if bad == 0:
        assert int(res["area"][i]) == int((m > 0).sum())
        pre = cv2.morphologyEx(cv2.morphologyEx(aux["raw"], cv2.MORPH_CLOSE, oseg.CROSS), cv2.MORPH_OPEN, oseg.CROSS)
        nc, lab, stats, _ = cv2.connectedComponentsWithStats(pre, connectivity=8, ltype=cv2.CV_32S)
        assert int(res["ncomp"][i]) == nc
        keep = stats[:, cv2.CC_STAT_AREA] >= 20
        keep[0] = False
        labf = np.where(keep[lab], lab, 0)
        assert np.array_equal(res["labels"][o:o + w * h].view(h, w).cpu().numpy(), labf)
        _check_contours(res, i, m)
        # K9: cv2.Canny(roi, 50, 150) and its overlap with the mask, bit for bit; the confidence of :557-578
        edges = cv2.Canny(np.ascontiguousarray(roi), 50, 150)
        sm = [int(v) for v in res["sums"][i].cpu().numpy()]
        assert sm == [int(((edges > 0) & (m > 0)).sum()), int((edges > 0).sum()), 0, roi.size], f"region {i}: Canny counts {sm}"
        from qcgraft.detector import ImageSegmentator
        assert abs(ImageSegmentator._adaptive_confidence_from_counts(sm, int((m > 0).sum())) - conf) <= 1e-12
    return bad


def test_adaptive_method_random_rois(qc):
    """Mismatch budget vs cv2.adaptiveThreshold + morphology + CC filter: 0 pixels (count printed)."""
    img = _frame(3, 480, 640)
    rng = np.random.default_rng(11)
    regs = _rand_rois(rng, 120, 480, 640, 260, 260)
    classes = [("crack", "scratch", "anything")[i % 3] for i in range(len(regs))]
    res = _run_adaptive(qc, img, regs, classes)
    bad = sum(_check_adaptive(res, i, img, r, c) for i, (r, c) in enumerate(zip(regs, classes)))
    print(f"adaptive: {bad} mismatching mask pixels over {sum(r['w'] * r['h'] for r in regs)}")


def test_adaptive_method_widths_and_blocks(qc):
    """Every width residue mod 8 (cv2's vector / 4-group / scalar column groups), block sizes 11 .. 65, RGB input,
    ROIs smaller than the block, noise frames (many values near the rounding boundaries)."""
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (660, 700, 3), dtype=np.uint8)
    regs = [dict(img=0, x=3 + k, y=2 * k, w=40 + k, h=30 + 3 * k) for k in range(16)]
    regs += [dict(img=0, x=0, y=0, w=700, h=660), dict(img=0, x=10, y=10, w=5, h=5), dict(img=0, x=1, y=1, w=3, h=200),
             dict(img=0, x=0, y=300, w=200, h=4), dict(img=0, x=17, y=40, w=333, h=331), dict(img=0, x=0, y=0, w=1, h=1)]
    classes = [("crack", "other")[i % 2] for i in range(len(regs))]
    res = _run_adaptive(qc, img, regs, classes)
    for i, (r, c) in enumerate(zip(regs, classes)):
        _check_adaptive(res, i, img, r, c)


def test_adaptive_method_large_block(qc):
    """2048x2048 line-scan frame: block size 205 (the largest tile halo)."""
    img = synth.make_linescan(9, 2048, 2048)
    regs = [dict(img=0, x=0, y=0, w=2048, h=2048)]
    res = _run_adaptive(qc, img, regs, ["crack"])
    _check_adaptive(res, 0, img, regs[0], "crack")


def test_dropin_segment_defects_matches_reference_golden(qc):
    """The acceptance test of row f3: ImageSegmentator.segment_defects (drop-in, CUDA) on the golden set's frames
    returns the areas / masks / methods the REFERENCE's own segment_defects recorded -- bit-exact for every class."""
    from qcgraft.detector import ImageSegmentator
    gold = json.load(open(os.path.join(G, "refseg_golden.json")))
    masks = np.load(os.path.join(G, "refseg_masks.npz"))
    seg = ImageSegmentator({})
    n = 0
    for fi, (frame, dets) in enumerate(refseg_cases.cases()):
        out = seg.segment_defects(frame, dets)
        assert "error" not in out, out
        got = {r["detection_id"]: r for r in out["segmented_regions"]}
        g_rows = gold["frames"][fi]["regions"]
        assert sorted(got) == [g["id"] for g in g_rows]
        for g in g_rows:
            r = got[g["id"]]
            b = g["bbox"]
            h = min(b["y2"], frame.shape[0]) - b["y1"]
            w = min(b["x2"], frame.shape[1]) - b["x1"]
            gm = np.unpackbits(masks[f"f{fi}_d{g['id']}"])[:h * w].reshape(h, w)
            assert np.array_equal(r["local_mask"] > 0, gm > 0), (fi, g["id"], g["class"])
            assert r["area_pixels"] == g["area_pixels"], (fi, g["id"])
            assert r["segmentation_method"] == g["method"], (fi, g["id"])
            assert abs(r["confidence_score"] - g["confidence_score"]) <= 1e-9, (fi, g["id"])
            # perimeter = n_axis + sqrt(2) n_diag from K8; cv2 sums float32 segment lengths: <= 1e-7 relative apart
            assert abs(r["perimeter"] - g["perimeter"]) <= 1e-6 * max(1.0, g["perimeter"])
            assert abs(r["compactness"] - g["compactness"]) <= 1e-6
            assert len(r["contours"]) == g["n_contours"]
            assert np.array_equal(np.asarray(r["mask"])[b["y1"]:b["y1"] + h, b["x1"]:b["x1"] + w] > 0, gm > 0)
            n += 1
        ref_ids = [g["id"] for g in g_rows if g["method"] != "threshold_fallback"]
        assert sum(got[i]["area_pixels"] for i in ref_ids) == gold["frames"][fi]["total_defect_area"]
    assert n == gold["n_rois"] >= 20


def test_contour_statistics_on_structured_masks(qc):
    """K8 on shapes that stress border following: one-pixel lines, diagonal chains (pinches), rings with nested
    components (RETR_EXTERNAL ignores what sits in a hole), shapes touching every ROI border, checkerboards."""
    H = W = 200
    img = np.full((H, W), 60, np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    img[20, 10:190] = 230; img[30:180, 15] = 230                          # 1-px lines
    img[(yy - xx == 20) & (yy > 40) & (yy < 160)] = 230                   # diagonal chain
    d = np.maximum(np.abs(yy - 110), np.abs(xx - 120))
    img[(d >= 40) & (d < 46)] = 230; img[(d >= 15) & (d < 25)] = 230      # ring with a ring inside its hole
    img[105:116, 115:126] = 230                                           # ... and a blob inside that
    img[0:8, 60:140] = 230; img[H - 6:, 30:90] = 230; img[70:130, W - 5:] = 230
    img[160:190, 150:190][((yy[160:190, 150:190] // 2) + (xx[160:190, 150:190] // 2)) % 2 == 0] = 230
    regs = [dict(img=0, x=0, y=0, w=W, h=H), dict(img=0, x=5, y=5, w=190, h=150), dict(img=0, x=60, y=50, w=120, h=121),
            dict(img=0, x=140, y=150, w=60, h=50), dict(img=0, x=0, y=0, w=33, h=200)]
    # composite-free check of the statistics: the adaptive / threshold methods produce the masks, cv2 the reference
    res = _run_threshold(qc, img, regs, ["x"] * len(regs))
    for i, r in enumerate(regs):
        _check_threshold(res, i, img, r, "x")
    res = _run_adaptive(qc, img, regs, ["other"] * len(regs))
    for i, r in enumerate(regs):
        _check_adaptive(res, i, img, r, "other")
