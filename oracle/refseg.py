"""Oracle for the reference's own per-class segmentation methods (SURVEY.md rows a9, a10, a11 as the reference
computes them -- not the K4 composite of oracle/segment.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows inference/segmentation.py:
  * :185-195  method per class: crack/scratch -> 'adaptive', discoloration/unknown -> 'threshold';
              dent -> 'watershed' and contamination -> 'region_growing' are OUT OF SCOPE: the
              drop-in routes them to the threshold method under the name 'threshold_fallback', which is what the
              reference itself returns when a method raises (:175-183).
  * :197-246  _threshold_segmentation: skimage gaussian(sigma=1) -> threshold_otsu -> x class factor -> '<' (dark
              classes) / '>' -> _clean_mask -> uint8*255; confidence :527-555.
  * :449-462  _clean_mask: _remove_small_components(10) -> ndimage.binary_fill_holes -> disk(2) closing, opening.
  * :248-298  _adaptive_segmentation: cv2.adaptiveThreshold(GAUSSIAN_C, BINARY_INV for crack/scratch, block
              max(11, min(h,w)//10) | 1, C=2) -> 3x3 ELLIPSE (= cross) CLOSE, OPEN -> _remove_small_components(20);
              confidence from Canny overlap :557-578.
  * :464-484  _remove_small_components: cv2.connectedComponents (8-conn) + bincount + isin.

Two forms of every stage: the LIBRARY form calls scipy / numpy / cv2 (and the skimage shim, oracle/skimage_shim.py)
exactly as the reference does; the RESTATED form spells the arithmetic out operation by operation in the order the
CUDA kernels follow.  tests/test_refseg_cpu.py pins restated == library bit for bit (float64 blur, histogram bins,
Otsu threshold, masks), and both against the reference's own file run through the shim (tests/golden).
"""
from __future__ import annotations

import cv2
import numpy as np
from scipy import ndimage

from . import skimage_shim as sk
from .segment import CROSS, morph_close_open_restated

THRESHOLD_FACTORS = {"crack": 0.9, "scratch": 0.9, "dent": 1.1, "discoloration": 0.8, "contamination": 0.85}   # :218-224
DARK_CLASSES = ("crack", "scratch")                                                                          # :229, :278
METHOD = {"crack": "adaptive", "scratch": "adaptive", "dent": "watershed", "discoloration": "threshold",
          "contamination": "region_growing"}                                                                 # :187-193
# what the drop-in runs: the two out-of-scope methods take the reference's own fallback path (:175-183)
DROPIN_METHOD = {"adaptive": "adaptive", "threshold": "threshold", "watershed": "threshold_fallback",
                 "region_growing": "threshold_fallback"}

# scipy.ndimage._filters._gaussian_kernel1d(1.0, 0, 4): exp(-x^2/2) / sum, x = -4..4 -- w[0] is the centre tap.
# Hard-coded (hex float64) so the device and this file share the very same nine numbers; pinned against scipy.
GAUSS_W = tuple(float.fromhex(h) for h in (
    "0x1.9884a307594fbp-2", "0x1.ef8eb9ad499bap-3", "0x1.ba4b99d1799abp-5", "0x1.22724cb7eb269p-8",
    "0x1.18a9c4fd536c6p-13"))
U8_TO_FLOAT = np.multiply(np.arange(256, dtype=np.uint8), 1.0 / 255, dtype=np.float64)        # img_as_float


def method_for(defect_class: str) -> str:
    return METHOD.get(defect_class, "threshold")


def dropin_method_for(defect_class: str) -> str:
    return DROPIN_METHOD[method_for(defect_class)]


# --------------------------------------------------------------------------------------------------------------
# sigma = 1 Gaussian, float64 (skimage.filters.gaussian -> scipy.ndimage.gaussian_filter, mode='nearest')
def _correlate9_nearest(a: np.ndarray, axis: int) -> np.ndarray:
    """scipy's NI_Correlate1D, symmetric branch (ni_filters.c): out = x[0]*w0, then for jj = -4..-1:
    out += (x[jj] + x[-jj]) * w[jj]; line extended by edge replication; no fused multiply-add."""
    n = a.shape[axis]
    idx = np.clip(np.arange(-4, n + 4), 0, n - 1)
    p = np.take(a, idx, axis=axis)

    def sl(o):        # samples at l + o for every l
        return np.take(p, np.arange(4 + o, 4 + o + n), axis=axis)
    out = sl(0) * GAUSS_W[0]
    for jj in (-4, -3, -2, -1):
        out = out + (sl(jj) + sl(-jj)) * GAUSS_W[-jj]
    return out


def gaussian_sigma1(roi: np.ndarray) -> np.ndarray:
    return sk.gaussian(roi, sigma=1.0)


def gaussian_sigma1_restated(roi: np.ndarray) -> np.ndarray:
    x = U8_TO_FLOAT[roi] if roi.dtype == np.uint8 else roi.astype(np.float64)
    return _correlate9_nearest(_correlate9_nearest(x, 0), 1)        # axis 0 first, then axis 1


# --------------------------------------------------------------------------------------------------------------
# skimage.filters.threshold_otsu on the float64 image
def histogram256_restated(a: np.ndarray):
    """np.histogram(a, 256, range=(min, max)) -> (counts int64[256], edges f64[257]); the bin of a value is fixed by
    the comparisons against np.linspace's edges (numpy corrects its multiplicative estimate by one step)."""
    a = a.reshape(-1)
    first, last = a.min(), a.max()
    step = (last - first) / 256
    edges = np.arange(257, dtype=np.float64) * step + first
    edges[-1] = last
    est = ((a - first) / (last - first)) * 256
    idx = est.astype(np.intp)
    idx[idx == 256] -= 1
    idx[a < edges[idx]] -= 1
    inc = (a >= edges[idx + 1]) & (idx != 255)
    idx[inc] += 1
    return np.bincount(idx, minlength=256), edges


def otsu_float_restated(blurred: np.ndarray):
    """Sequential form of threshold_otsu (float32 counts / weights, float64 means).  Returns the threshold."""
    flat = blurred.reshape(-1)
    if np.all(flat == flat[0]):
        return flat[0]
    counts, edges = histogram256_restated(flat)
    centers = (edges[:-1] + edges[1:]) / 2.0
    c32 = counts.astype(np.float32)
    w1 = np.zeros(256, np.float32); w2 = np.zeros(256, np.float32)
    s1 = np.zeros(256, np.float64); s2 = np.zeros(256, np.float64)
    acc32, acc64 = np.float32(0), np.float64(0)
    for i in range(256):
        acc32 = np.float32(acc32 + c32[i]); w1[i] = acc32
        acc64 = acc64 + np.float64(c32[i]) * centers[i]; s1[i] = acc64
    acc32, acc64 = np.float32(0), np.float64(0)
    for i in range(255, -1, -1):
        acc32 = np.float32(acc32 + c32[i]); w2[i] = acc32
        acc64 = acc64 + np.float64(c32[i]) * centers[i]; s2[i] = acc64
    best, bidx = -1.0, 0
    for i in range(255):
        m1 = s1[i] / np.float64(w1[i])
        m2 = s2[i + 1] / np.float64(w2[i + 1])
        d = m1 - m2
        v = np.float64(np.float32(w1[i] * w2[i + 1])) * (d * d)
        if v > best:
            best, bidx = v, i
    return centers[bidx]


def remove_small_components(mask_bool: np.ndarray, min_size: int) -> np.ndarray:
    """:464-484 on a bool mask."""
    n, labels = cv2.connectedComponents(mask_bool.astype(np.uint8) * 255)
    sizes = np.bincount(labels.ravel())
    large = sizes >= min_size
    large[0] = False
    return np.isin(labels, np.where(large)[0])


def clean_mask(mask_bool: np.ndarray, restated: bool = False) -> np.ndarray:
    """:449-462."""
    m = remove_small_components(mask_bool, 10)
    if restated:
        from .fill_holes import fill_holes
        m = fill_holes(m)
        return morph_close_open_restated(m.astype(np.uint8) * 255, 2) > 0
    m = ndimage.binary_fill_holes(m)
    k = sk.disk(2)
    m = sk.binary_closing(m, k)
    return sk.binary_opening(m, k)


def threshold_confidence(roi: np.ndarray, mask_bool: np.ndarray) -> float:
    """:527-555."""
    if not np.any(mask_bool):
        return 0.0
    fg, bg = roi[mask_bool], roi[~mask_bool]
    if len(fg) == 0 or len(bg) == 0:
        return 0.0
    sep = abs(np.mean(fg) - np.mean(bg)) / (np.std(roi) + 1e-6)
    return float(min(sep / 3.0, 1.0))


def threshold_segmentation(roi: np.ndarray, defect_class: str, restated: bool = False):
    """:197-246.  Returns (mask uint8 {0,255}, confidence, aux dict)."""
    blurred = gaussian_sigma1_restated(roi) if restated else gaussian_sigma1(roi)
    thr = otsu_float_restated(blurred) if restated else sk.threshold_otsu(blurred)
    adj = thr * THRESHOLD_FACTORS.get(defect_class, 1.0)
    raw = blurred < adj if defect_class in DARK_CLASSES else blurred > adj
    m = clean_mask(raw, restated)
    return m.astype(np.uint8) * 255, threshold_confidence(roi, m), dict(blurred=blurred, threshold=float(thr), raw=raw)


# --------------------------------------------------------------------------------------------------------------
# cv2.adaptiveThreshold(ADAPTIVE_THRESH_GAUSSIAN_C)
def adaptive_block_size(h: int, w: int) -> int:
    b = max(11, min(h, w) // 10)                                                   # :263-265
    return b + 1 if b % 2 == 0 else b


def gaussian_kernel_f32(bs: int) -> np.ndarray:
    return cv2.getGaussianKernel(bs, -1, cv2.CV_32F).ravel().astype(np.float32)


def _fma32(a, b, c):
    # a*b is exact in float64 (24+24 bits); the sum rounds to float64, then to float32: a double rounding that can
    # differ from a true float32 fma only when the float64 sum lands exactly on a float32 tie -- never observed
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def adaptive_mean_restated(roi: np.ndarray, bs: int) -> np.ndarray:
    """uint8 'mean' image of cv2.adaptiveThreshold = cvRound(float32 separable Gaussian, BORDER_REPLICATE).
    Model of cv2 4.13's AVX2 code path (probed, tools/probe_adaptive_threshold.py): row pass s = x[0]*k[0], then
    s = fma(x[t], k[t], s) for t = 1..bs-1; column pass s = x[c]*k[c], then s = fma(x[c+t] + x[c-t], k[c+t], s) for
    t = 1..bs//2.  That holds for output columns below (w & ~7); the next 4-column group has the fused row pass and an
    unfused column pass; the last w % 4 columns are unfused in both passes (compiler-vectorised scalar tail: exact
    for all but a handful of float values, none of which survives the rounding to uint8 in the pinned cases)."""
    k = gaussian_kernel_f32(bs)
    pad = bs // 2
    h, w = roi.shape
    src = roi.astype(np.float32)
    P = np.pad(src, ((0, 0), (pad, pad)), mode="edge")

    def rows(fused):
        acc = (P[:, 0:w] * k[0]).astype(np.float32)
        for t in range(1, bs):
            x = P[:, t:t + w]
            kk = np.full_like(x, k[t])
            acc = _fma32(x, kk, acc) if fused else (acc + (x * kk).astype(np.float32)).astype(np.float32)
        return acc

    def cols(acc, fused):
        Pc = np.pad(acc, ((pad, pad), (0, 0)), mode="edge")
        out = (Pc[pad:pad + h] * k[pad]).astype(np.float32)
        for t in range(1, pad + 1):
            s = (Pc[pad + t:pad + t + h] + Pc[pad - t:pad - t + h]).astype(np.float32)
            kk = np.full_like(s, k[pad + t])
            out = _fma32(s, kk, out) if fused else (out + (s * kk).astype(np.float32)).astype(np.float32)
        return out
    rf, ru = rows(True), rows(False)
    blur = cols(rf, True)
    t0 = w & ~7
    t1 = t0 + 4 * ((w - t0) // 4)
    blur[:, t0:t1] = cols(rf, False)[:, t0:t1]
    blur[:, t1:] = cols(ru, False)[:, t1:]
    return np.clip(np.rint(blur), 0, 255).astype(np.uint8)


def adaptive_threshold_restated(roi: np.ndarray, bs: int, inv: bool) -> np.ndarray:
    """BINARY: src - mean > -2 ; BINARY_INV: src - mean <= -2 (cv2's 768-entry table, delta = 2)."""
    d = roi.astype(np.int32) - adaptive_mean_restated(roi, bs).astype(np.int32)
    return np.where(d <= -2 if inv else d > -2, 255, 0).astype(np.uint8)


def adaptive_confidence(roi: np.ndarray, mask: np.ndarray) -> float:
    """:557-578."""
    ratio = np.sum(mask > 0) / mask.size
    if ratio < 0.01 or ratio > 0.8:
        return 0.3
    edges = cv2.Canny(roi, 50, 150)
    overlap = np.sum((edges > 0) & (mask > 0)) / np.sum(mask > 0)
    return float(min(0.5 + 0.5 * overlap, 1.0))


def adaptive_segmentation(roi: np.ndarray, defect_class: str, restated: bool = False):
    """:248-298.  Returns (mask uint8 {0,255}, confidence, aux dict)."""
    h, w = roi.shape
    bs = adaptive_block_size(h, w)
    inv = defect_class in DARK_CLASSES
    if restated:
        raw = adaptive_threshold_restated(roi, bs, inv)
        m = morph_close_open_restated(raw, 1)
    else:
        raw = cv2.adaptiveThreshold(np.ascontiguousarray(roi), 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C,
                                    cv2.THRESH_BINARY_INV if inv else cv2.THRESH_BINARY, bs, 2)
        m = cv2.morphologyEx(raw, cv2.MORPH_CLOSE, CROSS)
        m = cv2.morphologyEx(m, cv2.MORPH_OPEN, CROSS)
    m = remove_small_components(m > 0, 20).astype(np.uint8) * 255
    return m, adaptive_confidence(roi, m), dict(raw=raw, block=bs)


# --------------------------------------------------------------------------------------------------------------
def segment_region(roi: np.ndarray, defect_class: str, restated: bool = False):
    """What the drop-in returns for one ROI: (mask, confidence, method name)."""
    method = dropin_method_for(defect_class)
    if method == "adaptive":
        m, c, _ = adaptive_segmentation(roi, defect_class, restated)
    else:
        m, c, _ = threshold_segmentation(roi, defect_class, restated)
    return m, c, method


def find_contours(mask: np.ndarray):
    """:486-506."""
    contours, _ = cv2.findContours(np.ascontiguousarray(mask), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return [c for c in contours if cv2.contourArea(c) >= 10]


def segment_defects(image: np.ndarray, detections, restated: bool = False):
    """:39-146 reduced to what is compared: per kept detection (non-empty ROI) a dict with detection_id, mask,
    area_pixels, method, confidence_score, perimeter, compactness."""
    gray = cv2.cvtColor(image, cv2.COLOR_RGB2GRAY) if image.ndim == 3 else image
    out = []
    for i, d in enumerate(detections):
        b = d["bbox"]
        roi = gray[b["y1"]:b["y2"], b["x1"]:b["x2"]]
        if roi.size == 0:
            continue
        m, c, method = segment_region(roi, d["class"], restated)
        cs = find_contours(m)
        area = int(np.sum(m > 0))
        per = float(sum(cv2.arcLength(c_, True) for c_ in cs))
        comp = 0.0 if (not cs or area == 0 or per == 0) else min((4 * np.pi * area) / (per ** 2), 1.0)
        out.append(dict(detection_id=i, mask=m, area_pixels=area, method=method, confidence_score=c,
                        perimeter=per, compactness=comp, n_contours=len(cs)))
    return out
