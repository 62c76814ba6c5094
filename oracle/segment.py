"""Oracle for K4: gray -> Otsu -> polarity compare -> 3x3-cross close/open -> 8-conn
connected components (cv2 numbering) -> per-component area/bbox -> min-area filter.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows inference/segmentation.py:
  * :66-69   cv2.cvtColor(RGB2GRAY)
  * :274-289 (_adaptive_segmentation order: threshold -> MORPH_CLOSE -> MORPH_OPEN with
              getStructuringElement(MORPH_ELLIPSE,(3,3)) -> _remove_small_components)
  * :212-234 (_threshold_segmentation: Otsu threshold x class factor, '<' for dark
              classes, '>' otherwise)
  * :464-484 (_remove_small_components: cv2.connectedComponents + bincount + isin)
  * :98      per-detection area = count of mask>0
K4's graded composite is the a11 stage order with Otsu as the thresholder
(SURVEY.md §8a note v); every stage is a cv2 call that exists in this image.
"""
from __future__ import annotations

import numpy as np
import cv2

CROSS = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], np.uint8)  # == getStructuringElement(MORPH_ELLIPSE,(3,3))

# inference/segmentation.py:218-226
THRESHOLD_FACTORS = {"crack": 0.9, "scratch": 0.9, "dent": 1.1, "discoloration": 0.8, "contamination": 0.85}
DARK_CLASSES = ("crack", "scratch")  # :229


def rgb2gray(img: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)


def rgb2gray_restated(img: np.ndarray) -> np.ndarray:
    """Exact fixed-point model of cv2 RGB2GRAY."""
    i = img.astype(np.int64)
    return ((9798 * i[..., 0] + 19235 * i[..., 1] + 3735 * i[..., 2] + 16384) >> 15).astype(np.uint8)


def otsu_u8(gray: np.ndarray) -> int:
    t, _ = cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
    return int(t)


def otsu_u8_restated(gray: np.ndarray) -> int:
    """Scalar restatement of cv2's getThreshVal_Otsu_8u (sequential double recurrence,
    strict '>' so the first maximum wins).  This is the op order the CUDA kernel follows."""
    h = np.bincount(gray.ravel(), minlength=256)
    n = gray.size
    scale = 1.0 / n
    mu = 0.0
    for i in range(256):
        mu += i * float(h[i])
    mu *= scale
    mu1 = q1 = 0.0
    max_sigma, max_val = 0.0, 0
    eps = 1.1920928955078125e-07  # FLT_EPSILON
    for i in range(256):
        p_i = h[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma, max_val = sigma, i
    return max_val


def morph_close_open(mask: np.ndarray, iters: int = 1) -> np.ndarray:
    m = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, CROSS, iterations=iters)
    return cv2.morphologyEx(m, cv2.MORPH_OPEN, CROSS, iterations=iters)


def _dilate(b):  # outside = 0
    p = np.pad(b, 1, constant_values=False)
    return p[1:-1, 1:-1] | p[:-2, 1:-1] | p[2:, 1:-1] | p[1:-1, :-2] | p[1:-1, 2:]


def _erode(b):   # outside = 1 (ignored)
    p = np.pad(b, 1, constant_values=True)
    return p[1:-1, 1:-1] & p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]


def morph_close_open_restated(mask: np.ndarray, iters: int = 1) -> np.ndarray:
    """close = erode^k(dilate^k), open = dilate^k(erode^k), cross SE, cv2 border rules."""
    b = mask > 0
    for _ in range(iters): b = _dilate(b)
    for _ in range(iters): b = _erode(b)
    for _ in range(iters): b = _erode(b)
    for _ in range(iters): b = _dilate(b)
    return b.astype(np.uint8) * 255


def ccl_restated(mask: np.ndarray):
    """8-connectivity labeling with cv2.connectedComponents' default numbering: components
    numbered in increasing order of their first 2x2 block in block-raster order
    (SURVEY.md App. B3).  Returns (labels int32, num_labels incl. background)."""
    from scipy import ndimage
    lab, n = ndimage.label(mask > 0, structure=np.ones((3, 3)))
    if n == 0:
        return lab.astype(np.int32), 1
    h, w = mask.shape
    yy, xx = np.mgrid[0:h, 0:w]
    key = (yy // 2) * ((w + 1) // 2) + (xx // 2)
    first = ndimage.minimum(key, labels=lab, index=np.arange(1, n + 1))
    remap = np.zeros(n + 1, np.int32)
    remap[np.argsort(first, kind="stable") + 1] = np.arange(1, n + 1)
    return remap[lab].astype(np.int32), n + 1


def segment_k4(img: np.ndarray, polarity: int = 1, thr_scale: float = 1.0, morph_iters: int = 1,
               min_area: int = 20):
    """K4 composite on one region (gray HW or RGB HWC uint8).

    polarity 1: foreground = v > t*scale (light defects); 0: foreground = v < t*scale... see below.
    With thr_scale == 1.0, polarity 1 == cv2 THRESH_BINARY (v > t) and polarity 0 ==
    THRESH_BINARY_INV (v <= t), i.e. the two cv2 modes the reference's adaptive path selects
    (inference/segmentation.py:278).  With thr_scale != 1.0 the compare is the reference's
    float compare of :226-234: light: v > t*scale ; dark: v < t*scale.

    Returns dict(thr, mask u8{0,255} (after min-area filter), labels int32 (cv2 numbering,
    removed components zeroed), ncomp (labels incl. background, before filter),
    stats [ncomp,5] int32 (x,y,w,h,area; row 0 = background), keep [ncomp] bool).
    """
    gray = rgb2gray(img) if img.ndim == 3 else img
    t = otsu_u8(gray)
    if thr_scale == 1.0:
        fg = gray > t if polarity else gray <= t
    else:
        ta = t * thr_scale
        fg = gray.astype(np.float64) > ta if polarity else gray.astype(np.float64) < ta
    m0 = fg.astype(np.uint8) * 255
    m1 = morph_close_open(m0, morph_iters) if morph_iters > 0 else m0
    ncomp, labels, stats, _ = cv2.connectedComponentsWithStats(m1, connectivity=8, ltype=cv2.CV_32S)
    keep = stats[:, cv2.CC_STAT_AREA] >= min_area
    keep[0] = False
    labels_f = np.where(keep[labels], labels, 0).astype(np.int32)
    mask = (labels_f > 0).astype(np.uint8) * 255
    return dict(thr=t, mask=mask, labels=labels_f, ncomp=int(ncomp), stats=stats.astype(np.int32),
                keep=keep, mask_premorph=m0, mask_morph=m1)
