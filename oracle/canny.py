"""Oracle for the adaptive method's confidence (inference/segmentation.py:557-578): ``cv2.Canny(roi, 50, 150)``
restated in integer arithmetic, and the edge-overlap confidence built on it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

cv2.Canny (aperture 3, L2gradient=False), OpenCV 4.x imgproc/canny.cpp:
  * dx, dy = 3x3 Sobel, int16, BORDER_REPLICATE;  mag = |dx| + |dy|, zero outside the image
  * a pixel with mag > low is a candidate when it is a maximum along its gradient direction: with x = |dx|,
    y = |dy| << 15, TG22 = 13573 (tan 22.5 deg in Q15): y < x*TG22 -> horizontal (mag > left and mag >= right);
    y > x*TG22 + (x << 16) -> vertical (mag > up and mag >= down); otherwise diagonal with s = -1 if dx, dy have
    opposite signs else +1 (mag > up[j - s] and mag > down[j + s])
  * candidates with mag > high are edges; the other candidates become edges when 8-connected to one (hysteresis).
Pinned bit for bit against cv2.Canny in tests/test_refseg_cpu.py."""
from __future__ import annotations

import cv2
import numpy as np

TG22 = 13573


def canny_restated(roi: np.ndarray, low: int = 50, high: int = 150) -> np.ndarray:
    h, w = roi.shape
    p = np.pad(roi.astype(np.int32), 1, mode="edge")
    dx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    dy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    mag = np.abs(dx) + np.abs(dy)
    M = np.pad(mag, 1, constant_values=0)
    c = M[1:-1, 1:-1]
    left, right, up, down = M[1:-1, :-2], M[1:-1, 2:], M[:-2, 1:-1], M[2:, 1:-1]
    x = np.abs(dx).astype(np.int64)
    y = np.abs(dy).astype(np.int64) << 15
    tg22x = x * TG22
    tg67x = tg22x + (x << 16)
    horiz = y < tg22x
    vert = ~horiz & (y > tg67x)
    diag = ~horiz & ~vert
    s_pos = (dx ^ dy) >= 0                     # s = +1: compare up[j-1], down[j+1]; s = -1: up[j+1], down[j-1]
    up_l, up_r, dn_l, dn_r = M[:-2, :-2], M[:-2, 2:], M[2:, :-2], M[2:, 2:]
    d_ok = np.where(s_pos, (c > up_l) & (c > dn_r), (c > up_r) & (c > dn_l))
    cand = (c > low) & ((horiz & (c > left) & (c >= right)) | (vert & (c > up) & (c >= down)) | (diag & d_ok))
    strong = cand & (c > high)
    n, lab = cv2.connectedComponents(cand.astype(np.uint8), connectivity=8)
    keep = np.zeros(n, bool)
    keep[np.unique(lab[strong])] = True
    keep[0] = False
    return np.where(keep[lab], 255, 0).astype(np.uint8)


def adaptive_confidence_restated(roi: np.ndarray, mask: np.ndarray) -> float:
    """:557-578 with the restated Canny."""
    area = int(np.sum(mask > 0))
    ratio = area / mask.size
    if ratio < 0.01 or ratio > 0.8:
        return 0.3
    overlap = int(np.sum((canny_restated(roi) > 0) & (mask > 0))) / area
    return float(min(0.5 + 0.5 * overlap, 1.0))
