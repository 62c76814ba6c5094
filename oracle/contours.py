"""Oracle for row a12: perimeter / contour count of a mask without tracing contours.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

inference/segmentation.py:486-525: ``cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)``, contours with
``cv2.contourArea < 10`` dropped, perimeter = sum of ``cv2.arcLength(c, closed=True)``, compactness = min(4*pi*area /
perimeter^2, 1).  Library form: exactly those cv2 calls.  Restated form (what the device computes): the external
contours are the outer borders of the 8-connected components of the hole-filled mask (RETR_EXTERNAL ignores holes and
everything nested in them; holes = background not 4-connected to the frame, oracle/fill_holes.py); Suzuki's border
following moves between pixel centres, and every move shows up exactly once in one 2x2 pixel window of the filled mask
F (window = the four pixels around a lattice vertex, outside the image = 0):

    two edge-adjacent foreground pixels    one axis move (length 1) along that edge
    three foreground pixels                one diagonal move (length sqrt 2) between the two pixels next to the gap
    two diagonal foreground pixels         two diagonal moves (the border passes the pinch twice)
    one or four foreground pixels          no move

so perimeter = n_axis + sqrt(2) * n_diag per component, and the polygon area of cv2.contourArea is half the sum of the
cross products p x p' over the same moves, oriented with the foreground on the right of the direction of travel.
Pinned against cv2 in tests/test_refseg_cpu.py (perimeter to float32 rounding of cv2's per-segment lengths, areas and
counts exactly)."""
from __future__ import annotations

import cv2
import numpy as np

from .fill_holes import fill_holes

MIN_CONTOUR_AREA = 10          # inference/segmentation.py:499


def contour_stats(mask: np.ndarray):
    """Library form: (perimeter, number of kept contours, sorted kept contour areas)."""
    cs, _ = cv2.findContours(np.ascontiguousarray((mask > 0).astype(np.uint8) * 255), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    kept = [c for c in cs if cv2.contourArea(c) >= MIN_CONTOUR_AREA]
    return float(sum(cv2.arcLength(c, True) for c in kept)), len(kept), sorted(float(cv2.contourArea(c)) for c in kept)


def contour_stats_restated(mask: np.ndarray):
    F = fill_holes(mask > 0)
    h, w = F.shape
    n, lab = cv2.connectedComponents(F.astype(np.uint8), connectivity=8)
    P = np.zeros((h + 2, w + 2), bool); P[1:-1, 1:-1] = F
    L = np.zeros((h + 2, w + 2), np.int64); L[1:-1, 1:-1] = lab
    # window at vertex (vy, vx), vy in 0..h, vx in 0..w: a = (vy-1, vx-1), b = (vy-1, vx), c = (vy, vx-1), d = (vy, vx)
    a, b, c, d = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    la, lb, lc, ld = L[:-1, :-1], L[:-1, 1:], L[1:, :-1], L[1:, 1:]
    vy, vx = np.mgrid[0:h + 1, 0:w + 1]
    ax, ay, bx, by, cx, cy, dx, dy = vx - 1, vy - 1, vx, vy - 1, vx - 1, vy, vx, vy      # pixel coordinates (x, y)
    cnt = a.astype(int) + b + c + d
    n_axis = np.zeros(n, np.int64); n_diag = np.zeros(n, np.int64); cross = np.zeros(n, np.int64)

    def add(sel, label, axis, diag, x0, y0, x1, y1):
        np.add.at(n_axis, label[sel], axis)
        np.add.at(n_diag, label[sel], diag)
        np.add.at(cross, label[sel], (x0 * y1 - x1 * y0)[sel])
    two = cnt == 2
    add(two & c & d, lc, 1, 0, cx, cy, dx, dy)            # (. . / c d): foreground south of the crack, travel east  c -> d
    add(two & a & b, la, 1, 0, bx, by, ax, ay)            # (a b / . .): foreground north, travel west               b -> a
    add(two & b & d, lb, 1, 0, dx, dy, bx, by)            # (. b / . d): foreground east, travel north               d -> b
    add(two & a & c, la, 1, 0, ax, ay, cx, cy)            # (a . / c .): foreground west, travel south               a -> c
    three = cnt == 3
    add(three & ~d, la, 0, 1, bx, by, cx, cy)             # gap at d: b -> c
    add(three & ~a, lb, 0, 1, cx, cy, bx, by)             # gap at a: c -> b
    add(three & ~b, la, 0, 1, ax, ay, dx, dy)             # gap at b: a -> d
    add(three & ~c, la, 0, 1, dx, dy, ax, ay)             # gap at c: d -> a
    pin = two & a & d                                      # pinch: a -> d and d -> a (cross products cancel)
    np.add.at(n_diag, la[pin], 2)
    pin = two & b & c
    np.add.at(n_diag, lb[pin], 2)
    areas = np.abs(cross[1:]) / 2.0
    keep = areas >= MIN_CONTOUR_AREA
    per = float((n_axis[1:][keep] + np.sqrt(2.0) * n_diag[1:][keep]).sum())
    return per, int(keep.sum()), sorted(areas[keep].tolist())
