"""Oracle for row f3 (preparation): hole filling of a binary mask.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

inference/segmentation.py:455-456 (`_clean_mask`) calls scipy.ndimage.binary_fill_holes(mask) with the default
structuring element (4-connectivity): the holes are the background pixels that cannot be reached from the image
border through 4-connected background.  Restated as: label the complement with 4-connectivity, every component
that touches the border is outside, everything else is filled.  Pinned bit-exactly against scipy in
tests/test_oracle_pins.py.  On the device this is K4's union-find run on the complement bit rows with a 4-neighbour
contact rule (no diagonal carries) plus a "touches the border" flag per root.
"""
from __future__ import annotations

import cv2
import numpy as np


def fill_holes(mask: np.ndarray) -> np.ndarray:
    """bool/uint8 [H,W] -> bool [H,W] == scipy.ndimage.binary_fill_holes(mask)."""
    m = np.asarray(mask) != 0
    if m.size == 0:
        return m.copy()
    n, lab = cv2.connectedComponents((~m).astype(np.uint8), connectivity=4)
    border = np.zeros(n, bool)
    for edge in (lab[0], lab[-1], lab[:, 0], lab[:, -1]):
        border[edge] = True
    border[0] = False                                   # label 0 = the mask itself
    outside = border[lab] & ~m
    return ~outside
