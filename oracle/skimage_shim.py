"""scikit-image stand-in for running the REFERENCE's own inference/segmentation.py in this container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

scikit-image (requirements.txt:35, ``>=0.21.0``, unpinned) is absent from the image and there is no network, so
the four functions inference/segmentation.py:12-14 uses on the graft path are restated from the published 0.21+
sources with scipy/numpy (present).  "parity unpinned" for this library: the restatement cannot be cross-checked
against skimage itself offline.  What IS pinned: the reference's own file runs unchanged on top of this shim
(tests/golden/make_golden.py), and oracle/refseg.py restates the same arithmetic operation by operation, bit for bit
against this shim (tests/test_refseg_cpu.py).

Published semantics followed:
  * ``filters.gaussian(image, sigma)``: ``img_as_float`` then ``scipy.ndimage.gaussian_filter(image, sigma,
    mode='nearest', truncate=4.0)``.  ``img_as_float`` on uint8 is ``np.multiply(image, 1. / 255, dtype=float64)``
    (util/dtype.py ``_convert``: a multiplication by the rounded reciprocal, not a division).
  * ``filters.threshold_otsu(image, nbins=256)``: constant image -> its value; else ``exposure.histogram`` =
    ``np.histogram(image, bins=nbins, range=(min, max))`` with bin centres ``(edges[:-1] + edges[1:]) / 2``;
    ``_validate_image_histogram`` hands the counts over as float32; class weights by cumulative sums, the
    between-class variance ``weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2`` and its first argmax.
  * ``morphology.disk(2)``: ``x**2 + y**2 <= 4`` (13 pixels == the L1 ball of radius 2).
  * ``morphology.binary_closing / binary_opening``: scipy binary dilation (outside = 0) / erosion (outside = 1).
"""
from __future__ import annotations

import sys
import types

import numpy as np
from scipy import ndimage


def img_as_float(img: np.ndarray) -> np.ndarray:
    img = np.asarray(img)
    if img.dtype == np.uint8:
        return np.multiply(img, 1.0 / 255, dtype=np.float64)
    return img.astype(np.float64, copy=False)


def gaussian(image, sigma=1.0, **_kw):
    return ndimage.gaussian_filter(img_as_float(image), sigma, mode="nearest", truncate=4.0)


def threshold_otsu(image=None, nbins=256):
    image = np.asarray(image)
    first_pixel = image.reshape(-1)[0]
    if np.all(image == first_pixel):
        return first_pixel
    counts, bin_edges = np.histogram(image.reshape(-1), bins=nbins, range=None)
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2.0
    counts = counts.astype("float32", copy=False)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    mean1 = np.cumsum(counts * bin_centers) / weight1
    mean2 = (np.cumsum((counts * bin_centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return bin_centers[np.argmax(variance12)]


def disk(radius):
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=np.uint8)


def binary_dilation(image, footprint=None):
    return ndimage.binary_dilation(image, structure=footprint)


def binary_erosion(image, footprint=None):
    return ndimage.binary_erosion(image, structure=footprint, border_value=True)


def binary_closing(image, footprint=None):
    return binary_erosion(binary_dilation(image, footprint), footprint)


def binary_opening(image, footprint=None):
    return binary_dilation(binary_erosion(image, footprint), footprint)


def install():
    """Registers ``skimage`` (+ an empty ``matplotlib``) in sys.modules so inference/segmentation.py imports."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    stub("matplotlib.patches", Polygon=object)
    morph = stub("skimage.morphology", disk=disk, binary_closing=binary_closing, binary_opening=binary_opening,
                 binary_dilation=binary_dilation, binary_erosion=binary_erosion)
    sk = stub("skimage", morphology=morph)
    sk.measure = stub("skimage.measure")
    sk.filters = stub("skimage.filters", threshold_otsu=threshold_otsu, gaussian=gaussian)
    sk.segmentation = stub("skimage.segmentation", watershed=None)
