"""Oracle for K3b: the whole-frame classifier input of the reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

models/resent_model.py:57-66,114-130 builds the ResNet-50 input as
    PIL.Image.fromarray(rgb) -> transforms.Resize((224, 224)) -> ToTensor() -> Normalize(mean, std)
transforms.Resize on a PIL image is Image.resize(size, BILINEAR): Pillow's *antialiased* separable
resampler (Pillow src/libImaging/Resample.c, not vendored in /root/reference; Pillow 12.2 is installed here
and on the GPU box, so this restatement is pinned bit-exact against it in tests/test_oracle_pins.py):

  precompute_coeffs(in_size, out_size), all in float64:
      scale = in/out ; filterscale = max(scale, 1) ; support = 1.0 * filterscale ; ksize = ceil(support)*2 + 1
      per output index xx: center = (xx + 0.5) * scale ; ss = 1/filterscale
          xmin = max(int(center - support + 0.5), 0) ; xmax = min(int(center + support + 0.5), in) - xmin
          w[x] = triangle((x + xmin - center + 0.5) * ss), x < xmax ; w /= sum(w)
  normalize_coeffs_8bpc: k = int(w * 2^22 +- 0.5)  (round half away from zero), PRECISION_BITS = 32 - 8 - 2
  horizontal pass then vertical pass, each into uint8:
      acc = 2^21 + sum(pixel * k) ; out = clip8(acc >> 22)      (arithmetic shift, clamp to 0..255)
  A pass whose in_size == out_size is skipped.
ToTensor = float32(u8) / 255 ; Normalize = (x - mean) / std in float32.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
RESNET_MEAN = (0.485, 0.456, 0.406)   # models/resent_model.py:63
RESNET_STD = (0.229, 0.224, 0.225)    # models/resent_model.py:64


def _triangle(x: float) -> float:
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """-> (ksize, bounds int32 [out,2] = (xmin, count), kk int32 [out, ksize]) exactly as Pillow builds them."""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_triangle((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(v * (1 << PRECISION_BITS) - 0.5) if v < 0 else int(v * (1 << PRECISION_BITS) + 0.5)
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One resampling pass along `axis` of a uint8 [H,W,C] image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for o in range(bounds.shape[0]):
        x0, cnt = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt):
            acc += src[x0 + x] * int(kk[o, x])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear(img_u8_hwc: np.ndarray, out_hw=(224, 224)) -> np.ndarray:
    """uint8 [H,W,C] -> uint8 [oh,ow,C], bit-exact with PIL.Image.resize((ow, oh), BILINEAR)."""
    h, w = img_u8_hwc.shape[:2]
    oh, ow = out_hw
    x = img_u8_hwc
    if w != ow:
        _, b, k = precompute_coeffs(w, ow)
        x = _pass(x, b, k, 1)
    if h != oh:
        _, b, k = precompute_coeffs(h, oh)
        x = _pass(x, b, k, 0)
    return x


def resnet_input(img_u8_hwc: np.ndarray, out_hw=(224, 224), mean=RESNET_MEAN, std=RESNET_STD) -> np.ndarray:
    """The tensor models/resent_model.py:128 feeds the net: float32 [3,oh,ow]."""
    r = pil_resize_bilinear(img_u8_hwc, out_hw)
    x = r.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    m = np.asarray(mean, np.float32).reshape(3, 1, 1)
    s = np.asarray(std, np.float32).reshape(3, 1, 1)
    return ((x - m) / s).astype(np.float32)
