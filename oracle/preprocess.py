"""Oracle for K1: decode-side colour swap, stretch resize, LetterBox, /255 NCHW pack.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  * inference/detector.py:226-254   (_preprocess_image: BGR2RGB swap, optional
    stretch ``cv2.resize(image,(W,H))``)
  * models/yolo_model.py:71-82      (second BGR2RGB swap)
  * ultralytics 8.x ``LetterBox`` + predictor ``preprocess`` invoked by
    models/yolo_model.py:105-110 (restated):
    resize -> pad 114 -> [..., ::-1] -> HWC->NCHW -> float32 -> /255
"""
from __future__ import annotations

import numpy as np
import cv2
import torch


# --------------------------------------------------------------------------
# exact integer model of cv2.resize(INTER_LINEAR) on uint8
# --------------------------------------------------------------------------
def _coeffs(dst: int, src: int, clamp_f: bool):
    scale = np.float64(src) / dst
    f = ((np.arange(dst) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_f:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), a0, a1


def resize_linear_u8(img: np.ndarray, W: int, H: int) -> np.ndarray:
    """Integer restatement of ``cv2.resize(img,(W,H),interpolation=INTER_LINEAR)``
    for uint8 HWC/HW input.  Pinned bit-exact against cv2 in
    tests/test_oracle_pins.py."""
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    h, w = img.shape[:2]
    if (h, w) == (H, W):
        out = img.copy()
        return out[:, :, 0] if squeeze else out
    sx0, sx1, ax0, ax1 = _coeffs(W, w, True)
    sy0, sy1, ay0, ay1 = _coeffs(H, h, False)
    I = img.astype(np.int64)
    rows = I[:, sx0] * ax0[None, :, None] + I[:, sx1] * ax1[None, :, None]
    r0, r1 = rows[sy0], rows[sy1]
    out = (((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    out = out.astype(np.uint8)
    return out[:, :, 0] if squeeze else out


# --------------------------------------------------------------------------
# LetterBox geometry (ultralytics data/augment.py::LetterBox, 8.x)
# --------------------------------------------------------------------------
def letterbox_geometry(h0: int, w0: int, H: int, W: int, auto: bool = False,
                       scaleup: bool = True, stride: int = 32):
    """Returns dict(new_w,new_h,top,bottom,left,right,out_h,out_w,gain)."""
    r = min(H / h0, W / w0)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = W - new_w, H - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return dict(new_w=new_w, new_h=new_h, top=top, bottom=bottom, left=left, right=right,
                out_h=new_h + top + bottom, out_w=new_w + left + right, gain=r)


def letterbox_u8(img: np.ndarray, H: int, W: int, auto: bool = False, pad_value: int = 114,
                 use_cv2: bool = True):
    """LetterBox on a uint8 HWC image -> (padded uint8 HWC, geometry)."""
    h0, w0 = img.shape[:2]
    g = letterbox_geometry(h0, w0, H, W, auto=auto)
    if (w0, h0) != (g["new_w"], g["new_h"]):
        if use_cv2:
            img = cv2.resize(img, (g["new_w"], g["new_h"]), interpolation=cv2.INTER_LINEAR)
        else:
            img = resize_linear_u8(img, g["new_w"], g["new_h"])
    if use_cv2:
        out = cv2.copyMakeBorder(img, g["top"], g["bottom"], g["left"], g["right"],
                                 cv2.BORDER_CONSTANT, value=(pad_value,) * 3)
    else:
        out = np.full((g["out_h"], g["out_w"], 3), pad_value, np.uint8)
        out[g["top"]:g["top"] + g["new_h"], g["left"]:g["left"] + g["new_w"]] = img
    return out, g


def stretch_u8(img: np.ndarray, H: int, W: int) -> np.ndarray:
    """inference/detector.py:240-242 -- ``cv2.resize(image,(W,H))`` (not letterbox)."""
    return cv2.resize(img, (W, H))


def pack_nchw(imgs_u8_hwc, swap_rb: bool = True, mean=None, std=None, dtype=torch.bfloat16):
    """ultralytics predictor.preprocess: stack -> [..., ::-1] -> NCHW -> float32 -> /255.

    ``mean``/``std`` given -> torchvision ToTensor+Normalize order
    (models/resent_model.py:59-66): (v/255 - mean)/std in float32.
    The bf16 result is the RNE rounding of that float32 tensor (torch ``.to``).
    """
    x = np.stack(list(imgs_u8_hwc))
    if swap_rb:
        x = x[..., ::-1]
    x = np.ascontiguousarray(x.transpose(0, 3, 1, 2))
    t = torch.from_numpy(x).float() / 255
    if mean is not None:
        m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
        s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
        t = (t - m) / s
    return t if dtype == torch.float32 else t.to(dtype)


def yolo_preprocess(imgs_u8_hwc, H: int, W: int, auto: bool = False, swap_rb: bool = True,
                    dtype=torch.bfloat16):
    """Full a2 chain for a same-shape batch: LetterBox each -> pack.  Returns (tensor, geometry)."""
    outs, g = [], None
    for im in imgs_u8_hwc:
        o, g = letterbox_u8(im, H, W, auto=auto)
        outs.append(o)
    return pack_nchw(outs, swap_rb=swap_rb, dtype=dtype), g
