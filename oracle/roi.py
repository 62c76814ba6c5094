"""Oracle for K3: ROI-align crop-resize of kept boxes into 224x224 classifier inputs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference only *mocks* this hand-off (models/ensemble.py:129-142,181-195:
"would normally extract the region and run ResNet on it").  The intended
behaviour is fixed here as:
    torchvision.ops.roi_align(img_f32[N,3,H,W] (0..255), rois[K,5], (224,224),
                              spatial_scale=1, sampling_ratio=2, aligned=True)
    -> /255 -> Normalize(mean,std)        (mean/std: models/resent_model.py:62-65)
with rois = the int-truncated xyxy boxes of models/yolo_model.py:138.
"""
from __future__ import annotations

import numpy as np
import torch
import torchvision

RESNET_MEAN = (0.485, 0.456, 0.406)   # models/resent_model.py:63
RESNET_STD = (0.229, 0.224, 0.225)    # models/resent_model.py:64


def roi_align_normalize(imgs_u8_nhwc: np.ndarray, rois: np.ndarray, out_hw=(224, 224),
                        sampling_ratio: int = 2, aligned: bool = True, swap_rb: bool = False,
                        mean=RESNET_MEAN, std=RESNET_STD, dtype=torch.float32):
    """imgs: [N,H,W,3] uint8; rois: [K,5] float (img_idx,x1,y1,x2,y2) -> [K,3,oh,ow]."""
    x = np.asarray(imgs_u8_nhwc)
    if swap_rb:
        x = x[..., ::-1]
    t = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2))).float()
    r = torch.as_tensor(np.asarray(rois, np.float32))
    if r.numel() == 0:
        return torch.zeros((0, 3) + tuple(out_hw), dtype=dtype)
    o = torchvision.ops.roi_align(t, r, out_hw, spatial_scale=1.0,
                                  sampling_ratio=sampling_ratio, aligned=aligned)
    o = o / 255
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    o = (o - m) / s
    return o if dtype == torch.float32 else o.to(dtype)


def roi_align_restated(img_chw: np.ndarray, roi, out_hw=(224, 224), sr: int = 2, aligned=True):
    """Scalar restatement of torchvision's CPU roi_align for ONE roi (float32 math) --
    documents the sampling rule the CUDA kernel follows; pinned against torchvision in
    tests/test_oracle_pins.py.  img_chw float32 [C,H,W]; roi=(x1,y1,x2,y2)."""
    C, H, W = img_chw.shape
    PH, PW = out_hw
    f = np.float32
    off = f(0.5) if aligned else f(0)
    x1, y1, x2, y2 = [f(v) * f(1.0) - off for v in roi]
    rw, rh = f(x2 - x1), f(y2 - y1)
    if not aligned:
        rw, rh = max(rw, f(1)), max(rh, f(1))
    bh, bw = f(rh / f(PH)), f(rw / f(PW))
    gh = sr if sr > 0 else int(np.ceil(rh / PH))
    gw = sr if sr > 0 else int(np.ceil(rw / PW))
    count = f(max(gh * gw, 1))
    out = np.zeros((C, PH, PW), np.float32)
    for ph in range(PH):
        for pw in range(PW):
            acc = np.zeros(C, np.float32)
            for iy in range(gh):
                y = f(y1 + f(ph) * bh + f(f(iy) + f(.5)) * bh / f(gh))
                for ix in range(gw):
                    x = f(x1 + f(pw) * bw + f(f(ix) + f(.5)) * bw / f(gw))
                    if y < -1.0 or y > H or x < -1.0 or x > W:
                        continue
                    yy, xx = max(y, f(0)), max(x, f(0))
                    yl, xl = int(yy), int(xx)
                    if yl >= H - 1:
                        yh = yl = H - 1; yy = f(yl)
                    else:
                        yh = yl + 1
                    if xl >= W - 1:
                        xh = xl = W - 1; xx = f(xl)
                    else:
                        xh = xl + 1
                    ly, lx = f(yy - yl), f(xx - xl)
                    hy, hx = f(1) - ly, f(1) - lx
                    w1, w2, w3, w4 = f(hy * hx), f(hy * lx), f(ly * hx), f(ly * lx)
                    acc = acc + (w1 * img_chw[:, yl, xl] + w2 * img_chw[:, yl, xh]
                                 + w3 * img_chw[:, yh, xl] + w4 * img_chw[:, yh, xh]).astype(np.float32)
            out[:, ph, pw] = acc / count
    return out
