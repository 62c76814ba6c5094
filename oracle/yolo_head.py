"""Oracle for K2: YOLOv8 Detect/DFL decode, confidence filter, class-aware NMS,
scale_boxes, and the reference's result parsing.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference reaches this arithmetic through ``self.model(img, conf=, iou=)``
at models/yolo_model.py:105-110 (ultralytics, un-vendored: restated from the
8.1-8.3 sources; the wall-clock ``time_limit`` break is
deliberately excluded) and ``torchvision.ops.nms`` (present: called directly).
``parse_results``/``classify_severity`` follow models/yolo_model.py:128-172.
"""
from __future__ import annotations

import numpy as np
import torch
import torchvision

CLASS_NAMES = ["crack", "scratch", "dent", "discoloration", "contamination"]  # models/yolo_model.py:59-61


def make_anchors(level_shapes, strides, offset: float = 0.5):
    """ultralytics utils/tal.py::make_anchors: (x+0.5, y+0.5), y outer / x inner, levels in order."""
    pts, st = [], []
    for (h, w), s in zip(level_shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32))
    return torch.cat(pts).transpose(0, 1), torch.cat(st).transpose(0, 1)  # (2,A), (1,A)


def detect_decode(levels, nc: int, reg_max: int = 16, strides=(8, 16, 32)):
    """Detect._inference: levels = list of [B, 4*reg_max+nc, Hi, Wi] float32 -> [B, 4+nc, A]."""
    B = levels[0].shape[0]
    no = 4 * reg_max + nc
    x = torch.cat([l.float().reshape(B, no, -1) for l in levels], 2)
    box, cls = x.split((4 * reg_max, nc), 1)
    A = box.shape[-1]
    # DFL: view(b,4,c1,a).transpose(2,1).softmax(1) -> 1x1 conv with arange weights
    p = box.view(B, 4, reg_max, A).transpose(2, 1).softmax(1)
    w = torch.arange(reg_max, dtype=torch.float32).view(1, reg_max, 1, 1)
    dist = (p * w).sum(1)  # (B,4,A)  (upstream: conv2d; sum order may differ in the last ulp)
    anchors, st = make_anchors([l.shape[2:] for l in levels], strides)
    lt, rb = dist.chunk(2, 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    c_xy = (x1y1 + x2y2) / 2
    wh = x2y2 - x1y1
    dbox = torch.cat((c_xy, wh), 1) * st
    return torch.cat((dbox, cls.sigmoid()), 1)


def xywh2xyxy(x):
    y = torch.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def nms_restated(boxes: np.ndarray, scores: np.ndarray, iou_thr: float) -> np.ndarray:
    """Plain restatement of torchvision's CPU ``nms_kernel_impl`` (float32 IoU, no FMA,
    IoU compared against the *double* threshold, stable descending sort).  Used to pin
    the semantics the CUDA kernel reproduces; the graded oracle is torchvision itself."""
    b = boxes.astype(np.float32)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = ((x2 - x1).astype(np.float32) * (y2 - y1).astype(np.float32)).astype(np.float32)
    n = len(b)
    sup = np.zeros(n, bool)
    keep = []
    for _i in range(n):
        i = order[_i]
        if sup[i]:
            continue
        keep.append(i)
        js = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[js]); yy1 = np.maximum(y1[i], y1[js])
        xx2 = np.minimum(x2[i], x2[js]); yy2 = np.minimum(y2[i], y2[js])
        w = np.maximum(np.float32(0), (xx2 - xx1).astype(np.float32))
        h = np.maximum(np.float32(0), (yy2 - yy1).astype(np.float32))
        inter = (w * h).astype(np.float32)
        den = ((areas[i] + areas[js]).astype(np.float32) - inter).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = (inter / den).astype(np.float32)
        sup[js] |= ovr.astype(np.float64) > float(iou_thr)
    return np.asarray(keep, np.int64)


def non_max_suppression(pred, conf_thres=0.7, iou_thres=0.5, max_det=300, max_nms=30000,
                        max_wh=7680, return_idx=False):
    """ultralytics utils/ops.py::non_max_suppression (agnostic=False, multi_label=False).

    pred: [B, 4+nc, A] (output of detect_decode).  Returns list of [k,6] tensors
    (x1,y1,x2,y2,conf,cls) in descending score order, and optionally the anchor indices.
    """
    nc = pred.shape[1] - 4
    xc = pred[:, 4:4 + nc].amax(1) > conf_thres
    pred = pred.transpose(-1, -2).clone()
    pred[..., :4] = xywh2xyxy(pred[..., :4])
    out, idxs = [], []
    for xi, x in enumerate(pred):
        aidx = torch.nonzero(xc[xi]).flatten()
        x = x[xc[xi]]
        if not x.shape[0]:
            out.append(torch.zeros((0, 6))); idxs.append(torch.zeros((0,), dtype=torch.int64))
            continue
        box, cls = x[:, :4], x[:, 4:4 + nc]
        conf, j = cls.max(1, keepdim=True)
        m = conf.view(-1) > conf_thres
        x = torch.cat((box, conf, j.float()), 1)[m]
        aidx = aidx[m]
        n = x.shape[0]
        if not n:
            out.append(torch.zeros((0, 6))); idxs.append(torch.zeros((0,), dtype=torch.int64))
            continue
        if n > max_nms:
            o = x[:, 4].argsort(descending=True)[:max_nms]
            x, aidx = x[o], aidx[o]
        c = x[:, 5:6] * max_wh
        keep = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        out.append(x[keep]); idxs.append(aidx[keep])
    return (out, idxs) if return_idx else out


def scale_boxes(net_shape, boxes, orig_shape):
    """ultralytics ops.scale_boxes + clip_boxes (padding=True, xywh=False)."""
    H, W = net_shape
    h0, w0 = orig_shape
    gain = min(H / h0, W / w0)
    pad_x = round((W - w0 * gain) / 2 - 0.1)
    pad_y = round((H - h0 * gain) / 2 - 0.1)
    b = boxes.clone()
    b[..., 0] -= pad_x
    b[..., 1] -= pad_y
    b[..., 2] -= pad_x
    b[..., 3] -= pad_y
    b[..., :4] /= gain
    b[..., 0].clamp_(0, w0)
    b[..., 1].clamp_(0, h0)
    b[..., 2].clamp_(0, w0)
    b[..., 3].clamp_(0, h0)
    return b


def yolo_postprocess(levels, nc, net_shape, orig_shape, conf=0.7, iou=0.5, max_det=300,
                     strides=(8, 16, 32), return_idx=False, max_nms=30000):
    """a3+a4+a5: decode -> NMS -> scale_boxes.  Returns list of [k,6] float32 tensors."""
    y = detect_decode(levels, nc, strides=strides)
    dets, idxs = non_max_suppression(y, conf, iou, max_det=max_det, max_nms=max_nms, return_idx=True)
    dets = [torch.cat((scale_boxes(net_shape, d[:, :4], orig_shape), d[:, 4:]), 1) if len(d) else d
            for d in dets]
    return (dets, idxs) if return_idx else dets


# ---- models/yolo_model.py:128-172 ----------------------------------------
def classify_severity(confidence: float, area: float) -> str:
    normalized_area = area / (1024 * 1024)
    if confidence > 0.9 or normalized_area > 0.1:
        return "critical"
    elif confidence > 0.8 or normalized_area > 0.05:
        return "major"
    return "minor"


def parse_results(det: np.ndarray, class_names=CLASS_NAMES):
    """det: [k,6] float32 rows (x1,y1,x2,y2,conf,cls) -> the reference's detection dicts."""
    out = []
    for i, row in enumerate(np.asarray(det, np.float32)):
        x1, y1, x2, y2 = map(int, row[:4])
        conf = row[4]
        cid = int(row[5])
        out.append({
            "id": i,
            "class": class_names[cid] if cid < len(class_names) else f"class_{cid}",
            "confidence": float(conf),
            "bbox": {"x1": x1, "y1": y1, "x2": x2, "y2": y2, "width": x2 - x1, "height": y2 - y1,
                     "center_x": (x1 + x2) / 2, "center_y": (y1 + y2) / 2},
            "area": (x2 - x1) * (y2 - y1),
            "severity": classify_severity(conf, (x2 - x1) * (y2 - y1)),
        })
    return out
