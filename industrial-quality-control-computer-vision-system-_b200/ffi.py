"""ctypes binding of include/qcgraft.h + thin tensor-level wrappers.

The library is the product; this module only marshals ``tensor.data_ptr()`` and the current
CUDA stream into the C ABI.  There is NO CPU fallback: a missing library or device raises
``QcGraftError`` (the drop-in classes turn it into the reference's ``{"error": ...}`` dicts,
inference/detector.py:124-126).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import build as _build

QC_DT_F32, QC_DT_BF16 = 0, 1
QC_PRE_STRETCH, QC_PRE_LETTERBOX, QC_PRE_LETTERBOX_AUTO = 0, 1, 2
QC_PRE_SWAP_RB, QC_PRE_LUT_UNIFORM = 1, 2
QC_MAX_LEVELS = 4
QC_SEG_COMPOSITE, QC_SEG_REF_THRESHOLD, QC_SEG_REF_ADAPTIVE = 0, 1, 2


class QcGraftError(RuntimeError):
    pass


class PrePlan(C.Structure):
    _fields_ = [("src_h", C.c_int32), ("src_w", C.c_int32), ("new_h", C.c_int32), ("new_w", C.c_int32),
                ("top", C.c_int32), ("left", C.c_int32), ("dst_h", C.c_int32), ("dst_w", C.c_int32),
                ("gain", C.c_float), ("pad_x", C.c_int32), ("pad_y", C.c_int32)]


class YoloParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("nc", C.c_int32), ("reg_max", C.c_int32), ("num_levels", C.c_int32),
                ("level_h", C.c_int32 * QC_MAX_LEVELS), ("level_w", C.c_int32 * QC_MAX_LEVELS),
                ("level_stride", C.c_float * QC_MAX_LEVELS),
                ("level_batch_stride", C.c_int64 * QC_MAX_LEVELS),
                ("level_chan_stride", C.c_int64 * QC_MAX_LEVELS),
                ("in_dtype", C.c_int32), ("conf_thres", C.c_double), ("iou_thres", C.c_double),
                ("max_det", C.c_int32), ("max_nms", C.c_int32), ("max_wh", C.c_float),
                ("gain", C.c_float), ("pad_x", C.c_int32), ("pad_y", C.c_int32),
                ("orig_h", C.c_int32), ("orig_w", C.c_int32), ("scale_boxes", C.c_int32)]


class SegRegion(C.Structure):
    _fields_ = [("img", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("polarity", C.c_int32), ("min_area", C.c_int32), ("max_comp", C.c_int32),
                ("thr_scale", C.c_double),
                ("out_offset", C.c_int64), ("stats_offset", C.c_int32), ("band_offset", C.c_int32),
                ("chunk_offset", C.c_int32), ("bits_offset", C.c_int32), ("blk_offset", C.c_int32)]


class SegParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("channels", C.c_int32),
                ("img_stride", C.c_int64), ("row_stride", C.c_int32), ("morph_iters", C.c_int32),
                ("swap_rb", C.c_int32), ("num_regions", C.c_int32), ("total_pixels", C.c_int64), ("total_stats_rows", C.c_int32),
                ("total_bits_words", C.c_int32), ("total_blocks", C.c_int32), ("total_bands", C.c_int32),
                ("total_chunks", C.c_int32), ("max_region_w", C.c_int32), ("max_region_h", C.c_int32)]


class JpegInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("components", C.c_int32), ("subsampling", C.c_int32),
                ("restart_interval", C.c_int32), ("reserved", C.c_int32), ("blocks", C.c_int64)]


QC_JPEG_GRAY, QC_JPEG_444, QC_JPEG_422, QC_JPEG_420 = 0, 1, 2, 3

# K6 whole-chain tables (qc_rules_row / qc_rules_refined / qc_rules_out of include/qcgraft.h)
RULES_ROW_DTYPE = np.dtype([("x1", "i4"), ("y1", "i4"), ("x2", "i4"), ("y2", "i4"), ("cls", "i4"), ("severity", "i4"),
                            ("has_seg", "i4"), ("reserved", "i4"), ("ext", "i4", (4,)), ("area_pixels", "i8"),
                            ("conf", "f8"), ("seg_conf", "f8"), ("area_pct", "f8"), ("compactness", "f8")], align=True)
RULES_REFINED_DTYPE = np.dtype([("x1", "i4"), ("y1", "i4"), ("x2", "i4"), ("y2", "i4"), ("severity", "i4"), ("kept", "i4"),
                                ("area_pixels", "i8")], align=True)
RULES_OUT_DTYPE = np.dtype([("base", "i4"), ("merged_from", "i4"), ("x1", "i4"), ("y1", "i4"), ("x2", "i4"), ("y2", "i4"),
                            ("severity", "i4"), ("flags", "i4"), ("area_pixels", "i8"), ("conf", "f8"),
                            ("validation_score", "f8")], align=True)

SEG_REGION_DTYPE = np.dtype([("img", "i4"), ("x", "i4"), ("y", "i4"), ("w", "i4"), ("h", "i4"),
                             ("polarity", "i4"), ("min_area", "i4"), ("max_comp", "i4"), ("thr_scale", "f8"),
                             ("out_offset", "i8"), ("stats_offset", "i4"), ("band_offset", "i4"),
                             ("chunk_offset", "i4"), ("bits_offset", "i4"), ("blk_offset", "i4")], align=True)
assert SEG_REGION_DTYPE.itemsize == C.sizeof(SegRegion)

_SIGS = {
    "qc_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "qc_destroy": (None, [C.c_void_p]),
    "qc_last_error": (C.c_char_p, [C.c_void_p]),
    "qc_version": (C.c_char_p, []),
    "qc_sm_count": (C.c_int, [C.c_void_p]),
    "qc_launch_count": (C.c_int64, [C.c_void_p]),
    "qc_sizeof": (C.c_size_t, [C.c_int]),
    "qc_rules_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                 C.c_uint64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qc_jpeg_info_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(JpegInfo)]),
    "qc_jpeg_entropy_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "qc_jpeg_staging_bytes": (C.c_size_t, [C.POINTER(JpegInfo), C.c_int]),
    "qc_jpeg_workspace_bytes": (C.c_size_t, [C.POINTER(JpegInfo), C.c_int]),
    "qc_jpeg_decode": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.POINTER(JpegInfo),
                                 C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_int, C.c_void_p]),
    "qc_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "qc_profile_collect": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]),
    "qc_pre_make_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(PrePlan)]),
    "qc_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.POINTER(PrePlan),
                                C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "qc_yolo_workspace_bytes": (C.c_size_t, [C.POINTER(YoloParams)]),
    "qc_yolo_decode_nms": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(YoloParams), C.c_void_p,
                                     C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qc_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "qc_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                         C.c_double, C.c_float, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                         C.c_void_p]),
    "qc_roi_align": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int32,
                               C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.c_int, C.c_void_p, C.c_void_p]),
    "qc_roi_align_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    "qc_segment_plan": (C.c_int, [C.POINTER(SegRegion), C.c_int, C.POINTER(SegParams)]),
    "qc_pil_resize_table_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "qc_pil_resize_make_tables": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "qc_pil_resize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "qc_segment_workspace_bytes": (C.c_size_t, [C.POINTER(SegParams)]),
    "qc_segment": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SegParams), C.c_void_p, C.c_void_p, C.c_size_t,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qc_segment_ref_workspace_bytes": (C.c_size_t, [C.POINTER(SegParams), C.c_int]),
    "qc_segment_ref": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(SegParams), C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qc_spc_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "qc_rules_cluster": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_double, C.c_void_p, C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None
_lib_lock = threading.Lock()


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen lib/libqcgraft.so (building it first if the sources are newer)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if build_if_missing and _build.needs_build():
            try:
                _build.build()
            except Exception as e:  # no nvcc on the box: use the shipped .so if there is one
                if not os.path.exists(path):
                    raise QcGraftError(f"libqcgraft.so missing and cannot be built: {e}") from e
        if not os.path.exists(path):
            raise QcGraftError(f"{path} not found: run __graft_entry__.build()")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)     # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        for which, st in enumerate((PrePlan, YoloParams, SegRegion, SegParams)):
            if lib.qc_sizeof(which) != C.sizeof(st):
                raise QcGraftError(f"ABI mismatch: {st.__name__} is {C.sizeof(st)} B here, "
                                   f"{lib.qc_sizeof(which)} B in libqcgraft.so")
        _lib = lib
        return lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Context:
    """One qc_ctx per (device, thread): the reference calls predict() from thread pools
    (inference/detector.py:143, models/ensemble.py:88) with no locks."""

    _tls = threading.local()

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise QcGraftError("CUDA device required: qcgraft has no CPU fallback")
        self.lib = load_library()
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream()          # make sure the primary context exists
            rc = self.lib.qc_create(device, C.byref(h))
        if rc != 0:
            raise QcGraftError(f"qc_create({device}) failed with status {rc}")
        self.h = h
        self._lut_cache = {}
        self._ws = {}

    @classmethod
    def current(cls, device: int = 0) -> "Context":
        d = getattr(cls._tls, "ctxs", None)
        if d is None:
            d = cls._tls.ctxs = {}
        if device not in d:
            d[device] = Context(device)
        return d[device]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.qc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.qc_last_error(self.h)
            raise QcGraftError(f"{what}: status {rc}: {msg.decode() if msg else ''}")

    @property
    def launch_count(self) -> int:
        return int(self.lib.qc_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(self.lib.qc_sm_count(self.h))

    def profile(self, on: bool) -> None:
        """Per-kernel CUDA-event timing inside the library (measurement only; see include/qcgraft.h)."""
        self._check(self.lib.qc_profile_enable(self.h, int(bool(on))), "qc_profile_enable")

    def profile_collect(self, cap: int = 4096):
        """[(kernel name, ms)] in launch order since the last collection; waits for the recorded events."""
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = min(int(self.lib.qc_profile_collect(self.h, names, ms, cap)), cap)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def workspace(self, key: str, nbytes: int) -> torch.Tensor:
        """Grow-only cached scratch buffer (caller-owned memory as far as the C ABI goes)."""
        t = self._ws.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = t
        return t

    def norm_lut(self, mean=None, std=None) -> Tuple[torch.Tensor, bool]:
        """float32 [3][256] table built with the reference's own float32 ops:
        ultralytics: v.float()/255 ; torchvision: ToTensor (/255) then (t-mean)/std."""
        key = (tuple(mean) if mean is not None else None, tuple(std) if std is not None else None)
        if key not in self._lut_cache:
            t = (torch.arange(256, dtype=torch.float32) / 255).repeat(3, 1)
            if mean is not None:
                m = torch.tensor(mean, dtype=torch.float32).view(3, 1)
                s = torch.tensor(std, dtype=torch.float32).view(3, 1)
                t = (t - m) / s
            uniform = bool(torch.equal(t[0], t[1]) and torch.equal(t[0], t[2]))
            self._lut_cache[key] = (t.contiguous().to(self.device), uniform)
        return self._lut_cache[key]

    # -- K1 --------------------------------------------------------------------------------
    def make_pre_plan(self, mode: int, src_h: int, src_w: int, H: int, W: int, stride: int = 32) -> PrePlan:
        pl = PrePlan()
        rc = self.lib.qc_pre_make_plan(mode, src_h, src_w, H, W, stride, C.byref(pl))
        if rc != 0:
            raise QcGraftError(f"qc_pre_make_plan failed: {rc}")
        return pl

    def preprocess(self, src: torch.Tensor, plan: PrePlan, swap_rb: bool = True, mean=None, std=None,
                   pad_value: int = 114, out_dtype=torch.bfloat16, out: Optional[torch.Tensor] = None,
                   events=None):
        """src: uint8 CUDA tensor [n,h,w,3] (any row/image pitch, channel-contiguous)."""
        assert src.is_cuda and src.dtype == torch.uint8 and src.dim() == 4 and src.shape[3] == 3
        assert src.stride(3) == 1 and src.stride(2) == 3, "pixels must be packed RGB triplets"
        n = src.shape[0]
        assert (src.shape[1], src.shape[2]) == (plan.src_h, plan.src_w)
        lut, uniform = self.norm_lut(mean, std)
        if out is None:
            out = torch.empty((n, 3, plan.dst_h, plan.dst_w), dtype=out_dtype, device=src.device)
        assert out.is_contiguous() and out.dtype in (torch.bfloat16, torch.float32)
        flags = (QC_PRE_SWAP_RB if swap_rb else 0) | (QC_PRE_LUT_UNIFORM if uniform else 0)
        dt = QC_DT_BF16 if out.dtype == torch.bfloat16 else QC_DT_F32
        args = (self.h, src.data_ptr(), n, src.stride(0) if n > 1 else src.shape[1] * src.stride(1),
                src.stride(1), C.byref(plan), lut.data_ptr(), pad_value, flags, dt, out.data_ptr(),
                _stream(src.device))
        if events is not None:          # CUDA events hugging the launch itself (bench.py roofline)
            events[0].record()
        rc = self.lib.qc_preprocess(*args)
        if events is not None:
            events[1].record()
        self._check(rc, "qc_preprocess")
        return out

    # -- K2 --------------------------------------------------------------------------------
    def yolo_params(self, levels: Sequence[torch.Tensor], nc: int, strides=(8, 16, 32), conf=0.7, iou=0.5,
                    max_det=300, max_nms=30000, max_wh=7680.0, plan: Optional[PrePlan] = None,
                    reg_max: int = 16) -> YoloParams:
        p = YoloParams()
        p.n = levels[0].shape[0]
        p.nc, p.reg_max, p.num_levels = nc, reg_max, len(levels)
        for i, l in enumerate(levels):
            assert l.is_cuda and l.dim() == 4 and l.shape[1] == 4 * reg_max + nc
            assert l.stride(3) == 1 and l.stride(2) == l.shape[3], "level planes must be contiguous"
            p.level_h[i], p.level_w[i] = l.shape[2], l.shape[3]
            p.level_stride[i] = float(strides[i])
            p.level_batch_stride[i] = l.stride(0)
            p.level_chan_stride[i] = l.stride(1)
        dt = levels[0].dtype
        assert all(l.dtype == dt for l in levels) and dt in (torch.float32, torch.bfloat16)
        p.in_dtype = QC_DT_F32 if dt == torch.float32 else QC_DT_BF16
        p.conf_thres, p.iou_thres = float(conf), float(iou)
        p.max_det, p.max_nms, p.max_wh = max_det, max_nms, float(max_wh)
        if plan is not None:
            p.gain, p.pad_x, p.pad_y = plan.gain, plan.pad_x, plan.pad_y
            p.orig_h, p.orig_w, p.scale_boxes = plan.src_h, plan.src_w, 1
        else:
            p.gain, p.scale_boxes = 1.0, 0
        return p

    def yolo_decode_nms(self, levels: Sequence[torch.Tensor], params: YoloParams, out=None):
        n, md = params.n, params.max_det
        dev = levels[0].device
        if out is None:
            det = torch.empty((n, md, 6), dtype=torch.float32, device=dev)
            anchor = torch.empty((n, md), dtype=torch.int32, device=dev)
            count = torch.empty((n,), dtype=torch.int32, device=dev)
            ncand = torch.empty((n,), dtype=torch.int32, device=dev)
        else:
            det, anchor, count, ncand = out
        need = self.lib.qc_yolo_workspace_bytes(C.byref(params))
        ws = self.workspace("yolo", need)
        ptrs = (C.c_void_p * QC_MAX_LEVELS)(*[l.data_ptr() for l in levels] + [None] * (QC_MAX_LEVELS - len(levels)))
        rc = self.lib.qc_yolo_decode_nms(self.h, ptrs, C.byref(params), ws.data_ptr(), ws.numel(), det.data_ptr(),
                                         anchor.data_ptr(), count.data_ptr(), ncand.data_ptr(), _stream(dev))
        self._check(rc, "qc_yolo_decode_nms")
        return det, anchor, count, ncand

    def nms(self, boxes: torch.Tensor, scores: torch.Tensor, cls: torch.Tensor, counts: torch.Tensor,
            iou: float = 0.5, max_wh: float = 7680.0, max_det: int = 300):
        """boxes [n,cap,4] f32, scores [n,cap] f32, cls [n,cap] i32, counts [n] i32 -> (keep [n,max_det], count [n])."""
        n, cap = scores.shape
        assert boxes.is_contiguous() and scores.is_contiguous() and cls.is_contiguous()
        assert boxes.dtype == torch.float32 and scores.dtype == torch.float32 and cls.dtype == torch.int32
        keep = torch.empty((n, max_det), dtype=torch.int32, device=boxes.device)
        cnt = torch.empty((n,), dtype=torch.int32, device=boxes.device)
        need = self.lib.qc_nms_workspace_bytes(n, cap)
        ws = self.workspace("nms", need)
        rc = self.lib.qc_nms(self.h, boxes.data_ptr(), scores.data_ptr(), cls.data_ptr(), counts.data_ptr(), n, cap,
                             float(iou), float(max_wh), max_det, ws.data_ptr(), ws.numel(), keep.data_ptr(),
                             cnt.data_ptr(), _stream(boxes.device))
        self._check(rc, "qc_nms")
        return keep, cnt

    # -- K3 --------------------------------------------------------------------------------
    def roi_align(self, imgs: torch.Tensor, rois: torch.Tensor, out_hw=(224, 224), sampling_ratio: int = 2,
                  aligned: bool = True, swap_rb: bool = False, mean=(0.485, 0.456, 0.406),
                  std=(0.229, 0.224, 0.225), out_dtype=torch.bfloat16, out: Optional[torch.Tensor] = None,
                  dst_rows: Optional[torch.Tensor] = None):
        """dst_rows (int32 CUDA [K]): ROI k goes to row dst_rows[k] of ``out`` (default: row k)."""
        assert imgs.is_cuda and imgs.dtype == torch.uint8 and imgs.dim() == 4 and imgs.shape[3] == 3
        assert imgs.stride(3) == 1 and imgs.stride(2) == 3
        assert rois.is_cuda and rois.dtype == torch.float32 and rois.is_contiguous() and rois.shape[-1] == 5
        if dst_rows is not None:
            assert out is not None and dst_rows.is_cuda and dst_rows.dtype == torch.int32 and dst_rows.numel() == rois.shape[0]
        K = rois.shape[0]
        key = ("norm", tuple(mean), tuple(std))
        if key not in self._lut_cache:
            t = torch.tensor([[255.0, m, s] for m, s in zip(mean, std)], dtype=torch.float32)
            self._lut_cache[key] = t.to(self.device)
        norm = self._lut_cache[key]
        if out is None:
            out = torch.empty((K, 3) + tuple(out_hw), dtype=out_dtype, device=imgs.device)
        if K == 0:
            return out
        n, h, w = imgs.shape[:3]
        dt = QC_DT_BF16 if out.dtype == torch.bfloat16 else QC_DT_F32
        rc = self.lib.qc_roi_align_rows(self.h, imgs.data_ptr(), n, h, w, imgs.stride(0) if n > 1 else h * imgs.stride(1),
                                        imgs.stride(1), rois.data_ptr(), _ptr(dst_rows), K, out_hw[0], out_hw[1],
                                        sampling_ratio, int(aligned), int(swap_rb), norm.data_ptr(), dt, out.data_ptr(),
                                        _stream(imgs.device))
        self._check(rc, "qc_roi_align_rows")
        return out

    # -- K3b -------------------------------------------------------------------------------
    def pil_resize_tables(self, h: int, w: int, out_hw=(224, 224)):
        """(host int32 table, device copy) of Pillow's resampling coefficients for h x w -> out_hw, cached."""
        key = ("pil", h, w, tuple(out_hw))
        if key not in self._lut_cache:
            nb = self.lib.qc_pil_resize_table_bytes(h, w, out_hw[0], out_hw[1])
            host = torch.empty(nb // 4, dtype=torch.int32)
            rc = self.lib.qc_pil_resize_make_tables(h, w, out_hw[0], out_hw[1], host.data_ptr())
            self._check(rc, "qc_pil_resize_make_tables")
            self._lut_cache[key] = (host, host.to(self.device))
        return self._lut_cache[key]

    def pil_resize(self, imgs: torch.Tensor, out_hw=(224, 224), swap_rb: bool = False, mean=(0.485, 0.456, 0.406),
                   std=(0.229, 0.224, 0.225), out_dtype=torch.bfloat16, out: Optional[torch.Tensor] = None):
        """uint8 [n,h,w,3] CUDA frames -> [n,3,oh,ow]: PIL BILINEAR (antialiased) resize + ToTensor + Normalize,
        the whole-frame input of models/resent_model.py:114-130."""
        assert imgs.is_cuda and imgs.dtype == torch.uint8 and imgs.dim() == 4 and imgs.shape[3] == 3
        assert imgs.stride(3) == 1 and imgs.stride(2) == 3
        n, h, w = imgs.shape[:3]
        key = ("norm", tuple(mean), tuple(std))
        if key not in self._lut_cache:
            t = torch.tensor([[255.0, m, s] for m, s in zip(mean, std)], dtype=torch.float32)
            self._lut_cache[key] = t.to(self.device)
        norm = self._lut_cache[key]
        host, dev = self.pil_resize_tables(h, w, out_hw)
        if out is None:
            out = torch.empty((n, 3) + tuple(out_hw), dtype=out_dtype, device=imgs.device)
        if n == 0:
            return out
        assert out.is_contiguous() and tuple(out.shape) == (n, 3) + tuple(out_hw)
        dt = QC_DT_BF16 if out.dtype == torch.bfloat16 else QC_DT_F32
        rc = self.lib.qc_pil_resize(self.h, imgs.data_ptr(), n, h, w, imgs.stride(0) if n > 1 else h * imgs.stride(1),
                                    imgs.stride(1), dev.data_ptr(), host.data_ptr(), int(swap_rb), norm.data_ptr(), dt,
                                    out.data_ptr(), _stream(imgs.device))
        self._check(rc, "qc_pil_resize")
        return out

    # -- K4 --------------------------------------------------------------------------------
    def segment_plan(self, regions, n: int, h: int, w: int, channels: int, img_stride: int,
                     row_stride: int, morph_iters: int = 1, swap_rb: bool = False):
        """regions: list of dicts, or a dict of equal-length arrays, with img,x,y,w,h and optional
        polarity(1), thr_scale(1.0), min_area(20), max_comp.  Returns (recarray of planned
        qc_seg_region, SegParams)."""
        if isinstance(regions, dict):
            cols = regions
            R = len(np.atleast_1d(cols["img"]))
        else:
            R = len(regions)
            keys = ("img", "x", "y", "w", "h", "polarity", "thr_scale", "min_area", "max_comp")
            cols = {k: [r[k] for r in regions] for k in keys if R and all(k in r for r in regions)}
            for k in keys:                      # per-region optional keys
                if k not in cols and R and any(k in r for r in regions):
                    dflt = {"polarity": 1, "thr_scale": 1.0, "min_area": 20}.get(k)
                    cols[k] = [r.get(k, dflt) for r in regions]
        arr = np.zeros(max(R, 1), SEG_REGION_DTYPE)
        if R:
            for k in ("img", "x", "y", "w", "h"):
                arr[k][:R] = np.asarray(cols[k], np.int64)
            arr["polarity"][:R] = np.asarray(cols.get("polarity", 1), np.int64)
            arr["thr_scale"][:R] = np.asarray(cols.get("thr_scale", 1.0), np.float64)
            arr["min_area"][:R] = np.asarray(cols.get("min_area", 20), np.int64)
            nblk = ((arr["w"][:R].astype(np.int64) + 1) // 2) * ((arr["h"][:R].astype(np.int64) + 1) // 2)
            if "max_comp" in cols and cols["max_comp"] is not None and not (
                    isinstance(cols["max_comp"], list) and any(v is None for v in cols["max_comp"])):
                arr["max_comp"][:R] = np.asarray(cols["max_comp"], np.int64)
            else:
                arr["max_comp"][:R] = np.minimum(nblk + 1, 1 << 16)
        p = SegParams()
        p.n, p.h, p.w, p.channels = n, h, w, channels
        p.img_stride, p.row_stride, p.morph_iters = img_stride, row_stride, morph_iters
        p.swap_rb = int(swap_rb)
        rc = self.lib.qc_segment_plan(arr.ctypes.data_as(C.POINTER(SegRegion)), R, C.byref(p))
        if rc != 0:
            raise QcGraftError(f"qc_segment_plan failed: {rc} (region outside the image?)")
        return arr.view(np.recarray), p

    def adaptive_kernel_table(self, max_block: int):
        """Device float32 table of cv2.getGaussianKernel(bs, -1, CV_32F) for every odd block size 11..max_block (row
        (bs-11)/2, pitch = max_block): cv2's own numbers, so K7's adaptive threshold sums exactly what
        cv2.adaptiveThreshold sums (inference/segmentation.py:274-281).  Returns (tensor, pitch, max_block)."""
        import cv2
        max_block = max(int(max_block) | 1, 65)
        key = ("akern", max_block)
        if key not in self._lut_cache:
            t = np.zeros(((max_block - 11) // 2 + 1, max_block), np.float32)
            for bs in range(11, max_block + 1, 2):
                t[(bs - 11) // 2, :bs] = cv2.getGaussianKernel(bs, -1, cv2.CV_32F).ravel()
            self._lut_cache[key] = torch.from_numpy(t).to(self.device)
        return self._lut_cache[key], max_block, max_block

    def segment(self, imgs: torch.Tensor, regions, morph_iters: int = 1, want_mask: bool = True,
                want_labels: bool = True, swap_rb: bool = False, method: int = QC_SEG_COMPOSITE,
                adaptive_c: int = 2, want_sums: bool = False, want_contours: bool = False):
        """imgs: uint8 CUDA [n,h,w] (gray) or [n,h,w,3] (RGB).  Returns dict of device tensors + the plan.
        method: QC_SEG_COMPOSITE (K4: Otsu -> morphology -> CCL), QC_SEG_REF_THRESHOLD / QC_SEG_REF_ADAPTIVE (K7: the
        reference's _threshold_segmentation / _adaptive_segmentation, see include/qcgraft.h)."""
        assert imgs.is_cuda and imgs.dtype == torch.uint8 and imgs.dim() in (3, 4)
        n, h, w = imgs.shape[:3]
        ch = 1 if imgs.dim() == 3 else imgs.shape[3]
        assert imgs.stride(-1) == 1 and (ch == 1 or imgs.stride(2) == ch)
        dev = imgs.device
        arr, p = self.segment_plan(regions, n, h, w, ch, imgs.stride(0) if n > 1 else h * imgs.stride(1),
                                   imgs.stride(1), morph_iters, swap_rb)
        R = p.num_regions
        if R == 0:
            return dict(regions=arr, params=p, mask=None, labels=None,
                        ncomp=torch.zeros(0, dtype=torch.int32, device=dev),
                        stats=torch.zeros((0, 5), dtype=torch.int32, device=dev),
                        thr=torch.zeros(0, dtype=torch.int32, device=dev),
                        area=torch.zeros(0, dtype=torch.int64, device=dev),
                        sums=torch.zeros((0, 4), dtype=torch.int64, device=dev),
                        contours=torch.zeros((0, 4), dtype=torch.int64, device=dev))
        regs_dev = torch.from_numpy(arr[:R].view(np.uint8).reshape(-1)).to(dev, non_blocking=False)
        mask = torch.empty(p.total_pixels, dtype=torch.uint8, device=dev) if want_mask else None
        labels = torch.empty(p.total_pixels, dtype=torch.int32, device=dev) if want_labels else None
        ncomp = torch.empty(R, dtype=torch.int32, device=dev)
        stats = torch.empty((p.total_stats_rows, 5), dtype=torch.int32, device=dev)
        thr = torch.empty(R, dtype=torch.int32, device=dev)
        area = torch.empty(R, dtype=torch.int64, device=dev)
        sums = contours = None
        if method == QC_SEG_COMPOSITE:
            need = self.lib.qc_segment_workspace_bytes(C.byref(p))
            ws = self.workspace("seg", need)
            rc = self.lib.qc_segment(self.h, imgs.data_ptr(), C.byref(p), regs_dev.data_ptr(), ws.data_ptr(), ws.numel(),
                                     _ptr(mask), _ptr(labels), ncomp.data_ptr(), stats.data_ptr(), thr.data_ptr(),
                                     area.data_ptr(), _stream(dev))
            self._check(rc, "qc_segment")
        else:
            sums = torch.empty((R, 4), dtype=torch.int64, device=dev) if want_sums else None
            contours = torch.empty((R, 4), dtype=torch.int64, device=dev) if want_contours else None
            ak, pitch, maxb = (None, 0, 0)
            if method == QC_SEG_REF_ADAPTIVE:
                ak, pitch, maxb = self.adaptive_kernel_table(max(11, min(p.max_region_w, p.max_region_h) // 10) | 1)
            need = self.lib.qc_segment_ref_workspace_bytes(C.byref(p), method)
            ws = self.workspace("seg", need)
            rc = self.lib.qc_segment_ref(self.h, method, imgs.data_ptr(), C.byref(p), regs_dev.data_ptr(), _ptr(ak), pitch,
                                         maxb, adaptive_c, ws.data_ptr(), ws.numel(), _ptr(mask), _ptr(labels),
                                         ncomp.data_ptr(), stats.data_ptr(), thr.data_ptr(), area.data_ptr(), _ptr(sums),
                                         _ptr(contours), _stream(dev))
            self._check(rc, "qc_segment_ref")
        return dict(regions=arr, params=p, regs_dev=regs_dev, mask=mask, labels=labels, ncomp=ncomp, stats=stats,
                    thr=thr, area=area, sums=sums, contours=contours)

    # -- SPC -------------------------------------------------------------------------------
    # -- K6 --------------------------------------------------------------------------------
    def rules_cluster(self, box: torch.Tensor, cls: torch.Tensor, count: torch.Tensor,
                      keep: Optional[torch.Tensor] = None, overlap_threshold: float = 0.3) -> torch.Tensor:
        """box int32 [n,cap,4], cls int32 [n,cap], count int32 [n], keep uint8 [n,cap] or None -> labels int32 [n,cap]
        (inference/postprocess.py:233-303: -1 isolated, >= 0 component of its class, -2 not clustered)."""
        assert box.is_cuda and box.dtype == torch.int32 and box.is_contiguous() and box.shape[-1] == 4
        assert cls.dtype == torch.int32 and cls.is_contiguous() and count.dtype == torch.int32
        n, cap = box.shape[:2]
        out = torch.empty((n, cap), dtype=torch.int32, device=box.device)
        if keep is not None:
            assert keep.dtype == torch.uint8 and keep.is_contiguous()
        rc = self.lib.qc_rules_cluster(self.h, box.data_ptr(), cls.data_ptr(), keep.data_ptr() if keep is not None else None,
                                       count.data_ptr(), n, cap, float(overlap_threshold), out.data_ptr(),
                                       _stream(box.device))
        self._check(rc, "qc_rules_cluster")
        return out

    def rules_apply(self, rows: np.ndarray, count: np.ndarray, shapes: np.ndarray, min_conf: float = 0.5,
                    max_area_pct: float = 50.0, elongated_classes: int = 0, overlap_threshold: float = 0.3) -> dict:
        """The whole post-NMS detection chain (refine -> filter -> merge -> validate) of a batch on the device.
        rows: RULES_ROW_DTYPE [n, cap]; count int32 [n]; shapes int32 [n, 2] = (height, width).  One upload, one
        launch, one download; returns host arrays {refined, rows, member, count}."""
        assert rows.dtype == RULES_ROW_DTYPE and rows.ndim == 2
        n, cap = rows.shape
        count = np.ascontiguousarray(count, np.int32)
        shapes = np.ascontiguousarray(shapes, np.int32).reshape(n, 2)
        nb_rows, nb_ref, nb_out = n * cap * RULES_ROW_DTYPE.itemsize, n * cap * RULES_REFINED_DTYPE.itemsize, \
            n * cap * RULES_OUT_DTYPE.itemsize
        # one packed upload (rows | count | shapes) and one packed download (refined | out rows | member | count)
        up = np.empty(nb_rows + 4 * n + 8 * n, np.uint8)
        up[:nb_rows] = np.ascontiguousarray(rows).view(np.uint8).reshape(-1)
        up[nb_rows:nb_rows + 4 * n] = count.view(np.uint8)
        up[nb_rows + 4 * n:] = shapes.view(np.uint8).reshape(-1)
        d_in = torch.from_numpy(up).to(self.device)
        d_out = torch.empty(nb_ref + nb_out + 4 * n * cap + 4 * n, dtype=torch.uint8, device=self.device)
        p_in, p_out = d_in.data_ptr(), d_out.data_ptr()
        rc = self.lib.qc_rules_apply(self.h, p_in, p_in + nb_rows, p_in + nb_rows + 4 * n, n, cap, float(min_conf),
                                     float(max_area_pct), int(elongated_classes), float(overlap_threshold),
                                     p_out, p_out + nb_ref, p_out + nb_ref + nb_out,
                                     p_out + nb_ref + nb_out + 4 * n * cap, _stream(self.device))
        self._check(rc, "qc_rules_apply")
        host = d_out.cpu().numpy()
        return {"refined": host[:nb_ref].view(RULES_REFINED_DTYPE).reshape(n, cap),
                "rows": host[nb_ref:nb_ref + nb_out].view(RULES_OUT_DTYPE).reshape(n, cap),
                "member": host[nb_ref + nb_out:nb_ref + nb_out + 4 * n * cap].view(np.int32).reshape(n, cap),
                "count": host[nb_ref + nb_out + 4 * n * cap:].view(np.int32)}

    # -- K10 -------------------------------------------------------------------------------
    def jpeg_info(self, buf: np.ndarray) -> JpegInfo:
        """Header of one encoded stream (host only).  Raises QcGraftError for anything K10 does not decode."""
        buf = np.ascontiguousarray(np.frombuffer(buf, np.uint8) if isinstance(buf, (bytes, bytearray)) else buf, np.uint8)
        info = JpegInfo()
        self._check(self.lib.qc_jpeg_info_host(self.h, buf.ctypes.data, buf.size, C.byref(info)), "qc_jpeg_info_host")
        return info

    def jpeg_decode(self, bufs, rgb: bool = False, threads: Optional[int] = None, infos=None):
        """cv2.imdecode(buf, IMREAD_COLOR) for a batch of baseline JPEG streams (1-D uint8 arrays or bytes), bit for
        bit; rgb=True also folds the reference's BGR2RGB in.  Returns one uint8 CUDA tensor [n,h,w,3] when all
        frames have the same size, else a list of [h,w,3] tensors (views of one buffer)."""
        arrs = [np.ascontiguousarray(np.frombuffer(b, np.uint8) if isinstance(b, (bytes, bytearray)) else b, np.uint8)
                for b in bufs]
        n = len(arrs)
        if n == 0:
            return []
        if infos is None:
            infos = [self.jpeg_info(a) for a in arrs]
        cinfo = (JpegInfo * n)(*infos)
        nstage = int(self.lib.qc_jpeg_staging_bytes(cinfo, n))
        nws = int(self.lib.qc_jpeg_workspace_bytes(cinfo, n))
        ev = getattr(self, "_jpeg_event", None)
        if ev is not None:
            ev.synchronize()                               # the previous call's upload has left the staging buffer
        st = getattr(self, "_jpeg_staging", None)
        if st is None or st.numel() < nstage:
            st = self._jpeg_staging = torch.empty(nstage, dtype=torch.uint8).pin_memory()
        ws = self.workspace("jpeg", nws)
        sizes = [(i.height, i.width) for i in infos]
        offs, o = [], 0
        for h, w in sizes:
            offs.append(o)
            o += h * w * 3
        dst = torch.empty(o, dtype=torch.uint8, device=self.device)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_size_t * n)(*[a.size for a in arrs])
        coffs = (C.c_int64 * n)(*offs)
        cstr = (C.c_int32 * n)(*[3 * w for _, w in sizes])
        if threads is None:
            threads = max(1, min(n, (os.cpu_count() or 1), 32))
        rc = self.lib.qc_jpeg_decode(self.h, ptrs, lens, n, cinfo, int(threads), st.data_ptr(), st.numel(),
                                     ws.data_ptr(), ws.numel(), dst.data_ptr(), coffs, cstr, int(bool(rgb)),
                                     _stream(self.device))
        self._check(rc, "qc_jpeg_decode")
        self._jpeg_event = torch.cuda.Event()
        self._jpeg_event.record(torch.cuda.current_stream(self.device))
        if len(set(sizes)) == 1:
            return dst.view(n, sizes[0][0], sizes[0][1], 3)
        return [dst[o_:o_ + h * w * 3].view(h, w, 3) for o_, (h, w) in zip(offs, sizes)]

    def spc_stats(self, det: torch.Tensor, count: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n, md = det.shape[0], det.shape[1]
        if out is None:
            out = torch.empty(9, dtype=torch.float64, device=det.device)
        rc = self.lib.qc_spc_stats(self.h, det.data_ptr(), count.data_ptr(), n, md, out.data_ptr(),
                                   _stream(det.device))
        self._check(rc, "qc_spc_stats")
        return out
