"""Batched device pipeline for the inspection hot path:

    uint8 HWC frames --K1--> bf16 NCHW --YOLOv8 (torch)--> head tensors --K2--> kept boxes
        --K3--> 224x224 crops --ResNet-50 (torch)--> class probabilities
        --K4--> per-detection defect masks / areas / component stats
        --K5--> SPC 9-vector (payload of the one all-reduce per step)

It is the batched form of the reference's per-image chain
``QualityControlDetector.predict`` (inference/detector.py:72-126) ->
``EnsemblePredictor.predict`` (models/ensemble.py:71-105) -> ``YOLODetector.predict``
(models/yolo_model.py:84-126) / ``ResNetClassifier.predict`` (models/resent_model.py:132-179) ->
``ImageSegmentator.segment_defects`` (inference/segmentation.py:39-146).  One device->host copy of
the packed detections sits in the middle (the reference syncs there too, models/yolo_model.py:133)
because the ROI tables for K3/K4 are planned on the host.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ffi, nets

CLASS_NAMES = ["crack", "scratch", "dent", "discoloration", "contamination"]   # models/yolo_model.py:59-61
DARK_CLASSES = (0, 1)            # crack, scratch -> THRESH_BINARY_INV / '<' (inference/segmentation.py:229,278)
ADAPTIVE_CLASSES = (0, 1)        # crack, scratch -> 'adaptive'; every other class -> 'threshold' (:187-195; dent and
                                 # contamination through the reference's threshold fallback, :175-183)
SEG_FACTORS = np.array([0.9, 0.9, 1.1, 0.8, 0.85], np.float64)                   # :218-224 by class id
RESNET_MEAN = (0.485, 0.456, 0.406)
RESNET_STD = (0.229, 0.224, 0.225)


def chunk_plan(total: int, sizes: Sequence[int]) -> List[tuple]:
    """Static-shape launches for ``total`` classifier rows: [(graph batch size, rows used)].  ``sizes`` is sorted
    largest first; every launch but the last is filled completely, the remainder rounds up to the smallest size."""
    out, left = [], int(total)
    while left > 0:
        cs = next((c for c in sizes if c <= left), sizes[-1])
        out.append((cs, min(cs, left)))
        left -= min(cs, left)
    return out


def classifier_layout(total: int, sizes: Sequence[int]):
    """Where the ``total`` classifier rows of a batch live in the input arena, so that every static-shape graph always
    reads the same place (graphs are keyed by buffer address):

        arena[0 : CH)                 the remainder buckets at fixed offsets: 128 at 0, 64 at 128, 32 at 192, 16 at 224,
                                      8 at 240 and a second 8 at 248 (chunk_plan uses each size once, 8 at most twice)
        arena[CH*(j+1) : CH*(j+2))    full chunk j of CH = sizes[0] rows

    Returns (rows, views): rows[r] = arena row of classifier row r (int64 [total]); views = [(arena offset, graph batch
    size, rows used)] in classifier-row order."""
    CH = sizes[0]
    fixed, o = {}, 0
    for c in sizes[1:]:
        fixed[c] = o
        o += c
    second_small = o                    # the second bucket of the smallest size
    assert second_small + sizes[-1] <= CH
    rows = np.empty(total, np.int64)
    views, r, nfull, small_used = [], 0, 0, 0
    for cs, used in chunk_plan(total, sizes):
        if cs == CH:
            off = CH * (nfull + 1)
            nfull += 1
        elif cs == sizes[-1]:
            off = fixed[cs] if small_used == 0 else second_small
            small_used += 1
        else:
            off = fixed[cs]
        rows[r:r + used] = off + np.arange(used)
        views.append((off, cs, used))
        r += used
    return rows, views


@dataclass
class BatchResult:
    n: int
    det: np.ndarray                 # [n, max_det, 6] float32 (x1,y1,x2,y2,conf,cls), original-image coords
    count: np.ndarray               # [n] int32
    anchor: Optional[np.ndarray]    # [n, max_det] int32
    roi_img: np.ndarray             # [K] image index of each ROI (detections with a non-empty box)
    roi_det: np.ndarray             # [K] detection index inside its image
    roi_box: np.ndarray             # [K,4] int32 x1,y1,x2,y2 (int()-truncated, models/yolo_model.py:138)
    roi_probs: Optional[np.ndarray] = None      # [K, nc] float32 softmax of the ROI classifier
    global_probs: Optional[np.ndarray] = None   # [n, nc] whole-frame classifier
    seg_area: Optional[np.ndarray] = None       # [K] int64 kept mask pixels (== np.sum(mask>0))
    seg_thr: Optional[np.ndarray] = None        # [K] Otsu threshold / Otsu bin (threshold method) / block size (adaptive)
    seg_ncomp: Optional[np.ndarray] = None      # [K]
    seg_method: Optional[np.ndarray] = None     # [K] ffi.QC_SEG_* the ROI went through
    seg_sums: Optional[np.ndarray] = None       # [K,4] int64 confidence integers (threshold: gray sums; adaptive: Canny overlap)
    seg_contours: Optional[np.ndarray] = None   # [K,4] int64 (n_axis, n_diag, n_contours, 0): perimeter = n_axis + sqrt(2) n_diag
    spc: Optional[np.ndarray] = None            # [9] float64
    seg: Optional[dict] = None                  # device tensors of K4 (mask/labels/stats) when requested
    timings_ms: dict = field(default_factory=dict)


class InspectionPipeline:
    def __init__(self, yolo_scale: str = "s", nc: int = 5, device: int = 0, conf: float = 0.7, iou: float = 0.5,
                 imgsz: int = 640, max_det: int = 300, net_dtype=torch.bfloat16, seed: int = 0,
                 classify_rois: bool = True, classify_global: bool = True, letterbox_auto: bool = False,
                 use_graphs: bool = True, seg_mode: str = "reference", build_nets: bool = True):
        self.qc = ffi.Context.current(device)
        self.dev = torch.device("cuda", device)
        self.nc, self.conf, self.iou, self.imgsz, self.max_det = nc, conf, iou, imgsz, max_det
        self.net_dtype = net_dtype
        self.classify_rois, self.classify_global = classify_rois, classify_global
        self.letterbox_mode = ffi.QC_PRE_LETTERBOX_AUTO if letterbox_auto else ffi.QC_PRE_LETTERBOX
        torch.backends.cudnn.benchmark = True
        self.yolo_scale = yolo_scale
        self._plans = {}
        # nets replay as CUDA graphs at static shapes; K1 / K3 write straight into the graph inputs
        self.use_graphs = use_graphs
        self.yolo = self.resnet = self.yolo_g = self.resnet_g = None
        self.yolo_source = self.resnet_source = None
        if build_nets:          # seeded random-init nets (synthetic benchmarks and tests)
            self.install_yolo(nets.build_yolo(yolo_scale, nc, seed), f"synthetic:yolov8{yolo_scale}:seed{seed}")
            self.install_resnet(nets.build_resnet50(nc, seed), f"synthetic:resnet50:seed{seed}")
        self.crop_chunks = (256, 128, 64, 32, 16, 8)      # static ResNet batch sizes (largest first)
        self.seg_max_comp = 1024             # first-pass stats rows per ROI (overflowing ROIs are redone)
        # "reference": per-class methods of inference/segmentation.py:185-195 (K7); "composite": the north-star K4
        # chain (Otsu -> 3x3 morphology -> CCL) for every class
        assert seg_mode in ("reference", "composite")
        self.seg_mode = seg_mode
        self.seg_contours = True             # K8: perimeter / contour count of every ROI mask on the device
        self.record = False          # bench.py: CUDA-event pairs around every stage
        self.events = {}
        # host frames: H2D on a copy stream into one of two staging buffers, so the copy of batch i+1 overlaps the
        # compute of batch i (run_async is called before the previous batch is finalised)
        self.copy_stream = torch.cuda.Stream(self.dev)
        # the segmentation of a batch runs on its own stream, underneath the classifier (submit_rest)
        self.seg_stream = torch.cuda.Stream(self.dev)
        self.overlap_seg = True
        self._stage = {}             # (shape) -> [buffers, free events, next index]

    def install_yolo(self, module, source: str):
        """YOLOv8 module (nets.build_yolo / nets.load_yolo_checkpoint) -> device, net dtype, channels_last, graphs."""
        self.yolo = module.eval().to(self.dev, self.net_dtype).to(memory_format=torch.channels_last)
        self.yolo_g = nets.GraphedNet(self.yolo, self.dev, self.net_dtype)
        self.yolo_source = source
        self.yolo_scale = getattr(module, "scale", self.yolo_scale)

    def install_resnet(self, module, source: str):
        r = nets.fold_batchnorm(module.eval()).to(self.dev, self.net_dtype).to(memory_format=torch.channels_last)
        if self.net_dtype in (torch.bfloat16, torch.float16):
            r = nets.fuse_resnet_epilogues(r)          # cuDNN conv+bias(+add)+ReLU
        self.resnet = r
        self.resnet_g = nets.GraphedNet(self.resnet, self.dev, self.net_dtype)
        self.resnet_source = source

    class _Timed:
        def __init__(self, pipe, name):
            self.pipe, self.name = pipe, name

        def __enter__(self):
            if self.pipe.record:
                self.s = torch.cuda.Event(enable_timing=True)
                self.e = torch.cuda.Event(enable_timing=True)
                self.s.record()
            return self

        def __exit__(self, *a):
            if self.pipe.record:
                self.e.record()
                self.pipe.events.setdefault(self.name, []).append((self.s, self.e))

    def _timed(self, name):
        return InspectionPipeline._Timed(self, name)

    def stage_ms(self) -> dict:
        """mean milliseconds per call of every recorded stage (call after a synchronize)."""
        return {k: float(np.mean([s.elapsed_time(e) for s, e in v])) for k, v in self.events.items() if v}

    # thresholds can change at run time (api/routes.py:354-388 -> models/yolo_model.py:252-260)
    def update_thresholds(self, confidence: Optional[float] = None, nms: Optional[float] = None):
        if confidence is not None:
            self.conf = confidence
        if nms is not None:
            self.iou = nms

    def zero_detect_bias(self):
        """Synthetic-bench helper: remove Detect's bias_init so injected logits set the scores."""
        for seq in list(self.yolo.detect.cv2) + list(self.yolo.detect.cv3):
            seq[-1].bias.data.zero_()

    def plan_for(self, h: int, w: int) -> ffi.PrePlan:
        key = (h, w)
        if key not in self._plans:
            self._plans[key] = self.qc.make_pre_plan(self.letterbox_mode, h, w, self.imgsz, self.imgsz)
        return self._plans[key]

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def detect(self, frames: torch.Tensor, inject: Optional[Sequence[torch.Tensor]] = None,
               head_override: Optional[Sequence[torch.Tensor]] = None, swap_rb: bool = True):
        if self.yolo is None and head_override is None:
            raise ffi.QcGraftError("YOLO model not loaded")
        """K1 -> YOLOv8 -> K2 on a same-shape uint8 [n,h,w,3] CUDA batch.  The frames are in the
        channel order the caller handed to ``QualityControlDetector.predict``; the reference flips
        channels three times before the net sees them (inference/detector.py:234,
        models/yolo_model.py:80, ultralytics [..., ::-1]) = one net flip -> swap_rb=True."""
        n, h, w = frames.shape[:3]
        plan = self.plan_for(h, w)
        xshape = (n, 3, plan.dst_h, plan.dst_w)
        xbuf = self.yolo_g.input_buffer(xshape) if self.use_graphs else None
        ev = None
        if self.record:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.events.setdefault("k1_preprocess", []).append(ev)
        x = self.qc.preprocess(frames, plan, swap_rb=swap_rb, out_dtype=self.net_dtype, out=xbuf, events=ev)
        if head_override is not None:
            levels = list(head_override)
        else:
            with self._timed("yolo_net"):
                if self.use_graphs:
                    levels = self.yolo_g.run(xshape)
                else:
                    levels = self.yolo(x.contiguous(memory_format=torch.channels_last))
                if inject is not None:
                    levels = [l + j for l, j in zip(levels, inject)]
                levels = [l.contiguous() for l in levels]
        p = self.qc.yolo_params(levels, self.nc, conf=self.conf, iou=self.iou, max_det=self.max_det, plan=plan)
        with self._timed("k2_decode_nms"):
            det, anchor, count, ncand = self.qc.yolo_decode_nms(levels, p)
        return det, anchor, count, ncand

    @staticmethod
    def rois_from_detections(det: np.ndarray, count: np.ndarray):
        """int() truncation of models/yolo_model.py:138; empty boxes are skipped like
        inference/segmentation.py:82-83 (roi.size == 0)."""
        n = det.shape[0]
        k = np.arange(det.shape[1])[None, :] < count[:, None]
        box = det[..., :4].astype(np.int32)                       # truncation toward zero (coords >= 0)
        ok = k & (box[..., 2] > box[..., 0]) & (box[..., 3] > box[..., 1])
        img, di = np.nonzero(ok)
        return img.astype(np.int32), di.astype(np.int32), box[img, di]

    @torch.no_grad()
    def submit_detect(self, frames: torch.Tensor, inject=None, head_override=None, processed: bool = False) -> "DetectHandle":
        """First half of a batch: (upload) -> K1 -> YOLOv8 -> K2 -> K5, plus the asynchronous device->host copy of the
        packed detections into pinned memory.  Nothing here waits for the GPU."""
        qc = self.qc
        stage_slot = None
        if not frames.is_cuda:
            frames, stage_slot = self._upload(frames)
        det_d, anchor_d, count_d, _ = self.detect(frames, inject, head_override, swap_rb=not processed)
        spc_d = qc.spc_stats(det_d, count_d)
        det_h = torch.empty(det_d.shape, dtype=det_d.dtype, pin_memory=True)
        cnt_h = torch.empty(count_d.shape, dtype=count_d.dtype, pin_memory=True)
        det_h.copy_(det_d, non_blocking=True)
        cnt_h.copy_(count_d, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record()
        return DetectHandle(frames, stage_slot, det_h, cnt_h, spc_d, ready, processed)

    @torch.no_grad()
    def run_async(self, frames: torch.Tensor, inject=None, head_override=None, include_segmentation: bool = True,
                  keep_seg_tensors: bool = False, morph_iters: int = 1, min_area: int = 20,
                  processed: bool = False) -> "PendingResult":
        """frames: uint8 [n,h,w,3], CUDA or (pinned) host.  ``processed=False``: the frames are what the caller hands to
        ``QualityControlDetector.predict`` (the reference then flips channels at inference/detector.py:234,
        models/yolo_model.py:80, inside ultralytics, and at models/resent_model.py:123 -- folded into the kernels'
        swap flags); ``processed=True``: the frames are the output of ``_preprocess_image`` (already flipped once).
        = submit_detect + submit_rest; a caller that keeps several batches in flight calls the two halves itself, the
        detect half of batch i+1 BEFORE the rest half of batch i, so that the one host wait of a batch (its packed
        detections, needed to plan the ROI tables) never leaves the GPU without queued work."""
        if self.resnet is None and (self.classify_global or self.classify_rois):
            raise ffi.QcGraftError("ResNet model not loaded")
        return self.submit_rest(self.submit_detect(frames, inject, head_override, processed), include_segmentation,
                                keep_seg_tensors, morph_iters, min_area)

    @torch.no_grad()
    def submit_rest(self, h: "DetectHandle", include_segmentation: bool = True, keep_seg_tensors: bool = False,
                    morph_iters: int = 1, min_area: int = 20) -> "PendingResult":
        """Second half: wait for the batch's packed detections (the reference syncs there too,
        models/yolo_model.py:133-135), plan the ROI tables on the host, then K3b/K3 -> ResNet-50 -> K7."""
        qc = self.qc
        frames, stage_slot, processed, spc_d = h.frames, h.stage_slot, h.processed, h.spc_d
        n = frames.shape[0]
        h.ready.synchronize()
        det = h.det_h.numpy()
        count = h.cnt_h.numpy()
        roi_img, roi_det, roi_box = self.rois_from_detections(det, count)
        K = len(roi_img)
        res = BatchResult(n=n, det=det, count=count, anchor=None, roi_img=roi_img, roi_det=roi_det, roi_box=roi_box)

        # ---- K3b/K3 + ResNet-50: whole-frame classifier (models/resent_model.py:114-130, PIL-exact resize) and
        # ROI hand-off (models/ensemble.py:129-142).  Classifier rows: the n whole frames first, then the K ROIs.
        # All rows of the batch live in ONE arena [rows,3,224,224] whose static-shape views are the inputs of the
        # ResNet-50 graphs (classifier_layout); K3 fills it with ONE launch (a destination row per ROI), K3b with one
        # launch per contiguous span of whole-frame rows.
        # The segmentation needs the frames and the kept boxes only -- not the classifier -- and its kernels are
        # latency-bound chains on small ROIs: it is queued FIRST, on its own stream, so that it runs underneath the
        # ResNet-50 graphs instead of after them (joined again before the result copies).
        seg, seg_done, seg_ready, seg_host = None, None, None, {}
        if include_segmentation and K:
            cls = det[roi_img, roi_det, 5].astype(np.int32)
            side = self.seg_stream if self.overlap_seg else torch.cuda.current_stream(self.dev)
            if self.overlap_seg:
                side.wait_event(h.ready)           # K2 of THIS batch (and everything before it) is done
            with torch.cuda.stream(side):
                seg = self._segment_rois(frames, roi_img, roi_box, cls, keep_seg_tensors, morph_iters, min_area,
                                         swap_rb=not processed)
                # the segmentation's results go to pinned host memory right behind its kernels, on its own stream, with
                # their own event: a caller that assembles per-region dictionaries can start on them while the
                # classifier graphs are still running (PendingResult.seg_result)
                for gi, g in enumerate(seg):
                    for k in ("area", "thr", "ncomp", "sums", "contours") + (("mask",) if keep_seg_tensors else ()):
                        if g["seg"].get(k) is not None:
                            t = g["seg"][k]
                            hbuf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                            hbuf.copy_(t, non_blocking=True)
                            seg_host[f"{k}{gi}"] = hbuf
                seg_ready = torch.cuda.Event()
                seg_ready.record(side)
                if self.overlap_seg:
                    seg_done = seg_ready
            self.last_seg_pixels = int(sum(int(g["seg"]["params"].total_pixels) for g in seg))

        n_whole = n if self.classify_global else 0
        n_roi = K if self.classify_rois else 0
        nro = n_whole + n_roi
        self.last_num_crops = nro
        probs = None
        if nro:
            rows, views = classifier_layout(nro, self.crop_chunks)
            arena = self._classifier_arena(max(o + c for o, c, _ in views))
            # ResNet sees the caller's channel order (two flips: detector.py:234 + resent_model.py:123)
            rswap = bool(processed)
            if n_whole:
                cut = np.flatnonzero(np.diff(rows[:n_whole]) != 1) + 1          # contiguous spans of arena rows
                for a0, a1 in zip(np.concatenate(([0], cut)), np.concatenate((cut, [n_whole]))):
                    with self._timed("k3b_pil_resize"):
                        qc.pil_resize(frames[a0:a1], swap_rb=rswap, mean=RESNET_MEAN, std=RESNET_STD,
                                      out=arena[int(rows[a0]):int(rows[a0]) + int(a1 - a0)])
            if n_roi:
                r = np.concatenate((roi_img[:, None].astype(np.float32), roi_box.astype(np.float32),
                                    rows[n_whole:, None].astype(np.int32).view(np.float32)), 1)   # one upload: rois + rows
                # K3 hands its work items out in table order and the boxes of a batch differ 100x in area: biggest
                # first, so that no CTA starts a 400-pixel box when the others are finishing (-17 % on this batch);
                # every ROI carries its destination row, so the order of the table is free
                area = (roi_box[:, 2] - roi_box[:, 0]).astype(np.int64) * (roi_box[:, 3] - roi_box[:, 1])
                r = r[np.argsort(-area, kind="stable")]
                rd = torch.from_numpy(np.ascontiguousarray(r)).to(self.dev, non_blocking=True)
                rois = rd[:, :5].contiguous()
                drows = rd[:, 5].contiguous().view(torch.int32)
                with self._timed("k3_roi_align"):
                    qc.roi_align(frames, rois, swap_rb=rswap, mean=RESNET_MEAN, std=RESNET_STD, out=arena, dst_rows=drows)
            parts = []
            with self._timed("resnet_net"):
                for off, cs, used in views:
                    parts.append(torch.softmax(self._classify(arena[off:off + cs])[:used].float(), 1))
            probs = torch.cat(parts) if len(parts) > 1 else parts[0]

        if seg_done is not None:
            torch.cuda.current_stream(self.dev).wait_event(seg_done)

        # ---- final device->host reads: asynchronous copies into pinned buffers, finalised by .result() ------
        def stage(t):
            hbuf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            hbuf.copy_(t, non_blocking=True)
            return hbuf
        host = {"spc": stage(spc_d)}
        if probs is not None:
            host["probs"] = stage(probs)
        host.update(seg_host)
        done = torch.cuda.Event()
        done.record()
        if stage_slot is not None:
            stage_slot[1] = done         # the staging buffer may be overwritten once this batch's kernels are done
        return PendingResult(self, res, frames, host, done, dict(K=K, n=n, seg=seg, keep=keep_seg_tensors), seg_ready)

    def _segment_rois(self, frames, roi_img, roi_box, cls, keep: bool, morph_iters: int = 1, min_area: int = 20,
                      swap_rb: bool = True):
        """One launch sequence per segmentation method over all ROIs of the batch.  Returns the list of groups
        (ROI indices, method, device results); PendingResult.result() scatters them back into ROI order."""
        qc = self.qc
        K = len(roi_img)
        w, h = roi_box[:, 2] - roi_box[:, 0], roi_box[:, 3] - roi_box[:, 1]
        if self.seg_mode == "composite":
            specs = [("k4_segment", np.arange(K), ffi.QC_SEG_COMPOSITE, morph_iters, min_area)]
        else:
            ad = np.isin(cls, ADAPTIVE_CLASSES)
            specs = [("k7_adaptive", np.nonzero(ad)[0], ffi.QC_SEG_REF_ADAPTIVE, 1, 20),       # :284-289
                     ("k7_threshold", np.nonzero(~ad)[0], ffi.QC_SEG_REF_THRESHOLD, 2, 10)]    # :452-460
        groups = []
        for name, idx, code, iters, marea in specs:
            if len(idx) == 0:
                continue
            # biggest regions first: k7_fill hands regions out through a work counter, the band kernels stride over the
            # table -- either way the long regions should not be the last ones to start (results are scattered back by idx)
            idx = idx[np.argsort(-(w[idx].astype(np.int64) * h[idx]), kind="stable")]
            regs = dict(img=roi_img[idx], x=roi_box[idx, 0], y=roi_box[idx, 1], w=w[idx], h=h[idx],
                        polarity=(~np.isin(cls[idx], DARK_CLASSES)).astype(np.int32), thr_scale=1.0, min_area=marea)
            cap = None
            if code == ffi.QC_SEG_REF_THRESHOLD:
                c = cls[idx]
                regs["thr_scale"] = np.where((c >= 0) & (c < len(SEG_FACTORS)), SEG_FACTORS[np.clip(c, 0, len(SEG_FACTORS) - 1)], 1.0)
                regs["max_comp"] = np.full(len(idx), 2, np.int64)
            elif not keep:
                # stats rows are reserved per region; a first pass with a small reservation avoids initialising
                # ~100 MB of rows that almost never exist, regions that overflow it are redone by result()
                nblk = ((w[idx].astype(np.int64) + 1) // 2) * ((h[idx].astype(np.int64) + 1) // 2)
                cap = np.minimum(nblk + 1, self.seg_max_comp)
                regs["max_comp"] = cap
            # gray is read as the reference does AFTER inference/detector.py:234 swapped the channels
            with self._timed(name):
                seg = qc.segment(frames, regs, morph_iters=iters, swap_rb=swap_rb, want_mask=keep, want_labels=keep,
                                 method=code, want_sums=(code != ffi.QC_SEG_COMPOSITE),
                                 want_contours=(code != ffi.QC_SEG_COMPOSITE and self.seg_contours))
            groups.append(dict(name=name, idx=idx, code=code, iters=iters, seg=seg, cap=cap, regs=regs, swap_rb=swap_rb))
        return groups

    def _classifier_arena(self, rows: int) -> torch.Tensor:
        """Grow-only [rows,3,224,224] input arena of the classifier (sized in whole chunks, with headroom so that the
        graphs captured on its views stay valid from batch to batch)."""
        a = getattr(self, "_arena", None)
        if a is None or a.shape[0] < rows:
            CH = self.crop_chunks[0]
            rows = max(rows, 4 * CH) if a is None else max(rows, 2 * a.shape[0])
            self._arena = a = torch.zeros((rows, 3, 224, 224), dtype=self.net_dtype, device=self.dev)
        return a

    def _classify(self, x: torch.Tensor) -> torch.Tensor:
        if self.use_graphs:
            return self.resnet_g.run_on(x)
        return self.resnet(x.contiguous(memory_format=torch.channels_last))

    def _upload(self, host_frames: torch.Tensor):
        """uint8 [n,h,w,3] host tensor (pinned for a truly asynchronous copy) -> device staging buffer; the compute
        stream waits for the copy, the copy waits for the last batch that read the buffer."""
        key = tuple(host_frames.shape)
        st = self._stage.get(key)
        if st is None:
            st = self._stage[key] = {"slots": [[torch.empty(key, dtype=torch.uint8, device=self.dev), None]
                                               for _ in range(2)], "next": 0}
        slot = st["slots"][st["next"]]
        st["next"] ^= 1
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            if slot[1] is not None:
                self.copy_stream.wait_event(slot[1])
            slot[0].copy_(host_frames, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        cur.wait_event(ready)
        return slot[0], slot

    def run(self, frames: torch.Tensor, **kw) -> BatchResult:
        """Synchronous form: one batch in, host-side BatchResult out."""
        return self.run_async(frames, **kw).result()


@dataclass
class DetectHandle:
    """A batch between its two halves (InspectionPipeline.submit_detect / submit_rest)."""
    frames: torch.Tensor
    stage_slot: Optional[list]
    det_h: torch.Tensor
    cnt_h: torch.Tensor
    spc_d: torch.Tensor
    ready: "torch.cuda.Event"
    processed: bool


class PendingResult:
    """Handle of a batch whose result copies are still in flight; ``result()`` waits for them.  Letting the host
    run ahead (submit batch i+1, then finalise batch i) keeps the GPU busy between steps."""

    def __init__(self, pipe, res, frames, host, done, meta, seg_ready=None):
        self.pipe, self.res, self.frames, self.host, self.done, self.meta = pipe, res, frames, host, done, meta
        self.seg_ready = seg_ready
        self._final = None
        self._seg_final = False

    def seg_result(self) -> BatchResult:
        """Waits for the segmentation alone (it runs on its own stream underneath the classifier) and fills the
        ``seg_*`` fields of the batch result; ``result()`` later adds the classifier's part."""
        res, m, pipe = self.res, self.meta, self.pipe
        if self._seg_final or m["seg"] is None:
            return res
        (self.seg_ready or self.done).synchronize()
        K = m["K"]
        res.seg_area = np.zeros(K, np.int64)
        res.seg_thr = np.zeros(K, np.int32)
        res.seg_ncomp = np.zeros(K, np.int32)
        res.seg_method = np.zeros(K, np.int32)
        res.seg_sums = np.zeros((K, 4), np.int64)
        res.seg_contours = np.zeros((K, 4), np.int64)
        for gi, g in enumerate(m["seg"]):
            idx = g["idx"]
            area = self.host[f"area{gi}"].numpy().copy()
            ncomp = self.host[f"ncomp{gi}"].numpy()
            if g["cap"] is not None:
                over = np.nonzero(ncomp > g["cap"])[0]
                if len(over):   # components beyond the reservation escaped the min-area filter: redo exactly
                    sub = {k: (v[over] if isinstance(v, np.ndarray) else v) for k, v in g["regs"].items() if k != "max_comp"}
                    # on the segmentation stream: its cached workspaces may be in use by the next batch there
                    side = pipe.seg_stream if pipe.overlap_seg else torch.cuda.current_stream(pipe.dev)
                    with torch.cuda.stream(side):
                        seg2 = pipe.qc.segment(self.frames, sub, morph_iters=g["iters"], swap_rb=g["swap_rb"], want_mask=False,
                                               want_labels=False, method=g["code"])
                        area[over] = seg2["area"].cpu().numpy()
            res.seg_area[idx] = area
            res.seg_thr[idx] = self.host[f"thr{gi}"].numpy()
            res.seg_ncomp[idx] = ncomp
            res.seg_method[idx] = g["code"]
            if f"sums{gi}" in self.host:
                res.seg_sums[idx] = self.host[f"sums{gi}"].numpy()
            if f"contours{gi}" in self.host:
                res.seg_contours[idx] = self.host[f"contours{gi}"].numpy()
            if f"mask{gi}" in self.host:
                g["mask_host"] = self.host[f"mask{gi}"].numpy()      # packed mask bytes of the group, already on the host
        if m["keep"]:
            res.seg = m["seg"]
        self._seg_final = True
        return res

    def result(self) -> BatchResult:
        if self._final is not None:
            return self._final
        self.seg_result()
        self.done.synchronize()
        res, m, pipe = self.res, self.meta, self.pipe
        n, K = m["n"], m["K"]
        res.spc = self.host["spc"].numpy().copy()
        if "probs" in self.host:
            pr = self.host["probs"].numpy()
            o = 0
            if pipe.classify_global:
                res.global_probs = pr[:n].copy()
                o = n
            if pipe.classify_rois and K:
                res.roi_probs = pr[o:o + K].copy()
        self.frames = None
        self._final = res
        return res
