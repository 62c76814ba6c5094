#!/usr/bin/env python
"""bench.py -- end-to-end images/sec of the inspection hot path at 640^2 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference] [--config cfg1|cfg2|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json configs):
  cfg2 (default at N=1)  YOLOv8s + ResNet-50, 64 synthetic 640x640 RGB frames per GPU per step
  cfg5 (default at N>1)  the same per-frame work, 256 frames per GPU per step (camera-stream sharding, 8 x 256 at N=8)
  cfg1                   YOLOv8n + ResNet-50, ONE frame: latency of QualityControlDetector.predict (median / p95)
A step = K1 preprocess -> YOLOv8 -> K2 decode+NMS -> K3b PIL-exact whole-frame resize + K3 ROI-align of every kept box
-> ResNet-50 on both -> K7 per-class segmentation of every kept ROI (the reference's threshold / adaptive methods) ->
K5 SPC vector (+ NCCL all-reduce of that 72-byte vector when N > 1).  Detections come from injected head logits
(random-init Detect heads fire nothing).  Weak scaling: every rank processes its own frames, no
data-path collective.

One JSON line on stdout (rank 0):
  value        images/s with the step's frames already resident in HBM (3 rotating batches, larger than L2) -- CUDA
               events per step, max over ranks
  e2e          images/s through the same public call with HOST (pinned) frames: the H2D copy of the step's frames and
               the D2H reads of its results are inside the timed region
  roofline     the DOMINANT hand-written operator of the step (largest share of own-kernel time, from CUDA events in
               the timed region): algorithmic bytes of that call, computed from the step's actual ROI table, over its
               measured duration, vs MEASURED_PEAKS.json; `operators` lists every C-ABI call of the step the same way
  stage_split  GPU and CPU milliseconds per image and stage (pre | nets | decode_nms | roi | seg | rules) and
               `prepost_ratio` = CPU / GPU over everything but the conv nets
  kernels      per-kernel achieved GB/s on their own BASELINE configs (K1 cfg2, K2 cfg3, K3, K3b, K4 cfg4, K7),
               measured after the timed region with L2-defeating rotation
  cfg1_latency batch-1 latency of QualityControlDetector.predict (N=1 only)
  dropin_batch images/s of QualityControlDetector.predict_batch (host numpy frames -> reference result dictionaries, N=1 only)
  cpu_baseline the oracle port of the reference CPU path on the host cores (bounded sample): the reference's work
               (mock ROI classification, models/ensemble.py:181-195); `same_work_cpu` = with a real ResNet-50 pass per ROI
--impl reference times only that CPU path (rank 0; other ranks exit).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "end-to-end images/sec at 640^2 (YOLOv8s + ResNet-50 inspection hot path)"
UNIT = "images/s"
BATCH = 64
IMGSZ = 640
CANDS_PER_IMAGE = 30
NC = 5
CONFIGS = {
    "cfg1": dict(batch=1, yolo="n", name="cfg1: YOLOv8n + ResNet-50, one synthetic 640x640 RGB frame, batch=1, through "
                                         "QualityControlDetector.predict"),
    "cfg2": dict(batch=64, yolo="s", name="cfg2: YOLOv8s + ResNet-50, 64 synthetic 640x640 RGB frames per GPU per step, "
                                          "~30 injected candidates/image, K1->YOLOv8s->K2->K3b+K3->ResNet-50->K7->K5"),
    "cfg5": dict(batch=256, yolo="s", name="cfg5: camera-stream sharding, 256 synthetic 640x640 RGB frames per GPU per step "
                                           "(8 x 256 at N=8), YOLOv8s + ResNet-50, K1->YOLOv8s->K2->K3b+K3->ResNet-50->K7->K5, "
                                           "NCCL SPC-stats all-reduce"),
}
WORKLOAD = CONFIGS["cfg2"]["name"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_cpu_pipeline(seed=0, scale="s", classify_rois=False):
    from qcgraft import nets
    from qcgraft.rules import DetectionRules
    from oracle.pipeline import CpuReferencePipeline
    y = nets.build_yolo(scale, NC, seed)
    for seq in list(y.detect.cv2) + list(y.detect.cv3):
        seq[-1].bias.data.zero_()
    return CpuReferencePipeline(y, nets.build_resnet50(NC, seed), nc=NC, imgsz=IMGSZ, classify_rois=classify_rois,
                                rules=DetectionRules({}))


def cpu_reference_rate(n_images: int, warm: int = 1, seed0: int = 0, scale="s", classify_rois=False):
    """Sequential per-image loop, images_per_second = N / sum(t) (inference/detector.py:331-374).  Returns
    (images/s, per-image seconds, per-stage seconds per image)."""
    from qcgraft import synth
    cp = build_cpu_pipeline(scale=scale, classify_rois=classify_rois)
    imgs = synth.make_frames(seed0, n_images + warm)
    inj = synth.make_head(1000 + seed0, n_images + warm, IMGSZ, IMGSZ, CANDS_PER_IMAGE)
    times = []
    for k in range(n_images + warm):
        if k == warm:
            cp.stage_s = {kk: 0.0 for kk in cp.stage_s}
        t = time.time()
        cp.predict(imgs[k], inject=[l[k:k + 1] for l in inj])
        if k >= warm:
            times.append(time.time() - t)
    stages = {k: v / max(len(times), 1) for k, v in cp.stage_s.items()}
    return len(times) / sum(times), times, stages


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    per_step = 4
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it can
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)
    cores = torch.get_num_threads()
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_rate(1, warm=1, scale=cfg["yolo"])
    t0 = time.time()
    all_t, stages = [], {}
    for s_ in range(args.steps):
        _, ts, stg = cpu_reference_rate(per_step, warm=0, seed0=10 * s_, scale=cfg["yolo"])
        all_t += ts
        for k, v in stg.items():
            stages[k] = stages.get(k, 0.0) + v / args.steps
    total = time.time() - t0
    v = len(all_t) / sum(all_t)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * sum(all_t) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "reference_arm": "oracle port of the reference CPU path doing the "
                       "reference's work: sequential predict(), YOLOv8 + one whole-frame ResNet-50 pass per image, "
                       "MOCK ROI classification (models/ensemble.py:181-195), per-class segmentation, rule engine "
                       "(the pure-Python reference cannot travel to the GPU box; ultralytics/skimage absent)",
                       "images_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{len(all_t)} frames of {args.config}, sequential predict() loop, {total:.1f}s wall",
                             "stage_ms_per_image": {k: 1000 * x for k, x in stages.items()}},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def kernel_microbench(qc, peak, only=None):
    """Per-kernel achieved GB/s on the BASELINE config that stresses each kernel.  Every timed launch
    gets a different resident input set (rotation > L2) so nothing is served from L2."""
    from qcgraft import ffi, synth
    out = {}

    def time_call(fn, reps):
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps):
            ev[i][0].record(); fn(i); ev[i][1].record()
        torch.cuda.synchronize()
        ts = sorted(s.elapsed_time(e) for s, e in ev)
        return ts[len(ts) // 2], ts[0]

    rng = np.random.default_rng(0)
    want = lambda k: only is None or k in only
    if want("k1"):
        _mb_k1(qc, peak, rng, time_call, out)
    if want("k2"):
        _mb_k2(qc, peak, rng, time_call, out)
    if want("k3"):
        _mb_k3(qc, peak, rng, time_call, out)
    if want("k3b"):
        _mb_k3b(qc, peak, rng, time_call, out)
    if want("k4"):
        _mb_k4(qc, peak, rng, time_call, out)
    if want("k6"):
        _mb_k6(qc, peak, rng, time_call, out)
    if want("k7"):
        _mb_k7(qc, peak, rng, time_call, out)
    if want("k10"):
        _mb_k10(qc, peak, rng, time_call, out)
    return out


def _mb_k10(qc, peak, rng, time_call, out):
    """K10: 64 baseline JPEG uploads (640x640, 4:2:0, quality 90) -> packed BGR frames on the device, the
    cv2.imdecode of inference/detector.py:229 for a batch; next to it cv2.imdecode itself on one host core."""
    import cv2
    from qcgraft import synth
    frames = synth.make_frames(4000, BATCH)
    streams = [cv2.imencode(".jpg", np.ascontiguousarray(f[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90])[1] for f in frames]
    infos = [qc.jpeg_info(b) for b in streams]
    got = qc.jpeg_decode(streams, infos=infos)
    ok = all(np.array_equal(got[i].cpu().numpy(), cv2.imdecode(streams[i], cv2.IMREAD_COLOR)) for i in (0, BATCH // 2, BATCH - 1))
    t0 = time.perf_counter()
    for b in streams[:16]:
        cv2.imdecode(b, cv2.IMREAD_COLOR)
    cpu_ms_per_image = (time.perf_counter() - t0) * 1000 / 16
    walls = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        qc.jpeg_decode(streams, infos=infos)
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - t0) * 1000)
    qc.profile(True)
    qc.jpeg_decode(streams, infos=infos)
    torch.cuda.synchronize()
    kern = {}
    for name, ms in qc.profile_collect():
        kern[name] = kern.get(name, 0.0) + ms
    qc.profile(False)
    comp = sum(int(b.size) for b in streams)
    blocks = sum(int(i.blocks) for i in infos)
    dev_bytes = blocks * 64 * 2 + blocks * 64 * 2 + BATCH * IMGSZ * IMGSZ * 3      # coefficients in, planes out+in, frames out
    dev_ms = sum(v for k, v in kern.items() if k != "k10_h2d_copy")
    out["k10_jpeg_decode_64x640"] = {
        "images": BATCH, "compressed_bytes": comp, "bit_exact_vs_cv2": bool(ok),
        "ms_call_wall": sorted(walls)[len(walls) // 2], "images_per_s": BATCH / (sorted(walls)[len(walls) // 2] / 1000),
        "host_threads": max(1, min(BATCH, os.cpu_count() or 1, 32)),
        "device_kernels_ms": kern, "bytes": dev_bytes, "ms": dev_ms, "gbs": dev_bytes / dev_ms / 1e6 if dev_ms else None,
        "frac": dev_bytes / dev_ms / 1e6 / peak if dev_ms else None,
        "cv2_imdecode_ms_per_image_one_core": cpu_ms_per_image,
        "note": "wall time of the whole call: Huffman stage on the host threads + one H2D copy of the coefficient blocks + "
                "k10_idct + k10_color; gbs/frac are the two kernels alone (their first event also spans the H2D copy)"}


def _mb_k1(qc, peak, rng, time_call, out):
    from qcgraft import ffi, synth
    # K1 @ cfg2: 64 x 640^2, rotating 3 batches
    b = [torch.from_numpy(rng.integers(0, 256, (BATCH, IMGSZ, IMGSZ, 3), dtype=np.uint8)).cuda() for _ in range(3)]
    plan = qc.make_pre_plan(ffi.QC_PRE_LETTERBOX, IMGSZ, IMGSZ, IMGSZ, IMGSZ)
    outs = [torch.empty((BATCH, 3, IMGSZ, IMGSZ), dtype=torch.bfloat16, device="cuda") for _ in range(3)]
    for i in range(3):
        qc.preprocess(b[i], plan, out=outs[i])
    med, best = time_call(lambda i: qc.preprocess(b[i % 3], plan, out=outs[i % 3]), 12)
    by = BATCH * (IMGSZ * IMGSZ * 3 + 3 * IMGSZ * IMGSZ * 2)
    out["k1_preprocess_cfg2"] = {"bytes": by, "ms": med, "gbs": by / med / 1e6, "frac": by / med / 1e6 / peak}
    del b, outs
    # K1 with a real resize: 32 x 1080x1920 -> letterbox 640 (auto: 384x640), the non-identity path
    b = [torch.from_numpy(rng.integers(0, 256, (32, 1080, 1920, 3), dtype=np.uint8)).cuda() for _ in range(2)]
    plan = qc.make_pre_plan(ffi.QC_PRE_LETTERBOX_AUTO, 1080, 1920, IMGSZ, IMGSZ)
    outs = [torch.empty((32, 3, plan.dst_h, plan.dst_w), dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    for i in range(2):
        qc.preprocess(b[i], plan, out=outs[i])
    med, best = time_call(lambda i: qc.preprocess(b[i % 2], plan, out=outs[i % 2]), 8)
    by = 32 * (1080 * 1920 * 3 + 3 * plan.dst_h * plan.dst_w * 2)
    out["k1_letterbox_1080p"] = {"bytes": by, "ms": med, "gbs": by / med / 1e6, "frac": by / med / 1e6 / peak,
                                 "note": "exact cv2 INTER_LINEAR model; every source byte counted although a 3x downscale only taps 4/9 of them"}


def _mb_k2(qc, peak, rng, time_call, out):
    from qcgraft import ffi, synth
    # K2 @ cfg3: 32 x 1280^2 head, >= 1000 candidates/image (f32 head = reference dtype)
    heads = []
    for k in range(2):
        lv = synth.make_head(50 + k, 32, 1280, 1280, 1500)
        heads.append([l.cuda() for l in lv])
    p = qc.yolo_params(heads[0], NC)
    qc.yolo_decode_nms(heads[0], p); qc.yolo_decode_nms(heads[1], p)
    med, best = time_call(lambda i: qc.yolo_decode_nms(heads[i % 2], p), 8)
    A = 33600
    by = 32 * (69 * A * 4 + 300 * 28 + 4)
    out["k2_decode_nms_cfg3"] = {"bytes": by, "ms": med, "gbs": by / med / 1e6, "frac": by / med / 1e6 / peak,
                                 "note": "sparse decode: only the 5 class planes are streamed, box planes are "
                                         "gathered for candidates; sort+NMS phase is latency-bound"}


def _mb_k3(qc, peak, rng, time_call, out):
    from qcgraft import ffi, synth
    # K3: 512 ROIs (~80x80) of 64 frames -> 224^2 bf16
    frames = torch.from_numpy(synth.make_frames(7, 16)).cuda()
    K = 512
    r = np.zeros((K, 5), np.float32)
    r[:, 0] = rng.integers(0, 16, K)
    r[:, 1] = rng.integers(0, 500, K); r[:, 2] = rng.integers(0, 500, K)
    r[:, 3] = r[:, 1] + rng.integers(40, 120, K); r[:, 4] = r[:, 2] + rng.integers(40, 120, K)
    rois = torch.from_numpy(r).cuda()
    o3 = torch.empty((K, 3, 224, 224), dtype=torch.bfloat16, device="cuda")
    qc.roi_align(frames, rois, out=o3)
    med, best = time_call(lambda i: qc.roi_align(frames, rois, out=o3), 8)
    by = int(((r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2]) * 3).sum()) + K * 3 * 224 * 224 * 2
    out["k3_roi_align_512rois"] = {"bytes": by, "ms": med, "gbs": by / med / 1e6, "frac": by / med / 1e6 / peak}


def _mb_k3b(qc, peak, rng, time_call, out):
    from qcgraft import synth
    # K3b: 64 whole 640^2 frames -> 224^2 bf16 (PIL-exact antialiased resize + ToTensor + Normalize), rotating 3 batches
    b = [torch.from_numpy(rng.integers(0, 256, (BATCH, IMGSZ, IMGSZ, 3), dtype=np.uint8)).cuda() for _ in range(3)]
    o = torch.empty((BATCH, 3, 224, 224), dtype=torch.bfloat16, device="cuda")
    for i in range(3):
        qc.pil_resize(b[i], out=o)
    med, best = time_call(lambda i: qc.pil_resize(b[i % 3], out=o), 9)
    by = BATCH * (IMGSZ * IMGSZ * 3 + 3 * 224 * 224 * 2)
    out["k3b_pil_resize_cfg2"] = {"bytes": by, "ms": med, "gbs": by / med / 1e6, "frac": by / med / 1e6 / peak}


def _mb_k6(qc, peak, rng, time_call, out):
    # K6: same-class overlap clustering of 64 images x 300 post-NMS boxes (clustered: real merge work)
    n, cap = BATCH, 300
    ctr = rng.uniform(50, 590, (n, 30, 2))
    pick = rng.integers(0, 30, (n, cap))
    xy = np.take_along_axis(ctr, pick[..., None].repeat(2, 2), 1) + rng.normal(0, 6, (n, cap, 2))
    wh = rng.integers(8, 120, (n, cap, 2))
    box = np.concatenate((xy, xy + wh), 2).astype(np.int32)
    b = torch.from_numpy(np.ascontiguousarray(box)).cuda()
    c = torch.from_numpy(rng.integers(0, NC, (n, cap)).astype(np.int32)).cuda()
    cnt = torch.full((n,), cap, dtype=torch.int32).cuda()
    qc.rules_cluster(b, c, cnt)
    med, best = time_call(lambda i: qc.rules_cluster(b, c, cnt), 8)
    by = n * cap * (16 + 4 + 4 + 4)
    # the whole chain (refine -> filter -> cluster -> merge -> validate) on the same boxes, one launch
    from qcgraft import ffi as _ffi
    rows = np.zeros((n, cap), _ffi.RULES_ROW_DTYPE)
    rows["x1"], rows["y1"], rows["x2"], rows["y2"] = box[..., 0], box[..., 1], box[..., 2], box[..., 3]
    rows["cls"] = c.cpu().numpy()
    rows["severity"] = rng.integers(1, 4, (n, cap))
    rows["conf"] = rng.uniform(0.4, 0.99, (n, cap))
    rows["ext"] = -1
    shp = np.tile(np.array([[IMGSZ, IMGSZ]], np.int32), (n, 1))
    cn = np.full(n, cap, np.int32)
    qc.rules_apply(rows, cn, shp)
    t0 = time.perf_counter()
    for _ in range(5):
        qc.rules_apply(rows, cn, shp)
    wall = (time.perf_counter() - t0) * 1000 / 5
    qc.profile(True)
    qc.rules_apply(rows, cn, shp)
    chain_ms = sum(ms for name, ms in qc.profile_collect() if name == "k6_chain")
    qc.profile(False)
    out["k6_rules_chain_64x300"] = {"ms": chain_ms, "ms_call_wall_with_copies": wall, "bytes": n * cap * (88 + 32 + 56 + 4),
                                    "note": "one CTA per image, float64 rule arithmetic; latency-bound"}
    out["k6_rules_cluster_64x300"] = {"bytes": by, "ms": med, "gbs": by / med / 1e6, "frac": by / med / 1e6 / peak,
                                      "note": "quadratic pair test + shared-memory union-find per image: latency-bound, "
                                              "the bytes are negligible"}


def _mb_k4(qc, peak, rng, time_call, out):
    from qcgraft import ffi, synth
    # K4 @ cfg4: 16 x 2048^2 gray line-scan frames
    ls = [torch.from_numpy(np.stack([synth.make_linescan(900 + 16 * k + (i % 4)) for i in range(16)])).cuda()
          for k in range(2)]
    regs = [dict(img=i, x=0, y=0, w=2048, h=2048, min_area=20) for i in range(16)]
    res = qc.segment(ls[0], regs, morph_iters=1)
    qc.segment(ls[1], regs, morph_iters=1)
    ncomp = int(res["ncomp"].sum())
    med, best = time_call(lambda i: qc.segment(ls[i % 2], regs, morph_iters=1), 6)
    by = 16 * 6 * 2048 * 2048 + 20 * ncomp
    out["k4_segment_cfg4"] = {"bytes": by, "ms": med, "gbs": by / med / 1e6, "frac": by / med / 1e6 / peak,
                              "components": ncomp,
                              "note": "10 launches; union-find merge/rank phases are L2 / shared-memory latency bound"}


def _mb_k7(qc, peak, rng, time_call, out):
    """K7: the reference's per-class segmentation methods on 1400 detection ROIs of 64 frames (~cfg2's ROI table),
    stats-only outputs as the batched pipeline runs them."""
    from qcgraft import ffi, synth
    frames = torch.from_numpy(synth.make_frames(300, BATCH)).cuda()
    K = 1400
    w = rng.integers(16, 240, K); h = rng.integers(16, 240, K)
    x = (rng.random(K) * (IMGSZ - w)).astype(np.int64); y = (rng.random(K) * (IMGSZ - h)).astype(np.int64)
    img = rng.integers(0, BATCH, K)
    px = int((w * h).sum())
    for name, code, iters, marea in (("k7_threshold", ffi.QC_SEG_REF_THRESHOLD, 2, 10), ("k7_adaptive", ffi.QC_SEG_REF_ADAPTIVE, 1, 20)):
        regs = dict(img=img, x=x, y=y, w=w, h=h, polarity=rng.integers(0, 2, K), thr_scale=0.8, min_area=marea)
        if code == ffi.QC_SEG_REF_THRESHOLD:
            regs["max_comp"] = np.full(K, 2)
        else:
            regs["max_comp"] = np.minimum(((w + 1) // 2) * ((h + 1) // 2) + 1, 1024)
        call = lambda i: qc.segment(frames, regs, morph_iters=iters, swap_rb=True, want_mask=False, want_labels=False,
                                    method=code, want_sums=(code == ffi.QC_SEG_REF_THRESHOLD))
        call(0); call(1)
        l0 = qc.launch_count
        call(2)
        launches = qc.launch_count - l0
        med, best = time_call(call, 6)
        by = px * 3 + K * 64                      # RGB bytes of the ROIs read once + the per-ROI result rows
        out[f"{name}_1400rois"] = {"bytes": by, "ms": med, "gbs": by / med / 1e6, "frac": by / med / 1e6 / peak,
                                   "roi_pixels": px, "launches": launches,
                                   "note": "multi-pass integer / float64 work on small ROIs (blur, histogram, union-find, "
                                           "flood fill): latency- and issue-bound, the HBM roofline is reported as asked, "
                                           "it is not what limits these kernels"}


_REAL_STDOUT = None


def _capture_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, cuDNN notices) goes to stderr; the one JSON
    line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        os.write(1, data)


def cfg1_latency(calls: int = 60, warm: int = 8):
    """BASELINE cfg1: one 640x640 frame through QualityControlDetector.predict (YOLOv8n + ResNet-50), host numpy image
    in, reference result dict out -- the reference's own timing definition (inference/detector.py:331-374)."""
    from qcgraft import synth
    from qcgraft.detector import QualityControlDetector
    cfg = {"model": {"confidence_threshold": 0.7, "nms_threshold": 0.5},
           "quality_control": {"defect_classes": ["crack", "scratch", "dent", "discoloration", "contamination"]},
           "qcgraft": {"yolo_scale": "n", "device": torch.cuda.current_device(), "synthetic_weights": True}}
    d = QualityControlDetector("synthetic", "synthetic", cfg)
    d.ensemble_predictor.yolo_detector.model.zero_detect_bias()
    imgs = [synth.make_frame(7000 + i) for i in range(8)]
    inj = [l.cuda().to(torch.bfloat16) for l in synth.make_head(7100, 8, IMGSZ, IMGSZ, CANDS_PER_IMAGE)]
    ts, ndet = [], 0
    for k in range(warm + calls):
        i = k % 8
        t = time.perf_counter()
        r = d.predict(imgs[i], _inject=[l[i:i + 1] for l in inj])
        dt = time.perf_counter() - t
        if "error" in r:
            return {"error": r["error"]}
        if k >= warm:
            ts.append(dt * 1000)
            ndet += len(r["detections"])
    ts.sort()
    return {"calls": len(ts), "median_ms": statistics.median(ts), "p95_ms": ts[int(0.95 * (len(ts) - 1))],
            "min_ms": ts[0], "max_ms": ts[-1], "images_per_second": 1000 * len(ts) / sum(ts),
            "detections_per_image_after_rules": ndet / len(ts),
            "path": "QualityControlDetector.predict (YOLOv8n + ResNet-50, CUDA-graph nets, rule engine on)"}


def dropin_batch_rate(scale: str, batch: int = 64, calls: int = 4, warm: int = 2):
    """The same workload through the reference's OWN call signature: QualityControlDetector.predict_batch(list of host
    numpy frames) -> list of reference result dictionaries (inference/detector.py:128-180), i.e. host frames in,
    detections / classification / segmentation / rule engine / quality grade dictionaries out."""
    from qcgraft import synth
    from qcgraft.detector import QualityControlDetector
    cfg = {"model": {"confidence_threshold": 0.7, "nms_threshold": 0.5},
           "quality_control": {"defect_classes": ["crack", "scratch", "dent", "discoloration", "contamination"]},
           "qcgraft": {"yolo_scale": scale, "device": torch.cuda.current_device(), "synthetic_weights": True,
                       "max_batch": batch}}
    d = QualityControlDetector("synthetic", "synthetic", cfg)
    d.ensemble_predictor.yolo_detector.model.zero_detect_bias()
    imgs = [synth.make_frame(8000 + i) for i in range(batch)]
    inj = [l.cuda().to(torch.bfloat16) for l in synth.make_head(8100, batch, IMGSZ, IMGSZ, CANDS_PER_IMAGE)]
    ts, ndet = [], 0
    for k in range(warm + calls):
        t = time.perf_counter()
        rs = d.predict_batch(imgs, _inject=inj)
        dt = time.perf_counter() - t
        bad = [r["error"] for r in rs if "error" in r]
        if bad:
            return {"error": bad[0]}
        if k >= warm:
            ts.append(dt)
            ndet += sum(len(r["detections"]) for r in rs)
    return {"value": batch * len(ts) / sum(ts), "unit": UNIT, "batch": batch, "calls": len(ts),
            "ms_per_call_min": 1000 * min(ts), "ms_per_call_max": 1000 * max(ts),
            "detections_per_image_after_rules": ndet / (batch * len(ts)),
            "path": "QualityControlDetector.predict_batch: host numpy frames -> reference result dictionaries "
                    "(one device batch per call, rule chain on the device, dictionaries assembled on the host)"}


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=32)
    ap.add_argument("--kernels-only", default=None, help="comma list (k1,k2,k3,k3b,k4,k6,k7): only the micro-benchmarks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup
    if args.config is None:
        args.config = "cfg2" if max(args.gpus, 1) == 1 else "cfg5"
    cfg = CONFIGS[args.config]

    from qcgraft import dist as qdist
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", 0))
        run_reference(args, rank)
        return

    rank, local_rank, world = qdist.init()
    assert world == max(args.gpus, 1) or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    from qcgraft import ffi, synth
    from qcgraft.pipeline import InspectionPipeline

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peak, peak_src = measured_peaks()

    if args.kernels_only:
        qc0 = ffi.Context.current(local_rank)
        emit({"kernels": kernel_microbench(qc0, peak, set(args.kernels_only.split(","))), "peak": peak})
        qdist.shutdown()
        return

    if args.config == "cfg1":
        lat = cfg1_latency(max(50, args.steps * 10), max(args.warmup, 8))
        if rank == 0:
            line = {"metric": METRIC, "value": lat.get("images_per_second"), "unit": UNIT, "n_gpus": 1,
                    "steps": lat.get("calls"), "warmup": max(args.warmup, 8),
                    "ms_per_step": lat.get("median_ms"), "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "bf16 nets / u8+i32+f32+f64 kernels", "data": "synthetic",
                    "config": {"workload": cfg["name"], "l2": "8 rotating host frames, each uploaded inside predict()"},
                    "e2e": {"value": lat.get("images_per_second"), "unit": UNIT,
                            "h2d_bytes_per_step": IMGSZ * IMGSZ * 3, "d2h_bytes_per_step": None},
                    "cfg1_latency": lat}
            emit(line)
        qdist.shutdown()
        return

    B = cfg["batch"]
    pipe = InspectionPipeline(cfg["yolo"], NC, device=local_rank, imgsz=IMGSZ)
    pipe.zero_detect_bias()
    qc = pipe.qc

    # ---- synthetic inputs: 3 rotating batches per rank (>= 236 MB > 126 MB L2) ------------------
    NB = 3
    host, devb, inj = [], [], []
    for b in range(NB):
        seed = (rank * NB + b) * B
        fr = torch.from_numpy(synth.make_frames(seed, B)).pin_memory()
        host.append(fr)
        devb.append(fr.to(dev))
        lv = synth.make_head(5000 + seed, B, IMGSZ, IMGSZ, CANDS_PER_IMAGE)
        inj.append([l.to(dev, torch.bfloat16) for l in lv])
    spc_total = torch.zeros(9, dtype=torch.float64, device=dev)

    # Software pipelining: batch i+1 is submitted before batch i's result copies are awaited, so the GPU never
    # idles on the host between steps.  Every batch is still fully finalised (all D2H reads + the SPC
    # all-reduce) inside the timed region: the loop drains the last pending batch before the clock stops.
    pending = [None]

    def finish(h):
        r = h.result()
        v = torch.from_numpy(r.spc).to(dev)
        qdist.spc_allreduce(v)
        spc_total.add_(v)
        return r

    def make_step(src):
        # resident: frames already in HBM; e2e: pinned HOST frames straight into the public call -- the pipeline
        # uploads them on its copy stream (H2D of the next step's frames overlaps this step's compute).
        # Two batches are in flight: the detect half (K1 -> YOLOv8 -> K2) of batch i+1 is queued BEFORE the host waits
        # for batch i's packed detections, so the ROI planning on the host never leaves the GPU idle.
        state = {"det": None}

        def step(i, last=False):
            if state["det"] is None:
                state["det"] = pipe.submit_detect(src[i % NB], inject=inj[i % NB])
            cur = state["det"]
            state["det"] = None if last else pipe.submit_detect(src[(i + 1) % NB], inject=inj[(i + 1) % NB])
            h = pipe.submit_rest(cur)
            prev, pending[0] = pending[0], h
            r = finish(prev) if prev is not None else None
            if last:
                r = finish(h)
                pending[0] = None
            return r
        return step
    step_resident, step_e2e = make_step(devb), make_step(host)

    def timed(fn, steps, record=False):
        qdist.barrier(); torch.cuda.synchronize()
        pipe.record, pipe.events = record, {}
        l0 = qc.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.perf_counter()
        last = None
        for i in range(steps):
            ev[i][0].record()
            last = fn(i, last=(i == steps - 1))
            ev[i][1].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        qdist.barrier()
        pipe.record = False
        dev_ms = sum(s.elapsed_time(e) for s, e in ev)
        # device-side total (events) and host wall agree up to launch gaps; report the larger, max over ranks
        total_ms = qdist.max_over_ranks(max(dev_ms, 0.0), dev)
        wall_ms = qdist.max_over_ranks(wall * 1000, dev)
        return total_ms, wall_ms, qc.launch_count - l0, last

    for i in range(args.warmup):
        step_resident(i, last=(i == args.warmup - 1))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    tot_ms, wall_ms, launches, last = timed(step_resident, args.steps)
    crops = getattr(pipe, "last_num_crops", 0)
    for i in range(args.warmup):
        step_e2e(i, last=(i == args.warmup - 1))
    e2e_ms, e2e_wall_ms, _, last_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel durations of the same step, from CUDA events the library records around its own launches
    # (qc_profile_enable); a separate short pass so that the event records stay out of the timed regions ---------
    # Stage and kernel durations come from separate passes with the segmentation stream serialised behind the
    # classifier: in the timed regions above K7 runs UNDERNEATH the ResNet-50 graphs (pipeline.submit_rest), where a
    # stage's event pair would also span the other stream's work.
    pipe.overlap_seg = False
    STAGE_STEPS = max(3, min(args.steps, 6))
    step_resident(0, last=True)
    torch.cuda.synchronize()
    pipe.record, pipe.events = True, {}
    for i in range(STAGE_STEPS):
        step_resident(i, last=(i == STAGE_STEPS - 1))
    torch.cuda.synchronize()
    pipe.record = False
    # per-STEP milliseconds of every stage (sum of its calls / steps)
    stage_ms = {k: float(sum(s.elapsed_time(e) for s, e in v)) / STAGE_STEPS for k, v in pipe.events.items() if v}
    stage_calls = {k: len(v) / STAGE_STEPS for k, v in pipe.events.items() if v}
    PROF_STEPS = 3
    qc.profile(True)
    for i in range(PROF_STEPS):
        step_resident(i, last=(i == PROF_STEPS - 1))
    torch.cuda.synchronize()
    prof = qc.profile_collect()
    qc.profile(False)
    pipe.overlap_seg = True
    kernel_times = {}
    for name, ms in prof:
        k = kernel_times.setdefault(name, {"ms_per_step": 0.0, "launches_per_step": 0.0})
        k["ms_per_step"] += ms / PROF_STEPS
        k["launches_per_step"] += 1.0 / PROF_STEPS

    imgs_total = world * B * args.steps
    # events bracket every step on the launching stream; host-side gaps between steps are inside the
    # bracket of the next step only if the GPU is busy, so use the wall clock when it is larger
    value = imgs_total / (max(tot_ms, wall_ms) / 1000)
    e2e_value = imgs_total / (max(e2e_ms, e2e_wall_ms) / 1000)

    if rank != 0:
        qdist.shutdown()
        return

    # ---- operators of the step: algorithmic bytes from the step's own tables -------------------------------------
    # (every rotating batch carries the same number of candidates; the last step's ROI table stands for the step)
    roi_px = (last.roi_box[:, 2] - last.roi_box[:, 0]).astype(np.int64) * (last.roi_box[:, 3] - last.roi_box[:, 1])
    K = len(roi_px)
    A = sum((IMGSZ // s_) ** 2 for s_ in (8, 16, 32))
    adaptive = last.seg_method == ffi.QC_SEG_REF_ADAPTIVE if last.seg_method is not None else np.zeros(K, bool)
    op_bytes = {
        "k1_preprocess": B * (IMGSZ * IMGSZ * 3 + 3 * IMGSZ * IMGSZ * 2),
        "k2_decode_nms": B * ((64 + NC) * A * 2 + 300 * 28 + 4),                  # bf16 head in
        "k3b_pil_resize": B * (IMGSZ * IMGSZ * 3 + 3 * 224 * 224 * 2),
        "k3_roi_align": int(roi_px.sum()) * 3 + K * 3 * 224 * 224 * 2,
        "k7_adaptive": int(roi_px[adaptive].sum()) * 3 + int(adaptive.sum()) * 64,
        "k7_threshold": int(roi_px[~adaptive].sum()) * 3 + int((~adaptive).sum()) * 64,
    }
    operators = {}
    for k, by in op_bytes.items():
        ms = stage_ms.get(k)
        if ms:
            operators[k] = {"algorithmic_bytes_per_step": by, "ms_per_step": ms, "calls_per_step": stage_calls[k],
                            "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
    own_total = sum(v["ms_per_step"] for v in operators.values())
    for v in operators.values():
        v["share_of_own_kernel_time"] = v["ms_per_step"] / own_total if own_total else None
    dom = max(operators, key=lambda k: operators[k]["ms_per_step"]) if operators else None
    # per-kernel table (live CUDA events inside the library): algorithmic bytes of the kernels that can lead it
    thr_px, adp_px = int(roi_px[~adaptive].sum()), int(roi_px[adaptive].sum())
    k3_kernels = [k for k in kernel_times if k.startswith("k3_roi")]
    kern_bytes = {
        "k1_pack_stream": op_bytes["k1_preprocess"], "k1_resize_tma": op_bytes["k1_preprocess"],
        "k3b_resize": op_bytes["k3b_pil_resize"],
        "k7_blur": thr_px * (1 + 8),                       # u8 gray in, f64 blurred out
        "k7_hist": thr_px * 8, "k7_cmp": thr_px * 8 + thr_px // 8,
        "k7_pack_gray": int(roi_px.sum()) * (3 + 1),       # RGB in, gray out
        "k7_adaptive_small": adp_px + adp_px // 8, "k9_canny_nms": adp_px + adp_px // 4,
        "k7_fill": (2 * thr_px + adp_px) // 4,             # bit masks in and out, all calls of the step
        "k4_ccl_local": (2 * thr_px + 3 * adp_px) // 8 + (2 * thr_px + 3 * adp_px),   # bits in, one int32 per 2x2 block out
    }
    # K3 = k3_roi_win (every ROI whose taps fit three-pixel windows: bins <= 1.95) + k3_roi_tma (the rest; a scan that
    # finds nothing when no box is wider / taller than 436 px): the bytes go to the kernel that did the work
    if "k3_roi_win" in k3_kernels:
        big = (last.roi_box[:, 2] - last.roi_box[:, 0] > 436) | (last.roi_box[:, 3] - last.roi_box[:, 1] > 436)
        kern_bytes["k3_roi_win"] = int(roi_px[~big].sum()) * 3 + int((~big).sum()) * 3 * 224 * 224 * 2
        if big.any():
            kern_bytes["k3_roi_tma"] = int(roi_px[big].sum()) * 3 + int(big.sum()) * 3 * 224 * 224 * 2
    elif len(k3_kernels) == 1:
        kern_bytes[k3_kernels[0]] = op_bytes["k3_roi_align"]
    k_total = sum(v["ms_per_step"] for v in kernel_times.values())
    for name, v in kernel_times.items():
        v["share_of_own_kernel_time"] = v["ms_per_step"] / k_total if k_total else None
        by = kern_bytes.get(name)
        if by and v["ms_per_step"] > 0:
            v["algorithmic_bytes_per_step"] = by
            v["gbs"] = by / v["ms_per_step"] / 1e6
            v["frac"] = v["gbs"] / peak
    kernel_times = dict(sorted(kernel_times.items(), key=lambda kv: -kv[1]["ms_per_step"]))
    # dominant kernel = the longest single LAUNCH of the step (the per-name sums above add the 2-5 launches of the
    # latency-bound segmentation kernels on ~14 k-pixel regions, whose bound is shared memory / latency, not HBM)
    kdom = max(kernel_times, key=lambda k_: kernel_times[k_]["ms_per_step"] / max(kernel_times[k_]["launches_per_step"], 1e-9),
               default=None)
    roofline = None
    if kdom and "gbs" in kernel_times[kdom]:
        kv = kernel_times[kdom]
        roofline = {"bound": "hbm", "kernel": f"{kdom} (longest single kernel launch of the step: "
                    f"{100 * kv['share_of_own_kernel_time']:.0f} % of own-kernel time, CUDA events around every launch "
                    "inside the library; kernel_times lists every kernel)",
                    "achieved": kv["gbs"], "peak": peak, "unit": "GB/s", "frac": kv["frac"],
                    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, ncu --set full
                    # (profiles/r02_ncu_full_summary.md): k3_roi_win on 1400 boxes of the step's size distribution
                    # (tools/k3_case.py K3_MIX=1; 479 MB algorithmic there: 37.6 MB read = the tapped row segments once,
                    # 369.7 MB of the 421 MB of outputs written back by kernel end, the rest still in L2)
                    "traffic": 407300000 if kdom == "k3_roi_win" else None,
                    "traffic_note": "ncu capture of tools/k3_case.py K3_MIX=1 K3_ROIS=1400 (same kernel, same box-size "
                                    "distribution, 479 MB algorithmic): no re-reads" if kdom == "k3_roi_win" else None,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kv["algorithmic_bytes_per_step"] / kv["launches_per_step"],
                    "ms_per_launch": kv["ms_per_step"] / kv["launches_per_step"]}
    if dom:
        o = operators[dom]
        roofline = roofline or {"bound": "hbm", "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src,
                                "kernel": dom, "achieved": o["gbs"], "frac": o["frac"]}
        roofline.update({
                    "dominant_operator": {"operator": f"{dom} (one C-ABI call = one launch sequence; "
                                          f"{100 * o['share_of_own_kernel_time']:.0f} % of own-operator time)",
                                          "gbs": o["gbs"], "frac": o["frac"], "ms_per_call": o["ms_per_step"] / o["calls_per_step"],
                                          "algorithmic_bytes_per_call": o["algorithmic_bytes_per_step"] / o["calls_per_step"]},
                    "second": {"kernel": "k1_pack_stream (K1 preprocess, the step's pure streaming kernel)",
                               **{kk: operators["k1_preprocess"][kk] for kk in ("gbs", "frac", "ms_per_step")},
                               # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full
                               # (profiles/r01_f_ncu_full_summary.md)
                               "traffic": 180780032} if "k1_preprocess" in operators else None})
    gpu_stage = {"pre": stage_ms.get("k1_preprocess", 0) + stage_ms.get("k3b_pil_resize", 0),
                 "nets": stage_ms.get("yolo_net", 0) + stage_ms.get("resnet_net", 0),
                 "decode_nms": stage_ms.get("k2_decode_nms", 0), "roi": stage_ms.get("k3_roi_align", 0),
                 "seg": stage_ms.get("k7_adaptive", 0) + stage_ms.get("k7_threshold", 0) + stage_ms.get("k4_segment", 0),
                 "rules": 0.0}
    stage_split = {"gpu_ms_per_image": {k: v / B for k, v in gpu_stage.items()},
                   "note": "GPU: CUDA events around each stage, in a pass with the segmentation stream serialised (in the "
                           "timed regions K7 runs underneath the ResNet-50 graphs, so the stages sum to more than the "
                           "step); K5 SPC reduction and the host rule engine of the drop-in are not in the batched "
                           "step; CPU: wall time per stage of the port"}
    d2h = int(last.det.nbytes + last.count.nbytes + 72 + (last.roi_probs.nbytes if last.roi_probs is not None else 0)
              + (last.global_probs.nbytes if last.global_probs is not None else 0)
              + (last.seg_area.nbytes + last.seg_thr.nbytes + last.seg_ncomp.nbytes + last.seg_sums.nbytes
                 if last.seg_area is not None else 0))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max(tot_ms, wall_ms) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 nets / u8+i32+f32+f64 kernels", "data": "synthetic",
            "config": {"workload": cfg["name"], "per_gpu_batch": B, "global_batch": B * world,
                       "l2": f"3 rotating resident batches per rank ({3 * B * IMGSZ * IMGSZ * 3 // 2 ** 20} MiB > 126 MB L2)",
                       "parallelism": f"dp{world}", "detections_per_step": int(last.count.sum()),
                       "rois_classified_per_step": crops, "segmentation": pipe.seg_mode,
                       "note": "default workload: cfg2 at N=1, cfg5 (256 frames per GPU) at N>1"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * IMGSZ * IMGSZ * 3,
                    "d2h_bytes_per_step": d2h, "ms_per_step": max(e2e_ms, e2e_wall_ms) / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "operators": operators,
            "kernel_times": kernel_times,
            "stage_ms": stage_ms, "stage_split": stage_split, "device_ms_per_step": tot_ms / args.steps,
            "spc": {k: float(v) for k, v in zip(qdist.SPC_FIELDS, spc_total.cpu().numpy())}}
    if not args.no_kernels and world == 1:
        try:
            line["kernels"] = kernel_microbench(qc, peak)
        except Exception as e:   # the headline line must still be printed
            line["kernels"] = {"error": str(e)}
    if world == 1:
        try:
            del devb, host, inj
            torch.cuda.empty_cache()
            line["cfg1_latency"] = cfg1_latency()
        except Exception as e:
            line["cfg1_latency"] = {"error": str(e)}
        try:
            torch.cuda.empty_cache()
            line["dropin_batch"] = dropin_batch_rate(cfg["yolo"], B if B <= 64 else 64)
        except Exception as e:
            line["dropin_batch"] = {"error": str(e)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            pass
        t0 = time.time()
        v, ts, stg = cpu_reference_rate(args.cpu_sample, warm=1, scale=cfg["yolo"])
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{len(ts)} frames of {args.config} (+1 warm-up), sequential predict() loop doing the "
                                          f"reference's work (mock ROI classification), {time.time() - t0:.1f}s wall, "
                                          f"host cpu_count={os.cpu_count()}"}
        cpu_ms = {k: 1000 * x for k, x in stg.items()}
        stage_split["cpu_ms_per_image"] = cpu_ms
        gpu = stage_split["gpu_ms_per_image"]
        pp = [k for k in cpu_ms if k != "nets"]
        g_pp = sum(gpu.get(k, 0.0) for k in pp)
        stage_split["prepost_cpu_ms_per_image"] = sum(cpu_ms[k] for k in pp)
        stage_split["prepost_gpu_ms_per_image"] = g_pp
        stage_split["prepost_ratio"] = (sum(cpu_ms[k] for k in pp) / g_pp) if g_pp else None
        t0 = time.time()
        v2, ts2, _ = cpu_reference_rate(max(4, args.cpu_sample // 8), warm=1, scale=cfg["yolo"], classify_rois=True)
        line["same_work_cpu"] = {"value": v2, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{len(ts2)} frames, as cpu_baseline but with a real roi_align + ResNet-50 pass per "
                                           f"kept box (the work the GPU pipeline does), {time.time() - t0:.1f}s wall"}
    emit(line)
    qdist.shutdown()


if __name__ == "__main__":
    main()
