#!/usr/bin/env python
"""bench.py -- end-to-end images/sec of the inspection hot path at 640^2 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "cfg2"): YOLOv8s + ResNet-50, 64 synthetic 640x640 RGB frames
per GPU per step; a step = K1 preprocess -> YOLOv8s -> K2 decode+NMS -> K3b PIL-exact whole-frame resize +
K3 ROI-align of every kept box -> ResNet-50 on both -> K4 segmentation of every kept ROI -> K5 SPC vector
(+ NCCL all-reduce of that 72-byte vector when N > 1).  Detections come from injected head logits
(random-init Detect heads fire nothing).  Weak scaling: every rank processes its own
64 frames per step, no data-path collective.

One JSON line on stdout (rank 0):
  value     images/s with the step's frames already resident in HBM (3 rotating batches = 236 MB,
            larger than L2) -- CUDA events per step, max over ranks
  e2e       images/s through the same public call with HOST (pinned) frames: the H2D copy of the
            step's frames and the D2H reads of its results are inside the timed region
  roofline  the dominant hand-written kernel of the step (K1: most bytes), algorithmic bytes /
            its CUDA-event duration measured in the timed region, vs MEASURED_PEAKS.json
  kernels   per-kernel achieved GB/s on their own BASELINE configs (K1 cfg2, K2 cfg3, K3, K3b, K4 cfg4),
            measured after the timed region with L2-defeating rotation
  cpu_baseline  the oracle port of the reference CPU path on the host cores (bounded sample)
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
WORKLOAD = ("cfg2: YOLOv8s + ResNet-50, 64 synthetic 640x640 RGB frames per GPU per step, "
            "~30 injected candidates/image, K1->YOLOv8s->K2->K3b+K3->ResNet-50->K4->K5")


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


def build_cpu_pipeline(seed=0):
    from qcgraft import nets
    from oracle.pipeline import CpuReferencePipeline
    y = nets.build_yolo("s", NC, seed)
    for seq in list(y.detect.cv2) + list(y.detect.cv3):
        seq[-1].bias.data.zero_()
    return CpuReferencePipeline(y, nets.build_resnet50(NC, seed), nc=NC, imgsz=IMGSZ)


def cpu_reference_rate(n_images: int, warm: int = 1, seed0: int = 0):
    """Sequential per-image loop, images_per_second = N / sum(t) (inference/detector.py:331-374)."""
    from qcgraft import synth
    cp = build_cpu_pipeline()
    imgs = synth.make_frames(seed0, n_images + warm)
    inj = synth.make_head(1000 + seed0, n_images + warm, IMGSZ, IMGSZ, CANDS_PER_IMAGE)
    times = []
    for k in range(n_images + warm):
        t = time.time()
        cp.predict(imgs[k], inject=[l[k:k + 1] for l in inj])
        if k >= warm:
            times.append(time.time() - t)
    return len(times) / sum(times), times


def run_reference(args, rank):
    if rank != 0:
        return
    per_step = 4
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it can
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)
    cores = torch.get_num_threads()
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_rate(1, warm=1)
    t0 = time.time()
    all_t = []
    for s in range(args.steps):
        _, ts = cpu_reference_rate(per_step, warm=0, seed0=10 * s)
        all_t += ts
    total = time.time() - t0
    v = len(all_t) / sum(all_t)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * sum(all_t) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reference_arm": "oracle port of the reference CPU path "
                       "(pure-Python reference cannot travel to the GPU box; ultralytics/skimage absent)",
                       "images_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{len(all_t)} frames of cfg2, sequential predict() loop, {total:.1f}s wall"},
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
    return out


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
                              "note": "11 launches; union-find merge/rank phases are L2 / shared-memory latency bound"}


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


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=32)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup

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
    pipe = InspectionPipeline("s", NC, device=local_rank, imgsz=IMGSZ)
    pipe.zero_detect_bias()
    qc = pipe.qc
    peak, peak_src = measured_peaks()

    # ---- synthetic inputs: 3 rotating batches per rank (236 MB > 126 MB L2) --------------------
    NB = 3
    host, devb, inj = [], [], []
    for b in range(NB):
        seed = (rank * NB + b) * BATCH
        fr = torch.from_numpy(synth.make_frames(seed, BATCH)).pin_memory()
        host.append(fr)
        devb.append(fr.to(dev))
        lv = synth.make_head(5000 + seed, BATCH, IMGSZ, IMGSZ, CANDS_PER_IMAGE)
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

    def step_resident(i, last=False):
        h = pipe.run_async(devb[i % NB], inject=inj[i % NB])
        prev, pending[0] = pending[0], h
        r = finish(prev) if prev is not None else None
        if last:
            r = finish(h)
            pending[0] = None
        return r

    def step_e2e(i, last=False):
        # pinned HOST frames straight into the public call: the pipeline uploads them on its copy stream (H2D of
        # this step's frames overlaps the previous step's compute) -- the copy is inside the timed region
        h = pipe.run_async(host[i % NB], inject=inj[i % NB])
        prev, pending[0] = pending[0], h
        r = finish(prev) if prev is not None else None
        if last:
            r = finish(h)
            pending[0] = None
        return r

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
    tot_ms, wall_ms, launches, last = timed(step_resident, args.steps, record=True)
    stage_ms = pipe.stage_ms()
    crops = getattr(pipe, "last_num_crops", 0)
    for i in range(args.warmup):
        step_e2e(i, last=(i == args.warmup - 1))
    e2e_ms, e2e_wall_ms, _, last_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    imgs_total = world * BATCH * args.steps
    # events bracket every step on the launching stream; host-side gaps between steps are inside the
    # bracket of the next step only if the GPU is busy, so use the wall clock when it is larger
    value = imgs_total / (max(tot_ms, wall_ms) / 1000)
    e2e_value = imgs_total / (max(e2e_ms, e2e_wall_ms) / 1000)

    if rank != 0:
        qdist.shutdown()
        return

    k1_bytes = BATCH * (IMGSZ * IMGSZ * 3 + 3 * IMGSZ * IMGSZ * 2)
    k1_ms = stage_ms.get("k1_preprocess")
    roofline = {"bound": "hbm", "kernel": "k1_pack_stream (K1 preprocess, most bytes of the step)",
                "achieved": k1_bytes / k1_ms / 1e6 if k1_ms else None, "peak": peak, "unit": "GB/s",
                "frac": (k1_bytes / k1_ms / 1e6 / peak) if k1_ms else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full
                # (profiles/r01_f_ncu_full_summary.md: 78.66 MB read + 102.12 MB written inside the kernel)
                "traffic": 180780032,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes,
                "ms_per_launch": k1_ms}
    d2h = int(last.det.nbytes + last.count.nbytes + 72 + (last.roi_probs.nbytes if last.roi_probs is not None else 0)
              + (last.global_probs.nbytes if last.global_probs is not None else 0)
              + (last.seg_area.nbytes + last.seg_thr.nbytes + last.seg_ncomp.nbytes if last.seg_area is not None else 0))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max(tot_ms, wall_ms) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 nets / u8+i32+f32 kernels", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "global_batch": BATCH * world,
                       "l2": "3 rotating resident batches per rank (236 MB > 126 MB L2)",
                       "parallelism": f"dp{world}", "detections_per_step": int(last.count.sum()),
                       "rois_classified_per_step": crops},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * IMGSZ * IMGSZ * 3,
                    "d2h_bytes_per_step": d2h, "ms_per_step": max(e2e_ms, e2e_wall_ms) / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "stage_ms": stage_ms, "device_ms_per_step": tot_ms / args.steps,
            "spc": {k: float(v) for k, v in zip(qdist.SPC_FIELDS, spc_total.cpu().numpy())}}
    if not args.no_kernels:
        try:
            line["kernels"] = kernel_microbench(qc, peak)
        except Exception as e:   # the headline line must still be printed
            line["kernels"] = {"error": str(e)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            pass
        t0 = time.time()
        v, ts = cpu_reference_rate(args.cpu_sample, warm=1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{len(ts)} frames of cfg2 (+1 warm-up), sequential predict() loop, "
                                          f"{time.time() - t0:.1f}s wall, host cpu_count={os.cpu_count()}"}
    emit(line)
    qdist.shutdown()


if __name__ == "__main__":
    main()
