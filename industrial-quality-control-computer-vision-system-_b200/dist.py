"""One process per GPU: contiguous image sharding + the single per-step collective.

The hot path is embarrassingly parallel (inference/detector.py:72-126 keeps no cross-image state),
so ranks never exchange pixels, boxes or masks.  The only exchange is the all-reduce (sum) of the
9-element float64 SPC vector derived from analytics/sec_analysis.py:105-147 (see csrc/k5_spc.cu),
72 bytes per step over NCCL/NVLink -- latency-bound, issued asynchronously so it overlaps the
next stage.  ``gloo`` is used for the CPU tests of this logic.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

SPC_FIELDS = ("n_images", "sum_defects", "sum_defects_sq", "sum_confidence", "n_detections",
              "n_critical", "n_major", "n_minor", "sum_bbox_area")


def env_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split [r*N/R, (r+1)*N/R); remainders go to the low ranks."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def spc_allreduce(vec: torch.Tensor, async_op: bool = False):
    """Sum the 9-vector over ranks in place.  Returns the work handle when async_op."""
    assert vec.numel() == len(SPC_FIELDS) and vec.dtype == torch.float64
    if dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(vec, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


def spc_summary(vec) -> dict:
    """Derived SPC aggregates: c-bar of the c-chart (analytics/sec_analysis.py:182-193), mean
    confidence, severity rates."""
    v = np.asarray(vec.cpu() if torch.is_tensor(vec) else vec, np.float64)
    n_img, s_d, s_d2, s_conf, n_det, n_cr, n_ma, n_mi, s_area = v
    c_bar = s_d / max(n_img, 1.0)
    var = max(s_d2 / max(n_img, 1.0) - c_bar * c_bar, 0.0)
    return {"n_images": int(n_img), "c_bar": c_bar, "c_var": var,
            "ucl": c_bar + 3 * np.sqrt(c_bar), "lcl": max(0.0, c_bar - 3 * np.sqrt(c_bar)),
            "avg_confidence": s_conf / max(n_det, 1.0), "critical_rate": n_cr / max(n_det, 1.0),
            "major_rate": n_ma / max(n_det, 1.0), "minor_rate": n_mi / max(n_det, 1.0),
            "total_area_affected": s_area}


def max_over_ranks(value: float, device=None) -> float:
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown():
    if dist.is_initialized():
        try:
            dist.barrier()
        except Exception:
            pass
        dist.destroy_process_group()
