"""SPC batch statistics on top of K2's packed detections.

The reference's ``SPCAnalyzer.analyze`` (analytics/sec_analysis.py:55-103) is fed one result dict per image and
keeps a deque(100) of per-image metric rows.  Here the rows come straight from the packed ``det``/``count``
arrays of a whole batch (no dicts), ranks exchange them with one all-gather (or only their 9-vector sums with
the all-reduce of dist.py), and the chart arithmetic is the reference's:
  _extract_metrics :105-147, c/u/X-chart limits :182-228, control status :230-277, run rules 2 and 3 :279-318,
  Cp/Cpk/Cpm :320-370.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, List, Optional

import numpy as np

ROW_FIELDS = ("defect_count", "defect_rate", "avg_confidence", "critical_defects", "major_defects",
              "minor_defects", "total_area_affected")


def metric_rows(det: np.ndarray, count: np.ndarray) -> np.ndarray:
    """det [n,max_det,6] float32 rows (x1,y1,x2,y2,conf,cls), count [n] -> float64 [n,7] rows in ROW_FIELDS
    order == what _extract_metrics computes from the reference's detection dicts (int()-truncated boxes,
    models/yolo_model.py:138; severity :162-172; ensemble_confidence models/ensemble.py:186-208)."""
    n, md = det.shape[:2]
    valid = np.arange(md)[None, :] < np.asarray(count)[:, None]
    box = det[..., :4].astype(np.int32).astype(np.int64)
    w, h = box[..., 2] - box[..., 0], box[..., 3] - box[..., 1]
    area = w * h
    conf32 = det[..., 4].astype(np.float32)
    conf = conf32.astype(np.float64)
    ens = conf * 0.6 + np.minimum(conf * 1.1, 1.0) * 0.4
    na = area / (1024 * 1024)
    crit = (conf32 > np.float32(0.9)) | (na > 0.1)
    major = ~crit & ((conf32 > np.float32(0.8)) | (na > 0.05))
    minor = ~crit & ~major
    cnt = valid.sum(1).astype(np.float64)
    rows = np.zeros((n, len(ROW_FIELDS)), np.float64)
    rows[:, 0] = cnt
    rows[:, 1] = cnt
    with np.errstate(invalid="ignore", divide="ignore"):
        rows[:, 2] = np.where(cnt > 0, (ens * valid).sum(1) / np.maximum(cnt, 1), 0.0)
    rows[:, 3] = (crit & valid).sum(1)
    rows[:, 4] = (major & valid).sum(1)
    rows[:, 5] = (minor & valid).sum(1)
    rows[:, 6] = (area * valid).sum(1)
    return rows


def gather_rows(rows: np.ndarray) -> np.ndarray:
    """All ranks' rows in rank order (per-image SPC replay needs the series, not just the sums)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return rows
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(rows)).to(dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.cat(out).cpu().numpy()


class BatchSPC:
    """Rolling-window SPC state fed with metric rows (one per image)."""

    def __init__(self, window_size: int = 100):
        self.window_size = window_size
        self.buffer = deque(maxlen=window_size)          # rows (np.ndarray [7])

    def update(self, rows: np.ndarray):
        for r in np.asarray(rows, np.float64).reshape(-1, len(ROW_FIELDS)):
            self.buffer.append(r)

    def _col(self, name: str) -> List[float]:
        i = ROW_FIELDS.index(name)
        return [float(r[i]) for r in self.buffer]

    # ---- analytics/sec_analysis.py:159-228 -------------------------------------------------------
    def control_limits(self) -> Dict:
        if len(self.buffer) < 10:
            return {}
        out = {}
        c_bar = np.mean(self._col("defect_count"))
        out["defect_count"] = {"center_line": c_bar, "upper_control_limit": c_bar + 3 * np.sqrt(c_bar),
                               "lower_control_limit": max(0, c_bar - 3 * np.sqrt(c_bar)),
                               "upper_warning_limit": c_bar + 2 * np.sqrt(c_bar),
                               "lower_warning_limit": max(0, c_bar - 2 * np.sqrt(c_bar)), "chart_type": "c-chart"}
        u_bar = np.mean(self._col("defect_rate"))
        out["defect_rate"] = {"center_line": u_bar, "upper_control_limit": u_bar + 3 * np.sqrt(u_bar / 1),
                              "lower_control_limit": max(0, u_bar - 3 * np.sqrt(u_bar / 1)),
                              "upper_warning_limit": u_bar + 2 * np.sqrt(u_bar / 1),
                              "lower_warning_limit": max(0, u_bar - 2 * np.sqrt(u_bar / 1)), "chart_type": "u-chart"}
        data = self._col("avg_confidence")
        x_bar = np.mean(data)
        mr = [abs(data[i] - data[i - 1]) for i in range(1, len(data))]
        mr_bar = np.mean(mr) if mr else 0
        d2 = 1.128
        out["confidence"] = {"center_line": x_bar, "upper_control_limit": x_bar + 3 * (mr_bar / d2),
                             "lower_control_limit": x_bar - 3 * (mr_bar / d2),
                             "upper_warning_limit": x_bar + 2 * (mr_bar / d2),
                             "lower_warning_limit": x_bar - 2 * (mr_bar / d2), "chart_type": "X-chart"}
        return out

    # ---- :279-318 -----------------------------------------------------------------------------------
    def rule_violations(self) -> List[str]:
        if len(self.buffer) < 9:
            return []
        dc = self._col("defect_count")[-20:]
        centre = np.mean(dc)
        v = []
        for i in range(len(dc) - 8):
            s = dc[i:i + 9]
            if all(x > centre for x in s) or all(x < centre for x in s):
                v.append("rule2")
                break
        for i in range(len(dc) - 5):
            s = dc[i:i + 6]
            if all(s[j] < s[j + 1] for j in range(5)) or all(s[j] > s[j + 1] for j in range(5)):
                v.append("rule3")
                break
        return v

    # ---- :230-277 (metrics present in both dicts: defect_count, defect_rate) ---------------------------
    def control_status(self, row: Optional[np.ndarray] = None) -> Dict:
        limits = self.control_limits()
        cur = self.buffer[-1] if row is None else np.asarray(row, np.float64)
        status = {"in_control": True, "violations": [], "warnings": []}
        for name, lim in limits.items():
            if name not in ROW_FIELDS:
                continue                                   # 'confidence' is not a key of the metrics dict (:243)
            val = float(cur[ROW_FIELDS.index(name)])
            if val > lim["upper_control_limit"] or val < lim["lower_control_limit"]:
                status["in_control"] = False
                status["violations"].append({"metric": name, "value": val,
                                             "limit_violated": "upper" if val > lim["upper_control_limit"] else "lower"})
            elif val > lim["upper_warning_limit"] or val < lim["lower_warning_limit"]:
                status["warnings"].append({"metric": name, "value": val,
                                           "limit_violated": "upper" if val > lim["upper_warning_limit"] else "lower"})
        rv = self.rule_violations()
        status["rule_violations"] = rv
        if rv:
            status["in_control"] = False
        return status

    # ---- :320-370 -----------------------------------------------------------------------------------
    def capability(self) -> Dict:
        if len(self.buffer) < 30:
            return {"insufficient_data": True}
        rates = self._col("defect_rate")
        mean, std = np.mean(rates), np.std(rates, ddof=1)
        usl, lsl, target = 5.0, 0.0, 0.5
        if not std > 0:
            return {}
        cpu, cpl = (usl - mean) / (3 * std), (mean - lsl) / (3 * std)
        return {"cp": (usl - lsl) / (6 * std), "cpk": min(cpu, cpl),
                "cpm": (usl - lsl) / (6 * np.sqrt(std ** 2 + (mean - target) ** 2)), "cpu": cpu, "cpl": cpl,
                "mean": mean, "std": std}

    def analyze_batch(self, rows: np.ndarray) -> Dict:
        self.update(rows)
        return {"control_limits": self.control_limits(), "control_status": self.control_status(),
                "process_capability": self.capability(), "window": len(self.buffer)}
