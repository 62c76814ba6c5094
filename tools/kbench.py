#!/usr/bin/env python
"""Runs only the hand-written kernels on their BASELINE configs (same code as bench.py's
`kernels` section) -- the command profiled with ncu for profiles/*."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from qcgraft import ffi  # noqa: E402

if __name__ == "__main__":
    qc = ffi.Context.current(0)
    peak, _ = bench.measured_peaks()
    only = set(sys.argv[1:]) or None
    print(json.dumps(bench.kernel_microbench(qc, peak, only), indent=1))
