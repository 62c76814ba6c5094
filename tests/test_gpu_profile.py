"""qc_profile_enable / qc_profile_collect: per-kernel CUDA-event timing inside the library (bench.py's kernel_times)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_profile_lists_every_launch_in_order():
    from qcgraft import ffi
    qc = ffi.Context.current(0)
    rng = np.random.default_rng(0)
    src = torch.from_numpy(rng.integers(0, 256, (2, 480, 640, 3), dtype=np.uint8)).cuda()
    plan = qc.make_pre_plan(ffi.QC_PRE_LETTERBOX, 480, 640, 640, 640)
    qc.preprocess(src, plan)                                  # not recorded: profiling is off by default
    assert qc.profile_collect() == []
    qc.profile(True)
    l0 = qc.launch_count
    ref = qc.preprocess(src, plan)
    qc.preprocess(src, plan)
    n = qc.launch_count - l0
    recs = qc.profile_collect()
    qc.profile(False)
    assert len(recs) == n >= 2
    assert all(name.startswith("k1_") and 0.0 < ms < 50.0 for name, ms in recs), recs
    assert qc.profile_collect() == []
    # results are unchanged by the event records
    assert torch.equal(ref, qc.preprocess(src, plan))
