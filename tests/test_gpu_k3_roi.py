"""K3 parity: CUDA ROI-align + normalise (through the C ABI) vs torchvision.ops.roi_align.
Tolerance (north_star): float32 <= 1e-3, bf16 within 1 ulp of the RNE-rounded reference."""
import numpy as np
import pytest
import torch

from gpu_common import qc, bf16_ulp_diff  # noqa: F401
from oracle import roi as oroi

pytestmark = pytest.mark.gpu


def _rois(rng, n, H, W, K):
    r = np.zeros((K, 5), np.float32)
    r[:, 0] = rng.integers(0, n, K)
    x1 = rng.integers(0, W - 8, K); y1 = rng.integers(0, H - 8, K)
    r[:, 1], r[:, 2] = x1, y1
    r[:, 3] = np.minimum(W, x1 + rng.integers(4, 200, K))
    r[:, 4] = np.minimum(H, y1 + rng.integers(4, 200, K))
    return r


@pytest.mark.parametrize("swap", [False, True])
def test_roi_align_random(qc, swap):
    from qcgraft import synth
    rng = np.random.default_rng(21)
    imgs = synth.make_frames(40, 3)
    rois = _rois(rng, 3, 640, 640, 37)
    rois[0] = [0, 0, 0, 640, 640]            # whole frame
    rois[1] = [1, 630, 630, 640, 640]        # touching the border
    rois[2] = [2, 10, 10, 11, 11]            # 1x1 box
    rois[3] = [2, 100.5, 50.25, 180.75, 90.5]  # fractional
    ref = oroi.roi_align_normalize(imgs, rois, swap_rb=swap)
    out = qc.roi_align(torch.from_numpy(imgs).cuda(), torch.from_numpy(rois).cuda(), swap_rb=swap,
                       out_dtype=torch.float32)
    assert out.shape == ref.shape
    assert (out.cpu() - ref).abs().max() <= 1e-3
    out16 = qc.roi_align(torch.from_numpy(imgs).cuda(), torch.from_numpy(rois).cuda(), swap_rb=swap)
    assert bf16_ulp_diff(out16, ref.to(torch.bfloat16)) <= 1


def test_roi_align_adaptive_sampling_and_unaligned(qc):
    rng = np.random.default_rng(22)
    imgs = rng.integers(0, 256, (2, 97, 131, 3), dtype=np.uint8)
    rois = _rois(rng, 2, 97, 131, 9)
    for sr, aligned, hw in [(0, True, (7, 7)), (2, False, (14, 14)), (3, True, (5, 9))]:
        ref = oroi.roi_align_normalize(imgs, rois, out_hw=hw, sampling_ratio=sr, aligned=aligned)
        out = qc.roi_align(torch.from_numpy(imgs).cuda(), torch.from_numpy(rois).cuda(), out_hw=hw,
                           sampling_ratio=sr, aligned=aligned, out_dtype=torch.float32)
        assert (out.cpu() - ref).abs().max() <= 1e-3


def test_roi_align_empty(qc):
    imgs = torch.zeros((1, 32, 32, 3), dtype=torch.uint8).cuda()
    out = qc.roi_align(imgs, torch.zeros((0, 5), dtype=torch.float32).cuda())
    assert out.shape == (0, 3, 224, 224)
