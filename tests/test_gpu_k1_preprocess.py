"""K1 parity: CUDA preprocess (through the C ABI) vs the oracle (cv2 LetterBox + torch /255)."""
import numpy as np
import pytest
import torch

from gpu_common import qc, bf16_bits  # noqa: F401
from oracle import preprocess as opre

pytestmark = pytest.mark.gpu


def _run(qc, imgs, mode, H, W, swap=True, mean=None, std=None, dtype=torch.bfloat16):
    from qcgraft import ffi
    n, h, w = imgs.shape[:3]
    plan = qc.make_pre_plan(mode, h, w, H, W)
    src = torch.from_numpy(imgs).cuda()
    out = qc.preprocess(src, plan, swap_rb=swap, mean=mean, std=std, out_dtype=dtype)
    torch.cuda.synchronize()
    return out, plan


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_identity_pack_640(qc, dtype):
    from qcgraft import ffi, synth
    imgs = synth.make_frames(0, 5)
    out, plan = _run(qc, imgs, ffi.QC_PRE_LETTERBOX, 640, 640, dtype=dtype)
    ref, g = opre.yolo_preprocess(list(imgs), 640, 640, dtype=dtype)
    assert (plan.dst_h, plan.dst_w) == (640, 640) and plan.top == 0 and plan.left == 0
    assert out.shape == ref.shape
    if dtype == torch.bfloat16:
        assert torch.equal(bf16_bits(out), bf16_bits(ref))
    else:
        assert torch.equal(out.cpu(), ref)


def test_identity_random_bytes_all_values(qc):
    from qcgraft import ffi
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, (3, 96, 160, 3), dtype=np.uint8)   # P=15360 (tail chunk < 4096)
    for swap in (True, False):
        out, _ = _run(qc, imgs, ffi.QC_PRE_LETTERBOX, 96, 160, swap=swap)
        ref = opre.pack_nchw(list(imgs), swap_rb=swap)
        assert torch.equal(bf16_bits(out), bf16_bits(ref))


def test_mean_std_lut_three_tables(qc):
    from qcgraft import ffi
    from oracle.roi import RESNET_MEAN, RESNET_STD
    rng = np.random.default_rng(6)
    imgs = rng.integers(0, 256, (2, 224, 224, 3), dtype=np.uint8)
    for dtype in (torch.bfloat16, torch.float32):
        out, _ = _run(qc, imgs, ffi.QC_PRE_STRETCH, 224, 224, swap=False, mean=RESNET_MEAN, std=RESNET_STD, dtype=dtype)
        ref = opre.pack_nchw(list(imgs), swap_rb=False, mean=RESNET_MEAN, std=RESNET_STD, dtype=dtype)
        if dtype == torch.bfloat16:
            assert torch.equal(bf16_bits(out), bf16_bits(ref))
        else:
            assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("shape", [(1080, 1920), (480, 640), (333, 500), (500, 333), (1280, 1280), (100, 37)])
@pytest.mark.parametrize("auto", [False, True])
def test_letterbox_real_resize(qc, shape, auto):
    from qcgraft import ffi
    rng = np.random.default_rng(hash(shape) % 1000)
    imgs = rng.integers(0, 256, (2,) + shape + (3,), dtype=np.uint8)
    mode = ffi.QC_PRE_LETTERBOX_AUTO if auto else ffi.QC_PRE_LETTERBOX
    out, plan = _run(qc, imgs, mode, 640, 640, dtype=torch.float32)
    ref, g = opre.yolo_preprocess(list(imgs), 640, 640, auto=auto, dtype=torch.float32)
    assert (plan.dst_h, plan.dst_w) == (g["out_h"], g["out_w"])
    assert (plan.new_h, plan.new_w, plan.top, plan.left) == (g["new_h"], g["new_w"], g["top"], g["left"])
    assert out.shape == ref.shape
    # float32 /255 is injective on bytes -> equality here == bit-exact uint8 resize + border
    assert torch.equal(out.cpu(), ref)
    out16, _ = _run(qc, imgs, mode, 640, 640, dtype=torch.bfloat16)
    assert torch.equal(bf16_bits(out16), bf16_bits(ref.to(torch.bfloat16)))


def test_stretch_resize(qc):
    from qcgraft import ffi
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (2, 480, 720, 3), dtype=np.uint8)
    out, plan = _run(qc, imgs, ffi.QC_PRE_STRETCH, 640, 640, swap=False, dtype=torch.float32)
    ref = opre.pack_nchw([opre.stretch_u8(im, 640, 640) for im in imgs], swap_rb=False, dtype=torch.float32)
    assert torch.equal(out.cpu(), ref)


def test_pitched_rows_generic_path(qc):
    from qcgraft import ffi
    rng = np.random.default_rng(10)
    big = rng.integers(0, 256, (2, 200, 300, 3), dtype=np.uint8)
    src = torch.from_numpy(big).cuda()[:, 10:138, 7:263]       # 128x256 view, pitched rows
    plan = qc.make_pre_plan(ffi.QC_PRE_LETTERBOX, 128, 256, 128, 256)
    out = qc.preprocess(src, plan, swap_rb=True)
    torch.cuda.synchronize()
    ref = opre.pack_nchw(list(big[:, 10:138, 7:263]), swap_rb=True)
    assert torch.equal(bf16_bits(out), bf16_bits(ref))


def test_full_size_property_cfg2(qc):
    """BASELINE cfg2 size (64 x 640^2): plane sums equal the byte sums through the LUT (f32 out)."""
    from qcgraft import ffi
    rng = np.random.default_rng(11)
    imgs = rng.integers(0, 256, (64, 640, 640, 3), dtype=np.uint8)
    out, _ = _run(qc, imgs, ffi.QC_PRE_LETTERBOX, 640, 640, dtype=torch.float32)
    back = (out * 255).round().to(torch.uint8).cpu().numpy()          # inverse of the injective LUT
    assert np.array_equal(back, imgs[..., ::-1].transpose(0, 3, 1, 2))
