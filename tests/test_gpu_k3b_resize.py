"""K3b parity: CUDA PIL-exact resize + ToTensor + Normalize (through the C ABI) vs the oracle restatement of
Pillow's resampler (pinned against Pillow in tests/test_oracle_pins.py).  Integer resampling: bit-exact; the float32
output is bit-exact too (same IEEE ops); bf16 output == RNE rounding of that float32."""
import numpy as np
import pytest
import torch

from gpu_common import qc  # noqa: F401
from oracle import pil_resize as opr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(640, 640), (480, 640), (333, 500), (224, 224), (100, 150), (224, 500), (7, 9),
                                (1080, 1920), (2048, 2048)])
def test_f32_bit_exact(qc, hw):
    rng = np.random.default_rng(hw[0] + 3 * hw[1])
    imgs = rng.integers(0, 256, (2,) + hw + (3,), dtype=np.uint8)
    out = qc.pil_resize(torch.from_numpy(imgs).cuda(), out_dtype=torch.float32).cpu().numpy()
    for i in range(2):
        assert np.array_equal(out[i], opr.resnet_input(imgs[i])), hw


def test_bf16_and_swap_rb_cfg2_batch(qc):
    """BASELINE cfg2 frames (64 x 640^2): bf16 output == RNE of the exact float32; swap_rb flips the channels
    the way inference/detector.py:234 + models/resent_model.py:123 do."""
    from qcgraft import synth
    imgs = synth.make_frames(0, 64)
    d = torch.from_numpy(imgs).cuda()
    out = qc.pil_resize(d).cpu()
    outs = qc.pil_resize(d, swap_rb=True).cpu()
    for i in (0, 17, 63):
        ref = torch.from_numpy(opr.resnet_input(imgs[i]))
        assert torch.equal(out[i].view(torch.int16), ref.to(torch.bfloat16).view(torch.int16))
        refs = torch.from_numpy(opr.resnet_input(np.ascontiguousarray(imgs[i][..., ::-1])))
        assert torch.equal(outs[i].view(torch.int16), refs.to(torch.bfloat16).view(torch.int16))


def test_pitched_rows_and_odd_output(qc):
    rng = np.random.default_rng(5)
    big = rng.integers(0, 256, (3, 300, 420, 3), dtype=np.uint8)
    d = torch.from_numpy(big).cuda()
    view = d[:, 10:290, 7:407]                                   # 280 x 400 window: pitched rows / images
    out = qc.pil_resize(view, out_hw=(97, 131), out_dtype=torch.float32).cpu().numpy()
    for i in range(3):
        assert np.array_equal(out[i], opr.resnet_input(big[i, 10:290, 7:407], (97, 131)))


@pytest.mark.parametrize("hw,out_hw", [((640, 640), (300, 320)), ((480, 900), (256, 600)), ((97, 1200), (224, 257))])
def test_outputs_wider_than_one_cta(qc, hw, out_hw):
    """Output rows wider than the 256 threads of a CTA (a thread then owns several columns and reloads its packed
    coefficients per column), against Pillow itself."""
    from PIL import Image
    rng = np.random.default_rng(hw[1])
    imgs = rng.integers(0, 256, (2,) + hw + (3,), dtype=np.uint8)
    out = qc.pil_resize(torch.from_numpy(imgs).cuda(), out_hw=out_hw, out_dtype=torch.float32).cpu().numpy()
    for i in range(2):
        assert np.array_equal(out[i], opr.resnet_input(imgs[i], out_hw)), (hw, out_hw)
        pil = np.asarray(Image.fromarray(imgs[i]).resize((out_hw[1], out_hw[0]), Image.BILINEAR))
        assert np.array_equal(opr.pil_resize_bilinear(imgs[i], out_hw), pil)
