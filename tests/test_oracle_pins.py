"""Pins the oracle's restatements against the libraries the reference actually calls
(cv2 / torchvision, both present here and on the GPU box).  CPU only."""
import numpy as np
import cv2
import pytest
import torch
import torchvision

from oracle import preprocess as opre
from oracle import yolo_head as oyh
from oracle import roi as oroi
from oracle import segment as oseg


def test_resize_integer_model_matches_cv2():
    rng = np.random.default_rng(1)
    shapes = [((1080, 1920), (360, 640)), ((480, 640), (480, 640)), ((333, 500), (426, 640)),
              ((640, 640), (320, 320)), ((100, 37), (640, 237)), ((7, 5), (64, 48))]
    for _ in range(40):
        h, w = rng.integers(3, 300, 2)
        H, W = rng.integers(3, 300, 2)
        shapes.append(((int(h), int(w)), (int(H), int(W))))
    for (h, w), (H, W) in shapes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)
        got = opre.resize_linear_u8(img, W, H)
        assert np.array_equal(ref, got), ((h, w), (H, W))


def test_letterbox_restated_matches_cv2_path():
    rng = np.random.default_rng(2)
    for (h, w) in [(1080, 1920), (480, 640), (333, 500), (640, 640), (500, 333)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for auto in (False, True):
            a, ga = opre.letterbox_u8(img, 640, 640, auto=auto, use_cv2=True)
            b, gb = opre.letterbox_u8(img, 640, 640, auto=auto, use_cv2=False)
            assert ga == gb and np.array_equal(a, b)
            if not auto:
                assert a.shape[:2] == (640, 640)
            else:
                assert a.shape[0] % 32 == 0 and a.shape[1] % 32 == 0


def test_gray_model_matches_cv2():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (257, 311, 3), dtype=np.uint8)
    assert np.array_equal(oseg.rgb2gray(img), oseg.rgb2gray_restated(img))


def test_otsu_restated_matches_cv2():
    rng = np.random.default_rng(4)
    for t in range(120):
        h, w = rng.integers(5, 120, 2)
        kind = t % 4
        if kind == 0:
            g = rng.integers(0, 256, (h, w)).astype(np.uint8)
        elif kind == 1:
            g = np.clip(rng.normal(128, 20, (h, w)), 0, 255).astype(np.uint8)
        elif kind == 2:
            g = np.where(rng.random((h, w)) < 0.5, rng.integers(0, 100), rng.integers(150, 256)).astype(np.uint8)
        else:
            a, b = sorted(rng.integers(0, 256, 2))
            g = rng.choice([a, b, (a + b) // 2], (h, w)).astype(np.uint8)
        assert oseg.otsu_u8(g) == oseg.otsu_u8_restated(g)
    const = np.full((9, 9), 77, np.uint8)
    assert oseg.otsu_u8(const) == oseg.otsu_u8_restated(const)


def test_morph_restated_matches_cv2():
    rng = np.random.default_rng(5)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    assert np.array_equal(k, oseg.CROSS)
    for _ in range(60):
        h, w = rng.integers(1, 70, 2)
        m = (rng.random((h, w)) < rng.uniform(0.2, 0.8)).astype(np.uint8) * 255
        for it in (1, 2):
            assert np.array_equal(oseg.morph_close_open(m, it), oseg.morph_close_open_restated(m, it))


def test_ccl_numbering_matches_cv2():
    rng = np.random.default_rng(6)
    for _ in range(80):
        h, w = rng.integers(1, 90, 2)
        m = (rng.random((h, w)) < rng.uniform(0.1, 0.7)).astype(np.uint8) * 255
        n, lab = cv2.connectedComponents(m)
        lab2, n2 = oseg.ccl_restated(m)
        assert n == n2 and np.array_equal(lab, lab2)


def test_nms_restated_matches_torchvision():
    g = torch.Generator().manual_seed(7)
    for trial in range(20):
        n = 400
        c = torch.rand(40, 2, generator=g) * 600
        ctr = c[torch.randint(0, 40, (n,), generator=g)] + torch.randn(n, 2, generator=g) * 6
        wh = torch.rand(n, 2, generator=g) * 80 + 8
        boxes = torch.cat((ctr - wh / 2, ctr + wh / 2), 1)
        scores = torch.rand(n, generator=g)
        scores[::7] = scores[0]          # score ties -> stable order matters
        cls = torch.randint(0, 5, (n,), generator=g).float()
        ob = boxes + cls[:, None] * 7680
        for thr in (0.5, 0.3, 0.45):
            ref = torchvision.ops.nms(ob, scores, thr).numpy()
            got = oyh.nms_restated(ob.numpy(), scores.numpy(), thr)
            assert np.array_equal(ref, got)


def test_nms_threshold_equality_kept():
    # IoU exactly == thr is kept (strict '>')
    b = torch.tensor([[0., 0., 2., 2.], [0., 0., 2., 1.]])  # IoU = 0.5
    s = torch.tensor([0.9, 0.8])
    assert torchvision.ops.nms(b, s, 0.5).tolist() == [0, 1]
    assert oyh.nms_restated(b.numpy(), s.numpy(), 0.5).tolist() == [0, 1]


def test_class_offset_equals_batched_nms():
    g = torch.Generator().manual_seed(8)
    n = 600
    ctr = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 60 + 4
    boxes = torch.cat((ctr - wh / 2, ctr + wh / 2), 1)
    scores = torch.rand(n, generator=g)
    cls = torch.randint(0, 5, (n,), generator=g)
    a = torchvision.ops.nms(boxes + cls[:, None].float() * 7680, scores, 0.5)
    b = torchvision.ops.batched_nms(boxes, scores, cls, 0.5)
    assert torch.equal(a, b)


def test_roi_align_restated_matches_torchvision():
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (1, 40, 50, 3), dtype=np.uint8)
    rois = np.array([[0, 3, 4, 30, 33], [0, 0, 0, 50, 40], [0, 45, 35, 50, 40], [0, 10, 10, 11, 11]], np.float32)
    t = torch.from_numpy(img.transpose(0, 3, 1, 2).copy()).float()
    ref = torchvision.ops.roi_align(t, torch.from_numpy(rois), (14, 14), 1.0, 2, True).numpy()
    for k in range(len(rois)):
        got = oroi.roi_align_restated(t[0].numpy(), rois[k, 1:], (14, 14), 2, True)
        assert np.abs(ref[k] - got).max() < 2e-4


def test_decode_shapes_and_anchor_order():
    nc = 5
    levels = [torch.zeros(1, 64 + nc, s, s) for s in (8, 4, 2)]
    y = oyh.detect_decode(levels, nc)
    assert y.shape == (1, 4 + nc, 84)
    # zero logits: uniform DFL -> dist 7.5 -> centred boxes w=h=15*stride at anchor centres
    assert torch.allclose(y[0, :2, 0], torch.tensor([4.0, 4.0]))
    assert torch.allclose(y[0, :2, 1], torch.tensor([12.0, 4.0]))   # x inner
    assert torch.allclose(y[0, 2:4, 0], torch.tensor([120.0, 120.0]))
    assert torch.allclose(y[0, :2, 64], torch.tensor([8.0, 8.0]))   # first P4 anchor
    assert torch.allclose(y[0, 4:, 0], torch.full((nc,), 0.5))


# ---- K3b: Pillow's antialiased BILINEAR resize + torchvision ToTensor/Normalize ---------------------
@pytest.mark.parametrize("hw", [(640, 640), (480, 640), (333, 500), (224, 224), (100, 150), (224, 500), (7, 9),
                                (1080, 1920)])
def test_pil_resize_restatement_is_bit_exact(hw):
    from PIL import Image
    import torchvision.transforms as T
    from oracle import pil_resize as pr
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    img = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR))
    assert np.array_equal(pr.pil_resize_bilinear(img), ref)
    # the reference's transform chain (models/resent_model.py:57-66)
    tf = T.Compose([T.Resize((224, 224)), T.ToTensor(),
                    T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    assert np.array_equal(pr.resnet_input(img), tf(Image.fromarray(img)).numpy())


def test_fill_holes_restatement_matches_scipy():
    """Row f3 preparation: binary_fill_holes (inference/segmentation.py:455-456) == border-reachability of the
    4-connected background."""
    from scipy import ndimage
    from oracle.fill_holes import fill_holes
    rng = np.random.default_rng(9)
    for shape, dens in [((40, 60), 0.5), ((97, 131), 0.62), ((64, 64), 0.3), ((1, 17), 0.5), ((17, 1), 0.5), ((200, 300), 0.55)]:
        m = rng.random(shape) < dens
        assert np.array_equal(fill_holes(m), ndimage.binary_fill_holes(m)), shape
    ring = np.zeros((9, 9), bool); ring[2:7, 2:7] = True; ring[4, 4] = False; ring[3, 3] = False   # diagonal hole pair
    assert np.array_equal(fill_holes(ring), ndimage.binary_fill_holes(ring))
    assert fill_holes(np.zeros((5, 5), bool)).sum() == 0 and fill_holes(np.ones((5, 5), bool)).all()
