"""K2 parity: CUDA decode + filter + NMS (through the C ABI) vs ultralytics-restated decode +
torchvision.ops.nms.  Tier T1 = NMS only on identical boxes (indices exact, no excuses);
tier T2 = fused decode+NMS on margin-safe synthetic heads (indices exact, boxes/scores <= 1e-3)."""
import numpy as np
import pytest
import torch
import torchvision

from gpu_common import qc  # noqa: F401
from oracle import yolo_head as oyh

pytestmark = pytest.mark.gpu


def _clustered(g, n, ncl=40, span=600.0, ties=True):
    c = torch.rand(ncl, 2, generator=g) * span
    ctr = c[torch.randint(0, ncl, (n,), generator=g)] + torch.randn(n, 2, generator=g) * 6
    wh = torch.rand(n, 2, generator=g) * 80 + 8
    boxes = torch.cat((ctr - wh / 2, ctr + wh / 2), 1)
    scores = torch.rand(n, generator=g)
    if ties:
        scores[::7] = scores[0]
    cls = torch.randint(0, 5, (n,), generator=g, dtype=torch.int32)
    return boxes, scores, cls


@pytest.mark.parametrize("cap,iou", [(300, 0.5), (1500, 0.5), (1500, 0.3), (4096, 0.45), (6000, 0.5)])
def test_t1_nms_only_exact(qc, cap, iou):
    g = torch.Generator().manual_seed(cap)
    n = 3
    B, S, Cc, cnt = [], [], [], []
    for i in range(n):
        k = cap if i == 0 else int(cap * (0.3 + 0.3 * i))
        b, s, c = _clustered(g, k)
        pad = cap - k
        B.append(torch.cat((b, torch.zeros(pad, 4)))); S.append(torch.cat((s, torch.zeros(pad))))
        Cc.append(torch.cat((c, torch.zeros(pad, dtype=torch.int32)))); cnt.append(k)
    boxes, scores, cls = torch.stack(B), torch.stack(S), torch.stack(Cc)
    counts = torch.tensor(cnt, dtype=torch.int32)
    keep, kc = qc.nms(boxes.cuda(), scores.cuda(), cls.cuda(), counts.cuda(), iou=iou, max_det=300)
    keep, kc = keep.cpu(), kc.cpu()
    for i in range(n):
        k = cnt[i]
        ob = boxes[i, :k] + cls[i, :k, None].float() * 7680
        ref = torchvision.ops.nms(ob, scores[i, :k], iou)[:300]
        assert int(kc[i]) == len(ref)
        assert torch.equal(keep[i, :len(ref)].long(), ref), f"image {i}"
        assert (keep[i, len(ref):] == -1).all()


def test_t1_threshold_edges(qc):
    # IoU exactly == thr is kept; IoU just above is suppressed; zero-area boxes (0/0 = NaN) are kept
    boxes = torch.tensor([[[0., 0., 2., 2.], [0., 0., 2., 1.], [0., 0., 2., 1.0001], [5., 5., 5., 5.],
                           [5., 5., 5., 5.], [0., 0., 2., 2.]]])
    scores = torch.tensor([[0.9, 0.8, 0.7, 0.6, 0.5, 0.4]])
    cls = torch.tensor([[0, 0, 0, 1, 1, 2]], dtype=torch.int32)
    counts = torch.tensor([6], dtype=torch.int32)
    keep, kc = qc.nms(boxes.cuda(), scores.cuda(), cls.cuda(), counts.cuda(), iou=0.5)
    ob = boxes[0] + cls[0, :, None].float() * 7680
    ref = torchvision.ops.nms(ob, scores[0], 0.5)
    assert keep[0, :int(kc[0])].cpu().long().tolist() == ref.tolist()


def test_t1_near_threshold_offset_rounding(qc):
    """Pairs with IoU within ~1e-6 of the threshold, in classes 0..4 (the +cls*7680 float32
    rounding changes the decision -- SURVEY.md §7); must match torchvision on every pair."""
    g = torch.Generator().manual_seed(123)
    npairs = 1500
    x = torch.rand(npairs, generator=g) * 600
    y = torch.rand(npairs, generator=g) * 600
    w = torch.rand(npairs, generator=g) * 60 + 10
    h = torch.rand(npairs, generator=g) * 60 + 10
    a = torch.stack((x, y, x + w, y + h), 1)
    # second box: same box shifted in x so IoU ~= 0.5 -> shift d with (w-d)/(w+d) = 0.5 -> d = w/3
    d = w / 3 + (torch.rand(npairs, generator=g) - 0.5) * 2e-4
    b = torch.stack((x + d, y, x + w + d, y + h), 1)
    # far-apart placement of the pairs so that only the pair partner overlaps
    off = (torch.arange(npairs).float() * 0)  # pairs may overlap others; fine, oracle sees the same
    boxes = torch.stack((a, b), 1).reshape(1, -1, 4).contiguous()
    scores = torch.rand(1, 2 * npairs, generator=g)
    cls = torch.randint(0, 5, (npairs,), generator=g, dtype=torch.int32).repeat_interleave(2).view(1, -1)
    counts = torch.tensor([2 * npairs], dtype=torch.int32)
    keep, kc = qc.nms(boxes.cuda(), scores.cuda(), cls.cuda().contiguous(), counts.cuda(), iou=0.5, max_det=3000)
    ob = boxes[0] + cls[0, :, None].float() * 7680
    ref = torchvision.ops.nms(ob, scores[0], 0.5)
    assert int(kc[0]) == len(ref)
    assert torch.equal(keep[0, :len(ref)].cpu().long(), ref)


def _compare_full(qc, levels, H, W, orig=None, conf=0.7, iou=0.5, max_det=300, dtype=torch.float32, max_nms=30000):
    from qcgraft import ffi
    nc = 5
    n = levels[0].shape[0]
    orig = orig or (H, W)
    plan = qc.make_pre_plan(ffi.QC_PRE_LETTERBOX, orig[0], orig[1], H, W)
    assert (plan.dst_h, plan.dst_w) == (H, W)
    dev = [l.to(dtype).cuda().contiguous() for l in levels]
    p = qc.yolo_params(dev, nc, conf=conf, iou=iou, max_det=max_det, max_nms=max_nms, plan=plan)
    det, anchor, count, ncand = qc.yolo_decode_nms(dev, p)
    torch.cuda.synchronize()
    det, anchor, count, ncand = det.cpu(), anchor.cpu(), count.cpu(), ncand.cpu()
    ref_in = [l.to(dtype).float() for l in levels]          # bf16 mode: oracle sees the rounded values
    rdets, ridx = oyh.yolo_postprocess(ref_in, nc, (H, W), orig, conf=conf, iou=iou, max_det=max_det, return_idx=True,
                                       max_nms=max_nms)
    y = oyh.detect_decode(ref_in, nc)
    for i in range(n):
        k = len(rdets[i])
        assert int(ncand[i]) == int((y[i, 4:].amax(0) > conf).sum()), "candidate count"
        assert int(count[i]) == k, f"image {i}: kept {int(count[i])} vs {k}"
        assert torch.equal(anchor[i, :k].long(), ridx[i]), f"image {i}: kept anchor indices differ"
        assert (anchor[i, k:] == -1).all() and (det[i, k:] == 0).all()
        if k:
            assert torch.equal(det[i, :k, 5], rdets[i][:, 5]), "class ids"
            assert (det[i, :k, :4] - rdets[i][:, :4]).abs().max() <= 1e-3
            assert (det[i, :k, 4] - rdets[i][:, 4]).abs().max() <= 1e-3
    return det, count


def test_t2_fused_640_sparse(qc):
    from qcgraft import synth
    levels = synth.make_head(11, 4, 640, 640, 40)
    _compare_full(qc, levels, 640, 640)


def test_t2_fused_1280_dense_cfg3(qc):
    """BASELINE cfg3 shape: >= 1000 candidates/image at 1280^2 (NMS stress, max_det cut)."""
    from qcgraft import synth
    levels = synth.make_head(12, 3, 1280, 1280, 1500)
    det, count = _compare_full(qc, levels, 1280, 1280)
    assert int(count.min()) >= 100


def test_t2_fused_bf16_head(qc):
    from qcgraft import synth
    levels = synth.make_head(13, 2, 640, 640, 200)
    _compare_full(qc, levels, 640, 640, dtype=torch.bfloat16)


def test_t2_scale_boxes_letterboxed_source(qc):
    from qcgraft import synth
    levels = synth.make_head(14, 2, 640, 640, 60)
    _compare_full(qc, levels, 640, 640, orig=(480, 640))     # pad_y = 80, gain = 1
    _compare_full(qc, levels, 640, 640, orig=(1080, 1920))   # gain = 1/3


def test_t2_other_thresholds_and_maxdet(qc):
    from qcgraft import synth
    levels = synth.make_head(15, 2, 640, 640, 300)
    _compare_full(qc, levels, 640, 640, conf=0.8, iou=0.3, max_det=50)


def test_empty_image(qc):
    from qcgraft import synth
    levels = synth.make_head(16, 2, 640, 640, 0)
    det, count = _compare_full(qc, levels, 640, 640)
    assert int(count.sum()) == 0


def test_noncontiguous_concatenated_head(qc):
    """The head as ONE [B, 69, A] tensor (ultralytics' concatenated layout) addressed by strides."""
    from qcgraft import synth
    levels = synth.make_head(17, 2, 640, 640, 80)
    nc = 5
    B = 2
    cat = torch.cat([l.reshape(B, 69, -1) for l in levels], 2).cuda().contiguous()   # [B,69,8400]
    A_l = [6400, 1600, 400]
    starts = [0, 6400, 8000]
    views = [cat[:, :, s:s + a].unflatten(2, (hw, hw)) for s, a, hw in zip(starts, A_l, (80, 40, 20))]
    for v in views:
        assert v.stride(1) == 8400
    p = qc.yolo_params(views, nc)
    det, anchor, count, ncand = qc.yolo_decode_nms(views, p)
    rdets, ridx = oyh.yolo_postprocess(levels, nc, (640, 640), (640, 640), return_idx=True)
    for i in range(B):
        k = len(rdets[i])
        assert int(count[i]) == k and torch.equal(anchor[i, :k].cpu().long(), ridx[i])


def test_cfg3_full_batch_32(qc):
    """BASELINE cfg3 at its full size: 32 images x 1280^2 x >= 1000 candidates, every image index-exact."""
    from qcgraft import synth
    levels = synth.make_head(300, 32, 1280, 1280, 1200)
    det, count = _compare_full(qc, levels, 1280, 1280)
    assert int(count.min()) >= 100


def test_cfg2_full_batch_64_bf16(qc):
    """BASELINE cfg2 head shape (64 x 640^2, bf16 head as the bf16 net emits it)."""
    from qcgraft import synth
    levels = synth.make_head(301, 64, 640, 640, 30)
    _compare_full(qc, levels, 640, 640, dtype=torch.bfloat16)


def test_one_class_dominates_global_sort_path(qc):
    """> 1024 candidates of ONE class in an image: the per-class NMS CTA leaves its shared-memory sort for the
    global-memory regions; the other classes of the same image stay in shared memory."""
    from qcgraft import synth
    levels = synth.make_head(302, 2, 1280, 1280, 2600)
    for l in levels:
        cls = l[:, 64:69]
        top = cls.amax(1)
        move = torch.rand(top.shape, generator=torch.Generator().manual_seed(5)) < 0.8   # 80% of anchors -> class 2
        newc = cls.clone()
        newc[:, 2] = torch.where(move, top, cls[:, 2])
        for c in (0, 1, 3, 4):
            newc[:, c] = torch.where(move, torch.full_like(top, -8.0), cls[:, c])
        l[:, 64:69] = newc
    det, count = _compare_full(qc, levels, 1280, 1280)
    assert int(count.min()) >= 100


def test_max_det_cut_across_classes(qc):
    """max_det smaller than the kept total: the rank merge must return the global top-max_det by score."""
    from qcgraft import synth
    levels = synth.make_head(303, 3, 640, 640, 600, n_centres=300)
    for md in (7, 64, 300):
        _compare_full(qc, levels, 640, 640, max_det=md)


def test_max_nms_cut(qc):
    """More candidates than max_nms: only the max_nms best-scoring ones enter NMS (ultralytics' pre-NMS cut)."""
    from qcgraft import synth
    levels = synth.make_head(304, 2, 640, 640, 700, n_centres=120)
    for mn in (650, 257, 64):
        _compare_full(qc, levels, 640, 640, max_nms=mn)


def test_saturated_sigmoid_ties_take_the_first_class(qc):
    """float32 sigmoid saturates: logits (18, 20), (13.95, 14.0), (30, 17) ... give EQUAL scores, and torch.max returns
    the first class, not the class of the largest logit (class label and NMS class offset depend on it)."""
    from qcgraft import synth
    levels = synth.make_head(31, 2, 640, 640, 0)
    rng = np.random.default_rng(31)
    pairs = [(18.0, 20.0), (13.95, 14.0), (30.0, 17.0), (16.8, 25.0), (12.49, 12.5), (9.0, 9.0005), (11.0, 19.0)]
    for img in range(2):
        for l, lev in enumerate(levels):
            hh, ww = lev.shape[2:]
            for k, (a, b) in enumerate(pairs):
                y, x = int(rng.integers(0, hh)), int(rng.integers(0, ww))
                c = rng.permutation(5)
                lev[img, 64 + c[0], y, x] = a if c[0] < c[1] else b      # the smaller logit sits in the lower class index
                lev[img, 64 + c[1], y, x] = b if c[0] < c[1] else a
                for side in range(4):
                    lev[img, side * 16 + int(rng.integers(1, 6)), y, x] += 6.0
    det, count = _compare_full(qc, levels, 640, 640)
    assert int(count.min()) > 0
    _compare_full(qc, levels, 640, 640, dtype=torch.bfloat16)
