"""K4 parity: CUDA Otsu + morphology + union-find CCL + stats (through the C ABI) vs the cv2
composite (cv2.threshold OTSU -> morphologyEx CLOSE/OPEN cross -> connectedComponentsWithStats ->
min-area filter).  Threshold, mask, labels (cv2 numbering), component count, areas and bounding
boxes are all bit-exact."""
import numpy as np
import pytest
import torch

from gpu_common import qc  # noqa: F401
from oracle import segment as oseg

pytestmark = pytest.mark.gpu


def _check_region(res, i, reg, ref):
    o = int(res["regions"][i].out_offset)
    so = int(res["regions"][i].stats_offset)
    w, h = reg["w"], reg["h"]
    assert int(res["thr"][i]) == ref["thr"], f"region {i}: Otsu {int(res['thr'][i])} vs {ref['thr']}"
    nc = int(res["ncomp"][i])
    assert nc == ref["ncomp"], f"region {i}: ncomp {nc} vs {ref['ncomp']}"
    lab = res["labels"][o:o + w * h].view(h, w).cpu().numpy()
    assert np.array_equal(lab, ref["labels"]), f"region {i}: labels differ at {np.argwhere(lab != ref['labels'])[:5]}"
    mask = res["mask"][o:o + w * h].view(h, w).cpu().numpy()
    assert np.array_equal(mask, ref["mask"])
    maxc = int(res["regions"][i].max_comp)
    rows = min(nc, maxc)
    st = res["stats"][so:so + rows].cpu().numpy()
    assert np.array_equal(st[1:], ref["stats"][1:rows]), f"region {i}: stats differ"
    assert int(res["area"][i]) == int((ref["mask"] > 0).sum())


def _run_full_frames(qc, imgs, polarity=1, thr_scale=1.0, iters=1, min_area=20):
    n, h, w = imgs.shape[:3]
    regs = [dict(img=i, x=0, y=0, w=w, h=h, polarity=polarity, thr_scale=thr_scale, min_area=min_area)
            for i in range(n)]
    res = qc.segment(torch.from_numpy(imgs).cuda(), regs, morph_iters=iters)
    torch.cuda.synchronize()
    for i in range(n):
        ref = oseg.segment_k4(imgs[i], polarity, thr_scale, iters, min_area)
        _check_region(res, i, regs[i], ref)
    return res


@pytest.mark.parametrize("shape", [(64, 64), (37, 53), (1, 40), (40, 1), (33, 31), (300, 333), (256, 2050)])
@pytest.mark.parametrize("iters", [0, 1, 2])
def test_linescan_small(qc, shape, iters):
    from qcgraft import synth
    imgs = np.stack([synth.make_linescan(100 + k, *shape) for k in range(2)])
    _run_full_frames(qc, imgs, iters=iters, min_area=10)


@pytest.mark.parametrize("density", [0.1, 0.3, 0.45, 0.6, 0.8])
def test_random_masks_ccl_stress(qc, density):
    """Random binary patterns (iters=0): snake-like 8-connected components across band/word borders."""
    rng = np.random.default_rng(int(density * 100))
    for shape in [(130, 200), (67, 129), (512, 512)]:
        m = rng.random((2,) + shape) < density
        imgs = np.where(m, 200, 50).astype(np.uint8)
        _run_full_frames(qc, imgs, iters=0, min_area=5)


def test_structured_patterns(qc):
    h, w = 96, 160
    imgs = np.full((4, h, w), 40, np.uint8)
    imgs[0, ::2, :] = 220                                    # horizontal stripes
    imgs[1, :, ::2] = 220                                    # vertical stripes
    yy, xx = np.mgrid[0:h, 0:w]
    imgs[2][(yy + xx) % 4 == 0] = 220                        # diagonals (8-connectivity only)
    imgs[3][((yy // 3) + (xx // 3)) % 2 == 0] = 220          # checkerboard of 3x3 cells
    _run_full_frames(qc, imgs, iters=0, min_area=3)
    _run_full_frames(qc, imgs, iters=1, min_area=3)
    # nested rings + spiral
    im = np.full((1, 128, 128), 30, np.uint8)
    y2, x2 = np.mgrid[0:128, 0:128]
    d = np.maximum(np.abs(y2 - 64), np.abs(x2 - 64))
    for r in range(4, 60, 6):
        im[0][d == r] = 230
    _run_full_frames(qc, im, iters=0, min_area=1)


def test_polarity_and_class_factor(qc):
    from qcgraft import synth
    imgs = np.stack([synth.make_linescan(200 + k, 200, 260) for k in range(2)])
    _run_full_frames(qc, imgs, polarity=0, iters=1)                     # THRESH_BINARY_INV
    _run_full_frames(qc, imgs, polarity=1, thr_scale=0.8, iters=2, min_area=10)   # discoloration factor
    _run_full_frames(qc, imgs, polarity=0, thr_scale=0.9, iters=1)      # crack factor, dark
    const = np.full((1, 50, 60), 77, np.uint8)
    _run_full_frames(qc, const, iters=1)                                # constant image: Otsu -> 0, all fg


def test_rgb_input_gray_conversion(qc):
    from qcgraft import synth
    imgs = synth.make_frames(300, 2, 200, 280)
    _run_full_frames(qc, imgs, polarity=0, iters=1, min_area=10)


def test_roi_regions_mixed(qc):
    """Detection-ROI mode: many regions of different sizes/polarities over an RGB batch."""
    from qcgraft import synth
    rng = np.random.default_rng(31)
    imgs = synth.make_frames(310, 3)
    regs = []
    for k in range(25):
        w, h = int(rng.integers(5, 200)), int(rng.integers(5, 200))
        x, y = int(rng.integers(0, 640 - w)), int(rng.integers(0, 640 - h))
        regs.append(dict(img=int(rng.integers(0, 3)), x=x, y=y, w=w, h=h, polarity=int(rng.integers(0, 2)),
                         thr_scale=float(rng.choice([1.0, 0.8, 0.9, 1.1])), min_area=int(rng.choice([10, 20]))))
    res = qc.segment(torch.from_numpy(imgs).cuda(), regs, morph_iters=2)
    torch.cuda.synchronize()
    for i, r in enumerate(regs):
        roi = imgs[r["img"], r["y"]:r["y"] + r["h"], r["x"]:r["x"] + r["w"]]
        ref = oseg.segment_k4(np.ascontiguousarray(roi), r["polarity"], r["thr_scale"], 2, r["min_area"])
        _check_region(res, i, r, ref)


def test_cfg4_full_size_2048(qc):
    """BASELINE cfg4 size (2048^2 line-scan, >= 2000 components): exact vs cv2 on 2 frames."""
    from qcgraft import synth
    imgs = synth.make_linescans(400, 2)
    res = _run_full_frames(qc, imgs, iters=1, min_area=20)
    assert int(res["ncomp"].min()) >= 2000


def test_idempotence_property(qc):
    """Size-independent property: segmenting the produced mask again (iters=0, min_area=0) must
    reproduce the same labels (the mask is a fixed point of threshold+CCL)."""
    from qcgraft import synth
    imgs = synth.make_linescans(500, 1, 700, 900)
    n, h, w = imgs.shape
    regs = [dict(img=0, x=0, y=0, w=w, h=h, min_area=20)]
    r1 = qc.segment(torch.from_numpy(imgs).cuda(), regs, morph_iters=1)
    mask = r1["mask"][:h * w].view(1, h, w).contiguous()
    regs2 = [dict(img=0, x=0, y=0, w=w, h=h, min_area=0)]
    r2 = qc.segment(mask, regs2, morph_iters=0)
    l1 = r1["labels"][:h * w]
    l2 = r2["labels"][:h * w]
    assert torch.equal(l1 > 0, l2 > 0)
    # relabeling preserves order: label ids of l1 map monotonically onto 1..k in l2
    pairs = torch.unique(torch.stack((l1[l1 > 0], l2[l1 > 0])), dim=1)
    assert torch.equal(pairs[1], torch.arange(1, pairs.shape[1] + 1, device=pairs.device, dtype=pairs.dtype))
    assert (pairs[0][1:] > pairs[0][:-1]).all()


def test_cfg4_full_batch_16(qc):
    """BASELINE cfg4 at its full size: 16 line-scan frames of 2048^2 in one call, all exact vs cv2."""
    from qcgraft import synth
    imgs = synth.make_linescans(600, 16)
    _run_full_frames(qc, imgs, iters=1, min_area=20)


def test_dense_noise_rois(qc):
    """The pathological case of the bench: ROIs on a flat noisy surface (Otsu splits the noise, ~50% foreground)."""
    rng = np.random.default_rng(77)
    imgs = np.clip(rng.normal(150, 10, (2, 400, 600)), 0, 255).astype(np.uint8)
    regs = []
    for k in range(12):
        w, h = int(rng.integers(17, 300)), int(rng.integers(17, 300))
        regs.append(dict(img=k % 2, x=int(rng.integers(0, 600 - w)), y=int(rng.integers(0, 400 - h)), w=w, h=h,
                         polarity=k % 2, min_area=20))
    res = qc.segment(torch.from_numpy(imgs).cuda(), regs, morph_iters=1)
    for i, r in enumerate(regs):
        roi = np.ascontiguousarray(imgs[r["img"], r["y"]:r["y"] + r["h"], r["x"]:r["x"] + r["w"]])
        _check_region(res, i, r, oseg.segment_k4(roi, r["polarity"], 1.0, 1, 20))


def test_otsu_replay_many_sparse_histograms(qc):
    """Thousands of small ROIs with few gray levels: sparse histograms make exact ties between neighbouring splits, so
    nearly every region goes through the warp-cooperative replay of cv2's float64 recurrence (parallel reciprocals +
    Markstein division chain) -- the threshold must equal cv2.threshold(OTSU) on every one of them."""
    rng = np.random.default_rng(77)
    frames = np.zeros((4, 256, 256), np.uint8)
    levels = [np.array([10, 200]), np.array([0, 1, 2, 255]), np.arange(0, 256, 17), np.array([3, 4, 5, 6, 7, 250, 251])]
    for i in range(4):
        frames[i] = rng.choice(levels[i], size=(256, 256), p=None)
    frames[1, :128] = rng.integers(0, 256, (128, 256))             # dense part: decisive maxima next to tie cases
    regs = []
    for k in range(3000):
        w, h = int(rng.integers(3, 48)), int(rng.integers(3, 48))
        regs.append(dict(img=int(rng.integers(0, 4)), x=int(rng.integers(0, 256 - w)), y=int(rng.integers(0, 256 - h)),
                         w=w, h=h, polarity=int(rng.integers(0, 2)), min_area=1, max_comp=8))
    res = qc.segment(torch.from_numpy(frames).cuda(), regs, morph_iters=0, want_mask=False, want_labels=False)
    thr = res["thr"].cpu().numpy()
    for k, r in enumerate(regs):
        roi = np.ascontiguousarray(frames[r["img"], r["y"]:r["y"] + r["h"], r["x"]:r["x"] + r["w"]])
        assert int(thr[k]) == oseg.otsu_u8(roi), (k, r)
