"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/qcgraft.h
declares, the ctypes structs match the C layout, the host-side planners agree with the oracle, the
product never imports the oracle, and there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "industrial-quality-control-computer-vision-system-_b200")


def test_library_exports_every_declared_symbol():
    from qcgraft import ffi
    lib = ffi.load_library()
    hdr = open(os.path.join(ROOT, "include", "qcgraft.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(qc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"libqcgraft.so does not export {name}"
    assert declared == set(ffi.EXPORTED_SYMBOLS), declared ^ set(ffi.EXPORTED_SYMBOLS)
    assert lib.qc_version().startswith(b"qcgraft")


def test_struct_layouts_match():
    from qcgraft import ffi
    lib = ffi.load_library()
    for i, st in enumerate((ffi.PrePlan, ffi.YoloParams, ffi.SegRegion, ffi.SegParams)):
        assert lib.qc_sizeof(i) == C.sizeof(st)
    assert ffi.SEG_REGION_DTYPE.itemsize == C.sizeof(ffi.SegRegion)
    for f, _ in ffi.SegRegion._fields_:
        assert ffi.SEG_REGION_DTYPE.fields[f][1] == getattr(ffi.SegRegion, f).offset, f


def test_pre_make_plan_matches_letterbox_oracle():
    from qcgraft import ffi
    from oracle import preprocess as opre
    lib = ffi.load_library()
    rng = np.random.default_rng(0)
    shapes = [(1080, 1920), (480, 640), (333, 500), (640, 640), (1280, 1280), (100, 37), (37, 100), (641, 639)]
    shapes += [tuple(int(v) for v in rng.integers(8, 2500, 2)) for _ in range(200)]
    for h, w in shapes:
        for mode, auto in ((ffi.QC_PRE_LETTERBOX, False), (ffi.QC_PRE_LETTERBOX_AUTO, True)):
            pl = ffi.PrePlan()
            assert lib.qc_pre_make_plan(mode, h, w, 640, 640, 32, C.byref(pl)) == 0
            g = opre.letterbox_geometry(h, w, 640, 640, auto=auto)
            assert (pl.new_h, pl.new_w, pl.top, pl.left, pl.dst_h, pl.dst_w) == \
                   (g["new_h"], g["new_w"], g["top"], g["left"], g["out_h"], g["out_w"]), (h, w, auto)
            gain = min(pl.dst_h / h, pl.dst_w / w)                      # ultralytics scale_boxes
            assert pl.pad_x == round((pl.dst_w - w * gain) / 2 - 0.1) and pl.pad_y == round((pl.dst_h - h * gain) / 2 - 0.1)
            assert pl.gain == np.float32(gain)
    pl = ffi.PrePlan()
    assert lib.qc_pre_make_plan(ffi.QC_PRE_STRETCH, 480, 720, 640, 640, 32, C.byref(pl)) == 0
    assert (pl.dst_h, pl.dst_w, pl.top, pl.left) == (640, 640, 0, 0)
    assert lib.qc_pre_make_plan(ffi.QC_PRE_LETTERBOX, 0, 10, 640, 640, 32, C.byref(pl)) != 0


def test_segment_plan_offsets():
    from qcgraft import ffi
    lib = ffi.load_library()

    class Fake(ffi.Context):
        def __init__(self):
            self.lib = lib
    regs = [dict(img=0, x=0, y=0, w=2048, h=2048), dict(img=1, x=3, y=5, w=33, h=17, polarity=0, min_area=5),
            dict(img=1, x=0, y=0, w=1, h=1)]
    arr, p = Fake().segment_plan(regs, 2, 2048, 2048, 1, 2048 * 2048, 2048)
    assert p.num_regions == 3 and p.max_region_w == 2048
    assert arr.out_offset[0] == 0 and arr.out_offset[1] == 2048 * 2048 and arr.out_offset[2] % 16 == 0
    band = 16                 # k4_band_rows(2048); narrow regions get taller bands (33x17 and 1x1: one band each)
    assert arr.band_offset.tolist() == [0, 2048 // band, 2048 // band + 1] and p.total_bands == 2048 // band + 2
    assert arr.blk_offset.tolist() == [0, 1024 * 1024, 1024 * 1024 + 17 * 9]
    assert p.total_bits_words == 2048 * 64 + 17 * 2 + 1
    with pytest.raises(ffi.QcGraftError):                                   # region outside the image
        Fake().segment_plan([dict(img=0, x=2000, y=0, w=100, h=10)], 1, 2048, 2048, 1, 2048 * 2048, 2048)
    arr, p = Fake().segment_plan([], 1, 64, 64, 1, 4096, 64)
    assert p.num_regions == 0 and p.total_pixels == 0


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
    for f in ("bench.py",):      # bench may touch the oracle only inside the CPU-baseline helpers
        src = open(os.path.join(ROOT, f)).read()
        for m in re.finditer(r"^(\s*)(from|import)\s+oracle\b.*$", src, re.M):
            fn = re.findall(r"^def (\w+)", src[:m.start()], re.M)[-1]
            assert fn in ("build_cpu_pipeline", "cpu_reference_rate"), f"bench.py imports oracle in {fn}()"


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_no_cpu_fallback():
    from qcgraft import ffi
    with pytest.raises(ffi.QcGraftError):
        ffi.Context(0)
    from qcgraft.detector import QualityControlDetector
    d = QualityControlDetector("y.pt", "r.pt", {})
    out = d.predict(np.zeros((64, 64, 3), np.uint8))
    assert "error" in out                                        # the reference's error convention, never raises


def test_pil_resize_tables_match_pillow_coefficients():
    """qc_pil_resize_make_tables (host C++, float64) == the oracle's restatement of Pillow's precompute_coeffs +
    normalize_coeffs_8bpc, which tests/test_oracle_pins.py pins against Pillow itself."""
    from qcgraft import ffi
    from oracle import pil_resize as pr
    lib = ffi.load_library()
    for (h, w, oh, ow) in [(640, 640, 224, 224), (1080, 1920, 224, 224), (333, 500, 224, 224), (100, 150, 224, 224),
                           (224, 224, 224, 224), (2048, 2048, 224, 224), (7, 9, 224, 224), (480, 640, 96, 128)]:
        nb = lib.qc_pil_resize_table_bytes(h, w, oh, ow)
        t = np.zeros(nb // 4, np.int32)
        assert lib.qc_pil_resize_make_tables(h, w, oh, ow, t.ctypes.data) == 0
        kx, bx, kkx = pr.precompute_coeffs(w, ow)
        ky, by, kky = pr.precompute_coeffs(h, oh)
        assert t[:6].tolist() == [h, w, oh, ow, kx, ky]
        o = 8
        assert np.array_equal(t[o:o + ow * 2].reshape(ow, 2), bx); o += ow * 2
        assert np.array_equal(t[o:o + ow * kx].reshape(ow, kx), kkx); o += ow * kx
        assert np.array_equal(t[o:o + oh * 2].reshape(oh, 2), by); o += oh * 2
        assert np.array_equal(t[o:o + oh * ky].reshape(oh, ky), kky); o += oh * ky
        assert o == len(t)
        band, max_rows = int(t[6]), int(t[7])
        assert 1 <= band <= oh and max_rows * 3 * ((ow + 3) & ~3) <= 200 * 1024
        for o0 in range(0, oh, band):
            o1 = min(o0 + band, oh) - 1
            assert by[o1, 0] + by[o1, 1] - by[o0, 0] <= max_rows
    assert lib.qc_pil_resize_make_tables(0, 5, 224, 224, t.ctypes.data) != 0


def test_classifier_layout_is_a_fixed_place_partition():
    """Every static-shape classifier graph reads a fixed slice of the arena: rows are unique, views do not overlap, a
    size always sits at the same offset, and the views cover the rows in classifier order."""
    from qcgraft.pipeline import classifier_layout, chunk_plan
    sizes = (256, 128, 64, 32, 16, 8)
    seen = {}
    for total in list(range(1, 600)) + [1432, 2047, 4096, 5000]:
        rows, views = classifier_layout(total, sizes)
        assert len(set(rows.tolist())) == total and [(c, u) for _, c, u in views] == chunk_plan(total, sizes)
        iv = sorted((o, o + c) for o, c, _ in views)
        assert all(a[1] <= b[0] for a, b in zip(iv, iv[1:]))
        r = 0
        for off, cs, used in views:
            assert rows[r:r + used].tolist() == list(range(off, off + used))
            r += used
            if cs != 256:
                seen.setdefault(cs, set()).add(off)
    assert all(len(v) == (2 if cs == 8 else 1) for cs, v in seen.items()), seen


def test_classifier_chunk_plan():
    """Static ResNet batch sizes: every launch but the last is full, the tail rounds up to the smallest graph."""
    from qcgraft.pipeline import chunk_plan
    sizes = (256, 128, 64, 32, 16, 8)
    assert chunk_plan(1432, sizes) == [(256, 256)] * 5 + [(128, 128), (16, 16), (8, 8)]
    assert chunk_plan(3, sizes) == [(8, 3)] and chunk_plan(0, sizes) == [] and chunk_plan(8, sizes) == [(8, 8)]
    assert chunk_plan(265, sizes) == [(256, 256), (8, 8), (8, 1)]
    for total in range(0, 700, 37):
        plan = chunk_plan(total, sizes)
        assert sum(k for _, k in plan) == total and all(k == c for c, k in plan[:-1]) and all(c in sizes for c, _ in plan)


def test_entry_points_reject_null_context_without_touching_cuda():
    """Every compute entry point validates its context first: with a NULL qc_ctx it returns an error code and never
    reaches a CUDA call (this runs on a box without a GPU)."""
    from qcgraft import ffi
    lib = ffi.load_library()
    z = C.c_void_p(0)
    assert lib.qc_preprocess(None, z, 1, 0, 0, C.byref(ffi.PrePlan()), z, 114, 0, 0, z, z) != 0
    assert lib.qc_yolo_decode_nms(None, None, C.byref(ffi.YoloParams()), z, 0, z, z, z, z, z) != 0
    assert lib.qc_nms(None, z, z, z, z, 1, 8, 0.5, 7680.0, 300, z, 0, z, z, z) != 0
    assert lib.qc_roi_align(None, z, 1, 8, 8, 192, 24, z, 1, 224, 224, 2, 1, 0, z, 1, z, z) != 0
    assert lib.qc_pil_resize(None, z, 1, 8, 8, 192, 24, z, z, 0, z, 1, z, z) != 0
    assert lib.qc_segment(None, z, C.byref(ffi.SegParams()), z, z, 0, z, z, z, z, z, z, z) != 0
    assert lib.qc_spc_stats(None, z, z, 1, 300, z, z) != 0
    assert lib.qc_rules_cluster(None, z, z, z, z, 1, 300, 0.3, z, z) != 0
    assert lib.qc_last_error(None) == b"null context" and lib.qc_launch_count(None) == 0 and lib.qc_sm_count(None) == 0
    # geometry helpers are pure host code and do work without a device
    assert lib.qc_pil_resize_table_bytes(640, 640, 224, 224) > 0 and lib.qc_pil_resize_table_bytes(0, 640, 224, 224) == 0
    assert lib.qc_nms_workspace_bytes(2, 100) > 0 and lib.qc_nms_workspace_bytes(0, 100) == 0
