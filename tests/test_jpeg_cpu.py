"""K10 on the CPU: oracle/jpeg.py (numpy restatement of libjpeg-turbo's integer decode) == cv2.imdecode byte for byte,
and the library's host entropy stage == the oracle's bit-by-bit Huffman decoder (no GPU involved)."""
import ctypes as C

import cv2
import numpy as np
import pytest

from oracle import jpeg as J
from qcgraft import ffi

import jpeg_cases


def _lib():
    return ffi.load_library()


def _entropy(lib, buf):
    info = ffi.JpegInfo()
    rc = lib.qc_jpeg_info_host(None, buf.ctypes.data, buf.size, C.byref(info))
    assert rc == 0, rc
    coefs = np.empty(info.blocks * 64, np.int16)
    quant = np.zeros(3 * 64, np.uint16)
    assert lib.qc_jpeg_entropy_host(None, buf.ctypes.data, buf.size, coefs.ctypes.data, quant.ctypes.data) == 0
    return info, coefs, quant


def test_oracle_equals_cv2_imdecode():
    n = 0
    for desc, buf in jpeg_cases.small_streams(0, 36):
        assert np.array_equal(cv2.imdecode(buf, cv2.IMREAD_COLOR), J.decode(buf.tobytes())), desc
        n += 1
    assert n == 144


def test_host_entropy_stage_equals_oracle():
    lib = _lib()
    for desc, buf in jpeg_cases.small_streams(1, 14, 1, 80):
        H = J.parse(buf.tobytes())
        ref = J.huffman_blocks(buf.tobytes(), H)
        info, coefs, quant = _entropy(lib, buf)
        assert (info.width, info.height, info.components, info.subsampling, info.restart_interval) == \
            (H["w"], H["h"], H["nc"], H["mode"], H["restart"]), desc
        assert info.blocks == sum(bh * bw for bh, bw in H["blocks"])
        assert np.array_equal(np.concatenate([r.reshape(-1) for r in ref]), coefs), desc
        for c, comp in enumerate(H["comp"]):
            assert np.array_equal(quant[c * 64:(c + 1) * 64], H["qt"][comp["tq"]]), desc


@pytest.mark.parametrize("h,w,sampling,q", [(640, 640, "420", 90), (333, 500, "422", 95), (480, 640, "444", 60)])
def test_host_stage_plus_oracle_reconstruction_equals_cv2_at_frame_size(h, w, sampling, q):
    lib = _lib()
    rng = np.random.default_rng(h)
    buf = jpeg_cases.encode(jpeg_cases.picture(rng, h, w, 1), q, sampling)
    H = J.parse(buf.tobytes())
    info, coefs, quant = _entropy(lib, buf)
    cl, o = [], 0
    for bh, bw in H["blocks"]:
        cl.append(coefs[o:o + bh * bw * 64].reshape(bh, bw, 64))
        o += bh * bw * 64
    got = J.reconstruct(H, cl, [H["qt"][c["tq"]] for c in H["comp"]])
    assert np.array_equal(got, cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_unsupported_and_broken_streams_are_refused():
    lib = _lib()
    rng = np.random.default_rng(5)
    im = jpeg_cases.picture(rng, 40, 56, 1)
    info = ffi.JpegInfo()

    def rc_of(b):
        b = np.ascontiguousarray(b, np.uint8)
        return lib.qc_jpeg_info_host(None, b.ctypes.data, b.size, C.byref(info))
    assert rc_of(jpeg_cases.encode(im, progressive=True)) == -4          # QC_ERR_UNSUPPORTED
    ok, png = cv2.imencode(".png", im)
    assert rc_of(png) == -1                                              # not a JPEG
    good = jpeg_cases.encode(im)
    assert rc_of(good) == 0 and (info.width, info.height) == (56, 40)
    assert rc_of(good[:20]) == -1                                        # truncated header
    # EXIF orientation 6 (cv2.imdecode would rotate): APP1 segment spliced in after SOI
    exif = b"Exif\0\0" + b"II*\0" + (8).to_bytes(4, "little") + (1).to_bytes(2, "little") + \
        (0x0112).to_bytes(2, "little") + (3).to_bytes(2, "little") + (1).to_bytes(4, "little") + \
        (6).to_bytes(2, "little") + b"\0\0" + (0).to_bytes(4, "little")
    app1 = b"\xff\xe1" + (len(exif) + 2).to_bytes(2, "big") + exif
    rot = np.frombuffer(good.tobytes()[:2] + app1 + good.tobytes()[2:], np.uint8)
    assert cv2.imdecode(rot, cv2.IMREAD_COLOR).shape[:2] == (56, 40)     # cv2 does rotate this stream
    assert rc_of(rot) == -4
    # a scan cut short decodes with zero bits like libjpeg does not -- the stage must not crash or overrun
    cut = np.ascontiguousarray(good[:good.size // 2])
    assert rc_of(cut) == 0
    coefs = np.empty(info.blocks * 64, np.int16)
    quant = np.zeros(192, np.uint16)
    assert lib.qc_jpeg_entropy_host(None, cut.ctypes.data, cut.size, coefs.ctypes.data, quant.ctypes.data) in (0, -1)
