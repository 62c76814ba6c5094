"""CPU check of K10's host stage: qc_jpeg_entropy_host + oracle reconstruction == cv2.imdecode, with timings."""
import ctypes as C
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import jpeg as J  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "industrial-quality-control-computer-vision-system-_b200/lib/libqcgraft.so"))


class Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "components", "subsampling", "restart_interval", "reserved")] + \
               [("blocks", C.c_int64)]


lib.qc_jpeg_info_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Info)]
lib.qc_jpeg_entropy_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
SF = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420]
rng = np.random.default_rng(1)
for (h, w, sf, q) in [(640, 640, 2, 90), (640, 640, 2, 75), (1080, 1920, 2, 85), (333, 500, 1, 95), (480, 640, 0, 75)]:
    yy, xx = np.mgrid[0:h, 0:w]
    im = np.clip(np.stack([(xx * 3 + yy) % 256, (yy * 5) % 256, (xx + yy * 2) % 256], -1) + rng.normal(0, 10, (h, w, 3)),
                 0, 255).astype(np.uint8)
    ok, buf = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SF[sf]])
    tcv = min(_t for _t in [(lambda t0: (cv2.imdecode(buf, cv2.IMREAD_COLOR), time.perf_counter() - t0)[1])(time.perf_counter())
                            for _ in range(5)])
    ref = cv2.imdecode(buf, cv2.IMREAD_COLOR)
    info = Info()
    assert lib.qc_jpeg_info_host(None, buf.ctypes.data, buf.size, C.byref(info)) == 0
    coefs = np.empty(info.blocks * 64, np.int16)
    quant = np.zeros(3 * 64, np.uint16)
    dts = []
    for _ in range(5):
        t0 = time.perf_counter()
        assert lib.qc_jpeg_entropy_host(None, buf.ctypes.data, buf.size, coefs.ctypes.data, quant.ctypes.data) == 0
        dts.append(time.perf_counter() - t0)
    H = J.parse(buf.tobytes())
    cl, o = [], 0
    for bh, bw in H["blocks"]:
        cl.append(coefs[o:o + bh * bw * 64].reshape(bh, bw, 64))
        o += bh * bw * 64
    got = J.reconstruct(H, cl, [H["qt"][c["tq"]] for c in H["comp"]])
    print(f"{h}x{w} sf{sf} q{q}: {buf.size} bytes; entropy stage {min(dts) * 1e3:.2f} ms; cv2.imdecode {tcv * 1e3:.2f} ms; "
          f"equal: {np.array_equal(ref, got)}")
