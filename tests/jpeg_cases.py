"""Seeded JPEG streams shared by the CPU and GPU tests of K10 (written with cv2.imencode, the encoder the reference's
users upload from: quality, chroma sampling, restart intervals, optimised Huffman tables, grayscale, odd sizes)."""
import cv2
import numpy as np

SF = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
      "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


def picture(rng, h, w, kind):
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 3 + yy) % 256, (yy * 5) % 256, (xx + yy * 2) % 256], -1).astype(np.float64)
    return np.clip(base + rng.normal(0, 12, base.shape), 0, 255).astype(np.uint8)


def encode(img, quality=90, sampling="420", restart=0, optimize=False, progressive=False):
    params = [cv2.IMWRITE_JPEG_QUALITY, int(quality)]
    if img.ndim == 3:
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SF[sampling]]
    if restart:
        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(restart)]
    if optimize:
        params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
    if progressive:
        params += [cv2.IMWRITE_JPEG_PROGRESSIVE, 1]
    ok, buf = cv2.imencode(".jpg", img, params)
    assert ok
    return buf


def small_streams(seed=0, count=36, lo=1, hi=70):
    """(description, stream) pairs: every sampling at random sizes, plus one grayscale stream per size."""
    rng = np.random.default_rng(seed)
    for t in range(count):
        h, w = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
        im = picture(rng, h, w, t % 2)
        for sf in SF:
            q = int(rng.integers(5, 101))
            yield (t, h, w, sf, q), encode(im, q, sf, restart=int(rng.integers(1, 6)) if t % 3 == 0 else 0,
                                           optimize=t % 4 == 0)
        yield (t, h, w, "gray", 80), encode(im[:, :, 0], 80)
