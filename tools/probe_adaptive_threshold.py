#!/usr/bin/env python
"""CPU probe for SURVEY.md row f3 (cv2.adaptiveThreshold, ADAPTIVE_THRESH_GAUSSIAN_C -- the reference's crack /
scratch path, inference/segmentation.py:248-298): which parts can be pinned bit-exactly?

Findings (cv2 4.13, this container): the threshold stage is an exact integer table on `src - mean` given the uint8
`mean`; `mean` = cvRound (half-even) of a float32 separable Gaussian blur whose summation order is SIMD-dispatch
dependent -- four plain float32 orderings differ from cv2's float result by up to ~5e-5 on 30-50 % of the pixels,
yet round to the same uint8 unless the blurred value lies within ~1e-4 of a .5 boundary.  So a device version can be
pinned exactly on the integer stages and to "equal except pixels within 1e-4 of a rounding boundary" on the blur."""
import cv2
import numpy as np

rng = np.random.default_rng(0)
bs = 11
tot = bad = 0
for trial in range(20):
    img = rng.integers(0, 256, (int(rng.integers(40, 400)), int(rng.integers(40, 400))), dtype=np.uint8)
    ref = cv2.adaptiveThreshold(img, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, bs, 2)
    k = cv2.getGaussianKernel(bs, -1, cv2.CV_32F).ravel()
    src = img.astype(np.float32)
    blur = cv2.GaussianBlur(src, (bs, bs), 0, 0, borderType=cv2.BORDER_REPLICATE | cv2.BORDER_ISOLATED)
    mean = np.clip(np.rint(blur), 0, 255).astype(np.uint8)
    assert np.array_equal(np.where(img.astype(np.int32) - mean.astype(np.int32) > -2, 255, 0).astype(np.uint8), ref)
    pad = bs // 2
    P = np.pad(src, ((0, 0), (pad, pad)), mode="edge")
    acc = np.zeros_like(src)
    for t in range(bs):
        acc = acc + k[t] * P[:, t:t + src.shape[1]]
    Pc = np.pad(acc, ((pad, pad), (0, 0)), mode="edge")
    out = np.zeros_like(src)
    for t in range(bs):
        out = out + k[t] * Pc[t:t + src.shape[0], :]
    m2 = np.clip(np.rint(out), 0, 255).astype(np.uint8)
    tot += img.size
    bad += int((m2 != mean).sum())
print(f"threshold table exact on all trials; sequential float32 blur model: {bad} of {tot} uint8 means differ")
