"""Seeded frames + detection lists of the reference-segmentation golden set (tests/golden/refseg_golden.json).
Shared by the generator (make_refseg_golden.py, build container only) and the tests (CPU and GPU), so the frames never
need to be committed.  ROIs cover: both adaptive polarities, all widths mod 8 (cv2's vector / 4-group / scalar blur
tails), block sizes 11..33, frame borders, a constant ROI, ROIs below the 10- and 20-pixel component filters, thin
strips, a whole-frame ROI, a gray (2-D) frame, and an out-of-frame box that numpy slicing clips."""
import numpy as np

from qcgraft import synth


def _det(i, cls, x1, y1, x2, y2, conf=0.9):
    return {"id": i, "class": cls, "confidence": conf,
            "bbox": {"x1": x1, "y1": y1, "x2": x2, "y2": y2, "width": x2 - x1, "height": y2 - y1,
                     "center_x": (x1 + x2) / 2, "center_y": (y1 + y2) / 2}}


def cases():
    out = []
    # frame 0: the round-1 fixture's frame (seed 3) + more boxes
    f0 = synth.make_frame(3, 320, 320)
    f0[200:230, 40:90] = 150                               # a perfectly flat patch (constant ROI)
    d0 = [_det(0, "crack", 20, 30, 140, 150), _det(1, "discoloration", 150, 100, 300, 260),
          _det(2, "scratch", 0, 0, 97, 61), _det(3, "dent", 180, 0, 320, 90),
          _det(4, "discoloration", 45, 205, 85, 225), _det(5, "crack", 45, 205, 85, 225),
          _det(6, "unknown_defect", 0, 250, 133, 320), _det(7, "contamination", 100, 150, 220, 300),
          _det(8, "scratch", 250, 260, 320, 320), _det(9, "discoloration", 5, 5, 8, 8),
          _det(10, "crack", 300, 10, 304, 100)]
    out.append((f0, d0))
    # frame 1: denser defects, widths covering every residue mod 8, an empty (clipped-away) and a clipped box
    f1 = synth.make_frame(11, 320, 320)
    rng = np.random.default_rng(99)
    for _ in range(12):
        y, x = int(rng.integers(0, 300)), int(rng.integers(0, 300))
        f1[y:y + int(rng.integers(2, 25)), x:x + int(rng.integers(2, 40))] = rng.integers(10, 70) if rng.random() < .5 else rng.integers(190, 255)
    d1 = [_det(i, ("crack", "scratch")[i % 2], 10 + 3 * i, 20 + 5 * i, 10 + 3 * i + 40 + i, 20 + 5 * i + 50 + 2 * i) for i in range(8)]
    d1 += [_det(8, "discoloration", 160, 20, 310, 170), _det(9, "dent", 30, 180, 130, 300),
           _det(10, "contamination", 200, 200, 320, 320), _det(11, "crack", 0, 0, 320, 320),
           _det(12, "discoloration", 400, 400, 420, 420), _det(13, "scratch", 280, 280, 350, 360),
           _det(14, "unknown_defect", 140, 180, 200, 260)]
    out.append((f1, d1))
    # frame 2: gray line-scan frame (2-D input), blob defects
    f2 = synth.make_linescan(5, 256, 256)
    d2 = [_det(0, "discoloration", 0, 0, 256, 256), _det(1, "crack", 16, 16, 144, 200),
          _det(2, "discoloration", 100, 30, 240, 120), _det(3, "scratch", 30, 150, 250, 256),
          _det(4, "dent", 128, 128, 256, 256), _det(5, "unknown_defect", 7, 9, 70, 40),
          _det(6, "contamination", 60, 60, 190, 170)]
    out.append((f2, d2))
    return out
