"""Golden vectors for 16-bit images: the reference's own cv2 calls (opencv/benchmark.py:94-105, restated in
oracle/cv2_ref.py) run on CV_16U input in the build container (cv2 4.13.0).  Writes tests/golden/u16_cases.npz.
usage: python tests/golden/make_golden_u16.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cv2_ref as ref
from oracle import oracle as o8
assert ref.available()
rng = np.random.default_rng(16)
out = {}
shapes = [(1, 1), (2, 3), (5, 7), (3, 40), (40, 3), (33, 47), (21, 130)]
ops = [("copy", lambda a: ref.copy(a)), ("inversion", ref.invert), ("grayscale", ref.grayscale),
       ("threshold", lambda a: ref.threshold(a, 127, 255)), ("threshold16", lambda a: ref.threshold(a, 30000, 65535)),
       ("erosion-cross", lambda a: ref.erode(a, o8.CROSS_MASK)), ("erosion-square", lambda a: ref.erode(a, o8.SQUARE_MASK)),
       ("erosion-square-separated", lambda a: ref.erode(ref.erode(a, o8.SQUARE_MASK_1X3), o8.SQUARE_MASK_3X1)),
       ("dilate-5x3", lambda a: ref.dilate(a, np.array([[1, 0, 1, 1, 0], [0, 1, 1, 0, 1], [1, 1, 0, 0, 1]], np.uint8))),
       ("erode-7x7", lambda a: ref.erode(a, (np.arange(49).reshape(7, 7) % 3 != 0).astype(np.uint8))),
       ("erode-9x9", lambda a: ref.erode(a, np.ones((9, 9), np.uint8))),
       ("convolution-gaussian-blur-3x3", lambda a: ref.filter2d(a, o8.BLUR_3X3)),
       ("convolution-gaussian-blur-3x3-separated", lambda a: ref.filter2d(ref.filter2d(a, o8.BLUR_3X3_1X3), o8.BLUR_3X3_3X1)),
       ("convolution-gaussian-blur-5x5", lambda a: ref.filter2d(a, o8.BLUR_5X5)),
       ("convolution-gaussian-blur-5x5-separated", lambda a: ref.filter2d(ref.filter2d(a, o8.BLUR_5X5_1X5), o8.BLUR_5X5_5X1)),
       ("sharpen-saturating", lambda a: ref.filter2d(a, np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32))),
       ("conv-9x1-dyadic", lambda a: ref.filter2d(a, (np.array([1, 2, 4, 8, 2, 8, 4, 2, 1], np.float32) / 32).reshape(9, 1))),
       ("gaussian-blur-3x3", ref.gaussian_blur3x3)]
for (h, w) in shapes:
    for c in (1, 3, 4):
        img = rng.integers(0, 65536, (h, w, c), dtype=np.uint16)
        if h * w >= 9:   # flat / extreme regions: ties and saturation
            img[: h // 2, : w // 2] = (img[: h // 2, : w // 2] // 16384) * 21845
            img[0, 0] = 65535
        key = f"{h}x{w}x{c}"
        out[f"{key}/input"] = img
        for name, fn in ops:
            res = fn(img if c > 1 else img[:, :, 0])
            out[f"{key}/{name}"] = np.asarray(res).reshape(img.shape)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "u16_cases.npz"), **out)
print(len(out), "arrays")
