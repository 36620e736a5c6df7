"""tests/golden/fundus_crop.npz: a lossless 160x240 crop of the reference's input image as cv2 (libjpeg-turbo) decodes it
here, with the outputs of the reference's 12 cv2 calls (/root/reference/benchmarks/opencv/benchmark.py:94-105, restated in
oracle/cv2_ref.py) on that crop.  The GPU box may decode data/fundus.jpg with another libjpeg; this crop pins real
fundus pixels independently of the decoder.  Run in the build container:  python tests/golden/make_fundus_crop.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2 as cv  # noqa: E402

from oracle import cv2_ref as ref  # noqa: E402


def main() -> None:
    img = cv.imread(os.path.join(ROOT, "data", "fundus.jpg"))
    assert img is not None and img.shape == (2336, 3504, 3)
    crop = np.ascontiguousarray(img[1090:1250, 2350:2590])   # optic disc rim and vessels: edges, flat regions, saturation
    out = {"input": crop}
    for prefix, fn in ref.OPERATIONS:
        out[prefix] = fn(crop)
    np.savez_compressed(os.path.join(HERE, "fundus_crop.npz"), **out)
    print("fundus_crop.npz:", crop.shape, "mean", crop.mean(axis=(0, 1)))


if __name__ == "__main__":
    main()
