"""The reference's own OpenCV calls, restated one-for-one on ndarrays / cv.UMat.

TEST INFRASTRUCTURE ONLY.  This is the reference's CPU path: without OpenCL, ``cv.UMat``
executes OpenCV's CPU kernels.  It needs the third-party ``cv2`` wheel (the
reference pins opencv-python 4.11.0.86, benchmarks/opencv/uv.lock:62-63); when cv2 is not
importable ``available()`` is False and callers fall back to oracle/oracle.py.

Used (a) here in the build container to generate tests/golden/ and pin the C restatement,
(b) by bench.py as the "reference" CPU baseline.  Call sites: benchmarks/opencv/benchmark.py
(``opencv:LINE``).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    import cv2 as cv
except Exception:  # noqa: BLE001
    cv = None

from . import oracle as _o


def available() -> bool:
    return cv is not None


def version() -> str:
    return cv.__version__ if cv is not None else "absent"


def copy(img):                       # opencv:94 -- intended semantics dst = src
    return img.copy()


def invert(img):                     # opencv:95
    return cv.bitwise_not(img)


def grayscale(img):                  # opencv:96 (4-channel: BGRA2GRAY -> GRAY2BGRA)
    if img.ndim == 2 or img.shape[2] == 1:
        return img.copy()
    if img.shape[2] == 4:
        return cv.cvtColor(cv.cvtColor(img, cv.COLOR_BGRA2GRAY), cv.COLOR_GRAY2BGRA)
    return cv.cvtColor(cv.cvtColor(img, cv.COLOR_BGR2GRAY), cv.COLOR_GRAY2BGR)


def threshold(img, thresh=127, maxval=255):   # opencv:97
    return cv.threshold(img, thresh, maxval, cv.THRESH_BINARY)[1]


def erode(img, mask):                # opencv:98-99
    return cv.erode(img, mask)


def dilate(img, mask):               # opencv:81-82
    return cv.dilate(img, mask)


def filter2d(img, kernel):           # opencv:101,103
    return cv.filter2D(img, -1, kernel)


def gaussian_blur3x3(img):           # opencv:105
    return cv.GaussianBlur(img, (3, 3), 0)


def erode_separable(img):            # opencv:76-78
    return cv.erode(cv.erode(img, _o.SQUARE_MASK_1X3), _o.SQUARE_MASK_3X1)


def filter2d_separable(img, krow, kcol):   # opencv:84-90
    return cv.filter2D(cv.filter2D(img, -1, krow), -1, kcol)


def _keep_shape(fn):
    def wrapped(img):
        out = fn(img)
        return np.asarray(out).reshape(img.shape)
    return wrapped


OPERATIONS = [
    ("copy", copy),
    ("inversion", _keep_shape(invert)),
    ("grayscale", _keep_shape(grayscale)),
    ("threshold", _keep_shape(lambda im: threshold(im, 127, 255))),
    ("erosion-cross", _keep_shape(lambda im: erode(im, _o.CROSS_MASK))),
    ("erosion-square", _keep_shape(lambda im: erode(im, _o.SQUARE_MASK))),
    ("erosion-square-separated", _keep_shape(erode_separable)),
    ("convolution-gaussian-blur-3x3", _keep_shape(lambda im: filter2d(im, _o.BLUR_3X3))),
    ("convolution-gaussian-blur-3x3-separated", _keep_shape(lambda im: filter2d_separable(im, _o.BLUR_3X3_1X3, _o.BLUR_3X3_3X1))),
    ("convolution-gaussian-blur-5x5", _keep_shape(lambda im: filter2d(im, _o.BLUR_5X5))),
    ("convolution-gaussian-blur-5x5-separated", _keep_shape(lambda im: filter2d_separable(im, _o.BLUR_5X5_1X5, _o.BLUR_5X5_5X1))),
    ("gaussian-blur-3x3", _keep_shape(gaussian_blur3x3)),
]
