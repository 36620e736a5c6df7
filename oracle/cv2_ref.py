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


class ReferenceLoop:
    """The reference's ``perform_benchmark`` state and closures, as written (opencv:41-105): three ``cv.UMat`` images
    ``image`` / ``aux`` / ``sample`` allocated once, every operation called with its destination argument so that it
    writes into the pre-allocated ``sample`` (the separable ones through ``aux``), ``cv.copyTo(image, sample)`` with
    ``sample`` in the mask slot exactly like opencv:94.  bench.py's ``--impl reference`` arm and its ``cpu_baseline``
    time THIS; the functional wrappers above exist for the golden vectors.  One instance per worker thread."""

    def __init__(self, cpu_image: np.ndarray):
        self.image = cv.UMat(cpu_image)      # opencv:42-44
        self.aux = cv.UMat(cpu_image)
        self.sample = cv.UMat(cpu_image)
        image, aux, sample = self.image, self.aux, self.sample
        o = _o

        def erosion_separated():             # opencv:76-78
            cv.erode(image, o.SQUARE_MASK_1X3, aux)
            cv.erode(aux, o.SQUARE_MASK_3X1, sample)

        def convolution_3x3_separated():     # opencv:84-86
            cv.filter2D(image, -1, o.BLUR_3X3_1X3, aux)
            cv.filter2D(aux, -1, o.BLUR_3X3_3X1, sample)

        def convolution_5x5_separated():     # opencv:88-90
            cv.filter2D(image, -1, o.BLUR_5X5_1X5, aux)
            cv.filter2D(aux, -1, o.BLUR_5X5_5X1, sample)

        self.operations = [                  # opencv:94-105, same order, same calls
            ("copy", lambda: cv.copyTo(image, sample)),
            ("inversion", lambda: cv.bitwise_not(image, sample)),
            ("grayscale", lambda: cv.cvtColor(cv.cvtColor(image, cv.COLOR_BGR2GRAY), cv.COLOR_GRAY2BGR, sample)),
            ("threshold", lambda: cv.threshold(image, 127, 255, cv.THRESH_BINARY, sample)),
            ("erosion-cross", lambda: cv.erode(image, o.CROSS_MASK, sample)),
            ("erosion-square", lambda: cv.erode(image, o.SQUARE_MASK, sample)),
            ("erosion-square-separated", erosion_separated),
            ("convolution-gaussian-blur-3x3", lambda: cv.filter2D(image, -1, o.BLUR_3X3, sample)),
            ("convolution-gaussian-blur-3x3-separated", convolution_3x3_separated),
            ("convolution-gaussian-blur-5x5", lambda: cv.filter2D(image, -1, o.BLUR_5X5, sample)),
            ("convolution-gaussian-blur-5x5-separated", convolution_5x5_separated),
            ("gaussian-blur-3x3", lambda: cv.GaussianBlur(image, (3, 3), 0, sample)),
        ]

    def load(self, cpu_image: np.ndarray) -> None:
        """A new input frame into the pre-allocated ``image`` (the reference uploads once, opencv:42)."""
        cv.copyTo(cpu_image, None, self.image)

    def run_all(self) -> None:
        for _, fn in self.operations:
            fn()

    def result(self) -> np.ndarray:          # opencv:113 ``sample.get()``
        return self.sample.get()
