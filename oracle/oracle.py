"""ctypes front-end of oracle/dip_oracle.c plus the reference's fixed masks/kernels.

TEST INFRASTRUCTURE ONLY (see oracle/dip_oracle.c).  Each function cites the reference
call site it restates: /root/reference/benchmarks/opencv/benchmark.py (``opencv:LINE``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdip_oracle.so")
_lib = None

# Masks and kernels exactly as opencv:46-72 builds them.
CROSS_MASK = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=np.uint8)          # opencv:46
SQUARE_MASK = np.ones((3, 3), dtype=np.uint8)                                      # opencv:47
SQUARE_MASK_1X3 = np.ones((1, 3), dtype=np.uint8)                                  # opencv:48
SQUARE_MASK_3X1 = np.ones((3, 1), dtype=np.uint8)                                  # opencv:49
BLUR_3X3 = (np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float32) / np.float32(16))   # opencv:51-58
BLUR_3X3_1X3 = np.array([[0.25, 0.5, 0.25]], dtype=np.float32)                     # opencv:59
BLUR_3X3_3X1 = BLUR_3X3_1X3.T.copy()                                               # opencv:60
_B5 = np.array([1, 4, 6, 4, 1], dtype=np.float32)
BLUR_5X5 = (np.outer(_B5, _B5) / np.float32(256)).astype(np.float32)               # opencv:61-70
BLUR_5X5_1X5 = (_B5 / np.float32(16)).reshape(1, 5)                                # opencv:71
BLUR_5X5_5X1 = BLUR_5X5_1X5.T.copy()                                               # opencv:72


def build(force: bool = False) -> str:
    """Compile dip_oracle.c with gcc (``make -C oracle``)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "dip_oracle.c"))
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, sz, i = C.c_void_p, C.c_size_t, C.c_int
        common = [u8p, sz, u8p, sz, i, i, i]
        for name in ("orc_copy", "orc_invert", "orc_grayscale", "orc_gaussian_blur3x3"):
            getattr(L, name).argtypes = common
            getattr(L, name).restype = None
        L.orc_threshold.argtypes = common + [C.c_uint8, C.c_uint8]
        L.orc_threshold.restype = None
        for name in ("orc_erode", "orc_dilate"):
            getattr(L, name).argtypes = common + [u8p, i, i]
            getattr(L, name).restype = i
        L.orc_filter2d.argtypes = common + [C.c_void_p, i, i]
        L.orc_filter2d.restype = i
        L.orc_synth_fill.argtypes = [u8p, sz, i, i, i, C.c_uint32, C.c_uint64]
        L.orc_synth_fill.restype = None
        L.orc_checksum.argtypes = [u8p, sz, i, i, i]
        L.orc_checksum.restype = C.c_uint64
        L.orc_checksum_rows.argtypes = [u8p, sz, i, i, i, C.c_uint64]
        L.orc_checksum_rows.restype = C.c_uint64
        L.orc_synth_fill_rows.argtypes = [u8p, sz, i, i, i, C.c_uint32, C.c_uint64]
        L.orc_synth_fill_rows.restype = None
        L.orc_max_threads.restype = i
        L.orc_set_threads.argtypes = [i]
        _lib = L
    return _lib


def _as3(img: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(img, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    assert a.ndim == 3 and a.shape[2] in (1, 3, 4), a.shape
    return a


def _run(name: str, img: np.ndarray, *extra) -> np.ndarray:
    a = _as3(img)
    h, w, c = a.shape
    out = np.empty_like(a)
    if a.size:
        r = getattr(lib(), name)(a.ctypes.data, w * c, out.ctypes.data, w * c, w, h, c, *extra)
        if r not in (None, 0):
            raise ValueError(f"{name} failed with {r}")
    return out.reshape(img.shape)


def max_threads() -> int:
    return lib().orc_max_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(n)


def copy(img):                               # opencv:94 (intended dst = src; sycl:333)
    return _run("orc_copy", img)


def invert(img):                             # opencv:95
    return _run("orc_invert", img)


def grayscale(img):                          # opencv:96
    return _run("orc_grayscale", img)


def threshold(img, thresh=127, maxval=255):  # opencv:97
    return _run("orc_threshold", img, int(thresh), int(maxval))


def erode(img, mask):                        # opencv:98-99, :77-78
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    return _run("orc_erode", img, m.ctypes.data, m.shape[1], m.shape[0])


def dilate(img, mask):                       # opencv:81-82 (defined, unused by the reference)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    return _run("orc_dilate", img, m.ctypes.data, m.shape[1], m.shape[0])


def filter2d(img, kernel):                   # opencv:101,103, :85-86, :89-90
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    return _run("orc_filter2d", img, k.ctypes.data, k.shape[1], k.shape[0])


def gaussian_blur3x3(img):                   # opencv:105
    return _run("orc_gaussian_blur3x3", img)


def erode_separable(img):                    # opencv:76-78
    return erode(erode(img, SQUARE_MASK_1X3), SQUARE_MASK_3X1)


def filter2d_separable(img, krow, kcol):     # opencv:84-90 (u8 intermediate in `aux`)
    return filter2d(filter2d(img, krow), kcol)


def synth(w: int, h: int, c: int, seed: int = 0, frame: int = 0) -> np.ndarray:
    out = np.empty((h, w, c), dtype=np.uint8)
    lib().orc_synth_fill(out.ctypes.data, w * c, w, h, c, seed & 0xFFFFFFFF, frame)
    return out


def synth_rows(w: int, rows: int, c: int, seed: int, first_row: int) -> np.ndarray:
    out = np.empty((rows, w, c), dtype=np.uint8)
    lib().orc_synth_fill_rows(out.ctypes.data, w * c, w, rows, c, seed & 0xFFFFFFFF, first_row)
    return out


def checksum(img, first_row: int = 0) -> int:
    a = _as3(img)
    h, w, c = a.shape
    return int(lib().orc_checksum_rows(a.ctypes.data, w * c, w, h, c, first_row))


# The reference's operation list, opencv:94-105, as (prefix, function) in its order.
OPERATIONS = [
    ("copy", copy),
    ("inversion", invert),
    ("grayscale", grayscale),
    ("threshold", lambda im: threshold(im, 127, 255)),
    ("erosion-cross", lambda im: erode(im, CROSS_MASK)),
    ("erosion-square", lambda im: erode(im, SQUARE_MASK)),
    ("erosion-square-separated", erode_separable),
    ("convolution-gaussian-blur-3x3", lambda im: filter2d(im, BLUR_3X3)),
    ("convolution-gaussian-blur-3x3-separated", lambda im: filter2d_separable(im, BLUR_3X3_1X3, BLUR_3X3_3X1)),
    ("convolution-gaussian-blur-5x5", lambda im: filter2d(im, BLUR_5X5)),
    ("convolution-gaussian-blur-5x5-separated", lambda im: filter2d_separable(im, BLUR_5X5_1X5, BLUR_5X5_5X1)),
    ("gaussian-blur-3x3", gaussian_blur3x3),
]
