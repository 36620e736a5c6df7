"""CPU restatement of the operation layer on 16-bit unsigned images (numpy; TEST INFRASTRUCTURE ONLY, like oracle.py).

The reference only feeds 8-bit images (cv.imread without flags, opencv/benchmark.py:14); the 16-bit semantics restated here
are what the SAME cv2 calls (opencv:94-105) give for CV_16U input, pinned by tests/golden/u16_cases.npz (generated with
cv2 4.13 by tests/golden/make_golden_u16.py):
  bitwise_not   65535 - x
  cvtColor      Y = (3735 B + 19235 G + 9798 R + 16384) >> 15 (the 8-bit path's 15-bit coefficients), alpha := 65535
  threshold     x > t ? maxval : 0
  erode/dilate  min / max over the in-image taps
  filter2D      correlation, BORDER_REFLECT_101, round half to even, saturate to 65535 (exact sums here: float64)
  GaussianBlur  ([1 2 1] x [1 2 1] sum + 8) >> 4
"""
from __future__ import annotations

import numpy as np

from . import oracle as o8

TOP = 65535


def _as3(img):
    a = np.asarray(img)
    assert a.dtype == np.uint16
    return a[:, :, None] if a.ndim == 2 else a


def copy(img):
    return np.array(img, copy=True)


def invert(img):                                   # opencv:95
    return (TOP - np.asarray(img)).astype(np.uint16)


def grayscale(img):                                # opencv:96
    a = _as3(img)
    if a.shape[2] == 1:
        return np.array(img, copy=True)
    b, g, r = (a[..., i].astype(np.int64) for i in range(3))
    y = ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint16)
    out = np.empty_like(a)
    out[..., 0] = out[..., 1] = out[..., 2] = y
    if a.shape[2] == 4:
        out[..., 3] = TOP
    return out.reshape(np.asarray(img).shape)


def threshold(img, thresh=127, maxval=255):        # opencv:97
    return np.where(np.asarray(img) > thresh, maxval, 0).astype(np.uint16)


def _morph(img, mask, is_max: bool):
    a = _as3(img).astype(np.int64)
    m = np.asarray(mask)
    mh, mw = m.shape
    ay, ax = mh // 2, mw // 2
    h, w, c = a.shape
    fill = 0 if is_max else TOP   # a window without any in-image tap gives the border value itself, like cv2
    pad = np.full((h + mh, w + mw, c), fill, np.int64)
    pad[ay:ay + h, ax:ax + w] = a
    out = np.full((h, w, c), fill, np.int64)
    for j in range(mh):
        for i in range(mw):
            if m[j, i]:
                win = pad[j:j + h, i:i + w]
                out = np.maximum(out, win) if is_max else np.minimum(out, win)
    return out.astype(np.uint16).reshape(np.asarray(img).shape)


def erode(img, mask):                              # opencv:98-99, :77-78
    return _morph(img, mask, False)


def dilate(img, mask):                             # opencv:81-82
    return _morph(img, mask, True)


def _reflect101(idx, n):
    if n == 1:
        return np.zeros_like(idx)
    p = 2 * (n - 1)
    idx = np.mod(idx, p)
    return np.where(idx >= n, p - idx, idx)


def _correlate(img, kernel):
    a = _as3(img).astype(np.float64)
    k = np.atleast_2d(np.asarray(kernel, dtype=np.float32)).astype(np.float64)
    kh, kw = k.shape
    ay, ax = kh // 2, kw // 2
    h, w, c = a.shape
    acc = np.zeros((h, w, c), np.float64)
    ys, xs = np.arange(h), np.arange(w)
    for j in range(kh):
        yy = _reflect101(ys + j - ay, h)
        for i in range(kw):
            if k[j, i] != 0.0:
                xx = _reflect101(xs + i - ax, w)
                acc += k[j, i] * a[yy][:, xx]
    return acc


def filter2d(img, kernel):                         # opencv:101,103, :85-86, :89-90
    acc = _correlate(img, kernel)
    return np.clip(np.rint(acc), 0, TOP).astype(np.uint16).reshape(np.asarray(img).shape)


def gaussian_blur3x3(img):                         # opencv:105
    acc = _correlate(img, o8.BLUR_3X3)             # exact multiples of 1/16
    return np.clip(np.floor(acc + 0.5), 0, TOP).astype(np.uint16).reshape(np.asarray(img).shape)


def erode_separable(img):                          # opencv:76-78
    return erode(erode(img, o8.SQUARE_MASK_1X3), o8.SQUARE_MASK_3X1)


def filter2d_separable(img, krow, kcol):           # opencv:84-90 (16-bit intermediate in `aux`)
    return filter2d(filter2d(img, krow), kcol)


OPERATIONS = [
    ("copy", copy),
    ("inversion", invert),
    ("grayscale", grayscale),
    ("threshold", lambda im: threshold(im, 127, 255)),
    ("erosion-cross", lambda im: erode(im, o8.CROSS_MASK)),
    ("erosion-square", lambda im: erode(im, o8.SQUARE_MASK)),
    ("erosion-square-separated", erode_separable),
    ("convolution-gaussian-blur-3x3", lambda im: filter2d(im, o8.BLUR_3X3)),
    ("convolution-gaussian-blur-3x3-separated", lambda im: filter2d_separable(im, o8.BLUR_3X3_1X3, o8.BLUR_3X3_3X1)),
    ("convolution-gaussian-blur-5x5", lambda im: filter2d(im, o8.BLUR_5X5)),
    ("convolution-gaussian-blur-5x5-separated", lambda im: filter2d_separable(im, o8.BLUR_5X5_1X5, o8.BLUR_5X5_5X1)),
    ("gaussian-blur-3x3", gaussian_blur3x3),
]
