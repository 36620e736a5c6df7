"""Generate tests/golden/ from the reference's own OpenCV calls.

Run in the build container (where cv2 is installed):  python tests/golden/make_golden.py

It executes the calls of /root/reference/benchmarks/opencv/benchmark.py:94-105 (restated in
oracle/cv2_ref.py; when /root/reference is present the literal ``perform_benchmark`` of the
reference is imported as well and its written outputs' in-memory arrays are cross-checked
through cv.UMat) and stores
  * small_cases.npz : seeded inputs of edge shapes x {1,3,4} channels with every op's output,
  * kat.json        : sha256[:16] / crc32 of every op's output for the two full-size inputs
                      A = cv2.imread(data/fundus.jpg) and B = default_rng(0),
  * generic_cases.npz : outputs for argument shapes the benchmark does not use (odd masks,
                      dilate, thresholds, a non-dyadic float kernel),
  * pattern_cases.npz : saturated, maximum-gradient and rounding-tie patterns (every .5 case).
The fixtures travel to the GPU box; /root/reference and this script's cv2 dependency do not
have to.
"""
import hashlib
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2 as cv  # noqa: E402

from oracle import cv2_ref as ref  # noqa: E402
from oracle import oracle as orc  # noqa: E402

SMALL_SHAPES = [(1, 1), (1, 2), (2, 1), (2, 2), (1, 9), (9, 1), (3, 3), (5, 7), (7, 5), (16, 16),
                (33, 47), (64, 130)]
CHANNELS = [1, 3, 4]


def digest(a: np.ndarray) -> dict:
    b = np.ascontiguousarray(a).tobytes()
    return {"sha256_16": hashlib.sha256(b).hexdigest()[:16], "crc32": f"{zlib.crc32(b):08x}"}


def main() -> None:
    rng = np.random.default_rng(20261018)
    small = {}
    for (h, w) in SMALL_SHAPES:
        for c in CHANNELS:
            img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
            # mix in flat and extreme regions so ties / saturation are exercised
            if h * w >= 9:
                img[: h // 2, : w // 2] = (img[: h // 2, : w // 2] // 64) * 85
            key = f"{h}x{w}x{c}"
            small[f"{key}/input"] = img
            for prefix, fn in ref.OPERATIONS:
                small[f"{key}/{prefix}"] = fn(img)
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **small)

    generic = {}
    img = rng.integers(0, 256, (41, 53, 3), dtype=np.uint8)
    generic["input"] = img
    masks = {
        "m5x5_full": np.ones((5, 5), np.uint8),
        "m5x3_plus": np.array([[0, 1, 0], [0, 1, 0], [1, 1, 1], [0, 1, 0], [0, 1, 0]], np.uint8),
        "m1x5": np.ones((1, 5), np.uint8),
        "m7x1": np.ones((7, 1), np.uint8),
        "m3x3_diag": np.eye(3, dtype=np.uint8),
    }
    for name, m in masks.items():
        generic[f"mask/{name}"] = m
        generic[f"erode/{name}"] = ref.erode(img, m)
        generic[f"dilate/{name}"] = ref.dilate(img, m)
    kernels = {
        "box3": np.full((3, 3), 1.0 / 9.0, np.float32),                       # non-dyadic
        "sharpen": np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32),  # saturates
        "sobel_x": np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32),
        "k1x5": np.array([[1, 4, 6, 4, 1]], np.float32) / np.float32(16),
        "k5x1": (np.array([[1, 4, 6, 4, 1]], np.float32) / np.float32(16)).T.copy(),
        "k7x7": (np.outer([1, 6, 15, 20, 15, 6, 1], [1, 6, 15, 20, 15, 6, 1]) / 4096.0).astype(np.float32),
        "asym3x5": rng.integers(0, 8, (3, 5)).astype(np.float32) / np.float32(64),
    }
    for name, k in kernels.items():
        generic[f"kernel/{name}"] = k
        generic[f"filter2d/{name}"] = ref.filter2d(img, k)
    for t, mv in ((0, 255), (127, 255), (200, 99), (255, 255)):
        generic[f"threshold/{t}_{mv}"] = ref.threshold(img, t, mv)
    np.savez_compressed(os.path.join(HERE, "generic_cases.npz"), **generic)

    # saturated / maximum-gradient / rounding-tie patterns (regular, so they compress to a few KB)
    h, w = 67, 530
    yy, xx = np.mgrid[0:h, 0:w]
    pats = {
        "all255": np.full((h, w), 255), "columns": (xx % 2) * 255, "rows": (yy % 2) * 255, "checker": ((xx + yy) % 2) * 255,
        "ties8": ((xx * 3 + yy * 5) % 32) * 8, "ties2": ((xx + 2 * yy) % 4) * 2 + 1, "ramp": (xx + yy) % 256,
    }
    patterns = {}
    for name, plane in pats.items():
        img = np.ascontiguousarray(np.repeat(plane.astype(np.uint8)[:, :, None], 3, axis=2))
        patterns[f"{name}/input"] = img
        for prefix, fn in ref.OPERATIONS:
            patterns[f"{name}/{prefix}"] = fn(img)
    np.savez_compressed(os.path.join(HERE, "pattern_cases.npz"), **patterns)

    kat = {"cv2": cv.__version__, "inputs": {}, "outputs": {}}
    A = cv.imread(os.path.join(ROOT, "data", "fundus.jpg"))
    B = np.random.default_rng(0).integers(0, 256, (2336, 3504, 3), dtype=np.uint8)
    S = orc.synth(3504, 2336, 3, seed=0, frame=0)
    for name, img in (("A_fundus", A), ("B_rng0", B), ("S_synth_seed0_frame0", S)):
        kat["inputs"][name] = dict(shape=list(img.shape), **digest(img))
        kat["outputs"][name] = {prefix: digest(fn(img)) for prefix, fn in ref.OPERATIONS}
        # the same calls through cv.UMat, exactly as the reference holds its images (opencv:42-44)
        u = cv.UMat(img)
        assert digest(cv.bitwise_not(u).get()) == kat["outputs"][name]["inversion"]
        assert digest(cv.filter2D(u, -1, orc.BLUR_5X5).get()) == kat["outputs"][name]["convolution-gaussian-blur-5x5"]
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)

    # When the reference tree is present, run its literal perform_benchmark once as well so the
    # restated call list cannot drift from it (written as lossless PNG, so the files the
    # reference writes are compared byte-for-byte with the restated calls).
    refdir = "/root/reference/benchmarks/opencv"
    if os.path.isdir(refdir):
        import contextlib
        import importlib.util
        import io
        import tempfile

        spec = importlib.util.spec_from_file_location("ref_benchmark", os.path.join(refdir, "benchmark.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        tiny = rng.integers(0, 256, (32, 48, 3), dtype=np.uint8)
        with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()) as out:
            mod.perform_benchmark(tiny, "t.png", d, 1)
            written = sorted(os.listdir(d))
            files = {p: cv.imread(os.path.join(d, f"{p}-t.png"), cv.IMREAD_UNCHANGED) for p, _ in ref.OPERATIONS}
        rows = [l for l in out.getvalue().splitlines() if l.startswith("|")]
        assert len(rows) == 14, rows
        expect = sorted([f"{p}-t.png" for p, _ in ref.OPERATIONS] + ["-t.png"])
        assert written == expect, (written, expect)
        # PNG is lossless: the files the reference wrote must equal our restated calls
        for p, fn in ref.OPERATIONS:
            assert np.array_equal(files[p], fn(tiny)), p
        print("reference perform_benchmark cross-check: 14 rows, prefixes and all 12 output images match")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
