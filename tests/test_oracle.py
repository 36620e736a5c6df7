"""CPU: the oracle (oracle/dip_oracle.c) against the golden vectors generated from the
reference's own cv2 calls (tests/golden/make_golden.py), and its own invariants."""
import hashlib
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def digest(a):
    b = np.ascontiguousarray(a).tobytes()
    return {"sha256_16": hashlib.sha256(b).hexdigest()[:16], "crc32": f"{zlib.crc32(b):08x}"}


def test_small_cases_match_reference_outputs(orc, golden_small):
    keys = sorted({k.split("/")[0] for k in golden_small.files})
    assert len(keys) == 36
    for key in keys:
        img = golden_small[f"{key}/input"]
        for prefix, fn in orc.OPERATIONS:
            want = golden_small[f"{key}/{prefix}"]
            got = fn(img)
            assert np.array_equal(got, want), f"{key} {prefix}: {(got != want).sum()} bytes differ"


def test_generic_arguments_match_reference_outputs(orc, golden_generic):
    g = golden_generic
    img = g["input"]
    for name in [k.split("/", 1)[1] for k in g.files if k.startswith("mask/")]:
        assert np.array_equal(orc.erode(img, g[f"mask/{name}"]), g[f"erode/{name}"]), name
        assert np.array_equal(orc.dilate(img, g[f"mask/{name}"]), g[f"dilate/{name}"]), name
    for name in [k.split("/", 1)[1] for k in g.files if k.startswith("kernel/")]:
        got = orc.filter2d(img, g[f"kernel/{name}"]).astype(int)
        want = g[f"filter2d/{name}"].astype(int)
        # dyadic kernels are exact; the non-dyadic box filter may differ by float summation order
        tol = 1 if name == "box3" else 0
        assert np.abs(got - want).max() <= tol, name
    for key in [k for k in g.files if k.startswith("threshold/")]:
        t, mv = (int(v) for v in key.split("/")[1].split("_"))
        assert np.array_equal(orc.threshold(img, t, mv), g[key]), key


def test_full_size_known_answers(orc, kat):
    """hashes of every op's output on the full-size inputs."""
    B = np.random.default_rng(0).integers(0, 256, (2336, 3504, 3), dtype=np.uint8)
    assert digest(B) == {k: kat["inputs"]["B_rng0"][k] for k in ("sha256_16", "crc32")}
    S = orc.synth(3504, 2336, 3, seed=0, frame=0)
    assert digest(S)["sha256_16"] == kat["inputs"]["S_synth_seed0_frame0"]["sha256_16"]
    for prefix, fn in orc.OPERATIONS:
        assert digest(fn(B)) == kat["outputs"]["B_rng0"][prefix], prefix
    # the synthetic generator's frame: only the two cheapest and the two rounding-critical ops
    for prefix, fn in orc.OPERATIONS:
        if prefix in ("grayscale", "erosion-cross", "convolution-gaussian-blur-3x3-separated", "gaussian-blur-3x3"):
            assert digest(fn(S)) == kat["outputs"]["S_synth_seed0_frame0"][prefix], prefix


def test_fundus_known_answers(orc, kat):
    """The HRF fundus image (BASELINE.json configs 0-2).  The committed lossless crop of the reference decode
    (tests/golden/make_fundus_crop.py) always runs; the full-size hashes run when this box's JPEG decoder is the one that
    made them (the decoded pixels differ otherwise) -- nothing is skipped."""
    crop = np.load(os.path.join(ROOT, "tests", "golden", "fundus_crop.npz"))
    for prefix, fn in orc.OPERATIONS:
        assert np.array_equal(fn(crop["input"]), crop[prefix]), f"fundus crop {prefix}"
    import cv2
    img = cv2.imread(os.path.join(ROOT, "data", "fundus.jpg"))
    assert img is not None and img.shape == (2336, 3504, 3)
    if digest(img)["sha256_16"] == kat["inputs"]["A_fundus"]["sha256_16"]:
        for prefix, fn in orc.OPERATIONS:
            assert digest(fn(img)) == kat["outputs"]["A_fundus"][prefix], prefix
    else:  # another libjpeg: the crop of the reference decode must at least be close to what this one decodes
        assert np.abs(img[1090:1250, 2350:2590].astype(int) - crop["input"].astype(int)).max() <= 8


def test_grayscale_exhaustive_formula(orc):
    """Y = (3735 B + 19235 G + 9798 R + 16384) >> 15 over a dense colour lattice + numpy restatement."""
    v = np.arange(0, 256, 5, dtype=np.uint8)
    b, g, r = np.meshgrid(v, v, v, indexing="ij")
    img = np.stack([b.ravel(), g.ravel(), r.ravel()], axis=1).reshape(1, -1, 3)
    y = ((img[..., 0].astype(np.uint32) * 3735 + img[..., 1].astype(np.uint32) * 19235 + img[..., 2].astype(np.uint32) * 9798 + 16384) >> 15).astype(np.uint8)
    out = orc.grayscale(img)
    assert np.array_equal(out, np.repeat(y[..., None], 3, axis=2))


def test_invariants(orc):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (37, 29, 3), dtype=np.uint8)
    assert np.array_equal(orc.invert(orc.invert(img)), img)
    assert np.array_equal(orc.erode_separable(img), orc.erode(img, orc.SQUARE_MASK))
    th = orc.threshold(img)
    assert np.array_equal(orc.threshold(th), th)
    flat = np.full((9, 11, 4), 77, np.uint8)
    for k in (orc.BLUR_3X3, orc.BLUR_5X5):
        assert np.array_equal(orc.filter2d(flat, k), flat)
    assert np.array_equal(orc.gaussian_blur3x3(flat), flat)
    # the full and the separable Gaussian differ only by double rounding: at most 1 LSB
    d = orc.filter2d(img, orc.BLUR_5X5).astype(int) - orc.filter2d_separable(img, orc.BLUR_5X5_1X5, orc.BLUR_5X5_5X1).astype(int)
    assert np.abs(d).max() <= 1
    assert orc.checksum(img) != orc.checksum(orc.invert(img))


def test_numpy_restatement_of_packed_rounding():
    """The identity the CUDA kernels use: RNE(S / 2^s) = (R + ((R >> s) & 1)) >> s with R = S + 2^(s-1) - 1."""
    for s, smax in ((2, 1020), (4, 4080), (8, 65280)):
        S = np.arange(0, smax + 1, dtype=np.int64)
        R = S + (1 << (s - 1)) - 1
        got = (R + ((R >> s) & 1)) >> s
        want = np.rint(S / float(1 << s)).astype(np.int64)  # numpy rint is round-half-even
        assert np.array_equal(got, want), s


def test_numpy_restatement_of_half_precision_cascade():
    """The fused 1x5+5x1 Gaussian's vertical pass in IEEE half precision (stencil_tiled.cuh, "vertical cascade"):
    rows arrive as halves (h - 128) u with u = 2^-24 (the subnormal ulp), [1 4 6 4 1] is four rounds of pairwise sums,
    every partial sum stays an integer multiple of u in [-2048, 2047] (exact), and one fused multiply-add
    o * 2^-4 + 128 u rounds to the ulp, i.e. to round-half-even(S / 16) as a byte pattern."""
    u = np.float16(2.0 ** -24)
    assert np.array_equal((np.arange(256) * np.float64(u)).astype(np.float16).view(np.uint16), np.arange(256))  # byte b IS b u
    rng = np.random.default_rng(5)
    h = rng.integers(0, 256, (5, 200000))
    h[:, :4] = [[0, 255, 0, 255], [0, 255, 255, 0], [0, 255, 0, 255], [0, 255, 255, 0], [0, 255, 0, 255]]
    hb = ((h - 128) * np.float64(u)).astype(np.float16)  # what the horizontal pass's rounding FMA leaves with a zero addend
    p = [hb[k] + hb[k + 1] for k in range(4)]            # numpy adds float16 with IEEE rounding, like HADD2
    v = [p[k] + p[k + 1] for k in range(3)]
    q = [v[k] + v[k + 1] for k in range(2)]
    o = q[0] + q[1]
    S = h[0] + 4 * h[1] + 6 * h[2] + 4 * h[3] + h[4]
    assert np.array_equal(o.astype(np.float64) / np.float64(u), S - 2048)  # exact, centred
    # the rounding FMA: exact product and sum (float64 holds them), ONE rounding to half
    res = (o.astype(np.float64) / 16 + 128 * np.float64(u)).astype(np.float16).view(np.uint16)
    assert np.array_equal(res, np.rint(S / 16.0).astype(np.int64))


def test_numpy_restatement_of_the_five_instruction_gaussian():
    """The horizontal [1 4 6 4 1] / 16 pass of the 2-warp stack shape (stencil_tiled.cuh, half_gauss5_v2):
    S - 2048 = (a + e + 2c - 2048) + 4 (b + c + d).  For 0 <= t <= 1020 the half (t - 2048) u has the bit pattern 0x8800 - t, so
    the first bracket is integer arithmetic on the byte patterns; 4 v + K is exact in half precision and one fused multiply-add
    rounds S / 16 half to even (addend 128 u) or delivers the centred row (addend 0)."""
    u = np.float64(2.0 ** -24)
    t = np.arange(0, 1021)
    assert np.array_equal((0x8800 - t).astype(np.uint16).view(np.float16).astype(np.float64) / u, t - 2048)
    rng = np.random.default_rng(11)
    a, b, c, d, e = rng.integers(0, 256, (5, 300000))
    a[:4], b[:4], c[:4], d[:4], e[:4] = [0, 255, 0, 255], [0, 255, 255, 0], [0, 255, 0, 255], [0, 255, 255, 0], [0, 255, 0, 255]
    K = (0x8800 - a - e - 2 * c).astype(np.uint16).view(np.float16)
    v = (b + c + d).astype(np.uint16).view(np.float16)                    # <= 765: a subnormal half, bit pattern = value
    y = (v.astype(np.float64) * 4 + K.astype(np.float64)).astype(np.float16)   # the FMA: exact product and sum, one rounding
    S = a + 4 * b + 6 * c + 4 * d + e
    assert np.array_equal(y.astype(np.float64) / u, S - 2048)
    assert np.array_equal((y.astype(np.float64) / 16 + 128 * u).astype(np.float16).view(np.uint16), np.rint(S / 16.0).astype(np.int64))
    assert np.array_equal((y.astype(np.float64) / 16).astype(np.float16).astype(np.float64) / u, np.rint(S / 16.0) - 128)


def test_row_addressed_synth_and_checksum_compose(orc):
    """Bands of a mosaic: synth_rows stitches to the whole, band checksums add up to the whole's."""
    whole = orc.synth(64, 50, 1, seed=3, frame=0)
    parts = [orc.synth_rows(64, 20, 1, 3, 0), orc.synth_rows(64, 30, 1, 3, 20)]
    assert np.array_equal(np.concatenate(parts, axis=0), whole)
    total = (orc.checksum(parts[0], first_row=0) + orc.checksum(parts[1], first_row=20)) & 0xFFFFFFFFFFFFFFFF
    assert total == orc.checksum(whole)
    # frame f of a stack starts at global row f * height
    assert np.array_equal(orc.synth(64, 50, 1, seed=3, frame=2), orc.synth_rows(64, 50, 1, 3, 100))


def test_pattern_cases_match_reference_outputs(orc):
    """Saturated / maximum-gradient / rounding-tie patterns: the reference's cv2 outputs (pattern_cases.npz)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "pattern_cases.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) == 7
    for name in names:
        img = g[f"{name}/input"]
        for prefix, fn in orc.OPERATIONS:
            assert np.array_equal(fn(img), g[f"{name}/{prefix}"]), f"{name} {prefix}"


_U16_EXTRA = {
    "threshold16": lambda o16, a: o16.threshold(a, 30000, 65535),
    "dilate-5x3": lambda o16, a: o16.dilate(a, np.array([[1, 0, 1, 1, 0], [0, 1, 1, 0, 1], [1, 1, 0, 0, 1]], np.uint8)),
    "erode-7x7": lambda o16, a: o16.erode(a, (np.arange(49).reshape(7, 7) % 3 != 0).astype(np.uint8)),
    "erode-9x9": lambda o16, a: o16.erode(a, np.ones((9, 9), np.uint8)),
    "sharpen-saturating": lambda o16, a: o16.filter2d(a, np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)),
    "conv-9x1-dyadic": lambda o16, a: o16.filter2d(a, (np.array([1, 2, 4, 8, 2, 8, 4, 2, 1], np.float32) / 32).reshape(9, 1)),
}


def test_16bit_restatement_matches_reference_outputs():
    """oracle/oracle16.py (numpy) against the reference's cv2 calls on CV_16U input (tests/golden/u16_cases.npz, written by
    make_golden_u16.py with cv2 4.13): 7 shapes x {1,3,4} channels x 18 operations, 0 mismatching samples."""
    from oracle import oracle16 as o16
    g = np.load(os.path.join(ROOT, "tests", "golden", "u16_cases.npz"))
    cases = sorted({k.split("/")[0] for k in g.files})
    assert len(cases) == 21
    fns = {name: (lambda a, f=fn: f(a)) for name, fn in o16.OPERATIONS}
    fns.update({name: (lambda a, f=fn: f(o16, a)) for name, fn in _U16_EXTRA.items()})
    for key in cases:
        img = g[f"{key}/input"]
        assert img.dtype == np.uint16
        for name, fn in fns.items():
            assert np.array_equal(fn(img), g[f"{key}/{name}"]), (key, name)
