"""Oracle checksum of BASELINE.json config 5 at FULL size (65536 x 65536 u8 mosaic, Gaussian 5x5 then
3x3 erosion), computed on the CPU in row strips with a 3-row halo (halo 3 reproduces the
whole-image rows exactly).  Takes a few minutes; writes tests/golden/mosaic_kat.json.

    python tests/golden/make_mosaic_kat.py [side=65536] [strip_rows=1024] [iterations=1]
(iterations = 2 writes the key "<side>_iter2": the pipeline applied twice, for MosaicBand(iterate=True))
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402


def mosaic_checksum(side: int, strip: int, seed: int = 0, iterations: int = 1) -> int:
    """Checksum of the pipeline applied `iterations` times (each result is the next input: MosaicBand iterate=True)."""
    total = 0
    halo = 3 * iterations
    for r0 in range(0, side, strip):
        r1 = min(side, r0 + strip)
        lo, hi = max(0, r0 - halo), min(side, r1 + halo)
        cur = orc.synth_rows(side, hi - lo, 1, seed, lo)
        # the strip's top/bottom are true mosaic edges only when lo == 0 / hi == side; interior strip edges carry
        # 3 real halo rows per iteration (every iteration invalidates 3 more rows inwards), so rows [r0, r1) are exact
        for _ in range(iterations):
            cur = orc.erode(orc.filter2d(cur, orc.BLUR_5X5), orc.SQUARE_MASK)
        total = (total + orc.checksum(cur[r0 - lo: r1 - lo], first_row=r0)) & 0xFFFFFFFFFFFFFFFF
        print(f"rows {r0}-{r1} done", flush=True)
    return total


if __name__ == "__main__":
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    strip = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    iterations = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    cs = mosaic_checksum(side, strip, iterations=iterations)
    path = os.path.join(HERE, "mosaic_kat.json")
    data = {}
    if os.path.exists(path):
        data = json.load(open(path))
    data[str(side) if iterations == 1 else f"{side}_iter{iterations}"] = f"{cs:016x}"
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(side, iterations, f"{cs:016x}")
