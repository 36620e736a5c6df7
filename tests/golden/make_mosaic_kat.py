"""Oracle checksum of BASELINE.json config 5 at FULL size (65536 x 65536 u8 mosaic, Gaussian 5x5 then
3x3 erosion), computed on the CPU in row strips with a 3-row halo (halo 3 reproduces the
whole-image rows exactly).  Takes a few minutes; writes tests/golden/mosaic_kat.json.

    python tests/golden/make_mosaic_kat.py [side=65536] [strip_rows=1024]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402


def mosaic_checksum(side: int, strip: int, seed: int = 0) -> int:
    total = 0
    for r0 in range(0, side, strip):
        r1 = min(side, r0 + strip)
        lo, hi = max(0, r0 - 3), min(side, r1 + 3)
        src = orc.synth_rows(side, hi - lo, 1, seed, lo)
        # the strip's top/bottom are true mosaic edges only when lo == 0 / hi == side; interior strip edges
        # carry 3 real halo rows, so the rows [r0, r1) are exact
        out = orc.erode(orc.filter2d(src, orc.BLUR_5X5), orc.SQUARE_MASK)
        total = (total + orc.checksum(out[r0 - lo: r1 - lo], first_row=r0)) & 0xFFFFFFFFFFFFFFFF
        print(f"rows {r0}-{r1} done", flush=True)
    return total


if __name__ == "__main__":
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    strip = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    cs = mosaic_checksum(side, strip)
    path = os.path.join(HERE, "mosaic_kat.json")
    data = {}
    if os.path.exists(path):
        data = json.load(open(path))
    data[str(side)] = f"{cs:016x}"
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(side, f"{cs:016x}")
