"""Upload / download of 3504x2336x3 frames from pinned and pageable host memory (ms per frame, GB/s).
usage: python tools/xfer_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _dip_pkg import load
dip = load()
ctx = dip.Context(0)
W, H, C = 3504, 2336, 3
FRAME = W * H * C
print("| transfer | frames per call | ms per frame | GB/s |")
print("|---|---|---|---|")
for nb in (1, 8):
    pin = ctx.pinned_empty((nb, H, W, C)); pin[...] = 7
    page = np.full((nb, H, W, C), 9, np.uint8)
    stack = ctx.image(W, H, C, frames=nb)
    for name, buf in (("pinned", pin), ("pageable", page)):
        for direction in ("upload", "download"):
            fn = (lambda: stack.upload(buf)) if direction == "upload" else (lambda: stack.download(buf, sync=False))
            best = None
            for rep in range(4):
                fn(); ctx.sync()
                t0 = time.perf_counter()
                for _ in range(5):
                    fn()
                ctx.sync()
                dt = (time.perf_counter() - t0) / 5
                best = dt if best is None else min(best, dt)
            print(f"| {direction} ({name}) | {nb} | {best / nb * 1e3:.3f} | {nb * FRAME / best / 1e9:.1f} |")
    got = stack.download().reshape(nb, H, W, C)
    assert (got == 9).all() or (got == 7).all()
    stack.close()
