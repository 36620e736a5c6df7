"""Erosion / 5x5 Gaussian on a one-channel mosaic band, per knob setting.
usage: python tools/band_bench.py ["knob=value" ...]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _dip_pkg import load
dip = load(); L = dip.lib()
ctx = dip.Context(0)
mk = dip.make_op
b5 = np.array([1, 4, 6, 4, 1], np.float32) / 16
ops = {"g5": mk("convolve", kernel=np.outer(b5, b5)), "erode": mk("erode", mask=np.ones((3, 3), np.uint8)), "blur3": mk("gaussian_blur3x3"), "inv": mk("invert")}
variants = sys.argv[1:] or ["no_pad=0"]
for shape in [(65536, 8192, 1, 1), (65536, 16384, 1, 1), (3504, 2336, 3, 64)]:
    w, h, c, n = shape
    src = ctx.image(w, h, c, frames=n).synth_fill(seed=1); dst = ctx.image(w, h, c, frames=n)
    for v in variants:
        for kv in v.split(","):
            L.dip_debug_set(kv.split("=")[0].encode(), int(kv.split("=")[1]))
        out = []
        for name, op in ops.items():
            ts = []
            for r in range(4):
                ctx.apply(op, src, dst)
                ctx.timer_start()
                for _ in range(3): ctx.apply(op, src, dst)
                ts.append(ctx.timer_stop() / 3)
            out.append(f"{name} {min(ts):.4f} ms ({2 * w * h * c * n / min(ts) / 1e6:.0f} GB/s)")
        print(shape, v, " | ".join(out))
    src.close(); dst.close()

# the two kernels alternating, as a mosaic step runs them (sustained)
w, h = 65536, 8192
src = ctx.image(w, h, 1).synth_fill(seed=1); tmp = ctx.image(w, h, 1); dst = ctx.image(w, h, 1)
for v in variants:
    for kv in v.split(","):
        L.dip_debug_set(kv.split("=")[0].encode(), int(kv.split("=")[1]))
    ts = []
    for r in range(5):
        for _ in range(200):
            ctx.apply(ops["g5"], src, tmp); ctx.apply(ops["erode"], tmp, dst)
        ctx.timer_start()
        for _ in range(50):
            ctx.apply(ops["g5"], src, tmp); ctx.apply(ops["erode"], tmp, dst)
        ts.append(ctx.timer_stop() / 50)
    print("alternating g5 -> erode on 65536x8192x1:", v, f"{min(ts):.4f} / {statistics.median(ts):.4f} ms per pair")
