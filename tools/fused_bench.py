"""Fused 5x5 Gaussian -> 3x3 erosion kernel against the two kernels, on a one-channel mosaic band and an RGB stack.
usage: python tools/fused_bench.py [rounds=5]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _dip_pkg import load
dip = load()
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ctx = dip.Context(0)
mk = dip.make_op
b5 = np.array([1, 4, 6, 4, 1], np.float32) / 16
g5, er = mk("convolve", kernel=np.outer(b5, b5)), mk("erode", mask=np.ones((3, 3), np.uint8))
fused = mk("gauss5x5_erode3x3")
print("| workload | fused ms | two kernels ms | fused Gpixel/s | fraction of 6459 GB/s (2 B/pixel) |")
print("|---|---|---|---|---|")
for label, (w, h, c, n) in {"65536x16384x1 band": (65536, 16384, 1, 1), "64 x 3504x2336x3": (3504, 2336, 3, 64)}.items():
    src = ctx.image(w, h, c, frames=n).synth_fill(seed=1)
    tmp = ctx.image(w, h, c, frames=n); dst = ctx.image(w, h, c, frames=n)
    tf, t2 = [], []
    for r in range(rounds):
        ctx.apply(fused, src, dst)
        ctx.timer_start()
        for _ in range(3): ctx.apply(fused, src, dst)
        tf.append(ctx.timer_stop() / 3)
        ctx.timer_start()
        for _ in range(3):
            ctx.apply(g5, src, tmp); ctx.apply(er, tmp, dst)
        t2.append(ctx.timer_stop() / 3)
    px = w * h * n
    print(f"| {label} | {min(tf):.3f} / {statistics.median(tf):.3f} | {min(t2):.3f} / {statistics.median(t2):.3f} | {px / min(tf) / 1e6:.0f} | {2 * px * c / min(tf) / 1e6 / 6459:.2f} |")
    src.close(); tmp.close(); dst.close()
