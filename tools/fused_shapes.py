"""The fused Gaussian 5x5 -> erosion 3x3 kernel on a one-channel mosaic band for several settings of the "fused_shape" knob
(band rows x warps per CTA, register caps): checksum against the default shape, then ms per launch in the sustained power
state (0.5 s of launches before the timed ones).  usage: python tools/fused_shapes.py [shape ...]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _dip_pkg import load
dip = load(); L = dip.lib()
ctx = dip.Context(0)
fused = dip.make_op("gauss5x5_erode3x3")
shapes = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 5, 6, 7]
WARM, TIMED = int(os.environ.get("FUSED_WARM", "60")), int(os.environ.get("FUSED_TIMED", "20"))   # 1000 / 400: the sustained power state
names = {0: "shipped (30 x 2 on one channel, 15 x 4 on RGB)", 14: "15 x 4", 1: "30 x 2", 2: "30 x 4", 3: "15 x 4, 5 CTAs/SM by registers", 4: "15 x 2", 5: "30 x 3", 6: "15 x 4, 6 CTAs/SM by registers", 7: "45 x 2", 8: "15 x 4, row loop not unrolled", 9: "15 x 4, row loop unrolled 15x", 10: "15 x 4, row loop unrolled 3x", 11: "60 x 2", 12: "60 x 1", 13: "45 x 3"}
RGB = os.environ.get("FUSED_RGB") == "1"      # a 64-frame 3504x2336 RGB stack instead of the one-channel band (shapes 0 / 14 and 1 only)
w, h, c, n = (3504, 2336, 3, 64) if RGB else (65536, 16384, 1, 1)
src = ctx.image(w, h, c, frames=n).synth_fill(seed=1); dst = ctx.image(w, h, c, frames=n)
L.dip_debug_set(b"fused_shape", 0)
ctx.apply(fused, src, dst)
want = dst.checksum()
res = {s: [] for s in shapes}
for rnd in range(4):
    for s in shapes:
        L.dip_debug_set(b"fused_shape", s)
        ctx.apply(fused, src, dst)
        assert dst.checksum() == want, f"shape {s} differs"
        for _ in range(WARM): ctx.apply(fused, src, dst)
        ctx.timer_start()
        for _ in range(TIMED): ctx.apply(fused, src, dst)
        res[s].append(ctx.timer_stop() / TIMED)
print(f"| fused_shape | ms per {'64 x 3504x2336x3 stack' if RGB else '65536x16384 band'}, min / median | GB/s (2 B per byte) | fraction of 6459 |")
print("|---|---|---|---|")
for s in shapes:
    t = min(res[s])
    print(f"| {s}: {names.get(s, '')} | {t:.4f} / {statistics.median(res[s]):.4f} | {2 * w * h * c * n / t / 1e6:.0f} | {2 * w * h * c * n / t / 1e6 / 6459:.3f} |")
