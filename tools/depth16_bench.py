"""The 12 reference operations on 16-bit images (dip_image_create16): us per 3504x2336 RGB frame on a stack, and the
fraction of the measured copy bandwidth (algorithmic bytes = 2 x W x H x C x 2 per frame).
usage: python tools/depth16_bench.py [frames=16]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _dip_pkg import load
dip = load()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    peak = 6459.0
ctx = dip.Context(0)
W, H, C = 3504, 2336, 3
src = ctx.image16(W, H, C, frames).synth_fill(seed=0)
dst = ctx.image16(W, H, C, frames)
alg = 2 * W * H * C * 2
print(f"| operation, 16-bit samples ({frames}-frame stack) | us per frame | GB/s | fraction of {peak:.0f} GB/s |")
print("|---|---|---|---|")
for prefix, name, args in dip.REFERENCE_OPS:
    fn = lambda: getattr(ctx, name)(src, dst, *args)  # noqa: E731
    fn(); fn()
    ctx.timer_start()
    for _ in range(3):
        fn()
    t = ctx.timer_stop() / 3 / frames
    print(f"| {prefix} | {t * 1e3:.1f} | {alg / t / 1e6:.0f} | {alg / t / 1e6 / peak:.2f} |")
