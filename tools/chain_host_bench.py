"""dip_pipeline_create_chain / dip_pipeline_run end to end: pinned host frames in, a 4-op chain on the device, final result back to pinned memory.
usage: python tools/chain_host_bench.py [frames=64] [chunk=4]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _dip_pkg import load
dip = load()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = dip.Context(0)
W, H, C = 3504, 2336, 3
mk = dip.make_op
ops = [mk("grayscale"), mk("gaussian_blur3x3"), mk("threshold", thresh=127, maxval=255), mk("erode", mask=np.ones((3, 3), np.uint8))]
src = ctx.pinned_empty((frames, H, W, C)); src[...] = np.random.default_rng(0).integers(0, 256, (1, H, W, C), dtype=np.uint8)
dst = ctx.pinned_empty((frames, H, W, C))
pipe = ctx.pipeline(W, H, C, chunk, ops, chain=True)   # device buffers and events kept across runs
pipe.run(src, frames, [dst])
best = None
for _ in range(3):
    t0 = time.perf_counter()
    pipe.run(src, frames, [dst])
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
pipe.close()
fb = W * H * C
print(f"{frames} frames, chunk {chunk}: {best / frames * 1e3:.3f} ms per frame, {frames * W * H / best / 1e6:.0f} Mpixel/s of frames "
      f"({4 * frames * W * H / best / 1e6:.0f} op-Mpixel/s), H2D {frames * fb / best / 1e9:.1f} GB/s + D2H {frames * fb / best / 1e9:.1f} GB/s")
