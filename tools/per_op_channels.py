"""Per-op timing of a 64-frame 3504x2336 stack for 1, 3 and 4 channels (us per frame, GB/s algorithmic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _dip_pkg import load
dip = load()
ctx = dip.Context(0)
W, H, N = 3504, 2336, 64
print("| op | " + " | ".join(f"C={c} us/frame (GB/s)" for c in (1, 3, 4)) + " |")
print("|---|---|---|---|")
rows = {}
for c in (1, 3, 4):
    src = ctx.image(W, H, c, frames=N).synth_fill(seed=0)
    dst = ctx.image(W, H, c, frames=N)
    for desc, prefix, op in ctx.reference_ops() + [("fused", "gauss5x5+erode3x3 fused", ctx.fused_chain_op())]:
        for _ in range(2):
            ctx.apply(op, src, dst)
        ctx.timer_start()
        for _ in range(5):
            ctx.apply(op, src, dst)
        t = ctx.timer_stop() / 5 / N * 1e3
        rows.setdefault(prefix, []).append(f"{t:.2f} ({2 * W * H * c / t / 1e3:.0f})")
    src.close(); dst.close()
for k, v in rows.items():
    print(f"| {k} | " + " | ".join(v) + " |")
