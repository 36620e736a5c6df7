"""Run ONE reference op a few times on a stack (the command ncu wraps: few launches, nothing else).
usage: python tools/prof_one.py <prefix|fused> [frames=64] [launches=3] [knob=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _dip_pkg import load
dip = load()
L = dip.lib()
prefix = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for kv in sys.argv[4:]:
    L.dip_debug_set(kv.split("=")[0].encode(), int(kv.split("=")[1]))
ctx = dip.Context(0)
c = 1 if prefix == "fused" else 3
src = ctx.image(3504, 2336, c, frames=frames).synth_fill(seed=0)
dst = ctx.image(3504, 2336, c, frames=frames)
op = ctx.fused_chain_op() if prefix == "fused" else {p: o for _, p, o in ctx.reference_ops()}[prefix]
ctx.timer_start()
for _ in range(launches):
    ctx.apply(op, src, dst)
ms = ctx.timer_stop()
print(f"{prefix}: {ms / launches / frames * 1e3:.2f} us per frame over {launches} launches of {frames} frames")
