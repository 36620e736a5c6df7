"""Run selected reference ops a few times on a synthetic stack (profiling helper).
usage: python tools/run_ops.py FRAMES REPS [knob=value ...] prefix [prefix...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _dip_pkg import load
dip = load()
frames, reps = int(sys.argv[1]), int(sys.argv[2])
rest = sys.argv[3:]
for a in [a for a in rest if "=" in a]:
    k, v = a.split("=")
    dip.lib().dip_debug_set(k.encode(), int(v))
ctx = dip.Context(0)
src = ctx.image(3504, 2336, 3, frames=frames).synth_fill(seed=0)
dst = ctx.image(3504, 2336, 3, frames=frames)
ops = {p: op for _, p, op in ctx.reference_ops()}
for prefix in [a for a in rest if "=" not in a]:
    for _ in range(reps):
        ctx.apply(ops[prefix], src, dst)
ctx.sync()
print("ok", dst.checksum()[0])
