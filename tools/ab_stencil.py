"""Interleaved timing of the stencil ops inside one process (same GPU, same clocks).
usage: python tools/ab_stencil.py [frames=256] [rounds=4] [knob=value,...;knob=value,...]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _dip_pkg import load
dip = load()
L = dip.lib()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
variants = [("default", {})]
if len(sys.argv) > 3:
    variants = [(v, dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in v.split(",") if kv)) for v in sys.argv[3].split(";")]
ctx = dip.Context(0)
src = ctx.image(3504, 2336, 3, frames=frames).synth_fill(seed=0)
dst = ctx.image(3504, 2336, 3, frames=frames)
ops = {p: op for _, p, op in ctx.reference_ops()}
names = ["inversion", "grayscale", "erosion-cross", "erosion-square", "erosion-square-separated", "gaussian-blur-3x3", "convolution-gaussian-blur-3x3",
         "convolution-gaussian-blur-3x3-separated", "convolution-gaussian-blur-5x5", "convolution-gaussian-blur-5x5-separated"]
res = {(v, n): [] for v, _ in variants for n in names}
for r in range(rounds):
    for vname, knobs in variants:
        for k, val in knobs.items():
            L.dip_debug_set(k.encode(), val)
        for n in names:
            ctx.apply(ops[n], src, dst)
            ctx.timer_start()
            for _ in range(3):
                ctx.apply(ops[n], src, dst)
            res[(vname, n)].append(ctx.timer_stop() / 3 / frames * 1e3)
print("| variant | " + " | ".join(n.replace("convolution-gaussian-blur-", "conv") for n in names) + " |")
print("|---|" + "---|" * len(names))
for vname, _ in variants:
    print(f"| {vname} | " + " | ".join(f"{min(res[(vname, n)]):.2f}/{statistics.median(res[(vname, n)]):.2f}" for n in names) + " |")
