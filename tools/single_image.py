"""Back-to-back single-image calls over a ring of buffers larger than L2 (the per-op table's first column),
for several knob settings.  usage: python tools/single_image.py ["knob=value,knob=value" ...]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _dip_pkg import load
dip = load()
L = dip.lib()
ctx = dip.Context(0)
W, H, C, ring = 3504, 2336, 3, 8
srcs = [ctx.image(W, H, C).synth_fill(seed=3, first_frame=i) for i in range(ring)]
dsts = [ctx.image(W, H, C) for _ in range(ring)]
ops = {p: op for _, p, op in ctx.reference_ops()}
names = ["inversion", "erosion-cross", "erosion-square", "gaussian-blur-3x3", "convolution-gaussian-blur-3x3",
         "convolution-gaussian-blur-3x3-separated", "convolution-gaussian-blur-5x5", "convolution-gaussian-blur-5x5-separated"]
variants = sys.argv[1:] or ["tile_rows=0"]
res = {(v, n): [] for v in variants for n in names}
for rnd in range(5):
    for v in variants:
        for kv in v.split(","):
            L.dip_debug_set(kv.split("=")[0].encode(), int(kv.split("=")[1]))
        for n in names:
            for i in range(ring):
                ctx.apply(ops[n], srcs[i], dsts[i])
            ctx.timer_start()
            for r in range(40):
                ctx.apply(ops[n], srcs[r % ring], dsts[r % ring])
            res[(v, n)].append(ctx.timer_stop() / 40 * 1e3)
print("| knobs | " + " | ".join(n.replace("convolution-gaussian-blur-", "conv") for n in names) + " |")
print("|---|" + "---|" * len(names))
for v in variants:
    print(f"| {v} | " + " | ".join(f"{min(res[(v, n)]):.2f}/{statistics.median(res[(v, n)]):.2f}" for n in names) + " |")
