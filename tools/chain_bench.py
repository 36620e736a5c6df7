"""dip_apply_chain against the same ops applied one by one (256-frame stack of 3504x2336x3, one B200).
usage: python tools/chain_bench.py [frames=256] [rounds=4]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _dip_pkg import load
dip = load()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = dip.Context(0)
src = ctx.image(3504, 2336, 3, frames=frames).synth_fill(seed=0)
dst = ctx.image(3504, 2336, 3, frames=frames)
tmp = ctx.image(3504, 2336, 3, frames=frames)
mk = dip.make_op
cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], np.uint8)
b5 = np.array([1, 4, 6, 4, 1], np.float32) / 16
ops = {"gray": mk("grayscale"), "inv": mk("invert"), "thr": mk("threshold", thresh=127, maxval=255), "blur": mk("gaussian_blur3x3"),
       "erode": mk("erode", mask=np.ones((3, 3), np.uint8)), "cross": mk("erode", mask=cross),
       "g5": mk("convolve", kernel=np.outer(b5, b5)), "g5sep": mk("convolve_separable", kernel=b5, kernel_col=b5)}
chains = [["gray", "blur", "thr", "erode"], ["thr", "cross", "inv"], ["inv", "thr", "inv"], ["thr", "g5sep", "inv"], ["g5", "erode", "thr"],
          ["gray", "thr"], ["inv", "g5sep"], ["g5", "erode"], ["g5", "erode", "inv"], ["inv", "g5", "erode"],
          ["g5sep"], ["g5sep", "inv"], ["g5sep", "thr"], ["blur"], ["blur", "thr"], ["thr", "blur"]]
print("| chain | passes | chain us/frame | one by one us/frame | ratio |")
print("|---|---|---|---|---|")
for names in chains:
    lst = [ops[n] for n in names]
    tc, ts = [], []
    for r in range(rounds):
        ctx.apply_chain(lst, src, dst)
        ctx.timer_start()
        for _ in range(3):
            ctx.apply_chain(lst, src, dst)
        tc.append(ctx.timer_stop() / 3 / frames * 1e3)
        def seq():
            a = src
            for i, op in enumerate(lst):
                b = dst if (len(lst) - 1 - i) % 2 == 0 else tmp
                ctx.apply(op, a, b)
                a = b
        seq()
        ctx.timer_start()
        for _ in range(3):
            seq()
        ts.append(ctx.timer_stop() / 3 / frames * 1e3)
    p = dip.Context.chain_passes(lst, 3)
    print(f"| {' > '.join(names)} | {p} | {min(tc):.2f} / {statistics.median(tc):.2f} | {min(ts):.2f} / {statistics.median(ts):.2f} | {min(ts) / min(tc):.2f}x |")
