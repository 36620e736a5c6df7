"""Parity of the 12 reference ops against the oracle with development knobs set.
usage: python tests/soak/check_knob.py knob=value[,knob=value]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from _dip_pkg import load
from oracle import oracle as orc
orc.build()
dip = load()
L = dip.lib()
for kv in sys.argv[1].split(","):
    assert L.dip_debug_set(kv.split("=")[0].encode(), int(kv.split("=")[1])) == 0
ctx = dip.Context(0)
rng = np.random.default_rng(9)
bad = 0
for (n, h, w, c) in [(1, 1, 1, 3), (1, 7, 5, 3), (2, 61, 171, 3), (1, 200, 1031, 3), (3, 130, 700, 3), (1, 2336, 3504, 3), (1, 97, 513, 1), (1, 64, 600, 4)]:
    stack = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    src = ctx.from_array(stack); dst = ctx.image_like(src)
    for k, (prefix, name, args) in enumerate(dip.REFERENCE_OPS):
        getattr(ctx, name)(src, dst, *args)
        got = dst.download().reshape(stack.shape)
        for f in range(n):
            if not np.array_equal(got[f], orc.OPERATIONS[k][1](stack[f])):
                bad += 1
                print("MISMATCH", prefix, (n, h, w, c), f)
    src.close(); dst.close()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
