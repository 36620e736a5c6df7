"""One reference operation on the 256-frame stack in the SUSTAINED power state, per setting of a development knob:
per setting, WARM launches (default 100, ~0.3 s) and then TIMED launches (default 50) under one CUDA-event pair; the
rounds are interleaved and the result's checksum is compared with the first setting's.
usage: python tools/sustained_op.py <op-prefix> <knob> <value> [<value> ...]      env: SUST_WARM, SUST_TIMED, SUST_FRAMES"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _dip_pkg import load
dip = load(); L = dip.lib()
opname, knob, values = sys.argv[1], sys.argv[2], [int(v) for v in sys.argv[3:]]
WARM, TIMED, FRAMES = int(os.environ.get("SUST_WARM", "100")), int(os.environ.get("SUST_TIMED", "50")), int(os.environ.get("SUST_FRAMES", "256"))
ctx = dip.Context(0)
src = ctx.image(3504, 2336, 3, frames=FRAMES).synth_fill(seed=0)
dst = ctx.image(3504, 2336, 3, frames=FRAMES)
op = {p: o for _, p, o in ctx.reference_ops()}[opname]
res = {v: [] for v in values}
want = None
for rnd in range(3):
    for v in values:
        L.dip_debug_set(knob.encode(), v)
        ctx.apply(op, src, dst)
        cs = sum(dst.checksum()) & (2**64 - 1)
        want = cs if want is None else want
        assert cs == want, f"{knob}={v} changes the result"
        for _ in range(WARM): ctx.apply(op, src, dst)
        ctx.timer_start()
        for _ in range(TIMED): ctx.apply(op, src, dst)
        res[v].append(ctx.timer_stop() / TIMED / FRAMES * 1e3)
L.dip_debug_set(knob.encode(), 0)
alg = 2 * 3504 * 2336 * 3
print(f"| {knob} | {opname}: us per frame, min / median of 3 | GB/s | fraction of 6459 |")
print("|---|---|---|---|")
for v in values:
    t = statistics.median(res[v])
    print(f"| {v} | {min(res[v]):.2f} / {t:.2f} | {alg / t / 1e3:.0f} | {alg / t / 1e3 / 6459:.3f} |")
