import sys, time
sys.path.insert(0, '/root/repo')
from _dip_pkg import load
import ctypes as C
dip = load(); L = dip.lib()
ctx = dip.Context(0)
src = ctx.image(3504, 2336, 3).synth_fill(seed=1); dst = ctx.image(3504, 2336, 3)
ops = {p: op for _, p, op in ctx.reference_ops()}
for name in ("inversion", "erosion-square", "convolution-gaussian-blur-5x5-separated"):
    op = ops[name]
    for _ in range(100): L.dip_apply(ctx._h, C.byref(op), src._h, dst._h)
    ctx.sync()
    n = 5000
    t0 = time.perf_counter()
    for _ in range(n): L.dip_apply(ctx._h, C.byref(op), src._h, dst._h)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    print(f"{name}: enqueue {1e6*(t1-t0)/n:.2f} us/call, total {1e6*(t2-t0)/n:.2f} us/call")
# empty-ish call cost
tiny_a = ctx.image(16, 4, 3); tiny_b = ctx.image(16, 4, 3)
op = ops["inversion"]
t0 = time.perf_counter()
for _ in range(5000): L.dip_apply(ctx._h, C.byref(op), tiny_a._h, tiny_b._h)
t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
print(f"tiny invert: enqueue {1e6*(t1-t0)/5000:.2f} us/call, total {1e6*(t2-t0)/5000:.2f}")
