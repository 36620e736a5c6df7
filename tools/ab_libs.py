"""A/B of compile-time variants of the library (make BUILD=build_x OUT=../libdip_b200_x.so EXTRA=-D...): every variant
runs in its own process (DIP_B200_LIB), the rounds are interleaved so that all variants see the same clocks.
usage: python tools/ab_libs.py <op-prefix> <rounds> <lib-suffix> [<lib-suffix> ...]     ('' = the shipped library)
child: python tools/ab_libs.py --child <op-prefix> <frames>
Per variant: us per frame of the op alone (10 back-to-back launches after 0.4 s of the 12-op chain as warm-up, i.e. in
the sustained power state), us per frame of the op inside the chain (its lap of 5 chain passes) and the result's checksum."""
import os, sys, subprocess, statistics, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(opname, frames):
    from _dip_pkg import load
    dip = load()
    ctx = dip.Context(0)
    src = ctx.image(3504, 2336, 3, frames=frames).synth_fill(seed=0)
    dst = ctx.image(3504, 2336, 3, frames=frames)
    ops = [(p, op) for _, p, op in ctx.reference_ops()]
    target = dict(ops)[opname]
    for _ in range(16):  # ~0.4 s of the whole chain
        for _, op in ops:
            ctx.apply(op, src, dst)
    ctx.sync()
    ctx.timer_start()
    for _ in range(10):
        ctx.apply(target, src, dst)
    alone = ctx.timer_stop() / 10 / frames * 1e3
    lap = 0.0
    for _ in range(5):
        for p, op in ops:
            if p == opname:
                ctx.sync(); ctx.timer_start(); ctx.apply(op, src, dst); lap += ctx.timer_stop()
            else:
                ctx.apply(op, src, dst)
    ctx.apply(target, src, dst)
    print(json.dumps({"alone": alone, "chain": lap / 5 / frames * 1e3, "checksum": "%x" % (sum(dst.checksum()) & (2**64 - 1))}))


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    opname, rounds, libs = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    res = {l: [] for l in libs}
    for r in range(rounds):
        for l in libs:
            env = dict(os.environ)
            if l:
                env["DIP_B200_LIB"] = os.path.join(ROOT, "dip-benchmark_b200", f"libdip_b200_{l}.so")
            out = subprocess.run([sys.executable, __file__, "--child", opname, "256"], env=env, capture_output=True, text=True)
            if out.returncode != 0:
                print(l, "FAILED", out.stderr[-400:])
                continue
            res[l].append(json.loads(out.stdout.strip().splitlines()[-1]))
    print(f"| variant | {opname}: alone us/frame min/median | in the chain min/median | checksum |")
    print("|---|---|---|---|")
    for l in libs:
        a = [x["alone"] for x in res[l]]; c = [x["chain"] for x in res[l]]
        if a:
            print(f"| {l or 'shipped'} | {min(a):.2f}/{statistics.median(a):.2f} | {min(c):.2f}/{statistics.median(c):.2f} | {res[l][0]['checksum']} |")
