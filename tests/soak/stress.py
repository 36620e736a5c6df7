"""Randomised parity soak against the CPU oracle (test infrastructure; not part of the product path).
usage: python tests/soak/stress.py [trials=400] [seed=1]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from _dip_pkg import load
from oracle import oracle as orc
dip = load()
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 400
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = dip.Context(0)
fused = lambda im: orc.erode(orc.filter2d(im, orc.BLUR_5X5), orc.SQUARE_MASK)  # noqa: E731
OPS = list(zip(dip.REFERENCE_OPS, [f for _, f in orc.OPERATIONS])) + [(("fused", "gauss5x5_erode3x3", ()), fused)]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_parity_gpu import _chain_pool  # the op pool of the chain tests: name -> (dip_op, oracle function)
POOL = _chain_pool(dip, orc)
POOL_NAMES = sorted(POOL)
t0 = time.time()
fails = 0
for t in range(trials):
    kind = rng.integers(0, 13)
    c = int(rng.choice([1, 3, 4]))
    if kind >= 10:    # op chain (dip_apply_chain) against the ops applied one by one by the oracle
        h = int(rng.integers(1, 200)); w = int(rng.integers(1, 1100)); n = int(rng.choice([1, 1, 2]))
        img = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
        if rng.random() < 0.3:
            img = (img // 32 * 36).astype(np.uint8)
        names = [POOL_NAMES[int(i)] for i in rng.integers(0, len(POOL_NAMES), int(rng.integers(1, 8)))]
        src = ctx.from_array(img); dst = ctx.image_like(src)
        ctx.apply_chain([POOL[nm][0] for nm in names], src, dst)
        got = dst.download().reshape(img.shape)
        ok = True
        for f in range(n):
            want = img[f]
            for nm in names:
                want = POOL[nm][1](want)
            ok = ok and np.array_equal(got[f], want)
        desc = f"chain {names} on {n}x{h}x{w}x{c}"
        src.close(); dst.close()
    elif kind < 6:      # plain op on a random stack
        h = int(rng.integers(1, 260)); w = int(rng.integers(1, 1300)); n = int(rng.choice([1, 1, 1, 2, 5]))
        img = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
        if rng.random() < 0.3:
            img = (img // 32 * 36).astype(np.uint8)      # many ties
        (prefix, name, args), ofn = OPS[int(rng.integers(0, len(OPS)))]
        src = ctx.from_array(img); dst = ctx.image_like(src)
        getattr(ctx, name)(src, dst, *args)
        got = dst.download().reshape(img.shape)
        ok = all(np.array_equal(got[f], ofn(img[f])) for f in range(n))
        desc = f"{prefix} {n}x{h}x{w}x{c}"
        src.close(); dst.close()
    elif kind < 8:    # generic mask / kernel
        h = int(rng.integers(1, 90)); w = int(rng.integers(1, 200))
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        src = ctx.from_array(img); dst = ctx.image_like(src)
        if rng.random() < 0.5:
            mh, mw = int(rng.integers(1, 8)), int(rng.integers(1, 8))
            mask = (rng.random((mh, mw)) < 0.6).astype(np.uint8)
            if mask.sum() == 0:
                mask[mh // 2, mw // 2] = 1
            dil = rng.random() < 0.5
            (ctx.dilate if dil else ctx.erode)(src, dst, mask)
            want = (orc.dilate if dil else orc.erode)(img, mask)
            desc = f"{'dilate' if dil else 'erode'} mask {mh}x{mw} on {h}x{w}x{c}"
            ok = np.array_equal(dst.download().reshape(img.shape), want)
        else:
            kh, kw = int(rng.choice([1, 3, 5, 7])), int(rng.choice([1, 3, 5, 7]))
            k = (rng.integers(0, 9, (kh, kw)).astype(np.float32) / np.float32(64))
            ctx.convolve(src, dst, k)
            want = orc.filter2d(img, k)
            desc = f"convolve dyadic {kh}x{kw} on {h}x{w}x{c}"
            ok = np.array_equal(dst.download().reshape(img.shape), want)
        src.close(); dst.close()
    else:             # row bands of a chain with recomputed overlap
        h = int(rng.integers(40, 400)); w = int(rng.integers(8, 900)); world = int(rng.integers(1, 6))
        img = rng.integers(0, 256, (h, w, 1), dtype=np.uint8)
        want = fused(img)
        use_fused = rng.random() < 0.5
        out = np.zeros_like(img)
        if h // world < 3:
            continue
        bands = [dip.mosaic.MosaicBand(ctx, w, h, world, r, fused=use_fused) for r in range(world)]
        for b in bands:
            b.bufs[0].rows(b.band.own0, b.band.rows).upload(img[b.band.row0: b.band.row0 + b.band.rows])
        ctx.sync()
        for b in bands:
            b.connect_local(bands)
        for b in bands:
            b.step(overlap=bool(rng.integers(0, 2)))
        out = np.concatenate([b.result().download() for b in bands], axis=0)
        ok = np.array_equal(out, want)
        desc = f"mosaic {h}x{w} world={world} fused={use_fused}"
        for b in bands:
            b.close()
    if not ok:
        fails += 1
        print("MISMATCH", t, desc, flush=True)
print(f"{trials} trials, {fails} mismatches, {time.time() - t0:.1f} s")
sys.exit(1 if fails else 0)
