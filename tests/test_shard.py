"""CPU: batch sharding and mosaic band planning, including a world_size-2 gloo run of the
band + halo-exchange logic with the oracle standing in for the GPU kernels."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_batch_shard_partitions_exactly(dip):
    sh = dip.shard
    for total in (1, 7, 256, 1000):
        for world in (1, 2, 3, 4, 8):
            got = [sh.batch_shard(total, world, r) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        sh.batch_shard(8, 2, 2)


def test_band_plan_reproduces_whole_image(dip, orc):
    """Row-band chain (Gaussian 5x5 -> erode 3x3, BASELINE.json config 5) == whole-image result."""
    sh = dip.shard
    rng = np.random.default_rng(3)
    mosaic = rng.integers(0, 256, (96, 40, 1), dtype=np.uint8)
    want = orc.erode(orc.filter2d(mosaic, orc.BLUR_5X5), orc.SQUARE_MASK)
    radii = [2, 1]
    halo = sh.chain_halo(radii)
    assert halo == 3
    for world in (1, 2, 3, 4):
        out = np.zeros_like(mosaic)
        for rank in range(world):
            b = sh.mosaic_band(mosaic.shape[0], world, rank, halo)
            buf = np.zeros((b.buffer_rows,) + mosaic.shape[1:], np.uint8)
            lo = b.row0 - (halo if b.has_above else 0)
            hi = b.row0 + b.rows + (halo if b.has_below else 0)
            buf[b.own0 - (b.row0 - lo): b.own0 + b.rows + (hi - b.row0 - b.rows)] = mosaic[lo:hi]
            plan = sh.chain_row_plan(b, radii)
            stage_in = buf
            for k, st in enumerate(plan):
                v0, v1 = st["valid_row0"], st["valid_row0"] + st["valid_rows"]
                fn = (lambda im: orc.filter2d(im, orc.BLUR_5X5)) if k == 0 else (lambda im: orc.erode(im, orc.SQUARE_MASK))
                full = fn(stage_in[v0:v1])  # the op applied to the valid rows only (border rule at v0/v1)
                nxt = np.zeros_like(buf)
                o0 = st["out_row0"]
                nxt[st["dst_row0"]: st["dst_row0"] + st["out_rows"]] = full[o0 - v0: o0 - v0 + st["out_rows"]]
                stage_in = nxt
            out[b.row0: b.row0 + b.rows] = stage_in[b.own0: b.own0 + b.rows]
        assert np.array_equal(out, want), world


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from _dip_pkg import load
    from oracle import oracle as orc
    sh = load().shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        mosaic = rng.integers(0, 256, (64, 24, 1), dtype=np.uint8)  # every rank derives the same mosaic
        radii = [2, 1]
        halo = sh.chain_halo(radii)
        b = sh.mosaic_band(mosaic.shape[0], world, rank, halo)
        buf = np.zeros((b.buffer_rows,) + mosaic.shape[1:], np.uint8)
        buf[b.own0: b.own0 + b.rows] = mosaic[b.row0: b.row0 + b.rows]  # only the owned rows are local
        # halo exchange: send my edge rows, receive the neighbours' (the GPU path does this over NVLink)
        reqs = []
        t = torch.from_numpy(buf)
        if b.has_above:
            r0, n = b.send_up_rows()
            reqs.append(dist.isend(t[r0: r0 + n].clone(), rank - 1))
            up = torch.empty_like(t[:halo])
            reqs.append(dist.irecv(up, rank - 1))
        if b.has_below:
            r0, n = b.send_down_rows()
            reqs.append(dist.isend(t[r0: r0 + n].clone(), rank + 1))
            down = torch.empty_like(t[:halo])
            reqs.append(dist.irecv(down, rank + 1))
        for r in reqs:
            r.wait()
        if b.has_above:
            t[b.recv_above_row(): b.recv_above_row() + halo] = up
        if b.has_below:
            t[b.recv_below_row(): b.recv_below_row() + halo] = down
        stage_in = buf
        for k, st in enumerate(sh.chain_row_plan(b, radii)):
            v0, v1 = st["valid_row0"], st["valid_row0"] + st["valid_rows"]
            full = orc.filter2d(stage_in[v0:v1], orc.BLUR_5X5) if k == 0 else orc.erode(stage_in[v0:v1], orc.SQUARE_MASK)
            nxt = np.zeros_like(buf)
            nxt[st["dst_row0"]: st["dst_row0"] + st["out_rows"]] = full[st["out_row0"] - v0: st["out_row0"] - v0 + st["out_rows"]]
            stage_in = nxt
        mine = stage_in[b.own0: b.own0 + b.rows]
        want = orc.erode(orc.filter2d(mosaic, orc.BLUR_5X5), orc.SQUARE_MASK)[b.row0: b.row0 + b.rows]
        ok = bool(np.array_equal(mine, want))
        # checksum of checksums: all-reduce the per-band checksums, compare with the whole image
        cs = torch.tensor([orc.checksum(mine) & 0x7FFFFFFF], dtype=torch.int64)
        dist.all_reduce(cs)
        q.put((rank, ok, int(cs.item())))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_band_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == res[1][2]


def test_gloo_world2_batch_shards_cover_batch():
    """Two ranks process disjoint frame blocks; the union of their per-frame checksums equals the batch's."""
    from oracle import oracle as orc
    from _dip_pkg import load
    sh = load().shard
    total = 5
    frames = [orc.synth(32, 16, 3, seed=9, frame=f) for f in range(total)]
    want = [orc.checksum(orc.invert(f)) for f in frames]
    got = {}
    for rank in range(2):
        f0, n = sh.batch_shard(total, 2, rank)
        for f in range(f0, f0 + n):
            got[f] = orc.checksum(orc.invert(orc.synth(32, 16, 3, seed=9, frame=f)))
    assert [got[f] for f in range(total)] == want
