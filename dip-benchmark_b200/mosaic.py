"""Row-band driver for very large single-channel mosaics (BASELINE.json config 5).

One rank (process, GPU) owns rows [row0, row0+rows) of the mosaic in a band buffer with `halo`
extra rows above and below.  A step of a stencil chain (default: Gaussian 5x5 then erosion 3x3,
SURVEY.md 8d) is
  1. halo exchange: every rank PULLS the rows it needs from its neighbours' band buffers with
     peer copies over NVLink (the buffers are shared between the processes through CUDA IPC
     handles); the copies run on the context's halo stream,
  2. while they are in flight the first stage runs on the interior rows that do not depend on
     them; after a fence the two edge strips follow,
  3. the remaining stages run on the band (overlap rows recomputed locally, so one exchange per
     step suffices).
True mosaic edges (rank 0 top, last rank bottom) use the operation's own border rule.
The reference has no multi-GPU path; this is new work defined by BASELINE.json only.
"""
from __future__ import annotations

from . import shard


class MosaicBand:
    def __init__(self, ctx, width: int, total_rows: int, world: int, rank: int, radii=(2, 1), fused: bool = False,
                 chain=None, channels: int = 1):
        """fused=True runs the default chain as ONE kernel (dip_gauss5x5_erode3x3, halo 3): half the HBM traffic.
        chain=[(dip_op, vertical radius), ...] runs any sequence of operations instead (radius 0 for point ops, 1 for the
        3-tap stencils, 2 for the 5-tap ones): one halo exchange of sum(radii) rows per step, overlap recomputed."""
        self.ctx = ctx
        self.width, self.total_rows, self.world, self.rank = width, total_rows, world, rank
        self.fused = fused
        self.radii = [r for _, r in chain] if chain is not None else [3] if fused else list(radii)
        self.halo = shard.chain_halo(self.radii)
        self.band = shard.mosaic_band(total_rows, world, rank, self.halo)
        b = self.band
        self.channels = channels
        self.bufs = [ctx.image(width, b.buffer_rows, channels) for _ in range(len(self.radii) + 1)]
        self.plan = shard.chain_row_plan(b, self.radii)
        self.above = None   # neighbours' input band buffers
        self.below = None
        self._imported = []  # IPC-mapped images this band opened (closed with it)
        ops = {p: op for _, p, op in ctx.reference_ops()}
        if chain is not None:
            self.chain = [op for op, _ in chain]
        else:
            self.chain = [ctx.fused_chain_op()] if fused else [ops["convolution-gaussian-blur-5x5"], ops["erosion-square"]][: len(self.radii)]

    # -- setup ---------------------------------------------------------------------------------
    def fill_synthetic(self, seed: int = 0):
        b = self.band
        self.bufs[0].rows(b.own0, b.rows).synth_fill_rows(seed, b.row0)

    def export_handle(self) -> dict:
        b = self.band
        return {"handle": self.bufs[0].ipc_export(), "buffer_rows": b.buffer_rows, "rows": b.rows, "own0": b.own0}

    def connect(self, handles: list[dict]):
        """handles[r] = export_handle() of rank r (gathered by the caller, e.g. all_gather_object)."""
        if self.band.has_above:
            h = handles[self.rank - 1]
            self.above = (self.ctx.ipc_import(h["handle"], self.width, h["buffer_rows"], self.channels), h)
            self._imported.append(self.above[0])
        if self.band.has_below:
            h = handles[self.rank + 1]
            self.below = (self.ctx.ipc_import(h["handle"], self.width, h["buffer_rows"], self.channels), h)
            self._imported.append(self.below[0])

    def connect_local(self, bands: list["MosaicBand"]):
        """Single-process variant (several bands on one GPU): neighbours are plain images."""
        if self.band.has_above:
            n = bands[self.rank - 1]
            self.above = (n.bufs[0], {"own0": n.band.own0, "rows": n.band.rows})
        if self.band.has_below:
            n = bands[self.rank + 1]
            self.below = (n.bufs[0], {"own0": n.band.own0, "rows": n.band.rows})

    # -- one step --------------------------------------------------------------------------------
    def exchange(self):
        """Pull halo rows from the neighbours (halo stream; returns immediately)."""
        b = self.band
        if self.halo == 0:   # a chain of point operations needs no neighbour rows
            return
        if self.above is not None:
            img, h = self.above
            self.bufs[0].copy_rows_from(img, h["own0"] + h["rows"] - self.halo, b.recv_above_row(), self.halo)
        if self.below is not None:
            img, h = self.below
            self.bufs[0].copy_rows_from(img, h["own0"], b.recv_below_row(), self.halo)

    def halo_bytes(self) -> int:
        return self.halo * self.width * self.channels * (int(self.band.has_above) + int(self.band.has_below))

    def step(self, overlap: bool = True):
        ctx, b = self.ctx, self.band
        self.exchange()
        st0 = self.plan[0]
        r0 = self.radii[0]
        lo, hi = st0["out_row0"], st0["out_row0"] + st0["out_rows"]
        # rows of stage 0 that read halo rows: within r0 of the exchanged regions
        ilo = max(lo, b.own0 + r0) if b.has_above else lo
        ihi = min(hi, b.own0 + b.rows - r0) if b.has_below else hi
        if overlap and (b.has_above or b.has_below) and ihi > ilo:
            ctx.apply_rows(self.chain[0], self.bufs[0], st0["valid_row0"], st0["valid_rows"], ilo, ihi - ilo, self.bufs[1], ilo)
            ctx.halo_fence()
            if ilo > lo:
                ctx.apply_rows(self.chain[0], self.bufs[0], st0["valid_row0"], st0["valid_rows"], lo, ilo - lo, self.bufs[1], lo)
            if hi > ihi:
                ctx.apply_rows(self.chain[0], self.bufs[0], st0["valid_row0"], st0["valid_rows"], ihi, hi - ihi, self.bufs[1], ihi)
        else:
            ctx.halo_fence()
            ctx.apply_rows(self.chain[0], self.bufs[0], st0["valid_row0"], st0["valid_rows"], lo, hi - lo, self.bufs[1], st0["dst_row0"])
        for k in range(1, len(self.chain)):
            st = self.plan[k]
            ctx.apply_rows(self.chain[k], self.bufs[k], st["valid_row0"], st["valid_rows"], st["out_row0"], st["out_rows"],
                           self.bufs[k + 1], st["dst_row0"])

    def result(self):
        """View of the owned rows of the last stage's output."""
        b = self.band
        return self.bufs[-1].rows(b.own0, b.rows)

    def checksum(self) -> int:
        return self.result().checksum(first_row=self.band.row0)[0]

    def close(self):
        self.ctx.sync()
        for im in self._imported:
            im.close()
        for im in self.bufs:
            im.close()
