"""Row-band driver for very large single-channel mosaics (BASELINE.json config 5).

One rank (process, GPU) owns rows [row0, row0+rows) of the mosaic in a band buffer with `halo` extra rows above and
below.  A step of a stencil chain (default: Gaussian 5x5 then erosion 3x3) needs sum(radii) rows of each
neighbour's band (the overlap is recomputed locally, so one exchange per step suffices).  Two ways to get them:

  halo="kernel" (default for the built-in chains): the first stage's kernel reads the neighbours' rows ITSELF
      (dip_apply_rows_peer): the CTAs at the band's rim load those rows from the neighbour's band buffer -- mapped into
      this process through a CUDA IPC handle, i.e. over NVLink -- into their shared-memory tile.  No exchange step, no
      copy into the local halo rows.
  halo="copy": every rank first PULLS the rows with peer copies on the halo stream (dip_copy_rows) into its own
      buffer's halo rows, runs the interior meanwhile and the rim strips after a fence.  The A/B for the above, and the
      path for chains whose first stage is not a tiled stencil kernel.

iterate=True makes each step's result the next step's input (two exported band buffers used in turn) under the step-flag
protocol of include/dip.h ("Step flags"): before step k a rank's stream waits until both neighbours have signalled k-1,
after it the rank signals k into the neighbours' flag words.  With iterate=False (the benchmark's step: same input
every time) the input never changes after the initial fill + barrier, so no flags are needed.

True mosaic edges (rank 0 top, last rank bottom) use the operation's own border rule.
The reference has no multi-GPU path; this is new work defined by BASELINE.json only.
"""
from __future__ import annotations

from . import shard


class MosaicBand:
    def __init__(self, ctx, width: int, total_rows: int, world: int, rank: int, radii=(2, 1), fused: bool = False,
                 chain=None, channels: int = 1, halo: str | None = None, iterate: bool = False):
        """fused=True runs the default chain as ONE kernel (dip_gauss5x5_erode3x3, halo 3): half the HBM traffic.
        chain=[(dip_op, vertical radius), ...] runs any sequence of operations instead (radius 0 for point ops, 1 for the
        3-tap stencils, 2 for the 5-tap ones)."""
        self.ctx = ctx
        self.width, self.total_rows, self.world, self.rank = width, total_rows, world, rank
        self.fused = fused
        self.radii = [r for _, r in chain] if chain is not None else [3] if fused else list(radii)
        self.halo = shard.chain_halo(self.radii)
        self.band = shard.mosaic_band(total_rows, world, rank, self.halo)
        b = self.band
        self.channels = channels
        self.halo_mode = halo or ("copy" if chain is not None else "kernel")
        if self.halo_mode not in ("kernel", "copy"):
            raise ValueError("halo must be 'kernel' or 'copy'")
        self.iterate = iterate
        self.bufs = [ctx.image(width, b.buffer_rows, channels) for _ in range(len(self.radii) + 1)]
        self.plan = shard.chain_row_plan(b, self.radii)
        self.above = None   # (neighbour's mapped band buffers [input, last], its geometry)
        self.below = None
        self._imported = []  # IPC-mapped images this band opened (closed with it)
        self.flags = ctx.flags(2) if iterate else None   # word 0: signalled by the band above, word 1: by the band below
        self.step_no = 0
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
        h = {"handle": self.bufs[0].ipc_export(), "buffer_rows": b.buffer_rows, "rows": b.rows, "own0": b.own0}
        if self.iterate:   # the result buffer is the next step's input, and the neighbours signal into the flag words
            h["handle_last"] = self.bufs[-1].ipc_export()
            h["flags"] = self.flags.ipc_export()
        return h

    def _import(self, h: dict) -> tuple:
        first = self.ctx.ipc_import(h["handle"], self.width, h["buffer_rows"], self.channels)
        self._imported.append(first)
        last = flags = None
        if self.iterate:
            last = self.ctx.ipc_import(h["handle_last"], self.width, h["buffer_rows"], self.channels)
            flags = self.ctx.ipc_import(h["flags"], 8, 1, 1)
            self._imported += [last, flags]
        return ([first, last], h, flags)

    def connect(self, handles: list[dict]):
        """handles[r] = export_handle() of rank r (gathered by the caller, e.g. all_gather_object)."""
        if self.band.has_above:
            self.above = self._import(handles[self.rank - 1])
        if self.band.has_below:
            self.below = self._import(handles[self.rank + 1])

    def connect_local(self, bands: list["MosaicBand"]):
        """Single-process variant (several bands on one GPU): neighbours are plain images."""
        def local(n):
            return ([n.bufs[0], n.bufs[-1]], {"own0": n.band.own0, "rows": n.band.rows}, n.flags)
        if self.band.has_above:
            self.above = local(bands[self.rank - 1])
        if self.band.has_below:
            self.below = local(bands[self.rank + 1])

    # -- one step --------------------------------------------------------------------------------
    def _io(self):
        """(index into a neighbour's [first, last] pair, this step's buffers in stage order)."""
        if self.iterate and self.step_no % 2 == 1:   # odd steps read what the previous step wrote
            return 1, self.bufs[::-1]
        return 0, self.bufs

    def exchange(self, which: int = 0, into=None):
        """halo="copy": pull the halo rows from the neighbours (halo stream; returns immediately)."""
        b = self.band
        dst = into if into is not None else self.bufs[0]
        if self.halo == 0:   # a chain of point operations needs no neighbour rows
            return
        if self.above is not None:
            imgs, h, _ = self.above
            dst.copy_rows_from(imgs[which], h["own0"] + h["rows"] - self.halo, b.recv_above_row(), self.halo)
        if self.below is not None:
            imgs, h, _ = self.below
            dst.copy_rows_from(imgs[which], h["own0"], b.recv_below_row(), self.halo)

    def halo_bytes(self) -> int:
        return self.halo * self.width * self.channels * (int(self.band.has_above) + int(self.band.has_below))

    def _peers(self, which: int):
        b = self.band
        above = below = None
        if self.above is not None and self.halo:
            imgs, h, _ = self.above
            above = (imgs[which], b.recv_above_row(), h["own0"] + h["rows"] - self.halo, self.halo)
        if self.below is not None and self.halo:
            imgs, h, _ = self.below
            below = (imgs[which], b.recv_below_row(), h["own0"], self.halo)
        return above, below

    def _wait_neighbours(self):
        """iterate: both neighbours have finished the previous step (their rows of this step's input are written and
        they no longer read the buffer this step overwrites)."""
        if not self.iterate or self.step_no == 0:
            return
        if self.above is not None:
            self.ctx.flag_wait(self.flags, 0, self.step_no)
        if self.below is not None:
            self.ctx.flag_wait(self.flags, 1, self.step_no)

    def _signal_neighbours(self):
        if not self.iterate:
            return
        if self.above is not None:
            self.ctx.flag_signal(self.above[2], 1, self.step_no)   # I am the band below it
        if self.below is not None:
            self.ctx.flag_signal(self.below[2], 0, self.step_no)

    def step(self, overlap: bool = True):
        ctx, b = self.ctx, self.band
        which, bufs = self._io()
        st0 = self.plan[0]
        r0 = self.radii[0]
        lo, hi = st0["out_row0"], st0["out_row0"] + st0["out_rows"]
        has_nb = (b.has_above and self.above is not None) or (b.has_below and self.below is not None)
        v0, vn = st0["valid_row0"], st0["valid_rows"]
        if self.halo_mode == "kernel":
            above, below = self._peers(which)

            def stage0(row0, rows, dst_row0):
                ctx.apply_rows_peer(self.chain[0], bufs[0], v0, vn, row0, rows, bufs[1], dst_row0, above, below)
            # rim strips: the rows that read neighbour rows AND (iterate) the rows of the output buffer the neighbours
            # read in their previous step -- one 60-row tile each, the interior runs before the wait
            rim = 60
            ilo = min(max(lo, b.own0 + rim), hi) if b.has_above else lo
            ihi = max(min(hi, b.own0 + b.rows - rim), ilo) if b.has_below else hi
            if overlap and has_nb and self.iterate and ihi > ilo:
                stage0(ilo, ihi - ilo, ilo)
                self._wait_neighbours()
                if ilo > lo:
                    stage0(lo, ilo - lo, lo)
                if hi > ihi:
                    stage0(ihi, hi - ihi, ihi)
            else:
                self._wait_neighbours()
                stage0(lo, hi - lo, st0["dst_row0"])
        else:
            self._wait_neighbours()
            self.exchange(which, into=bufs[0])
            # rows of stage 0 that read halo rows: within r0 of the exchanged regions
            ilo = max(lo, b.own0 + r0) if b.has_above else lo
            ihi = min(hi, b.own0 + b.rows - r0) if b.has_below else hi
            if overlap and has_nb and ihi > ilo and not self.iterate:
                ctx.apply_rows(self.chain[0], bufs[0], v0, vn, ilo, ihi - ilo, bufs[1], ilo)
                ctx.halo_fence()
                if ilo > lo:
                    ctx.apply_rows(self.chain[0], bufs[0], v0, vn, lo, ilo - lo, bufs[1], lo)
                if hi > ihi:
                    ctx.apply_rows(self.chain[0], bufs[0], v0, vn, ihi, hi - ihi, bufs[1], ihi)
            else:
                ctx.halo_fence()
                ctx.apply_rows(self.chain[0], bufs[0], v0, vn, lo, hi - lo, bufs[1], st0["dst_row0"])
        for k in range(1, len(self.chain)):
            st = self.plan[k]
            ctx.apply_rows(self.chain[k], bufs[k], st["valid_row0"], st["valid_rows"], st["out_row0"], st["out_rows"],
                           bufs[k + 1], st["dst_row0"])
        self.step_no += 1
        self._signal_neighbours()

    def result(self):
        """View of the owned rows of the last step's output."""
        b = self.band
        out = self.bufs[-1] if not (self.iterate and self.step_no % 2 == 0 and self.step_no > 0) else self.bufs[0]
        return out.rows(b.own0, b.rows)

    def checksum(self) -> int:
        return self.result().checksum(first_row=self.band.row0)[0]

    def close(self):
        self.ctx.sync()
        for im in self._imported:
            im.close()
        for im in self.bufs:
            im.close()
        if self.flags is not None:
            self.flags.close()
