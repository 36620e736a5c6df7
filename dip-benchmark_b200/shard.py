"""Host-side partitioning for the multi-GPU configurations.

Pure arithmetic, no GPU: which frames of a batch and which row band of a mosaic a rank owns,
and which halo rows it exchanges with its neighbours.  The reference has no multi-GPU path;
these are the BASELINE.json configurations 4 (batch of 256 images sharded over 1/2/4/8 GPUs)
and 5 (65536 x 65536 mosaic in row bands with a halo exchange).
"""
from __future__ import annotations

from dataclasses import dataclass


def batch_shard(total_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [first, first+count) of the batch owned by ``rank``."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(total_frames, world)
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


@dataclass(frozen=True)
class Band:
    rank: int
    world: int
    row0: int            # first mosaic row owned
    rows: int            # rows owned
    halo: int            # halo rows kept either side of the band buffer
    has_above: bool
    has_below: bool

    @property
    def buffer_rows(self) -> int:
        return self.rows + 2 * self.halo

    # rows of the BAND BUFFER (0 = first halo row)
    @property
    def own0(self) -> int:
        return self.halo

    def send_up_rows(self) -> tuple[int, int]:
        """(buffer row, count) of the rows the upper neighbour needs as its lower halo."""
        return self.own0, self.halo

    def send_down_rows(self) -> tuple[int, int]:
        return self.own0 + self.rows - self.halo, self.halo

    def recv_above_row(self) -> int:
        return 0

    def recv_below_row(self) -> int:
        return self.own0 + self.rows


def mosaic_band(total_rows: int, world: int, rank: int, halo: int) -> Band:
    first, count = batch_shard(total_rows, world, rank)
    if count < halo:
        raise ValueError("band thinner than its halo")
    return Band(rank, world, first, count, halo, rank > 0, rank < world - 1)


def chain_halo(radii: list[int]) -> int:
    """Halo rows a chain of stencils needs when the overlap is recomputed (sum of radii)."""
    return sum(radii)


def chain_row_plan(band: Band, radii: list[int]) -> list[dict]:
    """Row ranges of every stage of a stencil chain run on one band buffer.

    Stage k reads buffer k and writes buffer k+1; all buffers have ``band.buffer_rows`` rows and
    share the band's coordinates.  With neighbours, stage k must still produce the rows the
    later stages need outside the owned rows (recomputed overlap); at a true mosaic edge the
    valid range stops at the edge so the op's own border rule applies there.
    """
    plan = []
    need_after = [sum(radii[k + 1:]) for k in range(len(radii))]
    valid0 = 0 if band.has_above else band.own0
    valid1 = band.buffer_rows if band.has_below else band.own0 + band.rows
    for k, _r in enumerate(radii):
        lo = band.own0 - (need_after[k] if band.has_above else 0)
        hi = band.own0 + band.rows + (need_after[k] if band.has_below else 0)
        plan.append(dict(valid_row0=valid0, valid_rows=valid1 - valid0, out_row0=lo, out_rows=hi - lo, dst_row0=lo))
        # the next stage may only treat as valid what this stage produced (or the true edge)
        valid0 = lo if band.has_above else band.own0
        valid1 = hi if band.has_below else band.own0 + band.rows
    return plan
