"""One rank of tests/test_parity_gpu.py::test_mosaic_bands_in_two_processes_over_cuda_ipc.

usage: python tests/ipc_band_worker.py RANK WORLD DIR WIDTH ROWS HALO_MODE
Every rank opens its own context on GPU 0, owns one row band, publishes its band buffer's CUDA IPC handle as a file in
DIR, maps the neighbours' handles (MosaicBand.connect) and runs one step of the Gaussian 5x5 -> erosion 3x3 chain, as two
kernels and as the fused kernel.  Ranks meet at file barriers on the host (no kernel ever waits for another process).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from _dip_pkg import load  # noqa: E402


def barrier(d: str, name: str, rank: int, world: int, timeout: float = 120.0) -> None:
    open(os.path.join(d, f"{name}.{rank}"), "w").close()
    t0 = time.time()
    while not all(os.path.exists(os.path.join(d, f"{name}.{r}")) for r in range(world)):
        if time.time() - t0 > timeout:
            raise TimeoutError(f"barrier {name}")
        time.sleep(0.01)


def main() -> None:
    rank, world, d, width, rows, halo = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    dip = load()
    ctx = dip.Context(0)
    for fused in (0, 1):
        band = dip.mosaic.MosaicBand(ctx, width, rows, world, rank, fused=bool(fused), halo=halo)
        band.fill_synthetic(seed=6)
        ctx.sync()
        h = band.export_handle()
        h["handle"] = h["handle"].hex()
        tmp = os.path.join(d, f"h{fused}_{rank}.tmp")
        with open(tmp, "w") as f:
            json.dump(h, f)
        os.replace(tmp, os.path.join(d, f"h{fused}_{rank}.json"))
        barrier(d, f"filled{fused}", rank, world)          # every band is filled and exported
        handles = []
        for r in range(world):
            with open(os.path.join(d, f"h{fused}_{r}.json")) as f:
                hr = json.load(f)
            hr["handle"] = bytes.fromhex(hr["handle"])
            handles.append(hr)
        band.connect(handles)
        band.step()
        ctx.sync()
        np.save(os.path.join(d, f"out{fused}_{rank}.npy"), band.result().download())
        barrier(d, f"done{fused}", rank, world)            # nobody unmaps a buffer a neighbour may still be reading
        band.close()
    ctx.close()


if __name__ == "__main__":
    main()
