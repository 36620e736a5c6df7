#!/usr/bin/env python
"""bench.py -- the dip-benchmark hot path on B200 (contract: see DESIGN.md "Measurement").

A "step" is one pass of the reference's full operation list (the 12 device ops of
/root/reference/benchmarks/opencv/benchmark.py:94-105, each input -> output) over a batch of
synthetic 3504x2336 RGB u8 frames (BASELINE.json configs[3]: 256 frames sharded over the ranks;
every frame has the configs[1]/[2] shape the metric is quoted on).

  python bench.py [--gpus N --steps K --warmup W]        # our arm (one process per GPU under torchrun)
  python bench.py --impl reference [...]                 # the reference's OpenCV CPU path, host cores
  python bench.py --per-op                               # per-op table (single image + batch) -> gpurun_out/

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, C = 3504, 2336, 3
PIXELS = W * H
FRAME_BYTES = PIXELS * C            # N = 24 556 032
ALG_BYTES_PER_FRAME_OP = 2 * FRAME_BYTES  # read every input byte once + write every output byte once
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per frame, from the committed
# `ncu --set full` capture (profiles/r1_ncu_stencil_full.md: 6.287 + 6.233 GB for a 256-frame launch)
NCU_TRAFFIC_PER_FRAME = (6.287e9 + 6.233e9) / 256
METRIC = "mpixel_per_s_full_op_chain_3504x2336_rgb_u8_batch"
UNIT = "Mpixel/s"


def measured_peak() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  Started before the warm-up so that
    samples exist even when the timed region is a few milliseconds; `mark()` opens the timed window and the
    summary uses the samples inside it (falling back to the nearest ones taken under the same load)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.t0 = None
        self.t1 = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self):
        if self.t0 is None:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def __exit__(self, *exc):
        if self.t1 is None:
            self.t1 = time.time()
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        rows = [(t, r) for t, r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        inside = [r for t, r in rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        window = "timed region"
        if not inside:  # timed region shorter than the sampling period: the samples taken under the warm-up load
            inside = [r for t, r in rows if self.t0 is None or t >= self.t0 - 1.0][-6:] or [r for _, r in rows][-6:]
            window = "nearest samples (timed region shorter than the sampling period)"
        sm = [float(r[0]) for r in inside]
        mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's OpenCV path (oracle/cv2_ref.py) or the C port (oracle/oracle.py)
# ------------------------------------------------------------------------------------------------
def cpu_chain_runner():
    """Returns (kind, fn(frame) -> None running the 12 ops, threads used)."""
    from oracle import cv2_ref, oracle as orc
    cores = os.cpu_count() or 1
    if cv2_ref.available():
        ops = [fn for _, fn in cv2_ref.OPERATIONS]
        kind = "reference"
    else:
        orc.build()
        orc.set_threads(1)
        ops = [fn for _, fn in orc.OPERATIONS]
        kind = "port"

    def chain(frame):
        for fn in ops:
            fn(frame)
    return kind, chain, cores


def cpu_sample(frames_total: int) -> dict:
    """The full op chain over `frames_total` frames, one frame per worker thread at a time, all host cores."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.build()
    kind, chain, cores = cpu_chain_runner()
    distinct = [orc.synth(W, H, C, seed=0, frame=f) for f in range(min(frames_total, 8))]
    chain(distinct[0])  # warm-up
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(lambda i: chain(distinct[i % len(distinct)]), range(frames_total)))
    dt = time.perf_counter() - t0
    mpix = frames_total * 12 * PIXELS / dt / 1e6
    return {"value": mpix, "unit": UNIT, "cores": cores, "kind": kind, "seconds": dt,
            "sample": f"{frames_total} frames x 12 ops of the 3504x2336 RGB chain, one frame per thread over {cores} threads"}


def workload_name(batch: int, world: int) -> str:
    return (f"batch of {batch} synthetic 3504x2336 RGB u8 frames, full 12-op chain (copy, inversion, grayscale, threshold, "
            f"3 erosions, 4 Gaussian convolutions, GaussianBlur), sharded over {world} GPU(s)")


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = args.cpu_frames
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_sample(frames)
        if i >= args.warmup:
            vals.append(info)
    total_px = sum(frames * 12 * PIXELS for _ in vals)
    total_s = sum(v["seconds"] for v in vals)
    value = total_px / total_s / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / len(vals), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args.batch, int(os.environ.get("WORLD_SIZE", str(args.gpus)))),
                   "frames_total": args.batch, "ops": 12,
                   "sample": f"each step = {frames} frames of that batch through the reference's OpenCV CPU calls (opencv/benchmark.py:94-105) on all host cores",
                   "frames_per_step": frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL prints its version banner to stdout when the communicator is created (these boxes export
        # NCCL_DEBUG=VERSION); stdout must carry the JSON line only, so file descriptor 1 points at stderr
        # until the first collective has run
        import torch
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        return world, rank, local, dist, torch
    return 1, 0, 0, None, None


def barrier_sync(ctx, dist, torch):
    ctx.sync()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    ctx.sync()


def max_over_ranks(x: float, dist, torch) -> float:
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_ours(args) -> None:
    world, rank, local, dist, torch = dist_setup(args.gpus)
    from _dip_pkg import load
    dip = load()
    ctx = dip.Context(local)
    first, count = dip.shard.batch_shard(args.batch, world, rank)
    ops = ctx.reference_ops()
    n_ops = len(ops)
    peak, peak_src = measured_peak()

    src = ctx.image(W, H, C, frames=count).synth_fill(seed=0, first_frame=first)
    dst = ctx.image(W, H, C, frames=count)
    ctx.sync()

    def step(marks_base=None):
        if marks_base is not None:
            ctx.mark(marks_base)
        for i, (_, _, op) in enumerate(ops):
            ctx.apply(op, src, dst)
            if marks_base is not None:
                ctx.mark(marks_base + 1 + i)

    stride = n_ops + 1
    with ClockSampler(local) as clk:
        for _ in range(args.warmup):
            step()
        barrier_sync(ctx, dist, torch)
        launches0 = ctx.launch_count()
        clk.mark()
        ctx.timer_start()
        for k in range(args.steps):
            step(k * stride)
        ms_local = ctx.timer_stop()
        clk.mark()
        barrier_sync(ctx, dist, torch)
    launches = ctx.launch_count() - launches0
    ms = max_over_ranks(ms_local, dist, torch)
    value = args.batch * n_ops * PIXELS / (ms / 1e3 / args.steps) / 1e6

    # per-kernel durations inside the timed region (rank 0's shard)
    per_op_ms = []
    for i in range(n_ops):
        per_op_ms.append(statistics.mean(ctx.elapsed(k * stride + i, k * stride + i + 1) for k in range(args.steps)))
    dom = max(range(n_ops), key=lambda i: per_op_ms[i])
    alg_bytes = ALG_BYTES_PER_FRAME_OP * count
    achieved = alg_bytes / (per_op_ms[dom] / 1e3) / 1e9
    per_op = {ops[i][1]: {"ms": per_op_ms[i], "gbs": alg_bytes / (per_op_ms[i] / 1e3) / 1e9,
                          "frac": alg_bytes / (per_op_ms[i] / 1e3) / 1e9 / peak} for i in range(n_ops)}

    # checksum of checksums of the last op's output: every rank produced real results
    cs = dst.checksum()
    cs_sum = sum(cs) & 0xFFFFFFFFFFFF

    # ---- end to end through the C-ABI with host buffers (pinned rings), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        chunk = args.chunk
        in_ring = chunk * max(1, min(8, (count + chunk - 1) // chunk))
        out_ring = chunk * 2
        host_in = ctx.pinned_empty((in_ring, H, W, C))
        for f in range(in_ring):  # the host copies of the synthetic frames come from the device generator
            src.view(f % count, 1).download(host_in[f])
        host_out = [ctx.pinned_empty((out_ring, H, W, C)) for _ in range(n_ops)]
        pipe = ctx.pipeline(W, H, C, chunk, [op for _, _, op in ops])
        pipe.run(host_in, min(count, 2 * chunk), host_out)  # warm-up
        barrier_sync(ctx, dist, torch)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.run(host_in, count, host_out)   # returns after the last download has landed
        dt_local = time.perf_counter() - t0
        barrier_sync(ctx, dist, torch)
        dt = max_over_ranks(dt_local, dist, torch)
        e2e = {"value": args.batch * n_ops * PIXELS / (dt / e2e_steps) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": count * FRAME_BYTES, "d2h_bytes_per_step": count * n_ops * FRAME_BYTES,
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps, "chunk_frames": chunk,
               "pcie_d2h_gbs_per_gpu": count * n_ops * FRAME_BYTES / (dt / e2e_steps) / 1e9,
               "bound": "PCIe device-to-host (12 result frames leave for every frame that enters); measured pinned D2H peak on this pool is ~56-58 GB/s",
               "note": "per rank: upload each frame once, 12 ops, download all 12 results; pinned host rings; host wall clock around the C-ABI call"}
        pipe.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_sample(args.cpu_frames)
        cpu.pop("seconds", None)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args.batch, world),
                       "frames_total": args.batch, "frames_per_gpu": count, "ops": n_ops, "launches_per_step": n_ops,
                       "l2": f"inputs larger than L2 ({count * FRAME_BYTES / 2**20:.0f} MiB in + same out per rank, streamed once per op)"},
            "clocks": clk.summary(),
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": ops[dom][1], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_TRAFFIC_PER_FRAME * count, "peak_source": peak_src,
                         "traffic_source": "ncu --set full capture in profiles/ (per launch, scaled to this rank's frames)",
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": per_op_ms[dom]},
            "per_op": per_op,
            "cpu_baseline": cpu,
            "checksum": f"{cs_sum:012x}",
        }
        print(json.dumps(line))
    src.close(); dst.close(); ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def run_mosaic(args) -> None:
    """BASELINE.json configs[4]: 65536 x 65536 u8 mosaic, Gaussian 5x5 + erosion 3x3, row bands + halo exchange."""
    world, rank, local, dist, torch = dist_setup(args.gpus)
    from _dip_pkg import load
    dip = load()
    ctx = dip.Context(local)
    side = args.mosaic_side
    band = dip.mosaic.MosaicBand(ctx, side, side, world, rank, fused=args.fused)
    band.fill_synthetic(seed=0)
    ctx.sync()
    if dist is not None:
        handles = [None] * world
        dist.all_gather_object(handles, band.export_handle())
        band.connect(handles)
    barrier_sync(ctx, dist, torch)
    with ClockSampler(local) as clk:
        t_warm = time.time()
        n_warm = 0
        while n_warm < args.warmup or time.time() - t_warm < 0.7:   # a step is < 3 ms: keep the GPU under load while clocks are sampled
            band.step(overlap=not args.no_overlap)
            n_warm += 1
            if n_warm % 16 == 0:
                ctx.sync()
        barrier_sync(ctx, dist, torch)
        launches0 = ctx.launch_count()
        clk.mark()
        ctx.timer_start()
        for _ in range(args.steps):
            band.step(overlap=not args.no_overlap)
        ms_local = ctx.timer_stop()
        clk.mark()
        barrier_sync(ctx, dist, torch)
    launches = ctx.launch_count() - launches0
    ms = max_over_ranks(ms_local, dist, torch) / args.steps
    # exchange alone (halo stream only), to quote the NVLink side of the roofline
    ex_ms = None
    if world > 1:
        barrier_sync(ctx, dist, torch)
        t0 = time.perf_counter()
        for _ in range(20):
            band.exchange()
        ctx.sync()
        ex_ms = max_over_ranks((time.perf_counter() - t0) / 20 * 1e3, dist, torch)
    cs = band.checksum()
    if dist is not None:
        parts = [None] * world
        dist.all_gather_object(parts, cs)
        cs_total = sum(parts) & 0xFFFFFFFFFFFFFFFF
    else:
        cs_total = cs
    # (after the checksum: the copy lands in a result buffer) the same peer path with a LARGE message (a quarter of the neighbour's band into scratch rows of the second
    # buffer): shows what the NVLink route sustains when it is not latency-bound
    bulk_gbs = None
    if world > 1:
        nb = band.below if band.below is not None else band.above
        rows_bulk = max(1, min(band.band.rows // 4, 4096))
        barrier_sync(ctx, dist, torch)
        for it in range(3):
            if it == 1:
                ctx.sync()
                t0 = time.perf_counter()
            band.bufs[1].copy_rows_from(nb[0], nb[1]["own0"], band.band.own0, rows_bulk)
        ctx.sync()
        bulk_s = (time.perf_counter() - t0) / 2
        bulk_gbs = max_over_ranks(-rows_bulk * side / bulk_s / 1e9, dist, torch) * -1.0   # min over ranks
    peak, peak_src = measured_peak()
    n_m = side * side
    if rank == 0:
        alg = (2 if args.fused else 4) * n_m  # fused: one pass; otherwise two materialised passes (read + write each)
        halo_bytes = 3 * side * 2 * (world - 1)  # all interior boundaries, both directions
        line = {
            "metric": "mpixel_per_s_mosaic_gauss5x5_erode3x3", "value": n_m / (ms / 1e3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"synthetic {side}x{side} u8 mosaic, filter2D 5x5 Gaussian then 3x3 erosion ({'one fused kernel' if args.fused else 'two kernels'}), {world} row band(s), halo 3 rows pulled from neighbours over NVLink (CUDA IPC peer copies), overlap recomputed",
                       "rows_per_gpu": band.band.rows, "overlap_exchange_with_interior": not args.no_overlap},
            "clocks": clk.summary(), "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "gauss5x5 + erode3x3 over the band", "achieved": alg / world / (ms / 1e3) / 1e9 * 1.0,
                         "peak": peak, "unit": "GB/s", "frac": alg / world / (ms / 1e3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "note": "per GPU: algorithmic bytes of the band (2 x bytes per pass) / step time"},
            "halo": None if world == 1 else {"bytes_per_step_total": halo_bytes, "bytes_per_gpu_per_direction": 3 * side,
                                             "exchange_ms": ex_ms, "achieved_gbs_per_gpu": 2 * 3 * side / (ex_ms / 1e3) / 1e9,
                                             "nvlink_peak_gbs": 770.0, "note": "196 608-byte messages: latency-bound, not bandwidth-bound",
                                             "bulk_peer_copy_gbs": bulk_gbs, "bulk_note": "same dip_copy_rows path, a quarter band (up to 268 MB) per copy, slowest rank"},
            "checksum": f"{cs_total:016x}",
        }
        print(json.dumps(line))
    band.close(); ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def run_per_op(args) -> None:
    """Per-op table: single 3504x2336 image (ring of buffers > L2) and one-launch batch."""
    from _dip_pkg import load
    dip = load()
    ctx = dip.Context(0)
    peak, peak_src = measured_peak()
    ops = ctx.reference_ops()
    ring = 8  # 8 x (in + out) x 24.5 MB = 393 MB >> 126 MB L2
    srcs = [ctx.image(W, H, C).synth_fill(seed=3, first_frame=i) for i in range(ring)]
    dsts = [ctx.image(W, H, C) for _ in range(ring)]
    bsrc = ctx.image(W, H, C, frames=args.batch).synth_fill(seed=0)
    bdst = ctx.image(W, H, C, frames=args.batch)
    rows = []
    reps = 40
    # singles first (best of 3 x 40 back-to-back calls), stacks afterwards: a 256-frame launch leaves the GPU at its power
    # cap, which would tax the next op's single-image number with lower clocks
    single = {}
    for desc, prefix, op in ops:
        best = None
        for _ in range(3):
            for i in range(ring):
                ctx.apply(op, srcs[i], dsts[i])
            ctx.timer_start()
            for r in range(reps):
                ctx.apply(op, srcs[r % ring], dsts[r % ring])
            t = ctx.timer_stop() / reps
            best = t if best is None else min(best, t)
        single[prefix] = best
    for desc, prefix, op in ops:
        t_single = single[prefix]
        for _ in range(2):
            ctx.apply(op, bsrc, bdst)
        ctx.timer_start()
        for _ in range(5):
            ctx.apply(op, bsrc, bdst)
        t_batch = ctx.timer_stop() / 5 / args.batch
        rows.append({"op": prefix, "single_us": t_single * 1e3, "single_gbs": ALG_BYTES_PER_FRAME_OP / t_single / 1e6,
                     "batch_us_per_frame": t_batch * 1e3, "batch_gbs": ALG_BYTES_PER_FRAME_OP / t_batch / 1e6,
                     "batch_mpix_s": PIXELS / t_batch / 1e3})
    for r in rows:
        r["single_frac"] = r["single_gbs"] / peak
        r["batch_frac"] = r["batch_gbs"] / peak
    # Upload / Download (opencv:92-93): PCIe-bound; pinned = one async copy, pageable = staged slices
    import numpy as np
    xfer = []
    nb = 8
    pin = ctx.pinned_empty((nb, H, W, C)); pin[...] = 7
    page = np.full((nb, H, W, C), 9, np.uint8)
    stack = ctx.image(W, H, C, frames=nb)
    for name, buf in (("pinned", pin), ("pageable", page)):
        for direction in ("upload", "download"):
            fn = (lambda: stack.upload(buf)) if direction == "upload" else (lambda: stack.download(buf, sync=False))
            fn(); ctx.sync()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            ctx.sync()
            dt = (time.perf_counter() - t0) / 3
            xfer.append({"op": f"{direction} ({name} host memory, {nb} frames per call)", "ms_per_frame": dt / nb * 1e3,
                         "gbs": nb * FRAME_BYTES / dt / 1e9})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "per_op.json"), "w") as f:
        json.dump({"peak_gbs": peak, "peak_source": peak_src, "rows": rows, "batch": args.batch, "transfers": xfer}, f, indent=1)
    print(f"| op | single us | GB/s | frac | batch us/frame | GB/s | frac | Mpx/s |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['op']} | {r['single_us']:.2f} | {r['single_gbs']:.0f} | {r['single_frac']:.2f} | {r['batch_us_per_frame']:.2f} | {r['batch_gbs']:.0f} | {r['batch_frac']:.2f} | {r['batch_mpix_s']:.0f} |")
    print()
    print("| transfer | ms per 3504x2336 RGB frame | GB/s |")
    print("|---|---|---|")
    for x in xfer:
        print(f"| {x['op']} | {x['ms_per_frame']:.3f} | {x['gbs']:.1f} |")
    ctx.close()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="total frames of the batch (sharded over the ranks)")
    ap.add_argument("--chunk", type=int, default=4, help="frames per pipeline chunk of the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-frames", type=int, default=48, help="frames of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--per-op", action="store_true")
    ap.add_argument("--workload", default="batch", choices=["batch", "mosaic"])
    ap.add_argument("--mosaic-side", type=int, default=65536)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--fused", action="store_true", help="mosaic: run Gaussian 5x5 + erosion 3x3 as one fused kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.per_op:
        run_per_op(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "mosaic":
        run_mosaic(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
