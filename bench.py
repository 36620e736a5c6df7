#!/usr/bin/env python
"""bench.py -- the dip-benchmark hot path on B200 (contract: see DESIGN.md "Measurement").

A "step" is `inner_reps` passes of the reference's full operation list (the 12 device ops of
/root/reference/benchmarks/opencv/benchmark.py:94-105, each input -> output) over a batch of synthetic 3504x2336 RGB u8
frames (BASELINE.json configs[3]: 256 frames sharded over the ranks; every frame has the configs[1]/[2] shape the metric
is quoted on).  inner_reps is chosen so that the timed region lasts at least a second at every N.

One default run also measures, and puts into the same JSON line:
  e2e / e2e_chain  the same work through the C-ABI with HOST buffers (every result downloaded / one result per frame),
                   next to a bare pinned-copy measurement of what this box's host fabric gives (`host_fabric`);
  mosaic           BASELINE.json configs[4]: 65536 x 65536 u8, Gaussian 5x5 -> erosion 3x3 on row bands, neighbour rows
                   read inside the kernel over NVLink (and the copy / two-kernel / iterated variants), checksum checked
                   against tests/golden/mosaic_kat.json;
  per_op_single    configs[0..2]: every op on ONE image (the decoded data/fundus.jpg), back-to-back calls over a ring of
                   buffers larger than L2, plain and replayed from a CUDA graph;
  cpu_baseline     the reference's OpenCV CPU path on this box's host cores (N = 1 only).

  python bench.py [--gpus N --steps K --warmup W]        # our arm (one process per GPU under torchrun)
  python bench.py --impl reference [...]                 # the reference's OpenCV CPU path, host cores
  python bench.py --per-op                               # per-op table (single image + batch) -> gpurun_out/

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import faulthandler
import glob
import json
import math
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
METRIC = "mpixel_per_s_full_op_chain_3504x2336_rgb_u8_batch"
UNIT = "Mpixel/s"
NVLINK_PEER_COPY_GBS = 770.0        # B200_PROFILING.md: measured peer copy per direction on this pool


def leg(name: str, limit_s: int = 240) -> None:
    """Progress line on stderr and a watchdog: a leg that does not finish within `limit_s` dumps every thread's stack and
    ends the process (torchrun then ends the other ranks), so a multi-rank protocol bug costs minutes, not the box."""
    faulthandler.cancel_dump_traceback_later()
    if name != "done":
        faulthandler.dump_traceback_later(limit_s, exit=True)
    print(f"[bench rank {os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {name}", file=sys.stderr, flush=True)


def measured_peak() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_frame(op_prefix: str) -> tuple[float | None, str]:
    """dram__bytes_read.sum + dram__bytes_write.sum per frame of the kernel that computes `op_prefix`, from the newest
    committed `ncu --set full` summary (profiles/r*_ncu_full.json, written by tools/summarize_ncu.py)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full.json")))
    for path in reversed(files):
        try:
            with open(path) as f:
                data = json.load(f)
            for k in data["kernels"]:
                if k.get("op") == op_prefix:
                    return (k["dram_read_bytes"] + k["dram_write_bytes"]) / data["frames"], os.path.relpath(path, ROOT)
        except Exception:  # noqa: BLE001
            continue
    return None, "no committed ncu summary names this kernel"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  Started before the warm-up so that
    samples exist even when the timed region is short; `mark()` opens / closes the timed window and the summary uses the
    samples inside it."""

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
        if not inside:
            inside = [r for t, r in rows if self.t0 is None or t >= self.t0 - 1.0][-6:] or [r for _, r in rows][-6:]
            window = "nearest samples (none fell inside the timed region)"
        sm = [float(r[0]) for r in inside]
        mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def workload_name(batch: int) -> str:
    return (f"batch of {batch} synthetic 3504x2336 RGB u8 frames, full 12-op chain (copy, inversion, grayscale, threshold, "
            f"3 erosions, 4 Gaussian convolutions, GaussianBlur), sharded over the GPUs of one box")


def config_dict(batch: int) -> dict:
    """Identical in both arms (the driver compares them)."""
    return {"workload": workload_name(batch), "frames_total": batch, "ops": 12, "frame": "3504x2336x3 u8"}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's OpenCV path (oracle/cv2_ref.py ReferenceLoop) or the C port (oracle/oracle.py)
# ------------------------------------------------------------------------------------------------
def cpu_info() -> dict:
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"cpu_model": model, "os_cpu_count": os.cpu_count(), "affinity_cpus": len(os.sched_getaffinity(0))}


class CpuReference:
    """Times the reference's calls as written: cv.UMat image / aux / sample allocated once per stream, every op writing
    into the pre-allocated sample (opencv:41-105).  Two variants (BASELINE.md 3 / SURVEY.md 8d):
      faithful  -- ONE stream of calls, OpenCV's own internal threading (what `python benchmark.py` does);
      all_cores -- one stream per host core, one frame per stream at a time (OpenCV's filter2D / erode / bitwise_not do
                   not parallelise internally, so this is what uses the box)."""

    def __init__(self):
        from oracle import cv2_ref, oracle as orc
        orc.build()
        self.orc = orc
        self.cv2_ref = cv2_ref
        self.cores = len(os.sched_getaffinity(0)) or (os.cpu_count() or 1)
        self.kind = "reference" if cv2_ref.available() else "port"
        self.frames = [orc.synth(W, H, C, seed=0, frame=f) for f in range(8)]
        self.cv_threads = None
        self._pool = None
        if self.kind == "reference":
            import cv2
            self.cv_threads = cv2.getNumThreads()
            self._tls = threading.local()
        else:
            orc.set_threads(1)
            self.ops = [fn for _, fn in orc.OPERATIONS]

    def _chain(self, i: int):
        frame = self.frames[i % len(self.frames)]
        if self.kind == "reference":
            loop = getattr(self._tls, "loop", None)
            if loop is None:
                loop = self._tls.loop = self.cv2_ref.ReferenceLoop(frame)
            loop.load(frame)      # the frame becomes `image` (opencv:42); part of the per-frame CPU work
            loop.run_all()
        else:
            for fn in self.ops:
                fn(frame)

    def all_cores(self, frames_total: int) -> dict:
        if self._pool is None:   # one pool for the run: every worker keeps its image / aux / sample UMats, like the reference does
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.cores)
            list(self._pool.map(self._chain, range(self.cores)))   # first touch of every worker's UMats, untimed
        t0 = time.perf_counter()
        list(self._pool.map(self._chain, range(frames_total)))
        dt = time.perf_counter() - t0
        return {"value": frames_total * 12 * PIXELS / dt / 1e6, "unit": UNIT, "cores": self.cores, "kind": self.kind, "seconds": dt,
                "sample": f"{frames_total} frames x 12 ops of the 3504x2336 RGB chain, the reference's cv.UMat calls with pre-allocated "
                          f"destinations (opencv:94-105), one frame per thread over {self.cores} threads"}

    def faithful(self, frames_total: int) -> dict:
        t0 = time.perf_counter()
        for i in range(frames_total):
            self._chain(i)
        dt = time.perf_counter() - t0
        return {"value": frames_total * 12 * PIXELS / dt / 1e6, "unit": UNIT, "cores": self.cv_threads or 1, "kind": self.kind, "seconds": dt,
                "sample": f"{frames_total} frames x 12 ops, ONE stream of the reference's calls, cv2.getNumThreads() = {self.cv_threads}"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = CpuReference()
    frames = args.batch   # the whole batch per step: same config as our arm
    vals = []
    for i in range(args.warmup + args.steps):
        info = cpu.all_cores(frames)
        if i >= args.warmup:
            vals.append(info)
    total_s = sum(v["seconds"] for v in vals)
    value = len(vals) * frames * 12 * PIXELS / total_s / 1e6
    faithful = cpu.faithful(6)
    faithful.pop("seconds")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / len(vals), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "impl": "reference",
        "config": config_dict(args.batch),
        "detail": {"step": f"each step = all {frames} frames of the batch through the reference's OpenCV CPU calls (opencv/benchmark.py:94-105, "
                           "cv.UMat operands, pre-allocated sample / aux) on all host cores",
                   "cv2_version": cpu.cv2_ref.version(), "cv2_num_threads": cpu.cv_threads, **cpu_info()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": vals[-1]["cores"], "kind": vals[-1]["kind"], "sample": vals[-1]["sample"],
                         "all_cores": {"value": value, "unit": UNIT, "cores": vals[-1]["cores"]}, "faithful": faithful},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local: int) -> dict:
    """Pin this rank's threads to the CPUs of its GPU's NUMA node BEFORE any pinned host memory is allocated, so that
    the host rings are first-touched next to the GPU's PCIe root port.  DIP_NO_AFFINITY=1 leaves the process alone."""
    info = {"applied": False}
    if os.environ.get("DIP_NO_AFFINITY") == "1":
        info["reason"] = "DIP_NO_AFFINITY=1"
        return info
    try:
        bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], text=True).strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus     # 00000000:1b:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            info["reason"] = "the platform reports no NUMA node for this GPU"
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(applied=True, cpus=len(cpus))
    except Exception as e:  # noqa: BLE001
        info["reason"] = f"{type(e).__name__}: {e}"
    return info


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and os.environ.get("DIP_BENCH_SHARE_GPUS") == "1":
        # development only: more ranks than GPUs (ranks share devices, gloo for the plumbing) to exercise the multi-rank
        # protocols of an 8-GPU run on a small box; the times of such a run mean nothing
        import torch
        import torch.distributed as dist
        local = local % torch.cuda.device_count()
        torch.cuda.set_device(local)
        dist.init_process_group("gloo")
        dist.barrier()
        return world, rank, local, dist, torch
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


class Ranks:
    """The collective plumbing of a run (torch.distributed over NCCL; nothing here is on the data path)."""

    def __init__(self, ctx, dist, torch):
        self.ctx, self.dist, self.torch = ctx, dist, torch

    def barrier(self):
        self.ctx.sync()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()
        self.ctx.sync()

    def _reduce(self, x: float, op) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cpu" if self.dist.get_backend() == "gloo" else "cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM if self.dist else None)

    def gather(self, obj):
        if self.dist is None:
            return [obj]
        out = [None] * self.dist.get_world_size()
        self.dist.all_gather_object(out, obj)
        return out


def batch_leg(args, ctx, dip, rk: Ranks, world: int, rank: int, local: int, count: int, src, dst, ops, peak: float, peak_src: str) -> dict:
    n_ops = len(ops)

    def chain(marks_base=None):
        if marks_base is not None:
            ctx.mark(marks_base)
        for i, (_, _, op) in enumerate(ops):
            ctx.apply(op, src, dst)
            if marks_base is not None:
                ctx.mark(marks_base + 1 + i)

    # how long is one pass of the chain on this rank's shard?  -> inner_reps so that K steps last >= 1 s on every rank
    chain(); rk.barrier()
    ctx.timer_start()
    chain()
    t_chain = rk.max(ctx.timer_stop())
    inner = max(1, math.ceil(args.min_timed_s * 1e3 / (args.steps * t_chain)))
    stride = n_ops + 1
    with ClockSampler(local) as clk:
        for _ in range(args.warmup * inner):
            chain()
        rk.barrier()
        launches0 = ctx.launch_count()
        clk.mark()
        ctx.timer_start()
        for k in range(args.steps * inner):
            chain(k * stride)
        ms_local = ctx.timer_stop()
        clk.mark()
        rk.barrier()
    launches = ctx.launch_count() - launches0
    ms = rk.max(ms_local)
    passes = args.steps * inner
    value = args.batch * n_ops * PIXELS * passes / (ms / 1e3) / 1e6
    per_op_ms = [statistics.mean(ctx.elapsed(k * stride + i, k * stride + i + 1) for k in range(passes)) for i in range(n_ops)]
    dom = max(range(n_ops), key=lambda i: per_op_ms[i])
    alg_bytes = ALG_BYTES_PER_FRAME_OP * count
    achieved = alg_bytes / (per_op_ms[dom] / 1e3) / 1e9
    per_op = {ops[i][1]: {"ms": per_op_ms[i], "gbs": alg_bytes / (per_op_ms[i] / 1e3) / 1e9,
                          "frac": alg_bytes / (per_op_ms[i] / 1e3) / 1e9 / peak} for i in range(n_ops)}
    traffic_pf, traffic_src = ncu_traffic_per_frame(ops[dom][1])
    cs = dst.checksum()
    return {"value": value, "ms_per_step": ms / args.steps, "inner_reps": inner, "timed_region_s": ms / 1e3, "launches": launches,
            "clocks": clk.summary(), "per_op": per_op, "checksum": f"{sum(cs) & 0xFFFFFFFFFFFF:012x}",
            "roofline": {"bound": "hbm", "kernel": ops[dom][1], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_pf * count if traffic_pf is not None else None, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": per_op_ms[dom],
                         "step_frac": sum(alg_bytes for _ in range(n_ops)) / (sum(per_op_ms) / 1e3) / 1e9 / peak}}


def e2e_legs(args, ctx, dip, rk: Ranks, world: int, count: int, src, ops) -> tuple[dict, dict, dict]:
    """Host buffers in, host buffers out, through dip_pipeline_run; and what the box's host fabric gives bare."""
    n_ops = len(ops)
    chunk = args.chunk
    in_ring = chunk * max(1, min(8, (count + chunk - 1) // chunk))
    out_ring = chunk * 2
    host_in = ctx.pinned_empty((in_ring, H, W, C))
    for f in range(in_ring):  # the host copies of the synthetic frames come from the device generator
        src.view(f % count, 1).download(host_in[f])
    host_out = [ctx.pinned_empty((out_ring, H, W, C)) for _ in range(n_ops)]

    # -- bare copies: every rank at once, pinned memory, no kernels: D2H alone, then both directions at once
    dev = ctx.image(W, H, C, frames=chunk)
    dev2 = ctx.image(W, H, C, frames=chunk)

    def bare(both: bool) -> float:
        reps = 24
        for _ in range(2):
            dev.download(host_out[0][:chunk], sync=False)
        rk.barrier()
        t0 = time.perf_counter()
        for r in range(reps):
            dev.download(host_out[r % n_ops][:chunk], sync=False)
            if both:
                dev2.upload(host_in[:chunk])
        ctx.sync()
        dt = rk.max(time.perf_counter() - t0)
        rk.barrier()
        return reps * chunk * FRAME_BYTES / dt / 1e9   # per rank, per direction
    d2h_only = bare(False)
    d2h_both = bare(True)
    dev.close(); dev2.close()
    fabric = {"d2h_gbs_per_gpu": d2h_only, "d2h_gbs_aggregate": d2h_only * world, "bidirectional_gbs_per_gpu_per_direction": d2h_both,
              "bidirectional_gbs_aggregate": 2 * d2h_both * world,
              "note": f"bare cudaMemcpyAsync of {chunk}-frame pinned buffers on all {world} rank(s) at once, no kernels; slowest rank"}

    # -- every op's result comes back (the reference's loop, taken end to end)
    pipe = ctx.pipeline(W, H, C, chunk, [op for _, _, op in ops])
    pipe.run(host_in, min(count, 2 * chunk), host_out)  # warm-up
    rk.barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.run(host_in, count, host_out)   # returns after the last download has landed
    dt = rk.max(time.perf_counter() - t0)
    rk.barrier()
    d2h_rate = count * n_ops * FRAME_BYTES / (dt / e2e_steps) / 1e9
    e2e = {"value": args.batch * n_ops * PIXELS / (dt / e2e_steps) / 1e6, "unit": UNIT,
           "h2d_bytes_per_step": count * FRAME_BYTES, "d2h_bytes_per_step": count * n_ops * FRAME_BYTES,
           "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps, "chunk_frames": chunk,
           "pcie_d2h_gbs_per_gpu": d2h_rate, "host_fabric_gbs": d2h_only * world, "frac_of_host_fabric": d2h_rate / d2h_only,
           "bound": "host fabric, device-to-host (12 result frames leave for every frame that enters)",
           "note": "per rank: upload each frame once, 12 ops, download all 12 results; pinned host rings; host wall clock around the C-ABI call"}
    pipe.close()

    # -- a pipeline: the ops applied in sequence, one result per frame (dip_pipeline_create_chain)
    by = {p: op for _, p, op in ops}
    chain = [by["grayscale"], by["gaussian-blur-3x3"], by["threshold"], by["erosion-cross"]]
    cpipe = ctx.pipeline(W, H, C, chunk, chain, chain=True)
    cpipe.run(host_in, min(count, 2 * chunk), [host_out[0]])
    rk.barrier()
    t0 = time.perf_counter()
    cpipe.run(host_in, count, [host_out[0]])
    t1 = rk.max(time.perf_counter() - t0)
    reps = max(1, math.ceil(args.min_timed_s / t1))
    rk.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        cpipe.run(host_in, count, [host_out[0]])
    dt = rk.max(time.perf_counter() - t0)
    rk.barrier()
    e2e_chain = {"value": args.batch * len(chain) * PIXELS * reps / dt / 1e6, "unit": UNIT, "ops": len(chain),
                 "chain": "grayscale -> GaussianBlur 3x3 -> threshold -> erosion (cross): 3 passes over HBM (dip_apply_chain)",
                 "frames_per_s": args.batch * reps / dt, "h2d_bytes_per_step": count * FRAME_BYTES, "d2h_bytes_per_step": count * FRAME_BYTES,
                 "pcie_gbs_per_gpu_per_direction": count * FRAME_BYTES * reps / dt / 1e9, "passes_over_batch": reps,
                 "frac_of_host_fabric": count * FRAME_BYTES * reps / dt / 1e9 / d2h_both,
                 "note": "every frame crosses PCIe once in each direction; value counts pixels x ops of the chain"}
    cpipe.close()
    for arr in [host_in] + host_out:
        dip.ops.free_pinned(arr)
    return e2e, e2e_chain, fabric


def mosaic_leg(args, ctx, dip, rk: Ranks, world: int, rank: int, peak: float) -> dict:
    """BASELINE.json configs[4]: 65536 x 65536 u8, Gaussian 5x5 then erosion 3x3, one row band per GPU."""
    side = args.mosaic_side
    try:
        with open(os.path.join(ROOT, "tests", "golden", "mosaic_kat.json")) as f:
            kat = json.load(f)
    except OSError:
        kat = {}
    n_m = side * side

    def run(name: str, fused: bool, halo: str, iterate: bool = False) -> dict:
        leg(f"mosaic: {name[:40]}")
        band = dip.mosaic.MosaicBand(ctx, side, side, world, rank, fused=fused, halo=halo, iterate=iterate)
        band.fill_synthetic(seed=0)
        ctx.sync()
        if world > 1:
            band.connect(rk.gather(band.export_handle()))
        rk.barrier()
        out = {"variant": name}
        if iterate:   # two steps from the fresh fill: the oracle's checksum of the pipeline applied twice
            band.step(); band.step()
            cs2 = sum(rk.gather(band.checksum())) & 0xFFFFFFFFFFFFFFFF
            out["checksum_after_2_steps"] = f"{cs2:016x}"
            out["kat_match"] = (kat.get(f"{side}_iter2") == f"{cs2:016x}") if f"{side}_iter2" in kat else None
        # Every rank runs the SAME number of steps in every phase (the counts come from times that are maxima over the
        # ranks): with iterate=True a rank may be at most one step ahead of its neighbours, so a rank that stopped earlier
        # than its neighbour would leave that neighbour's last step waiting for a flag that never comes.
        def run_steps(n: int):
            for i in range(n):
                band.step()
                if (i + 1) % 64 == 0:
                    ctx.sync()   # bounds the work queued on the stream

        rk.barrier()
        ctx.timer_start()
        run_steps(8)
        est = rk.max(ctx.timer_stop()) / 8
        run_steps(max(8, int(math.ceil(0.3e3 / est / 8.0)) * 8))   # ~0.3 s into the sustained power state
        rk.barrier()
        ctx.timer_start()
        run_steps(8)
        est = rk.max(ctx.timer_stop()) / 8
        steps = max(8, int(math.ceil(args.mosaic_timed_s * 1e3 / est / 8.0)) * 8)
        rk.barrier()
        launches0 = ctx.launch_count()
        ctx.timer_start()
        for _ in range(steps):
            band.step()
        ms = rk.max(ctx.timer_stop()) / steps
        rk.barrier()
        alg = (2 if fused else 4) * n_m
        out.update(ms_per_step=ms, steps=steps, mpixel_per_s=n_m / (ms / 1e3) / 1e6, launches_per_step=(ctx.launch_count() - launches0) / steps,
                   frac_hbm=alg / world / (ms / 1e3) / 1e9 / peak, algorithmic_bytes_per_gpu=alg / world)
        if not iterate:
            cs = sum(rk.gather(band.checksum())) & 0xFFFFFFFFFFFFFFFF
            out["checksum"] = f"{cs:016x}"
            out["kat_match"] = (kat.get(str(side)) == f"{cs:016x}") if str(side) in kat else None
        if halo == "copy" and world > 1:   # the exchange alone (halo stream only): the NVLink side of the roofline
            rk.barrier()
            t0 = time.perf_counter()
            for _ in range(50):
                band.exchange()
            ctx.sync()
            ex_ms = rk.max((time.perf_counter() - t0) / 50 * 1e3)
            per_gpu = 2 * 3 * side
            out["exchange"] = {"exchange_ms": ex_ms, "bytes_per_gpu": per_gpu, "gbs": per_gpu / (ex_ms / 1e3) / 1e9,
                               "frac_nvlink": per_gpu / (ex_ms / 1e3) / 1e9 / NVLINK_PEER_COPY_GBS}
        rk.barrier()
        band.close()
        return out

    res = {"side": side, "rows_per_gpu": side // world,
           "workload": f"synthetic {side}x{side} u8 mosaic, filter2D 5x5 Gaussian then 3x3 erosion, {world} row band(s), 3 halo rows per side"}
    res["fused"] = run("one fused kernel; rim CTAs read the neighbours' rows in the kernel over NVLink", True, "kernel")
    res["two_kernel"] = run("two kernels (Gaussian, erosion); neighbour rows read in the first kernel", False, "kernel")
    res["fused_copy_halo"] = run("one fused kernel; halo rows copied first (dip_copy_rows on the halo stream), interior overlapped", True, "copy")
    res["fused_iterate"] = run("one fused kernel, each result is the next input; steps ordered between neighbours by flags (dip_flag_*)", True, "kernel", iterate=True)
    res["ms_per_step"] = res["fused"]["ms_per_step"]
    res["frac_hbm"] = res["fused"]["frac_hbm"]
    res["checksum"] = res["fused"]["checksum"]
    res["kat_match"] = bool(res["fused"]["kat_match"]) and bool(res["two_kernel"]["kat_match"]) and bool(res["fused_copy_halo"]["kat_match"]) \
        and res["fused_iterate"]["kat_match"] is not False
    halo_bytes = 2 * 3 * side
    if world > 1:
        ex = res["fused_copy_halo"].get("exchange", {})
        # what reading the rows in the kernel costs: the step against the same kernel on a band without neighbours is not
        # measurable in place, so the bound is quoted: 2 x 3 rows over NVLink at the measured peer-copy rate
        res["halo"] = {"bytes_per_gpu_per_step": halo_bytes, "exchange_ms": ex.get("exchange_ms"), "gbs": ex.get("gbs"),
                       "frac_nvlink": ex.get("frac_nvlink"), "nvlink_peer_copy_gbs": NVLINK_PEER_COPY_GBS,
                       "in_kernel_vs_copy_ms": res["fused"]["ms_per_step"] - res["fused_copy_halo"]["ms_per_step"],
                       "iterate_overhead_ms": res["fused_iterate"]["ms_per_step"] - res["fused"]["ms_per_step"],
                       "note": "196 608-byte messages are latency-bound; read inside the kernel they cost no step of their own"}
    else:
        res["halo"] = None
    return res


def load_fundus():
    import numpy as np
    path = os.path.join(ROOT, "data", "fundus.jpg")
    try:
        import cv2
        img = cv2.imread(path)
        if img is not None and img.shape == (H, W, C):
            return np.ascontiguousarray(img), "data/fundus.jpg decoded by cv2"
    except Exception:  # noqa: BLE001
        pass
    try:
        from PIL import Image
        img = np.asarray(Image.open(path).convert("RGB"))[:, :, ::-1]
        if img.shape == (H, W, C):
            return np.ascontiguousarray(img), "data/fundus.jpg decoded by PIL"
    except Exception:  # noqa: BLE001
        pass
    return None, "no JPEG decoder on this box: synthetic frame of the fundus shape"


def per_op_single_leg(args, ctx, dip, ops, peak: float) -> dict:
    """configs[0..2]: one 3504x2336 RGB image per call, the reference's own loop shape (opencv:33-35): back-to-back calls
    over a ring of 8 source / destination pairs (393 MB, larger than L2), plain and replayed from a CUDA graph."""
    img, how = load_fundus()
    ring, reps = 8, 40
    if img is not None:
        srcs = [ctx.from_array(img) for _ in range(ring)]
    else:
        srcs = [ctx.image(W, H, C).synth_fill(seed=3, first_frame=i) for i in range(ring)]
    dsts = [ctx.image(W, H, C) for _ in range(ring)]
    ctx.sync()
    rows = {}
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
        ctx.graph_begin()
        for r in range(reps):
            ctx.apply(op, srcs[r % ring], dsts[r % ring])
        g = ctx.graph_end()
        gbest = None
        for _ in range(3):
            g.launch()
            ctx.timer_start()
            g.launch()
            t = ctx.timer_stop() / reps
            gbest = t if gbest is None else min(gbest, t)
        g.close()
        rows[prefix] = {"us": best * 1e3, "gbs": ALG_BYTES_PER_FRAME_OP / best / 1e6, "frac": ALG_BYTES_PER_FRAME_OP / best / 1e6 / peak,
                        "mpixel_per_s": PIXELS / best / 1e3, "graph_us": gbest * 1e3, "graph_frac": ALG_BYTES_PER_FRAME_OP / gbest / 1e6 / peak}
    for im in srcs + dsts:
        im.close()
    worst = min(rows, key=lambda p: rows[p]["frac"])
    return {"input": how, "calls": f"best of 3 x {reps} back-to-back calls over a ring of {ring} image pairs (larger than L2); graph_* = the same {reps} calls replayed from one CUDA graph",
            "ops": rows, "min_frac": rows[worst]["frac"], "min_frac_op": worst, "min_graph_frac": min(r["graph_frac"] for r in rows.values())}


def run_ours(args) -> None:
    world, rank, local, dist, torch = dist_setup(args.gpus)
    affinity = bind_to_gpu_numa(local) if world > 1 else {"applied": False, "reason": "single rank"}
    from _dip_pkg import load
    dip = load()
    ctx = dip.Context(local)
    rk = Ranks(ctx, dist, torch)
    first, count = dip.shard.batch_shard(args.batch, world, rank)
    ops = ctx.reference_ops()
    n_ops = len(ops)
    peak, peak_src = measured_peak()

    src = ctx.image(W, H, C, frames=count).synth_fill(seed=0, first_frame=first)
    dst = ctx.image(W, H, C, frames=count)
    ctx.sync()
    leg("batch leg")
    b = batch_leg(args, ctx, dip, rk, world, rank, local, count, src, dst, ops, peak, peak_src)

    e2e = e2e_chain = fabric = None
    if not args.no_e2e:
        leg("end-to-end legs")
        e2e, e2e_chain, fabric = e2e_legs(args, ctx, dip, rk, world, count, src, ops)
    src.close(); dst.close()

    mosaic = None
    if not args.no_mosaic:
        leg("mosaic leg")
        mosaic = mosaic_leg(args, ctx, dip, rk, world, rank, peak)

    single = None
    leg("single-image leg")
    if rank == 0 and not args.no_single:
        single = per_op_single_leg(args, ctx, dip, ops, peak)
    rk.barrier()
    leg("cpu baseline", 600)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ref = CpuReference()
        cpu = ref.all_cores(args.cpu_frames)
        cpu.pop("seconds", None)
        cpu["all_cores"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"]}
        cpu["faithful"] = ref.faithful(4)
        cpu["faithful"].pop("seconds", None)
        cpu.update(cv2_num_threads=ref.cv_threads, **cpu_info())

    if rank == 0:
        line = {
            "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": config_dict(args.batch),
            "detail": {"frames_per_gpu": count, "inner_reps": b["inner_reps"], "timed_region_s": b["timed_region_s"],
                       "step": f"{b['inner_reps']} pass(es) of the 12-op chain over the rank's {count} frames (one launch per op per pass)",
                       "launches_per_step": n_ops * b["inner_reps"],
                       "l2": f"inputs larger than L2 ({count * FRAME_BYTES / 2**20:.0f} MiB in + same out per rank, streamed once per op)",
                       "affinity": affinity},
            "clocks": b["clocks"],
            "e2e": e2e, "e2e_chain": e2e_chain, "host_fabric": fabric,
            "gpu_launches": b["launches"],
            "roofline": b["roofline"],
            "per_op": b["per_op"],
            "mosaic": mosaic,
            "per_op_single": single,
            "cpu_baseline": cpu,
            "checksum": b["checksum"],
        }
        print(json.dumps(line))
    leg("done")
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def run_per_op(args) -> None:
    """Per-op table: single 3504x2336 image (ring of buffers > L2) and one-launch batch."""
    from _dip_pkg import load
    dip = load()
    ctx = dip.Context(0)
    peak, peak_src = measured_peak()
    ops = ctx.reference_ops()
    single = per_op_single_leg(args, ctx, dip, ops, peak)
    bsrc = ctx.image(W, H, C, frames=args.batch).synth_fill(seed=0)
    bdst = ctx.image(W, H, C, frames=args.batch)
    rows = []
    for desc, prefix, op in ops:
        for _ in range(2):
            ctx.apply(op, bsrc, bdst)
        ctx.timer_start()
        for _ in range(5):
            ctx.apply(op, bsrc, bdst)
        t_batch = ctx.timer_stop() / 5 / args.batch
        s = single["ops"][prefix]
        rows.append({"op": prefix, "single_us": s["us"], "single_gbs": s["gbs"], "single_frac": s["frac"], "graph_us": s["graph_us"], "graph_frac": s["graph_frac"],
                     "batch_us_per_frame": t_batch * 1e3, "batch_gbs": ALG_BYTES_PER_FRAME_OP / t_batch / 1e6,
                     "batch_frac": ALG_BYTES_PER_FRAME_OP / t_batch / 1e6 / peak, "batch_mpix_s": PIXELS / t_batch / 1e3})
    bsrc.close(); bdst.close()
    # arbitrary masks / kernels: the tiled generic kernel and the two-pass separable path
    import numpy as np
    gsrc = ctx.image(W, H, C, frames=16).synth_fill(seed=0)
    gdst = ctx.image(W, H, C, frames=16)
    rng = np.random.default_rng(0)
    generic = []
    L = dip.lib()
    for name, fn in (("erode, arbitrary 3x3 mask (6 taps)", lambda: ctx.erode(gsrc, gdst, np.array([[1, 1, 0], [0, 1, 1], [1, 0, 1]], np.uint8))),
                     ("filter2D, box 3x3 (non-dyadic)", lambda: ctx.convolve(gsrc, gdst, np.full((3, 3), 1 / 9, np.float32))),
                     ("filter2D, random 5x5", lambda: ctx.convolve(gsrc, gdst, (rng.random((5, 5)) / 13).astype(np.float32))),
                     ("filter2D, random 7x7", lambda: ctx.convolve(gsrc, gdst, (rng.random((7, 7)) / 25).astype(np.float32))),
                     ("separable, random 1x7 + 7x1", lambda: ctx.convolve_separable(gsrc, gdst, (rng.random(7) / 4).astype(np.float32), (rng.random(7) / 4).astype(np.float32)))):
        for knob in (0, 1):
            L.dip_debug_set(b"direct_generic", knob)
            fn()
            ctx.timer_start()
            for _ in range(3):
                fn()
            t = ctx.timer_stop() / 3 / 16
            generic.append({"op": name, "kernel": "direct" if knob else "tiled generic", "us_per_frame": t * 1e3, "frac": ALG_BYTES_PER_FRAME_OP / t / 1e6 / peak})
    L.dip_debug_set(b"direct_generic", 0)
    gsrc.close(); gdst.close()
    # Upload / Download (opencv:92-93): PCIe-bound; pinned = one async copy, pageable = staged slices
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
        json.dump({"peak_gbs": peak, "peak_source": peak_src, "rows": rows, "batch": args.batch, "transfers": xfer, "generic": generic,
                   "single_input": single["input"]}, f, indent=1)
    print(f"input of the single-image columns: {single['input']}")
    print("| op | single us | GB/s | frac | graph replay us | frac | batch us/frame | GB/s | frac | Mpx/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['op']} | {r['single_us']:.2f} | {r['single_gbs']:.0f} | {r['single_frac']:.2f} | {r['graph_us']:.2f} | {r['graph_frac']:.2f} | "
              f"{r['batch_us_per_frame']:.2f} | {r['batch_gbs']:.0f} | {r['batch_frac']:.2f} | {r['batch_mpix_s']:.0f} |")
    print()
    print("| arbitrary mask / kernel (16-frame stack) | kernel | us per frame | frac of HBM peak |")
    print("|---|---|---|---|")
    for r in generic:
        print(f"| {r['op']} | {r['kernel']} | {r['us_per_frame']:.1f} | {r['frac']:.3f} |")
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
    ap.add_argument("--chunk", type=int, default=4, help="frames per pipeline chunk of the end-to-end legs")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-frames", type=int, default=48, help="frames of the bounded CPU sample of our arm's cpu_baseline")
    ap.add_argument("--min-timed-s", type=float, default=1.0, help="minimum length of the batch leg's timed region")
    ap.add_argument("--mosaic-timed-s", type=float, default=0.7, help="minimum length of every mosaic variant's timed region")
    ap.add_argument("--mosaic-side", type=int, default=65536)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mosaic", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--per-op", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.per_op:
        run_per_op(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
