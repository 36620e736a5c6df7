"""perform_benchmark / measure_time of the reference harness on top of libdip_b200.so.

Mirror of /root/reference/benchmarks/opencv/benchmark.py:26-136 (same CLI, same table rows, same
``{prefix}-{filename}`` outputs) with the cv2 calls of its closures replaced by the C-ABI.  The
timing loop is the reference's (host wall clock around ``rounds`` calls); ``--sync`` adds the SYCL
backend's wait-after-every-call semantics (sycl/benchmark.cpp:335), without it the calls are only
enqueued, like the reference's OpenCV-UMat and CUDA.jl loops, and one sync closes the loop.

Image decode/encode is the harness's business, not the library's: cv2 when importable, else PIL.
"""
from __future__ import annotations

import os
from argparse import ArgumentParser, ArgumentTypeError
from time import perf_counter
from typing import Any, Callable

import numpy as np

from . import ops as _ops


def _imread(path: str) -> np.ndarray:
    try:
        import cv2
        img = cv2.imread(path)
        if img is not None:
            return img
    except ImportError:
        pass
    from PIL import Image as PILImage
    return np.ascontiguousarray(np.asarray(PILImage.open(path).convert("RGB"))[:, :, ::-1])  # BGR like cv.imread


def _imwrite(path: str, img: np.ndarray) -> None:
    try:
        import cv2
        cv2.imwrite(path, img)
        return
    except ImportError:
        pass
    from PIL import Image as PILImage
    PILImage.fromarray(img[:, :, ::-1] if img.ndim == 3 and img.shape[2] == 3 else img.squeeze()).save(path)


def parse_image(string: str) -> tuple[np.ndarray, str]:          # opencv:10-14
    if not os.path.isfile(string):
        raise ArgumentTypeError("Not a valid image file")
    return (_imread(string), os.path.basename(string))


def parse_dir(string: str) -> str:                                # opencv:17-23
    if os.path.exists(string) and not os.path.isdir(string):
        raise ArgumentTypeError("Not a valid directory")
    os.makedirs(string, exist_ok=True)
    return string


def measure_time(func: Callable[[], Any], rounds: int, sync: Callable[[], Any] | None = None,
                 record: Callable[[Callable[[], Any], int], Any] | None = None, block: int = 0) -> tuple[float, float]:
    """opencv:26-38; ``sync`` (if given) runs after the single call and after the loop.  With ``record`` and ``block`` the
    loop's calls are recorded ``block`` at a time into a CUDA graph (dip_graph_*) and replayed: ``rounds // block`` launches
    plus the remainder as plain calls -- the same ``rounds`` executions of ``func``'s device work, one host call per block."""
    time_start_once = perf_counter()
    func()
    if sync:
        sync()
    time_end_once = perf_counter()

    graph = record(func, block) if (record and block > 1 and rounds >= block) else None   # recorded outside the timed loop, like a warm-up
    _round = 0
    time_start_times = perf_counter()
    if graph is not None:
        while _round + block <= rounds:
            graph.launch()
            _round += block
    while _round < rounds:
        func()
        _round += 1
    if sync:
        sync()
    time_end_times = perf_counter()
    if graph is not None:
        graph.close()
    return (time_end_once - time_start_once, (time_end_times - time_start_times) / rounds)


def perform_benchmark(cpu_image: np.ndarray, filename: str, dir: str, rounds: int, device: int = 0,
                      sync_every_call: bool = False, out=print, graph_block: int = 0) -> dict[str, np.ndarray]:
    """opencv:41-114.  Returns {prefix: output array} in addition to writing the files.  ``graph_block`` > 1 replays the
    device operations' loops from CUDA graphs of that many calls (transfers cannot be recorded and stay plain calls)."""
    ctx = _ops.Context(device)
    image = ctx.from_array(cpu_image)            # cv.UMat(cpu_image)   opencv:42
    sample = ctx.from_array(cpu_image)           #                      opencv:44 (aux is not needed: fused separable ops)
    host = ctx.pinned_empty(cpu_image.shape if cpu_image.ndim == 3 else cpu_image.shape + (1,))
    host[...] = cpu_image.reshape(host.shape)
    result = ctx.pinned_empty(host.shape)

    def call(fn, *a):
        def f():
            fn(*a)
            if sync_every_call:
                ctx.sync()
        return f

    operations: list[tuple[str, str, Callable[[], Any]]] = []
    operations.append(("Upload", "", call(lambda: image.upload(host))))                     # opencv:92
    operations.append(("Download", "", call(lambda: image.download(result, sync=False))))  # opencv:93
    descriptions = {p: d for d, p, _ in ctx.reference_ops()}
    for prefix, name, args in _ops.REFERENCE_OPS:                                           # opencv:94-105
        operations.append((descriptions[prefix], prefix, call(getattr(ctx, name), image, sample, *args)))

    biggest_description_length = max(len(desc) for desc, _, _ in operations)
    outputs: dict[str, np.ndarray] = {}
    def record(func, block):
        try:
            ctx.graph_begin()
            for _ in range(block):
                func()
            return ctx.graph_end()
        except _ops.capi.DipError:      # not recordable (a transfer, or a sync after every call): plain calls
            try:
                ctx.graph_end()
            except _ops.capi.DipError:
                pass
            return None

    for description, prefix, func in operations:                                            # opencv:109-114
        time_once, time_rounds = measure_time(func, rounds, ctx.sync, record if (graph_block > 1 and prefix and not sync_every_call) else None, graph_block)
        out(f"| {description: <{biggest_description_length}} | {time_once:10.6f}s (once) | {time_rounds:10.6f}s ({rounds} times) |")
        output_image = sample.download().reshape(cpu_image.shape)
        outputs[prefix] = output_image
        if prefix:  # the SYCL/VisionGL guard (sycl:370); the OpenCV harness also writes "-<file>"
            _imwrite(os.path.join(dir, f"{prefix}-{filename}"), output_image)
    _ops.free_pinned(host)
    _ops.free_pinned(result)
    image.close(); sample.close(); ctx.close()
    return outputs


def main(argv=None):
    parser = ArgumentParser(prog="benchmark.py", description="Image processing algorithms benchmark on B200 (libdip_b200)")
    parser.add_argument("infile", type=parse_image, help="Path to image file")
    parser.add_argument("outdir", type=parse_dir, help="Path to image output directory")
    parser.add_argument("--rounds", type=int, default=10000, help="Times to be executed, default 10000")
    parser.add_argument("--device", type=int, default=0)
    parser.add_argument("--sync", action="store_true", help="wait after every call (SYCL backend semantics)")
    parser.add_argument("--graph", type=int, default=0, metavar="N", help="replay the device operations' loops from CUDA graphs of N calls")
    args = parser.parse_args(argv)
    perform_benchmark(args.infile[0], args.infile[1], args.outdir, args.rounds, args.device, args.sync, graph_block=args.graph)


if __name__ == "__main__":
    main()
