"""dip_benchmark_b200 -- B200-native image-op kernels behind the dip-benchmark operation API.

The product is ``libdip_b200.so`` (hand-written CUDA for sm_100a behind the C-ABI of
``include/dip.h``); this package is the thin Python host side used by the parity tests,
``bench.py`` and the Python mirror of the reference's ``perform_benchmark``
(/root/reference/benchmarks/opencv/benchmark.py:41-114).

There is no CPU fallback: importing works anywhere (so the CPU-only checks can verify the
exported symbols), but creating a :class:`Context` without a CUDA device raises.

The directory is named ``dip-benchmark_b200`` (not an identifier); load it with
``from _dip_pkg import load; dip = load()`` from the repository root.
"""
from .capi import DipError, lib, lib_path, build_library, declared_symbols  # noqa: F401
from .ops import Context, Image, REFERENCE_OPS, make_op  # noqa: F401
from . import shard  # noqa: F401
from . import mosaic  # noqa: F401

__all__ = ["Context", "Image", "DipError", "REFERENCE_OPS", "make_op", "lib", "lib_path", "build_library",
           "declared_symbols", "shard", "mosaic"]
