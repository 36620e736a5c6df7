"""ctypes binding of include/dip.h (the C-ABI of libdip_b200.so)."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# DIP_B200_LIB: another build of the same library (compile-time variants for A/B measurements, tools/README.md)
_LIB = os.environ.get("DIP_B200_LIB") or os.path.join(_HERE, "libdip_b200.so")
_HEADER = os.path.join(_ROOT, "include", "dip.h")
_lib = None


class DipError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"dip error {code}: {message}")
        self.code = code
        self.message = message


DIP_OK, DIP_ERR_ARG, DIP_ERR_SHAPE, DIP_ERR_UNSUPPORTED, DIP_ERR_CUDA, DIP_ERR_NOMEM, DIP_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6


class ImageInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("frames", C.c_int),
                ("pitch", C.c_size_t), ("frame_stride", C.c_size_t), ("device_ptr", C.c_void_p), ("device", C.c_int)]


class DeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("device", C.c_int), ("sm_count", C.c_int), ("cc_major", C.c_int),
                ("cc_minor", C.c_int), ("total_mem", C.c_size_t), ("l2_bytes", C.c_size_t)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int), ("thresh", C.c_uint8), ("maxval", C.c_uint8), ("mask", C.c_void_p),
                ("mask_w", C.c_int), ("mask_h", C.c_int), ("kernel", C.c_void_p), ("kw", C.c_int), ("kh", C.c_int),
                ("kernel_col", C.c_void_p), ("thresh16", C.c_uint16), ("maxval16", C.c_uint16)]


class IpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


class PeerRows(C.Structure):
    _fields_ = [("image", C.c_void_p), ("src_row0", C.c_int), ("image_row0", C.c_int), ("rows", C.c_int)]


def lib_path() -> str:
    return _LIB


def build_library(force: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (``make -C csrc``)."""
    args = ["make", "-s", "-j4", "-C", os.path.join(_HERE, "csrc")]
    if force:
        subprocess.check_call(args + ["clean"])
    subprocess.check_call(args)
    return _LIB


def declared_symbols() -> list[str]:
    """Every function include/dip.h declares with DIP_API."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"DIP_API\s+[\w\s\*]+?\b(dip_\w+)\s*\(", text)))


def lib() -> C.CDLL:
    """Load libdip_b200.so; there is deliberately no fallback when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise ImportError(f"{_LIB} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the product path has no CPU or PyTorch fallback)")
    L = C.CDLL(_LIB)
    vp, i, sz, u8 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint8
    P = C.POINTER
    sig = {
        "dip_version": ([], C.c_char_p),
        "dip_device_count": ([P(i)], i),
        "dip_init": ([P(vp), i], i),
        "dip_init_on_stream": ([P(vp), i, vp], i),
        "dip_shutdown": ([vp], i),
        "dip_sync": ([vp], i),
        "dip_last_error_string": ([vp], C.c_char_p),
        "dip_get_device_info": ([vp, P(DeviceInfo)], i),
        "dip_image_create": ([vp, i, i, i, P(vp)], i),
        "dip_image_create_batch": ([vp, i, i, i, i, P(vp)], i),
        "dip_image_wrap": ([vp, vp, i, i, i, i, sz, sz, P(vp)], i),
        "dip_image_view": ([vp, vp, i, i, P(vp)], i),
        "dip_image_rows": ([vp, vp, i, i, P(vp)], i),
        "dip_image_destroy": ([vp], i),
        "dip_image_get_info": ([vp, P(ImageInfo)], i),
        "dip_image_create16": ([vp, i, i, i, i, P(vp)], i),
        "dip_image_depth": ([vp, P(i)], i),
        "dip_threshold16": ([vp, vp, vp, C.c_uint16, C.c_uint16], i),
        "dip_host_alloc": ([P(vp), sz], i),
        "dip_host_free": ([vp], i),
        "dip_host_register": ([vp, sz], i),
        "dip_host_unregister": ([vp], i),
        "dip_upload": ([vp, vp, vp, sz], i),
        "dip_download": ([vp, vp, vp, sz], i),
        "dip_copy": ([vp, vp, vp], i),
        "dip_invert": ([vp, vp, vp], i),
        "dip_grayscale": ([vp, vp, vp], i),
        "dip_threshold": ([vp, vp, vp, u8, u8], i),
        "dip_erode": ([vp, vp, vp, vp, i, i], i),
        "dip_dilate": ([vp, vp, vp, vp, i, i], i),
        "dip_erode_separable": ([vp, vp, vp, i, i], i),
        "dip_convolve": ([vp, vp, vp, vp, i, i], i),
        "dip_convolve_separable": ([vp, vp, vp, vp, i, vp, i], i),
        "dip_gaussian_blur3x3": ([vp, vp, vp], i),
        "dip_gauss5x5_erode3x3": ([vp, vp, vp], i),
        "dip_apply": ([vp, P(Op), vp, vp], i),
        "dip_apply_chain": ([vp, P(Op), i, vp, vp], i),
        "dip_chain_passes": ([P(Op), i, i, P(i)], i),
        "dip_chain_lut": ([P(Op), i, P(u8)], i),
        "dip_reference_ops": ([P(Op), i], i),
        "dip_reference_op_prefix": ([i], C.c_char_p),
        "dip_reference_op_description": ([i], C.c_char_p),
        "dip_run_ops_host": ([vp, vp, i, i, i, i, P(Op), i, P(vp), i], i),
        "dip_pipeline_create": ([vp, i, i, i, i, P(Op), i, P(vp)], i),
        "dip_pipeline_create_chain": ([vp, i, i, i, i, P(Op), i, P(vp)], i),
        "dip_run_chain_host": ([vp, vp, i, i, i, i, P(Op), i, vp, i], i),
        "dip_pipeline_run": ([vp, vp, i, i, P(vp), i], i),
        "dip_pipeline_destroy": ([vp], i),
        "dip_graph_begin": ([vp], i),
        "dip_graph_end": ([vp, P(vp)], i),
        "dip_graph_launch": ([vp, vp], i),
        "dip_graph_destroy": ([vp], i),
        "dip_mark": ([vp, i], i),
        "dip_elapsed": ([vp, i, i, P(C.c_float)], i),
        "dip_timer_start": ([vp], i),
        "dip_timer_stop": ([vp, P(C.c_float)], i),
        "dip_launch_count": ([vp, P(C.c_uint64)], i),
        "dip_flush_l2": ([vp], i),
        "dip_synth_fill": ([vp, vp, C.c_uint32, C.c_uint64], i),
        "dip_checksum": ([vp, vp, P(C.c_uint64)], i),
        "dip_checksum_rows": ([vp, vp, C.c_uint64, P(C.c_uint64)], i),
        "dip_synth_fill_rows": ([vp, vp, C.c_uint32, C.c_uint64], i),
        "dip_image_ipc_export": ([vp, vp, P(IpcHandle)], i),
        "dip_image_ipc_import": ([vp, P(IpcHandle), i, i, i, i, sz, sz, P(vp)], i),
        "dip_copy_rows": ([vp, vp, i, vp, i, i], i),
        "dip_halo_fence": ([vp], i),
        "dip_apply_rows": ([vp, P(Op), vp, i, i, i, i, vp, i], i),
        "dip_apply_rows_peer": ([vp, P(Op), vp, i, i, i, i, vp, i, P(PeerRows), P(PeerRows)], i),
        "dip_flags_create": ([vp, i, P(vp)], i),
        "dip_flag_signal": ([vp, vp, i, C.c_uint32], i),
        "dip_flag_wait": ([vp, vp, i, C.c_uint32], i),
        "dip_flag_read": ([vp, vp, i, P(C.c_uint32)], i),
        "dip_jpeg_info": ([vp, vp, sz, P(i), P(i), P(i)], i),
        "dip_jpeg_decode": ([vp, vp, sz, vp, i], i),
        "dip_jpeg_decode_batch": ([vp, P(vp), P(sz), i, vp, i], i),
        "dip_jpeg_encode": ([vp, vp, i, i, vp, sz, P(sz)], i),
        "dip_jpeg_backend": ([vp], C.c_char_p),
        "dip_debug_set": ([C.c_char_p, i], i),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    L._dip_signatures = sig
    _lib = L
    return L


def check(rc: int, ctx=None) -> None:
    if rc >= 0:
        return
    msg = lib().dip_last_error_string(ctx)
    if (not msg) and ctx:
        msg = lib().dip_last_error_string(None)
    raise DipError(rc, (msg or b"").decode("utf-8", "replace"))
