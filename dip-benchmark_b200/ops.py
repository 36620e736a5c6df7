"""Host-side mirror of the reference's operation layer on top of the C-ABI.

``Context`` plays the role of ``sycl::queue`` (sycl/benchmark.cpp:428) and ``Image`` the role
of ``cv.UMat`` / ``malloc_device`` buffers (opencv/benchmark.py:42-44, sycl:260-262).  Method
names follow the reference's operation list (opencv/benchmark.py:92-105).  Everything runs on
the GPU through libdip_b200.so; nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib

# opencv/benchmark.py:46-72
CROSS_MASK = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=np.uint8)
SQUARE_MASK = np.ones((3, 3), dtype=np.uint8)
SQUARE_MASK_1X3 = np.ones((1, 3), dtype=np.uint8)
SQUARE_MASK_3X1 = np.ones((3, 1), dtype=np.uint8)
BLUR_3X3 = (np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float32) / np.float32(16))
BLUR_3X3_1X3 = np.array([[0.25, 0.5, 0.25]], dtype=np.float32)
BLUR_3X3_3X1 = BLUR_3X3_1X3.T.copy()
_B5 = np.array([1, 4, 6, 4, 1], dtype=np.float32)
BLUR_5X5 = (np.outer(_B5, _B5) / np.float32(256)).astype(np.float32)
BLUR_5X5_1X5 = (_B5 / np.float32(16)).reshape(1, 5)
BLUR_5X5_5X1 = BLUR_5X5_1X5.T.copy()


class Image:
    """A stack of ``frames`` device images (H x W x C interleaved u8)."""

    def __init__(self, ctx: "Context", handle: int, parent: "Image | None" = None):
        self.ctx = ctx
        self._h = C.c_void_p(handle)
        self._parent = parent  # keeps the owning allocation alive
        info = capi.ImageInfo()
        check(lib().dip_image_get_info(self._h, C.byref(info)), ctx._h)
        self.width, self.height, self.channels, self.frames = info.width, info.height, info.channels, info.frames
        self.pitch, self.frame_stride, self.device_ptr = info.pitch, info.frame_stride, info.device_ptr
        bits = C.c_int()
        check(lib().dip_image_depth(self._h, C.byref(bits)), ctx._h)
        self.dtype = np.dtype(np.uint16 if bits.value == 16 else np.uint8)   # 16-bit stacks: dip_image_create16

    @property
    def shape(self):
        base = (self.height, self.width, self.channels)
        return base if self.frames == 1 else (self.frames,) + base

    @property
    def nbytes(self) -> int:
        return self.frames * self.height * self.width * self.channels * self.dtype.itemsize

    def close(self):
        if self._h:
            lib().dip_image_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- transfers (opencv/benchmark.py:92-93) --------------------------------------------
    def upload(self, array: np.ndarray) -> "Image":
        a = np.ascontiguousarray(array, dtype=self.dtype)
        if a.nbytes != self.nbytes:
            raise ValueError(f"array of {a.shape} does not match image stack {self.shape}")
        self._keep = a  # borrowed until the next sync
        check(lib().dip_upload(self.ctx._h, self._h, a.ctypes.data, 0), self.ctx._h)
        return self

    def download(self, out: np.ndarray | None = None, sync: bool = True) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        if out.nbytes != self.nbytes or not out.flags.c_contiguous:
            raise ValueError("bad output array")
        check(lib().dip_download(self.ctx._h, self._h, out.ctypes.data, 0), self.ctx._h)
        if sync:
            self.ctx.sync()
        return out

    def view(self, first: int, count: int) -> "Image":
        h = C.c_void_p()
        check(lib().dip_image_view(self.ctx._h, self._h, first, count, C.byref(h)), self.ctx._h)
        return Image(self.ctx, h.value, parent=self)

    def rows(self, row0: int, rows: int) -> "Image":
        h = C.c_void_p()
        check(lib().dip_image_rows(self.ctx._h, self._h, row0, rows, C.byref(h)), self.ctx._h)
        return Image(self.ctx, h.value, parent=self)

    def checksum(self, first_row: int = 0) -> list[int]:
        out = (C.c_uint64 * self.frames)()
        check(lib().dip_checksum_rows(self.ctx._h, self._h, first_row, out), self.ctx._h)
        return [int(v) for v in out]

    def synth_fill_rows(self, seed: int, first_row: int) -> "Image":
        check(lib().dip_synth_fill_rows(self.ctx._h, self._h, seed & 0xFFFFFFFF, first_row), self.ctx._h)
        return self

    def ipc_export(self) -> bytes:
        h = capi.IpcHandle()
        check(lib().dip_image_ipc_export(self.ctx._h, self._h, C.byref(h)), self.ctx._h)
        return bytes(h.bytes)

    def copy_rows_from(self, src: "Image", src_row: int, dst_row: int, rows: int):
        """Stream-ordered (halo stream) copy of rows from `src` (possibly on a peer GPU) into this image."""
        check(lib().dip_copy_rows(self.ctx._h, src._h, src_row, self._h, dst_row, rows), self.ctx._h)

    def synth_fill(self, seed: int = 0, first_frame: int = 0) -> "Image":
        check(lib().dip_synth_fill(self.ctx._h, self._h, seed & 0xFFFFFFFF, first_frame), self.ctx._h)
        return self


class Graph:
    """A recorded sequence of operations (dip_graph_*): launch() replays it with one CUDA graph launch."""

    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self._h = ctx, C.c_void_p(handle)

    def launch(self):
        check(lib().dip_graph_launch(self.ctx._h, self._h), self.ctx._h)

    def close(self):
        if self._h:
            lib().dip_graph_destroy(self._h)
            self._h = C.c_void_p(None)


class Context:
    """One GPU, one compute stream, two copy streams (include/dip.h lifecycle group)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().dip_init(C.byref(self._h), device)
        if rc < 0:
            self._h = C.c_void_p(None)
            check(rc, None)
        self.device = device
        self._ops = None

    def close(self):
        if self._h:
            lib().dip_shutdown(self._h)
            self._h = C.c_void_p(None)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sync(self):
        check(lib().dip_sync(self._h), self._h)

    def device_info(self) -> dict:
        info = capi.DeviceInfo()
        check(lib().dip_get_device_info(self._h, C.byref(info)), self._h)
        return dict(name=info.name.decode(), device=info.device, sm_count=info.sm_count,
                    cc=(info.cc_major, info.cc_minor), total_mem=info.total_mem, l2_bytes=info.l2_bytes)

    # -- images ------------------------------------------------------------------------------
    def image(self, width: int, height: int, channels: int = 3, frames: int = 1) -> Image:
        h = C.c_void_p()
        check(lib().dip_image_create_batch(self._h, width, height, channels, frames, C.byref(h)), self._h)
        return Image(self, h.value)

    def ipc_import(self, handle: bytes, width: int, height: int, channels: int = 1, frames: int = 1) -> Image:
        h = capi.IpcHandle()
        C.memmove(h.bytes, handle, 64)
        out = C.c_void_p()
        check(lib().dip_image_ipc_import(self._h, C.byref(h), width, height, channels, frames, 0, 0, C.byref(out)), self._h)
        return Image(self, out.value)

    def halo_fence(self):
        check(lib().dip_halo_fence(self._h), self._h)

    def wrap(self, device_ptr: int, width: int, height: int, channels: int, frames: int = 1, pitch: int = 0,
             frame_stride: int = 0, keepalive=None) -> Image:
        """Borrow caller-owned device memory (dip_image_wrap): e.g. a region of interest of a larger image."""
        h = C.c_void_p()
        check(lib().dip_image_wrap(self._h, C.c_void_p(device_ptr), width, height, channels, frames, pitch, frame_stride, C.byref(h)), self._h)
        return Image(self, h.value, parent=keepalive)

    # -- JPEG on the GPU side (cv.imread opencv/benchmark.py:14, cv.imwrite :114) -------------------
    def jpeg_info(self, data: bytes) -> tuple[int, int, int]:
        """(width, height, components) of a JPEG byte stream."""
        w, h, n = C.c_int(), C.c_int(), C.c_int()
        buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
        check(lib().dip_jpeg_info(self._h, buf, len(data), C.byref(w), C.byref(h), C.byref(n)), self._h)
        return w.value, h.value, n.value

    def imdecode(self, streams, dst: Image | None = None, first_frame: int = 0, channels: int = 3) -> Image:
        """cv.imdecode on the GPU: `streams` (one bytes object or a list of same-sized JPEGs) decoded straight into the
        frames of a device stack (allocated when `dst` is None).  Only the compressed bytes cross PCIe."""
        if isinstance(streams, (bytes, bytearray, memoryview)):
            streams = [bytes(streams)]
        streams = [bytes(s) for s in streams]
        if dst is None:
            w, h, _ = self.jpeg_info(streams[0])
            dst = self.image(w, h, channels, len(streams))
        bufs = [(C.c_ubyte * len(s)).from_buffer_copy(s) for s in streams]
        if len(bufs) == 1:
            check(lib().dip_jpeg_decode(self._h, bufs[0], len(streams[0]), dst._h, first_frame), self._h)
        else:
            ptrs = (C.c_void_p * len(bufs))(*[C.addressof(b) for b in bufs])
            lens = (C.c_size_t * len(bufs))(*[len(s) for s in streams])
            check(lib().dip_jpeg_decode_batch(self._h, ptrs, lens, len(bufs), dst._h, first_frame), self._h)
        self.sync()  # the host copies of the streams are released on return
        return dst

    def imencode(self, src: Image, frame: int = 0, quality: int = 95) -> bytes:
        """cv.imencode('.jpg') on the GPU: the bitstream of one frame of a device stack (imwrite's defaults)."""
        cap = src.width * src.height * src.channels + 65536
        out = (C.c_ubyte * cap)()
        n = C.c_size_t()
        check(lib().dip_jpeg_encode(self._h, src._h, frame, quality, out, cap, C.byref(n)), self._h)
        return bytes(memoryview(out)[: n.value])

    def jpeg_backend(self) -> str:
        return lib().dip_jpeg_backend(self._h).decode()

    def image16(self, width: int, height: int, channels: int = 3, frames: int = 1) -> Image:
        """A stack of 16-bit unsigned images (dip_image_create16)."""
        h = C.c_void_p()
        check(lib().dip_image_create16(self._h, width, height, channels, frames, C.byref(h)), self._h)
        return Image(self, h.value)

    def image_like(self, other: Image) -> Image:
        if other.dtype == np.uint16:
            return self.image16(other.width, other.height, other.channels, other.frames)
        return self.image(other.width, other.height, other.channels, other.frames)

    def from_array(self, array: np.ndarray) -> Image:
        """cv.UMat(cpu_image) -- opencv/benchmark.py:42 (array: HxWxC, HxW, or NxHxWxC)."""
        a = np.asarray(array)
        a = a if a.dtype == np.uint16 else a.astype(np.uint8, copy=False)   # uint16 arrays become 16-bit stacks
        if a.ndim == 2:
            a = a[:, :, None]
        if a.ndim == 3:
            a = a[None]
        n, hgt, w, c = a.shape
        img = self.image16(w, hgt, c, n) if a.dtype == np.uint16 else self.image(w, hgt, c, n)
        img.upload(a)
        self.sync()
        return img

    # -- the operation list ------------------------------------------------------------------
    def copy(self, src: Image, dst: Image):
        check(lib().dip_copy(self._h, src._h, dst._h), self._h)

    def invert(self, src: Image, dst: Image):
        check(lib().dip_invert(self._h, src._h, dst._h), self._h)

    def grayscale(self, src: Image, dst: Image):
        check(lib().dip_grayscale(self._h, src._h, dst._h), self._h)

    def threshold(self, src: Image, dst: Image, thresh: int = 127, maxval: int = 255):
        check(lib().dip_threshold(self._h, src._h, dst._h, thresh, maxval), self._h)

    def threshold16(self, src: Image, dst: Image, thresh: int, maxval: int):
        """cv.threshold on 16-bit images with 16-bit arguments (dip_threshold16)."""
        check(lib().dip_threshold16(self._h, src._h, dst._h, thresh, maxval), self._h)

    def erode(self, src: Image, dst: Image, mask: np.ndarray):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        check(lib().dip_erode(self._h, src._h, dst._h, m.ctypes.data, m.shape[1], m.shape[0]), self._h)

    def dilate(self, src: Image, dst: Image, mask: np.ndarray):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        check(lib().dip_dilate(self._h, src._h, dst._h, m.ctypes.data, m.shape[1], m.shape[0]), self._h)

    def erode_separable(self, src: Image, dst: Image, mask_w: int = 3, mask_h: int = 3):
        check(lib().dip_erode_separable(self._h, src._h, dst._h, mask_w, mask_h), self._h)

    def convolve(self, src: Image, dst: Image, kernel: np.ndarray):
        k = np.ascontiguousarray(kernel, dtype=np.float32)
        if k.ndim == 1:
            k = k[None]
        check(lib().dip_convolve(self._h, src._h, dst._h, k.ctypes.data, k.shape[1], k.shape[0]), self._h)

    def convolve_separable(self, src: Image, dst: Image, krow: np.ndarray, kcol: np.ndarray):
        kr = np.ascontiguousarray(krow, dtype=np.float32).ravel()
        kc = np.ascontiguousarray(kcol, dtype=np.float32).ravel()
        check(lib().dip_convolve_separable(self._h, src._h, dst._h, kr.ctypes.data, kr.size, kc.ctypes.data, kc.size), self._h)

    def gaussian_blur3x3(self, src: Image, dst: Image):
        check(lib().dip_gaussian_blur3x3(self._h, src._h, dst._h), self._h)

    def gauss5x5_erode3x3(self, src: Image, dst: Image):
        """Fused chain: 5x5 Gaussian convolution then 3x3 square erosion in one kernel."""
        check(lib().dip_gauss5x5_erode3x3(self._h, src._h, dst._h), self._h)

    def fused_chain_op(self) -> capi.Op:
        op = capi.Op()
        op.kind = 10  # DIP_OP_GAUSS5X5_ERODE3X3
        return op

    # -- the reference's list as data --------------------------------------------------------
    def reference_ops(self):
        """[(description, prefix, Op)] for the 12 device operations, opencv/benchmark.py:94-105."""
        if self._ops is None:
            arr = (capi.Op * 12)()
            n = lib().dip_reference_ops(arr, 12)
            check(n, self._h)
            self._ops_arr = arr
            self._ops = [(lib().dip_reference_op_description(i).decode(), lib().dip_reference_op_prefix(i).decode(), arr[i])
                         for i in range(n)]
        return self._ops

    def apply(self, op: capi.Op, src: Image, dst: Image):
        check(lib().dip_apply(self._h, C.byref(op), src._h, dst._h), self._h)

    def apply_chain(self, ops, src: Image, dst: Image):
        """dst = ops[-1](...ops[0](src)) in as few passes over HBM as the chain allows (dip_apply_chain)."""
        arr = (capi.Op * len(ops))(*ops)
        check(lib().dip_apply_chain(self._h, arr, len(ops), src._h, dst._h), self._h)

    @staticmethod
    def chain_passes(ops, channels: int = 3) -> int:
        """Kernel launches dip_apply_chain makes for this chain (needs no GPU)."""
        arr = (capi.Op * len(ops))(*ops)
        n = C.c_int(0)
        check(lib().dip_chain_passes(arr, len(ops), channels, C.byref(n)), None)
        return n.value

    @staticmethod
    def chain_lut(ops) -> np.ndarray:
        """The per-byte map a run of copy / inversion / threshold ops folds into, as a 256-entry table (needs no GPU)."""
        arr = (capi.Op * len(ops))(*ops)
        lut = (C.c_uint8 * 256)()
        check(lib().dip_chain_lut(arr, len(ops), lut), None)
        return np.frombuffer(lut, dtype=np.uint8).copy()

    def apply_rows(self, op: capi.Op, src: Image, valid_row0: int, valid_rows: int, out_row0: int, out_rows: int,
                   dst: Image, dst_row0: int = 0):
        check(lib().dip_apply_rows(self._h, C.byref(op), src._h, valid_row0, valid_rows, out_row0, out_rows, dst._h,
                                   dst_row0), self._h)

    def apply_rows_peer(self, op: capi.Op, src: Image, valid_row0: int, valid_rows: int, out_row0: int, out_rows: int,
                        dst: Image, dst_row0: int = 0, above=None, below=None):
        """dip_apply_rows_peer: `above` / `below` = (neighbour's mapped band image, first src row it serves, its row, rows)."""
        def peer(t):
            if t is None:
                return None
            img, src_row0, image_row0, rows = t
            return capi.PeerRows(img._h, src_row0, image_row0, rows)
        pa, pb = peer(above), peer(below)
        check(lib().dip_apply_rows_peer(self._h, C.byref(op), src._h, valid_row0, valid_rows, out_row0, out_rows, dst._h, dst_row0,
                                        C.byref(pa) if pa else None, C.byref(pb) if pb else None), self._h)

    # -- step flags between neighbouring bands (include/dip.h "Step flags") -----------------------------
    def flags(self, count: int) -> Image:
        h = C.c_void_p()
        check(lib().dip_flags_create(self._h, count, C.byref(h)), self._h)
        return Image(self, h.value)

    def flag_signal(self, flags: Image, index: int, value: int):
        check(lib().dip_flag_signal(self._h, flags._h, index, value & 0xFFFFFFFF), self._h)

    def flag_wait(self, flags: Image, index: int, value: int):
        check(lib().dip_flag_wait(self._h, flags._h, index, value & 0xFFFFFFFF), self._h)

    def flag_read(self, flags: Image, index: int) -> int:
        v = C.c_uint32(0)
        check(lib().dip_flag_read(self._h, flags._h, index, C.byref(v)), self._h)
        return int(v.value)

    def run_ops_host(self, host_in: np.ndarray, ops, host_out: list[np.ndarray], chunk_frames: int = 4):
        """Upload -> every op -> download, pipelined over three streams (dip_run_ops_host)."""
        a = host_in
        assert a.dtype == np.uint8 and a.flags.c_contiguous and a.ndim == 4
        n, hgt, w, c = a.shape
        arr = (capi.Op * len(ops))(*ops)
        outs = (C.c_void_p * len(ops))(*[o.ctypes.data for o in host_out])
        check(lib().dip_run_ops_host(self._h, a.ctypes.data, w, hgt, c, n, arr, len(ops), outs, chunk_frames), self._h)

    def run_chain_host(self, host_in: np.ndarray, ops, host_out: np.ndarray, chunk_frames: int = 4):
        """Upload -> the ops in sequence (dip_apply_chain) -> download the final result only (dip_run_chain_host)."""
        a = host_in
        assert a.dtype == np.uint8 and a.flags.c_contiguous and a.ndim == 4
        assert host_out.dtype == np.uint8 and host_out.flags.c_contiguous and host_out.shape == a.shape
        n, hgt, w, c = a.shape
        arr = (capi.Op * len(ops))(*ops)
        check(lib().dip_run_chain_host(self._h, a.ctypes.data, w, hgt, c, n, arr, len(ops), host_out.ctypes.data, chunk_frames), self._h)

    def pipeline(self, width: int, height: int, channels: int, chunk_frames: int, ops, chain: bool = False) -> "Pipeline":
        return Pipeline(self, width, height, channels, chunk_frames, ops, chain=chain)

    # -- CUDA-graph replay (include/dip.h "CUDA-graph replay") -------------------------------------------
    def graph_begin(self):
        check(lib().dip_graph_begin(self._h), self._h)

    def graph_end(self) -> "Graph":
        h = C.c_void_p()
        check(lib().dip_graph_end(self._h, C.byref(h)), self._h)
        return Graph(self, h.value)

    # -- timing ---------------------------------------------------------------------------------
    def mark(self, slot: int):
        check(lib().dip_mark(self._h, slot), self._h)

    def elapsed(self, slot_a: int, slot_b: int) -> float:
        ms = C.c_float()
        check(lib().dip_elapsed(self._h, slot_a, slot_b, C.byref(ms)), self._h)
        return float(ms.value)

    def timer_start(self):
        check(lib().dip_timer_start(self._h), self._h)

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(lib().dip_timer_stop(self._h, C.byref(ms)), self._h)
        return float(ms.value)

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(lib().dip_launch_count(self._h, C.byref(n)), self._h)
        return int(n.value)

    def flush_l2(self):
        check(lib().dip_flush_l2(self._h), self._h)

    # -- pinned host memory -----------------------------------------------------------------------
    def pinned_empty(self, shape, dtype=np.uint8) -> np.ndarray:
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(lib().dip_host_alloc(C.byref(p), nbytes), self._h)
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        _PINNED[arr.ctypes.data] = p.value
        return arr


class Pipeline:
    """Persistent upload -> ops -> download pipeline over host rings (dip_pipeline_*)."""

    def __init__(self, ctx: Context, width: int, height: int, channels: int, chunk_frames: int, ops, chain: bool = False):
        """chain=True: the ops run in sequence and only the final result is downloaded (host_out has one entry)."""
        self.ctx = ctx
        self._keep = list(ops)
        self._ops = (capi.Op * len(ops))(*ops)
        self.n_ops = len(ops)
        self.n_out = 1 if chain else len(ops)
        self.chunk = chunk_frames
        self._h = C.c_void_p()
        create = lib().dip_pipeline_create_chain if chain else lib().dip_pipeline_create
        check(create(ctx._h, width, height, channels, chunk_frames, self._ops, len(ops), C.byref(self._h)), ctx._h)

    def run(self, host_in: np.ndarray, frames: int, host_out: list[np.ndarray]):
        """host_in: (in_ring, H, W, C); host_out[i]: (out_ring, H, W, C); rings are multiples of chunk."""
        assert len(host_out) == self.n_out
        outs = (C.c_void_p * self.n_out)(*[o.ctypes.data for o in host_out])
        check(lib().dip_pipeline_run(self._h, host_in.ctypes.data, host_in.shape[0], frames, outs, host_out[0].shape[0]), self.ctx._h)

    def close(self):
        if self._h:
            lib().dip_pipeline_destroy(self._h)
            self._h = C.c_void_p(None)


_PINNED: dict[int, int] = {}


def free_pinned(arr: np.ndarray) -> None:
    p = _PINNED.pop(arr.ctypes.data, None)
    if p is not None:
        lib().dip_host_free(C.c_void_p(p))


# (prefix, method name, extra args) in the reference's order -- opencv/benchmark.py:94-105
OP_KINDS = {"copy": 0, "invert": 1, "grayscale": 2, "threshold": 3, "erode": 4, "erode_separable": 5, "convolve": 6,
            "convolve_separable": 7, "gaussian_blur3x3": 8, "dilate": 9, "gauss5x5_erode3x3": 10}


def make_op(kind: str, *, thresh: int = 0, maxval: int = 0, mask=None, mask_wh=None, kernel=None, kernel_col=None) -> capi.Op:
    """A dip_op entry for dip_apply / dip_apply_chain (include/dip.h); the arrays stay referenced by the object."""
    op = capi.Op()
    op.kind = OP_KINDS[kind]
    op.thresh, op.maxval = thresh, maxval
    keep = []
    if mask is not None:
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        keep.append(m)
        op.mask, op.mask_h, op.mask_w = m.ctypes.data, m.shape[0], m.shape[1]
    if mask_wh is not None:
        op.mask_w, op.mask_h = mask_wh
    if kernel is not None:
        k = np.ascontiguousarray(kernel, dtype=np.float32)
        keep.append(k)
        op.kernel = k.ctypes.data
        if kernel_col is None:
            k2 = np.atleast_2d(k)
            op.kh, op.kw = k2.shape
        else:
            kc = np.ascontiguousarray(kernel_col, dtype=np.float32)
            keep.append(kc)
            op.kernel_col, op.kw, op.kh = kc.ctypes.data, k.size, kc.size
    op._keep = keep
    return op


REFERENCE_OPS = [
    ("copy", "copy", ()),
    ("inversion", "invert", ()),
    ("grayscale", "grayscale", ()),
    ("threshold", "threshold", (127, 255)),
    ("erosion-cross", "erode", (CROSS_MASK,)),
    ("erosion-square", "erode", (SQUARE_MASK,)),
    ("erosion-square-separated", "erode_separable", (3, 3)),
    ("convolution-gaussian-blur-3x3", "convolve", (BLUR_3X3,)),
    ("convolution-gaussian-blur-3x3-separated", "convolve_separable", (BLUR_3X3_1X3, BLUR_3X3_3X1)),
    ("convolution-gaussian-blur-5x5", "convolve", (BLUR_5X5,)),
    ("convolution-gaussian-blur-5x5-separated", "convolve_separable", (BLUR_5X5_1X5, BLUR_5X5_5X1)),
    ("gaussian-blur-3x3", "gaussian_blur3x3", ()),
]
