"""Which nvJPEG backends this GPU / driver accepts and what they cost on data/fundus.jpg (ctypes straight onto
libnvjpeg; device memory from the library).  usage: python tools/nvjpeg_probe.py"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _dip_pkg import load
dip = load()
nj = C.CDLL("libnvjpeg.so.12")
ctx = dip.Context(0)
data = open(os.path.join(ROOT, "data", "fundus.jpg"), "rb").read()
w, h, _ = ctx.jpeg_info(data)
B = 8
stack = ctx.image(w, h, 3, frames=B)


class NvImage(C.Structure):
    _fields_ = [("channel", C.c_void_p * 4), ("pitch", C.c_size_t * 4)]


buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
names = {0: "DEFAULT", 1: "HYBRID", 2: "GPU_HYBRID", 3: "HARDWARE", 4: "GPU_HYBRID_DEVICE", 5: "HARDWARE_DEVICE"}
BGRI = 6
for be, name in names.items():
    hd = C.c_void_p()
    st = nj.nvjpegCreateEx(be, None, None, 0, C.byref(hd))
    if st != 0:
        print(f"{name}: nvjpegCreateEx -> {st}")
        continue
    state = C.c_void_p()
    st = nj.nvjpegJpegStateCreate(hd, C.byref(state))
    if st != 0:
        print(f"{name}: nvjpegJpegStateCreate -> {st}")
        continue
    im = NvImage()
    im.channel[0] = stack.device_ptr
    im.pitch[0] = stack.pitch
    ts = []
    st = 0
    for _ in range(4):
        ctx.sync()
        t0 = time.perf_counter()
        st = nj.nvjpegDecode(hd, state, buf, C.c_size_t(len(data)), BGRI, C.byref(im), None)
        ctx.sync()
        import ctypes
        C.CDLL("libcudart.so.12").cudaDeviceSynchronize()
        ts.append(time.perf_counter() - t0)
        if st != 0:
            break
    single = f"nvjpegDecode -> {st}" if st else f"nvjpegDecode {min(ts) * 1e3:.2f} ms"
    bstate = C.c_void_p()
    nj.nvjpegJpegStateCreate(hd, C.byref(bstate))
    st = nj.nvjpegDecodeBatchedInitialize(hd, bstate, B, 1, BGRI)
    if st != 0:
        print(f"{name}: {single}; nvjpegDecodeBatchedInitialize -> {st}")
        continue
    ims = (NvImage * B)()
    for i in range(B):
        ims[i].channel[0] = stack.device_ptr + i * stack.frame_stride
        ims[i].pitch[0] = stack.pitch
    ptrs = (C.c_void_p * B)(*[C.addressof(buf)] * B)
    lens = (C.c_size_t * B)(*[len(data)] * B)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        st = nj.nvjpegDecodeBatched(hd, bstate, ptrs, lens, ims, None)
        C.CDLL("libcudart.so.12").cudaDeviceSynchronize()
        ts.append(time.perf_counter() - t0)
        if st != 0:
            break
    print(f"{name}: {single}; " + (f"nvjpegDecodeBatched -> {st}" if st else f"nvjpegDecodeBatched({B}) {min(ts) / B * 1e3:.2f} ms per frame"))

# decoupled API: the backend choice is honoured here (nvjpegDecoderCreate); phases timed separately
vp = C.c_void_p
cudart = C.CDLL("libcudart.so.12")
hd = vp(); nj.nvjpegCreateSimple(C.byref(hd))
for be in (1, 2):
    dec, dst_, params, js, pin, dev = vp(), vp(), vp(), vp(), vp(), vp()
    r = [nj.nvjpegDecoderCreate(hd, be, C.byref(dec)), nj.nvjpegDecoderStateCreate(hd, dec, C.byref(dst_)),
         nj.nvjpegBufferPinnedCreate(hd, None, C.byref(pin)), nj.nvjpegBufferDeviceCreate(hd, None, C.byref(dev)),
         nj.nvjpegStateAttachPinnedBuffer(dst_, pin), nj.nvjpegStateAttachDeviceBuffer(dst_, dev),
         nj.nvjpegJpegStreamCreate(hd, C.byref(js)), nj.nvjpegDecodeParamsCreate(hd, C.byref(params)),
         nj.nvjpegDecodeParamsSetOutputFormat(params, BGRI)]
    if any(r):
        print(f"decoupled {names[be]}: setup -> {r}")
        continue
    im = NvImage(); im.channel[0] = stack.device_ptr; im.pitch[0] = stack.pitch
    best = None
    for _ in range(5):
        cudart.cudaDeviceSynchronize()
        t0 = time.perf_counter()
        s1 = nj.nvjpegJpegStreamParse(hd, buf, C.c_size_t(len(data)), 0, 0, js)
        t1 = time.perf_counter()
        s2 = nj.nvjpegDecodeJpegHost(hd, dec, dst_, params, js)
        t2 = time.perf_counter()
        s3 = nj.nvjpegDecodeJpegTransferToDevice(hd, dec, dst_, js, None)
        s4 = nj.nvjpegDecodeJpegDevice(hd, dec, dst_, C.byref(im), None)
        t3 = time.perf_counter()
        cudart.cudaDeviceSynchronize()
        t4 = time.perf_counter()
        if s1 or s2 or s3 or s4:
            print(f"decoupled {names[be]}: status {s1} {s2} {s3} {s4}")
            break
        cur = (t4 - t0, t1 - t0, t2 - t1, t3 - t2, t4 - t3)
        best = cur if best is None or cur[0] < best[0] else best
    if best:
        print(f"decoupled {names[be]}: total {best[0] * 1e3:.2f} ms = parse {best[1] * 1e3:.2f} + host {best[2] * 1e3:.2f} + enqueue {best[3] * 1e3:.2f} + device wait {best[4] * 1e3:.2f}")
