"""JPEG on the GPU side of upload / download (dip_jpeg_*) against the reference's CPU calls
(cv.imread / cv.imwrite, opencv/benchmark.py:14,:114) on data/fundus.jpg (3504x2336, baseline 4:2:2, 2.4 MB).
usage: python tools/jpeg_bench.py [batch=32] [rounds=5]"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cv2
from _dip_pkg import load
dip = load()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
data = open(os.path.join(ROOT, "data", "fundus.jpg"), "rb").read()
arr = np.frombuffer(data, np.uint8)
ctx = dip.Context(0)
w, h, _ = ctx.jpeg_info(data)
mpx = w * h / 1e6
one = ctx.image(w, h, 3); stack = ctx.image(w, h, 3, frames=batch)


def best(f, n=rounds):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return min(ts), statistics.median(ts)


rows = []
cv2.setNumThreads(1)
t, m = best(lambda: cv2.imdecode(arr, cv2.IMREAD_COLOR))
rows.append(("cv.imdecode, 1 host thread (the reference's cv.imread minus the file read)", t, 1))
img = cv2.imdecode(arr, cv2.IMREAD_COLOR)
t, m = best(lambda: cv2.imencode(".jpg", img))
rows.append(("cv.imencode('.jpg'), 1 host thread (cv.imwrite minus the file write)", t, 1))
pinned = np.empty_like(img)
ctx.imdecode(data, dst=one)
t, m = best(lambda: ctx.imdecode(data, dst=one))
rows.append((f"dip_jpeg_decode, one frame into device memory [{ctx.jpeg_backend()}]", t, 1))
t, m = best(lambda: (cv2.imdecode(arr, cv2.IMREAD_COLOR), one.upload(img), ctx.sync()))
rows.append(("cv.imdecode + dip_upload (pageable): the reference's way into device memory", t, 1))
streams = [data] * batch
for threads in (1, 2, 4, 8, 16, 32):
    if threads > (os.cpu_count() or 1):
        break
    os.environ["DIP_JPEG_THREADS"] = str(threads)
    ctx.imdecode(streams, dst=stack)
    t, m = best(lambda: ctx.imdecode(streams, dst=stack))
    rows.append((f"dip_jpeg_decode_batch, {batch} frames into a device stack, {threads} host thread(s) [{ctx.jpeg_backend()}]", t, batch))
os.environ["DIP_JPEG_CPU_HUFFMAN"] = "1"
ctx2 = dip.Context(0)
stack2 = ctx2.image(w, h, 3, frames=batch)
ctx2.imdecode(streams, dst=stack2)
t, m = best(lambda: ctx2.imdecode(streams, dst=stack2))
rows.append((f"dip_jpeg_decode_batch, {batch} frames, {os.environ['DIP_JPEG_THREADS']} host threads, Huffman on the CPU [{ctx2.jpeg_backend()}]", t, batch))
import concurrent.futures as cf
cv2.setNumThreads(1)
with cf.ThreadPoolExecutor(os.cpu_count()) as ex:
    t, m = best(lambda: list(ex.map(lambda _: cv2.imdecode(arr, cv2.IMREAD_COLOR), range(batch))))
rows.append((f"cv.imdecode, {batch} frames over {os.cpu_count()} host threads (host memory only)", t, batch))
ctx.imencode(one)
t, m = best(lambda: ctx.imencode(one))
rows.append(("dip_jpeg_encode, one device frame to a bitstream on the host (quality 95, 4:2:0)", t, 1))
t, m = best(lambda: (one.download(pinned), cv2.imencode(".jpg", pinned)))
rows.append(("dip_download + cv.imencode: the reference's way out of device memory", t, 1))
print(f"| path ({w}x{h} RGB, {len(data)} B compressed) | ms per frame | frames/s | Mpixel/s |")
print("|---|---|---|---|")
for name, t, n in rows:
    print(f"| {name} | {t / n * 1e3:.2f} | {n / t:.1f} | {n * mpx / t:.0f} |")
a = stack.download(); b = stack2.download()
print(f"\nGPU vs CPU Huffman decode of the same stream: max |diff| {int(np.abs(a.astype(int) - b.astype(int)).max())}, PSNR vs cv.imdecode: "
      f"{10 * np.log10(255.0 ** 2 / max(np.mean((a[0].astype(float) - img) ** 2), 1e-12)):.1f} / {10 * np.log10(255.0 ** 2 / max(np.mean((b[0].astype(float) - img) ** 2), 1e-12)):.1f} dB")
