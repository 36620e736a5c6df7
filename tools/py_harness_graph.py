"""The Python mirror of the reference harness (dip-benchmark_b200/benchmark.py) on data/fundus.jpg: plain calls against
loops replayed from CUDA graphs of 100 calls (--graph 100).  usage: python tools/py_harness_graph.py [rounds=2000]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
from _dip_pkg import load
dip = load()
from dip_benchmark_b200 import benchmark
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
img = cv2.imread(os.path.join(ROOT, "data", "fundus.jpg"))
with tempfile.TemporaryDirectory() as d:
    a, b = [], []
    benchmark.perform_benchmark(img, "fundus.png", d, rounds, out=a.append)
    benchmark.perform_benchmark(img, "fundus.png", d, rounds, out=b.append, graph_block=100)
print(f"| operation ({rounds} rounds, enqueue-only timing like the reference's OpenCV harness) | plain calls, us per call | replayed from graphs of 100 calls, us per call |")
print("|---|---|---|")
for x, y in zip(a, b):
    name = x.split("|")[1].strip()
    tx = float(x.split("|")[3].split("s")[0]) * 1e6
    ty = float(y.split("|")[3].split("s")[0]) * 1e6
    print(f"| {name} | {tx:.2f} | {ty:.2f} |")
