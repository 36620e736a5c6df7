"""CPU: the C-ABI library loads and exports every symbol include/dip.h declares; the host-side
argument handling that needs no device."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(dip):
    syms = dip.declared_symbols()
    assert len(syms) >= 40
    L = dip.lib()
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/dip.h but not exported"
    # and the binding covers the header, no more, no less
    assert set(L._dip_signatures) == set(syms)


def test_exports_are_plain_c_symbols(dip):
    out = subprocess.check_output(["nm", "-D", "--defined-only", dip.lib_path()], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(dip.declared_symbols()) <= exported
    assert not [s for s in exported if s.startswith("_Z") and "dip" in s], "C++ symbols leak from the ABI"


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "dip.h"\nint main(void){ dip_op o; (void)o; return DIP_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_sass_is_blackwell_native(dip):
    """The stencil kernels must carry TMA loads (UTMALDG) and the packed min instruction."""
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", dip.lib_path()], text=True, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    assert "UTMALDG" in sass
    assert "VIMNMX3.U16x2" in sass or "VIMNMX3" in sass
    assert "IDP.4A" in sass


def test_reference_op_list_matches_the_reference(dip):
    L = dip.lib()
    from dip_benchmark_b200 import capi
    arr = (capi.Op * 12)()
    assert L.dip_reference_ops(arr, 12) == 12
    prefixes = [L.dip_reference_op_prefix(i).decode() for i in range(12)]
    assert prefixes == [p for p, _, _ in dip.REFERENCE_OPS]
    assert L.dip_reference_op_description(4) == b"Erosion (3x3 Cross Kernel)"
    assert L.dip_reference_ops(arr, 3) < 0


def test_no_device_fails_loudly(dip):
    n = C.c_int(-1)
    rc = dip.lib().dip_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(dip.DipError) as e:
        dip.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_measure_time_mirrors_the_reference_loop(dip):
    """benchmark.measure_time = opencv/benchmark.py:26-38: one cold call, then `rounds` calls; optional sync hook."""
    from dip_benchmark_b200 import benchmark
    calls, syncs = [], []
    once, avg = benchmark.measure_time(lambda: calls.append(1), 7, lambda: syncs.append(1))
    assert len(calls) == 8 and len(syncs) == 2 and once >= 0 and avg >= 0
    # recorded in blocks: 23 rounds = 1 cold call + 3 recorded while building the graph + 7 replays of 3 + 2 plain calls
    class G:
        launches, closed = 0, False
        def launch(self): G.launches += 1
        def close(self): G.closed = True
    calls.clear()
    def record(func, block):
        for _ in range(block):
            func()
        return G()
    benchmark.measure_time(lambda: calls.append(1), 23, None, record, 3)
    assert (len(calls), G.launches, G.closed) == (1 + 3 + 2, 7, True)
    calls.clear()
    benchmark.measure_time(lambda: calls.append(1), 5, None, lambda f, b: None, 3)   # not recordable: plain calls
    assert len(calls) == 6


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference runs here on the CPU (OpenCV path or the C port) and prints exactly one JSON line."""
    import json
    import sys
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                   "--batch", "8"], text=True)
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
    # both variants of BASELINE.md 3: all host cores (the line's value) and the reference-faithful single stream of calls
    assert d["cpu_baseline"]["all_cores"]["value"] == d["value"] and d["cpu_baseline"]["faithful"]["value"] > 0
    assert d["config"]["frames_total"] == 8 and d["detail"]["cv2_num_threads"] is not None


def test_both_bench_arms_describe_the_same_config():
    """The driver compares the two arms' `config`: they must be equal for equal arguments."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.config_dict(256) == bench.config_dict(256) and "frames_total" in bench.config_dict(256)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_dict(args.batch)') == 2   # one per arm, nothing arm-specific inside


def test_bench_reads_traffic_from_the_committed_ncu_capture_and_parses_clock_samples():
    """bench.py's roofline.traffic comes from the newest profiles/r*_ncu_full.json (not a constant), and its clock sampler
    turns nvidia-smi rows into the `clocks` object (median under load, reasons) for the timed window only."""
    import importlib.util
    import json
    import time
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    per_frame, src = bench.ncu_traffic_per_frame("convolution-gaussian-blur-5x5-separated")
    assert src.startswith("profiles/") and src.endswith("_ncu_full.json")
    cap = json.load(open(os.path.join(ROOT, src)))
    k = [x for x in cap["kernels"] if x.get("op") == "convolution-gaussian-blur-5x5-separated"][0]
    assert per_frame == (k["dram_read_bytes"] + k["dram_write_bytes"]) / cap["frames"]
    assert 0.9 < per_frame / (2 * bench.FRAME_BYTES) < 1.1           # measured DRAM traffic ~ algorithmic bytes
    assert bench.ncu_traffic_per_frame("no-such-op")[0] is None
    cs = bench.ClockSampler(0)
    now = time.time()
    cs.t0, cs.t1 = now - 1.0, now
    row = lambda mhz, cap_: [str(mhz), "1965", "998.5", "0x4", "Not Active", "Not Active", "Not Active", cap_]   # noqa: E731
    cs.rows = [(now - 5.0, row(1965, "Not Active")), (now - 0.8, row(1530, "Active")), (now - 0.5, row(1545, "Active")), (now - 0.2, row(1560, "Active"))]
    out = cs.summary()
    assert out == {"sm_mhz": 1545.0, "sm_max_mhz": 1965.0, "power_w_max": 998.5, "reasons": ["sw_power_cap"], "samples": 3, "window": "timed region"}


def test_chain_planner_counts_passes_without_a_gpu():
    """dip_chain_passes is host logic: byte-wise runs fold into one map that rides on a neighbouring tiled kernel."""
    import numpy as np
    from dip_benchmark_b200 import Context, make_op
    cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], np.uint8)
    g5 = np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1]).astype(np.float32) / 256
    ops = {"gray": make_op("grayscale"), "inv": make_op("invert"), "thr": make_op("threshold", thresh=127, maxval=255),
           "erode": make_op("erode", mask=np.ones((3, 3), np.uint8)), "cross": make_op("erode", mask=cross),
           "tall": make_op("erode", mask=np.ones((5, 3), np.uint8)), "g5": make_op("convolve", kernel=g5),
           "blur": make_op("gaussian_blur3x3")}
    def passes(names, c=3):
        return Context.chain_passes([ops[n] for n in names], c)
    assert passes(["gray", "blur", "thr", "erode"]) == 3          # gray | blur+threshold | erode (4 reference calls)
    assert passes(["thr", "cross", "inv"]) == 1
    assert passes(["inv", "thr", "inv", "thr"]) == 1
    assert passes(["g5", "erode"]) == 1                            # the fused Gaussian+erosion kernel
    assert passes(["inv", "g5", "erode", "thr"]) == 1              # ... which takes the maps on both sides
    assert passes(["inv", "tall", "thr"]) == 3                     # direct kernels take no maps
    assert passes(["gray", "inv", "gray"], 3) == 1 and passes(["gray", "inv", "gray"], 4) == 2 and passes(["gray", "inv", "gray"], 1) == 1


def test_byte_wise_runs_fold_into_one_exact_map():
    """ByteMap composition (the host logic behind dip_apply_chain) against numpy on all 256 values: 2000 random runs
    plus every pair of thresholds at the edge values."""
    import numpy as np
    import pytest
    from dip_benchmark_b200 import Context, DipError, make_op
    x = np.arange(256, dtype=np.uint8)
    rng = np.random.default_rng(11)
    edge = [0, 1, 126, 127, 128, 254, 255]
    def run(seq):
        ops, want = [], x.copy()
        for kind, t, m in seq:
            if kind == "inv":
                ops.append(make_op("invert")); want = 255 - want
            elif kind == "copy":
                ops.append(make_op("copy"))
            else:
                ops.append(make_op("threshold", thresh=t, maxval=m)); want = np.where(want > t, m, 0).astype(np.uint8)
        assert np.array_equal(Context.chain_lut(ops), want), seq
    for _ in range(2000):
        seq = []
        for _ in range(int(rng.integers(1, 8))):
            r = rng.random()
            seq.append(("inv", 0, 0) if r < 0.4 else ("copy", 0, 0) if r < 0.5 else
                       ("thr", int(rng.choice(edge + [int(rng.integers(0, 256))])), int(rng.choice(edge + [int(rng.integers(0, 256))]))))
        run(seq)
    for t1 in edge:
        for m1 in edge:
            for t2 in edge:
                run([("thr", t1, m1), ("inv", 0, 0), ("thr", t2, 200), ("inv", 0, 0)])
                run([("inv", 0, 0), ("thr", t1, m1), ("thr", t2, 77)])
    with pytest.raises(DipError):
        Context.chain_lut([make_op("grayscale")])


def _build_c_example(tmp_path):
    import subprocess
    lib_dir = os.path.join(ROOT, "dip-benchmark_b200")
    exe = str(tmp_path / "chain")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "chain.c"), "-L", lib_dir, "-ldip_b200", "-Wl,-rpath," + lib_dir, "-o", exe])
    return exe


def test_plain_c_caller_links_and_fails_loudly_without_a_gpu(dip, tmp_path):
    """examples/chain.c (C99, no C++ runtime on the caller's side) links against the C-ABI; on a box without a GPU
    dip_init reports DIP_ERR_NO_DEVICE and the program exits 4 -- there is no CPU fallback to fall into."""
    import ctypes as C
    import subprocess
    exe = _build_c_example(tmp_path)
    n = C.c_int(0)
    if dip.lib().dip_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe, "64", "48", str(tmp_path / "out.raw")], capture_output=True, text=True)
    assert r.returncode == 4 and "no GPU" in r.stderr, (r.returncode, r.stderr)


def test_staging_copy_pool_runs_every_task_exactly_once(tmp_path):
    """csrc/copy_pool.h (the host threads behind pageable uploads / downloads): 240 000 runs from four caller threads
    with varying task counts, default spin and a spin so short that workers keep falling asleep."""
    exe = str(tmp_path / "copy_pool_stress")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "dip-benchmark_b200", "csrc"),
                           os.path.join(ROOT, "tests", "copy_pool_stress.cpp"), "-o", exe])
    for env in ({}, {"DIP_COPY_THREADS": "5", "DIP_COPY_SPIN_US": "5"}):
        r = subprocess.run([exe], capture_output=True, text=True, env={**os.environ, **env}, timeout=300)
        assert r.returncode == 0 and "errors 0" in r.stdout, (env, r.stdout, r.stderr)
