"""Turn the ncu outputs in gpurun_out/ into the summaries committed under profiles/.
usage: python tools/summarize_ncu.py ROUND_TAG launches.csv report.ncu-rep [frames=256]
Writes profiles/TAG_ncu_launches.md, profiles/TAG_ncu_full.md and profiles/TAG_ncu_full.json (per-kernel DRAM bytes;
bench.py reads its roofline.traffic from the newest such file).  The full capture must hold exactly one step of
`python bench.py` (12 launches in the reference's op order) for the kernels to be labelled with their op."""
import collections, csv, json, os, re, subprocess, sys

tag, launches, report = sys.argv[1], sys.argv[2], sys.argv[3]
FRAMES = int(sys.argv[4]) if len(sys.argv) > 4 else 256
OPS = ["copy", "inversion", "grayscale", "threshold", "erosion-cross", "erosion-square", "erosion-square-separated",
       "convolution-gaussian-blur-3x3", "convolution-gaussian-blur-3x3-separated", "convolution-gaussian-blur-5x5",
       "convolution-gaussian-blur-5x5-separated", "gaussian-blur-3x3"]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAME = 24556032

def units_ms(u):
    return u == "ms"


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void unnamed>::", "").replace("unnamed>::", "")

# ---- launch list -------------------------------------------------------------------------------
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")) / 1e3
    agg.setdefault((short(row["Kernel Name"]), row["Grid Size"], row["Block Size"]), []).append(v)
tot = sum(sum(v) for v in agg.values())
out = [f"# {tag} ncu launch list: `python bench.py --no-e2e --no-cpu --no-mosaic --no-single --steps 1 --warmup 3 --min-timed-s 0 --batch {FRAMES}` (B200, --clock-control none)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv`: cold-cache, serialised launches; compare SHARES.", "",
       "| kernel | grid | block | launches | mean us | share |", "|---|---|---|---|---|---|"]
for (k, g, b), v in agg.items():
    out.append(f"| {k} | {g} | {b} | {len(v)} | {sum(v)/len(v):.1f} | {100*sum(v)/tot:.1f}% |")
out += ["", "stencil_kernel<OP,C,R,W>: OP 0 erode square (also the fused 1x3+3x1), 1 erode cross, 4 conv 3x3, 5 conv 1x3+3x1 fused, "
        "6 GaussianBlur 3x3, 7 conv 5x5, 8 conv 1x5+5x1 fused.  point_flat_kernel<OP,U>: 0 copy, 1 invert, 2 threshold.",
        "synth_kernel / checksum_kernel run once outside the timed region."]
open(os.path.join(root, "profiles", f"{tag}_ncu_launches.md"), "w").write("\n".join(out) + "\n")

# ---- full set ----------------------------------------------------------------------------------
raw = subprocess.check_output(["ncu", "-i", report, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "launch__grid_size"]
stall = [h for h in hdr if "issue_stalled" in h and "pcsamp" in h and "not_issued" not in h]
kernels_json = []
out = [f"# {tag} ncu --set full: every kernel of one step of the {FRAMES}-frame batch (B200, --clock-control none)", "",
       f"Algorithmic bytes per launch = 2 x {FRAME} x {FRAMES} = {2*FRAME*FRAMES/1e9:.3f} GB.  dram % is of ncu's sustained DRAM peak (~8.2 TB/s).", "",
       "| kernel | ms | dram read GB | dram write GB | traffic / algorithmic | alg. GB/s | dram % | issue active % | alu pipe % | fma pipe % | warps active % | regs | warp-instr | top stalls |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for d in data:
    name = short(d[idx["Kernel Name"]])
    v = [d[idx[k]].replace(",", "") for k in keys]
    ms = float(v[0]) if units_ms(rows[1][idx[keys[0]]]) else float(v[0]) / 1e3 if rows[1][idx[keys[0]]] == "us" else float(v[0])
    rd, wr = float(v[1]), float(v[2])
    if rows[1][idx[keys[1]]].lower().startswith("mbyte"): rd /= 1e3
    if rows[1][idx[keys[2]]].lower().startswith("mbyte"): wr /= 1e3
    alg = 2 * FRAME * FRAMES / 1e9
    sv = sorted(((float(d[idx[k]].replace(",", "") or 0), k.split("stalled_")[1]) for k in stall), reverse=True)
    st = sum(x for x, _ in sv) or 1
    tops = ", ".join(f"{n} {100*x/st:.0f}" for x, n in sv[:3])
    out.append(f"| {name} | {ms:.3f} | {rd:.3f} | {wr:.3f} | {(rd+wr)/alg:.3f} | {alg/ms*1e3:.0f} | {float(v[3]):.1f} | {float(v[4]):.1f} | {float(v[5]):.1f} | "
               f"{float(v[6]):.1f} | {float(v[7]):.1f} | {v[8]} | {float(v[9])/1e6:.0f}M | {tops} |")
    kernels_json.append({"kernel": name, "ms": ms, "dram_read_bytes": rd * 1e9, "dram_write_bytes": wr * 1e9, "registers": int(float(v[8])),
                         "warp_instructions": float(v[9]), "issue_active_pct": float(v[4]), "alu_pipe_pct": float(v[5]), "fma_pipe_pct": float(v[6])})
if len(kernels_json) == len(OPS):
    for k, op in zip(kernels_json, OPS):
        k["op"] = op
open(os.path.join(root, "profiles", f"{tag}_ncu_full.md"), "w").write("\n".join(out) + "\n")
json.dump({"tag": tag, "frames": FRAMES, "algorithmic_bytes_per_launch": 2 * FRAME * FRAMES, "kernels": kernels_json,
           "how": "ncu --set full --clock-control none, one step of bench.py's batch leg; per launch"},
          open(os.path.join(root, "profiles", f"{tag}_ncu_full.json"), "w"), indent=1)
print("\n".join(out))
