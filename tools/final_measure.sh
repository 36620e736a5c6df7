# The measurements a round's profiles/ are regenerated from, on ONE 1-GPU box.  usage: gpurun -- bash tools/final_measure.sh [tag]
tag=${1:-final}
o=gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > $o/${tag}_gputest.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_gputest.log); tail -3 $o/${tag}_gputest.log
timeout 300 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --per-op > $o/${tag}_per_op.md 2> $o/${tag}_per_op.err; echo "per-op rc=$?"
timeout 200 python tools/depth16_bench.py 16 > $o/${tag}_depth16.md 2>&1; echo "depth16 rc=$?"
timeout 300 python tools/jpeg_bench.py 32 5 > $o/${tag}_jpeg_bench.md 2>&1; echo "jpeg rc=$?"
B="python bench.py --no-e2e --no-cpu --no-mosaic --no-single --steps 1 --warmup 3 --min-timed-s 0 --batch 64"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv $B > $o/${tag}_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"stencil_kernel|point_flat_kernel|gray3" --launch-skip 60 --launch-count 12 -f -o $o/${tag}_prof $B > $o/${tag}_ncu2.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"fused_g5e3" --launch-skip 2 --launch-count 1 -f -o $o/${tag}_prof_fused python tools/prof_one.py fused 64 4 > $o/${tag}_ncu3.log 2>&1; echo "ncu fused rc=$?"
# the reports are too large to travel (64 MiB limit on gpurun_out/): summarise here, keep the raw pages as CSV
python tools/summarize_ncu.py ${tag} $o/${tag}_launches.csv $o/${tag}_prof.ncu-rep 64 > /dev/null 2> $o/${tag}_summarize.err; cp profiles/${tag}_ncu_* $o/ 2>/dev/null
ncu -i $o/${tag}_prof.ncu-rep --page raw --csv > $o/${tag}_prof_raw.csv 2>/dev/null
ncu -i $o/${tag}_prof_fused.ncu-rep --page raw --csv > $o/${tag}_prof_fused_raw.csv 2>/dev/null
rm -f $o/${tag}_prof.ncu-rep $o/${tag}_prof_fused.ncu-rep
mkdir -p $o/cli_out
timeout 300 ./dip-benchmark_b200/dip_bench data/fundus.jpg $o/cli_out/ 10000 --csv $o/${tag}_cli_results.csv > $o/${tag}_cli_sync.txt 2>&1; echo "cli rc=$?"
timeout 300 ./dip-benchmark_b200/dip_bench data/fundus.jpg $o/cli_out/ 10000 --no-sync --csv $o/${tag}_cli_results.csv > $o/${tag}_cli_nosync.txt 2>&1; echo "cli nosync rc=$?"
rm -rf $o/cli_out
ls -la $o | tail -20
