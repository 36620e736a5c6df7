tag=r2j; o=gpurun_out
B="python bench.py --no-e2e --no-cpu --no-mosaic --no-single --steps 1 --warmup 3 --min-timed-s 0 --batch 64"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv $B > $o/${tag}_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"stencil_kernel|point_flat_kernel|gray3" --launch-skip 60 --launch-count 12 -f -o /tmp/${tag}_prof $B > $o/${tag}_ncu2.log 2>&1; echo "ncu full rc=$?"
python tools/summarize_ncu.py ${tag} $o/${tag}_launches.csv /tmp/${tag}_prof.ncu-rep 64 > /dev/null 2> $o/${tag}_summarize.err; cp profiles/${tag}_ncu_* $o/ 2>/dev/null
grep "stencil_kernel<8" $o/${tag}_ncu_full.md
