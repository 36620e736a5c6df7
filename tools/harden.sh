# Hardening batch on one GPU: the concurrent-context test, the randomised soak, compute-sanitizer (memcheck, racecheck) on
# the small-shape parity tests.  usage: gpurun -- bash tools/harden.sh [tag]
tag=${1:-harden}; o=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "one_context_per_thread" > $o/${tag}_threads.log 2>&1; echo "threads rc=$?"; tail -3 $o/${tag}_threads.log
timeout 600 python tests/soak/stress.py 4000 20261018 > $o/${tag}_soak.log 2>&1; echo "soak rc=$?"; tail -3 $o/${tag}_soak.log
K="all_reference_ops or golden_small or extreme_and_tie or fused_gauss5x5 or one_dimensional or arbitrary_small or random_chains or recorded_random"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q -k "$K" > $o/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $o/${tag}_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q -k "all_reference_ops or fused_gauss5x5 or arbitrary_small" > $o/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $o/${tag}_racecheck.log | tail -3
