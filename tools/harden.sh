# Hardening batch on one GPU: the concurrent-context test and the randomised soak (compute-sanitizer is closed on this pool:
# out-of-bounds accesses are looked for with small cases, foreign-pitch regions of interest and comparison with the CPU oracle).  usage: gpurun -- bash tools/harden.sh [tag]
tag=${1:-harden}; o=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "one_context_per_thread" > $o/${tag}_threads.log 2>&1; echo "threads rc=$?"; tail -3 $o/${tag}_threads.log
timeout 600 python tests/soak/stress.py 4000 20261018 > $o/${tag}_soak.log 2>&1; echo "soak rc=$?"; tail -3 $o/${tag}_soak.log
