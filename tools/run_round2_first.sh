#!/usr/bin/env bash
# First gpurun call of round 2 (1 GPU): what round 1 could not measure any more.
#   parity suite, default bench (both arms), scaled-twin parity of the 3 Gbp configurations at k = 31 and 63
set -u
TAG="${1:-r02a}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -4 "$OUT/pytest_gpu.log"
timeout 300 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 2500 "$OUT/bench_n1.json"
timeout 600 python tools/twin_check.py --kmer-size 31 > "$OUT/twin_k31.json" 2> "$OUT/twin_k31.err"; cat "$OUT/twin_k31.json"; tail -3 "$OUT/twin_k31.err"
timeout 600 python tools/twin_check.py --kmer-size 63 > "$OUT/twin_k63.json" 2> "$OUT/twin_k63.err"; cat "$OUT/twin_k63.json"; tail -3 "$OUT/twin_k63.err"
