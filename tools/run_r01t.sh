#!/usr/bin/env bash
# Runs under gpurun on 1 GPU (tag = $1): multi-bank / -histo2D parity subset + -histo2D (C5 shape) bench line
set -u
TAG="${1:-r01t}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -x -q -k "histo2d or perbank or c123 or forced or tiny or cli" > "$OUT/pytest_gpu_subset.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu_subset.log"
tail -5 "$OUT/pytest_gpu_subset.log"
timeout 600 python bench.py --steps 2 --warmup 2 --histo2d --genome 100000000 --coverage 50 --device-synth --no-e2e > "$OUT/bench_c5_histo2d.json" 2> "$OUT/bench_c5_histo2d.err"; tail -c 2500 "$OUT/bench_c5_histo2d.json"; tail -5 "$OUT/bench_c5_histo2d.err"
