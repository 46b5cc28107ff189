#!/usr/bin/env bash
# Runs under gpurun on 1 GPU (tag = $1): GPU parity suite, default bench line, -histo2D (C5 shape) bench line
set -u
TAG="${1:-r01s}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
tail -15 "$OUT/pytest_gpu.log"
timeout 300 python bench.py --no-cpu-baseline > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 2600 "$OUT/bench_n1.json"; tail -3 "$OUT/bench_n1.err"
timeout 600 python bench.py --steps 2 --warmup 2 --histo2d --genome 100000000 --coverage 50 --device-synth --no-e2e > "$OUT/bench_c5_histo2d.json" 2> "$OUT/bench_c5_histo2d.err"; tail -c 2500 "$OUT/bench_c5_histo2d.json"; tail -5 "$OUT/bench_c5_histo2d.err"
ls "$OUT"
