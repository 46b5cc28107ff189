#!/usr/bin/env bash
# Runs under gpurun on 1 GPU (tag = $1): GPU parity suite at HEAD, default bench + reference arm, 3 G k-mer job and -histo2D C5 shape with the job-sized minimizer (m = 12), launch list
set -u
TAG="${1:-r01w}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
tail -12 "$OUT/pytest_gpu.log"
timeout 300 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 2800 "$OUT/bench_n1.json"; tail -3 "$OUT/bench_n1.err"
timeout 600 python bench.py --steps 3 --warmup 3 --genome 125000000 --coverage 30 --device-synth --no-e2e --no-cpu-baseline > "$OUT/bench_g125m.json" 2> "$OUT/bench_g125m.err"; tail -c 2500 "$OUT/bench_g125m.json"; tail -5 "$OUT/bench_g125m.err"
timeout 600 python bench.py --steps 2 --warmup 2 --histo2d --genome 100000000 --coverage 50 --device-synth --no-e2e > "$OUT/bench_c5_histo2d.json" 2> "$OUT/bench_c5_histo2d.err"; tail -c 2500 "$OUT/bench_c5_histo2d.json"; tail -5 "$OUT/bench_c5_histo2d.err"
timeout 300 python bench.py --steps 5 --warmup 3 --minimizer-size 12 --no-e2e --no-cpu-baseline > "$OUT/bench_n1_m12.json" 2> "$OUT/bench_n1_m12.err"; tail -c 1500 "$OUT/bench_n1_m12.json"
