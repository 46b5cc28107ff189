#!/usr/bin/env bash
# Runs under gpurun on 1 GPU (tag = $1): heavy-partition parity subset, 3 G k-mer job (bucket path) + its ncu launch list, -histo2D C5 shape
set -u
TAG="${1:-r01v}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -x -q -k "heavy or mixed or multi_rank or histo2d or forced" > "$OUT/pytest_gpu_subset.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu_subset.log"
tail -12 "$OUT/pytest_gpu_subset.log"
timeout 600 python bench.py --steps 3 --warmup 3 --genome 125000000 --coverage 30 --device-synth --no-e2e --no-cpu-baseline > "$OUT/bench_g125m.json" 2> "$OUT/bench_g125m.err"; tail -c 2500 "$OUT/bench_g125m.json"; tail -5 "$OUT/bench_g125m.err"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --genome 125000000 --coverage 30 --device-synth --no-e2e --no-cpu-baseline > "$OUT/launches.log" 2>&1
timeout 600 python bench.py --steps 2 --warmup 2 --histo2d --genome 100000000 --coverage 50 --device-synth --no-e2e > "$OUT/bench_c5_histo2d.json" 2> "$OUT/bench_c5_histo2d.err"; tail -c 2500 "$OUT/bench_c5_histo2d.json"; tail -5 "$OUT/bench_c5_histo2d.err"
