#!/usr/bin/env bash
# Runs under gpurun on 1 GPU (tag = $1): ncu launch list + --set full of the dominant kernel at HEAD (C2), and a 9 G k-mer job (375 Mbp x 30x)
set -u
TAG="${1:-r01y}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
KERNELS="k_count_smem" timeout 400 tools/profile_gpu.sh "$TAG" > "$OUT/profile.log" 2>&1
timeout 500 python bench.py --steps 2 --warmup 2 --genome 375000000 --coverage 30 --device-synth --no-e2e --no-cpu-baseline > "$OUT/bench_g375m.json" 2> "$OUT/bench_g375m.err"; tail -c 2500 "$OUT/bench_g375m.json"; tail -5 "$OUT/bench_g375m.err"
ls "$OUT"
