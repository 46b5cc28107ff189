#!/usr/bin/env bash
# Runs under gpurun on 1 GPU (tag = $1): GPU parity suite, the driver's two bench arms, k=63 line, launch list, ncu --set full of the two top kernels
set -u
TAG="${1:-r01q}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc >> "$OUT/gpu.txt"; free -g >> "$OUT/gpu.txt"
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -3 "$OUT/pytest_gpu.log"
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; tail -1 "$OUT/smoke.log"
timeout 600 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 3000 "$OUT/bench_n1.json"; tail -3 "$OUT/bench_n1.err"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; tail -c 1200 "$OUT/bench_ref.json"
timeout 600 python bench.py --steps 5 --warmup 3 --kmer-size 63 --no-cpu-baseline > "$OUT/bench_k63.json" 2> "$OUT/bench_k63.err"; tail -c 2500 "$OUT/bench_k63.json"
if [ "${SKIP_NCU:-0}" != 1 ]; then
  KERNELS="${KERNELS:-k_count_smem k_superkmers}" timeout 900 tools/profile_gpu.sh "$TAG" > "$OUT/profile.log" 2>&1
fi
ls "$OUT"
