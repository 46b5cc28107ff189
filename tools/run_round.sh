#!/usr/bin/env bash
# Runs under gpurun: GPU parity suite + the bench lines that go into profiles/ (tag = $1)
set -u
TAG="${1:-r01o}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc >> "$OUT/gpu.txt"; free -g >> "$OUT/gpu.txt"
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -3 "$OUT/pytest_gpu.log"
fi
timeout 600 python bench.py --steps 5 --warmup 3 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 3000 "$OUT/bench_n1.json"
DSKGPU_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/trace.json" 2> "$OUT/trace.err"
timeout 600 python bench.py --steps 5 --warmup 3 --kmer-size 63 > "$OUT/bench_k63.json" 2> "$OUT/bench_k63.err"; tail -c 3000 "$OUT/bench_k63.json"
timeout 600 python bench.py --steps 3 --warmup 3 --genome 125000000 --coverage 30 --device-synth --no-e2e --no-cpu-baseline > "$OUT/bench_g125m.json" 2> "$OUT/bench_g125m.err"; tail -c 3000 "$OUT/bench_g125m.json"; tail -5 "$OUT/bench_g125m.err"
timeout 600 python bench.py --steps 3 --warmup 3 --genome 125000000 --coverage 30 --device-synth --no-e2e --no-cpu-baseline --kmer-size 63 > "$OUT/bench_g125m_k63.json" 2> "$OUT/bench_g125m_k63.err"; tail -c 3000 "$OUT/bench_g125m_k63.json"; tail -5 "$OUT/bench_g125m_k63.err"
timeout 900 python bench.py --steps 2 --warmup 2 --histo2d --genome 100000000 --coverage 50 --device-synth --no-e2e > "$OUT/bench_c5_histo2d.json" 2> "$OUT/bench_c5_histo2d.err"; tail -c 2500 "$OUT/bench_c5_histo2d.json"; tail -5 "$OUT/bench_c5_histo2d.err"
