#!/usr/bin/env bash
# Runs under gpurun (tag = $1): GPU parity suite, default bench line, heavy-bin policy sweep on a 3 G k-mer job, launch list
set -u
TAG="${1:-r01p}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc >> "$OUT/gpu.txt"
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -3 "$OUT/pytest_gpu.log"
fi
timeout 600 python bench.py --steps 5 --warmup 3 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 2500 "$OUT/bench_n1.json"
for S in ${SWEEP:-0 1 2 4}; do
  DSKGPU_SMEM_MAX_SPLIT0=$S timeout 600 python bench.py --steps 3 --warmup 3 --genome 125000000 --coverage 30 --device-synth --no-e2e --no-cpu-baseline \
     > "$OUT/bench_g125m_split$S.json" 2> "$OUT/bench_g125m_split$S.err"
  python - "$OUT/bench_g125m_split$S.json" $S <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("split0<=%s: %.2f Gk/s  %s  smem_parts=%s splits=%s" % (sys.argv[2], d["value"], {k: round(v, 2) for k, v in d["stage_ms"].items()}, d["config"]["smem_partitions"], d["config"]["smem_splits"]))
except Exception as e:
    print("split", sys.argv[2], "failed", e)
PY
done
DSKGPU_SMEM_MAX_SPLIT0=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/bench_n1_split0.json" 2> "$OUT/bench_n1_split0.err"; tail -c 1200 "$OUT/bench_n1_split0.json"
if [ "${SKIP_NCU:-0}" != 1 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > "$OUT/launches.log" 2>&1
fi
ls "$OUT"
