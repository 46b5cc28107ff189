#!/usr/bin/env bash
# Runs under `gpurun --gpus N`:  tools/run_mgpu.sh <tag> <N> [steps...]   steps: tests | bench | bench63 | trace
set -u
TAG="${1:-r02_n2}"; N="${2:-2}"; shift; shift; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nvidia-smi topo -m >> "$OUT/gpu.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for step in "$@"; do
  case "$step" in
    tests)   timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_cli_dropin.py -x -q -k "real or torchrun" > "$OUT/pytest_mgpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_mgpu.log"; tail -6 "$OUT/pytest_mgpu.log" ;;
    bench)   timeout 1500 $TR --master-port 29512 bench.py --gpus $N ${BENCH_ARGS:-} > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench exit $?"; cut -c1-4000 "$OUT/bench_n$N.json"; tail -5 "$OUT/bench_n$N.err" | cut -c1-400 ;;
    bench63) timeout 1500 $TR --master-port 29514 bench.py --gpus $N --kmer-size 63 ${BENCH_ARGS:-} > "$OUT/bench_k63_n$N.json" 2> "$OUT/bench_k63_n$N.err"; echo "bench63 exit $?"; cut -c1-3000 "$OUT/bench_k63_n$N.json"; tail -5 "$OUT/bench_k63_n$N.err" | cut -c1-400 ;;
    trace)   DSKGPU_TRACE_XCHG=1 timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 2 --warmup 2 --no-e2e > "$OUT/trace_n$N.json" 2> "$OUT/trace_n$N.err"; grep xchg "$OUT/trace_n$N.err" | tail -24 ;;
    ref)     timeout 1500 $TR --master-port 29515 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > "$OUT/bench_ref_n$N.json" 2> "$OUT/bench_ref_n$N.err"; cat "$OUT/bench_ref_n$N.json" ;;
  esac
done
