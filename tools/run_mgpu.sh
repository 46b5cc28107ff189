#!/usr/bin/env bash
# Runs under `gpurun --gpus N` (tag = $1, N = $2): real multi-process parity check + the driver's N-GPU bench launch
set -u
TAG="${1:-r01q_n2}"; N="${2:-2}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nvidia-smi topo -m >> "$OUT/gpu.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/mgpu_check.py > "$OUT/mgpu_check.log" 2>&1; echo "mgpu_check exit $?" >> "$OUT/mgpu_check.log"
grep -E "mgpu_check|Error|error" "$OUT/mgpu_check.log" | tail -20
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; tail -c 2500 "$OUT/bench_n$N.json"; tail -5 "$OUT/bench_n$N.err"
[ "${TRACE:-0}" = 1 ] && DSKGPU_TRACE_XCHG=1 timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 2 --warmup 2 --no-e2e > "$OUT/trace_n$N.json" 2> "$OUT/trace_n$N.err"; grep xchg "$OUT/trace_n$N.err" | tail -24
