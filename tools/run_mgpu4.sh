#!/usr/bin/env bash
# Runs under `gpurun --gpus N` (tag = $1, N = $2): multi-process parity check (short list) + the driver's N-GPU bench launch
set -u
TAG="${1:-r01z_n4}"; N="${2:-4}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MGPU_SHORT=1 timeout 150 $TR --master-port 29541 tools/mgpu_check.py > "$OUT/mgpu_check.log" 2>&1; echo "mgpu_check exit $?" >> "$OUT/mgpu_check.log"
grep -E "mgpu_check" "$OUT/mgpu_check.log" | tail -12
timeout 150 $TR --master-port 29542 bench.py --gpus $N --no-e2e > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
echo "bench exit $?"; cut -c1-3500 "$OUT/bench_n$N.json"; tail -3 "$OUT/bench_n$N.err" | cut -c1-300
