#!/usr/bin/env bash
# Runs under `gpurun --gpus N` (tag = $1, N = $2): the driver's N-GPU bench launch only
set -u
TAG="${1:-r01z_n2}"; N="${2:-2}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
echo "exit $?"; cut -c1-3500 "$OUT/bench_n$N.json"; tail -3 "$OUT/bench_n$N.err" | cut -c1-300
