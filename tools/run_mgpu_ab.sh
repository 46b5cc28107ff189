#!/usr/bin/env bash
# Runs under `gpurun --gpus N`: A/B of the minimizer length on the N-GPU bench, same box, device-resident leg only
set -u
TAG="${1:-r01x_n2}"; N="${2:-2}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for M in 10 12 10 12; do
  timeout 200 $TR --master-port 2952$M bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --minimizer-size $M >> "$OUT/bench_m$M.json" 2>> "$OUT/bench_m$M.err"
done
timeout 300 $TR --master-port 29530 bench.py --gpus $N > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
grep -h '^{' "$OUT"/bench_m10.json "$OUT"/bench_m12.json "$OUT/bench_n$N.json" | cut -c1-1400
