#!/usr/bin/env bash
# Runs under gpurun: launch list of one bench step + one `ncu --set full` capture per hot kernel.
# usage: tools/profile_gpu.sh <tag> [extra bench args]
set -u
TAG="${1:-r01}"; shift || true
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
BENCH="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline $*"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file "$OUT/launches.csv" $BENCH > "$OUT/launches.log" 2>&1
for K in ${KERNELS:-k_count_smem k_superkmers k_scan_emit k_scan_tables k_part_scatter k_rs_onesweep}; do
  # kernels that run once per step are captured on their second launch (the timed step), the others on their third
  case $K in k_count_smem|k_part_scatter|k_rs_hist) S=1;; *) S=${SKIP:-2};; esac
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o "$OUT/prof_$K" $BENCH > "$OUT/prof_$K.log" 2>&1
done
ls -la "$OUT"
