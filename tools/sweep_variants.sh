#!/usr/bin/env bash
# runs bench.py over library variants (variants/*.so) and partition-size knobs; prints value + stage_ms per combination
OUT=${1:-gpurun_out/sweep}; mkdir -p $OUT
for lib in ${LIBS:-variants/*.so}; do
  for t in ${TS:-125}; do
    r=$(DSKGPU_LIB=$PWD/$lib DSKGPU_SMEM_T_PCT=$t timeout 120 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline ${BENCH_EXTRA:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%.2f G/s  %.2f ms ' % (d['value'], d['ms_per_step']), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'dom %.3f ms' % d['roofline']['avg_launch_ms'], 'distinct', d['checks']['distinct_kmers'], 'solid', d['checks']['solid_kmers'], 'splits', d['engine']['smem_splits'], 'parts', d['engine']['partitions'])")
    echo "$lib T=$t : $r" | tee -a $OUT/sweep.txt
  done
done
