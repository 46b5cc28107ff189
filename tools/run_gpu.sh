#!/usr/bin/env bash
# One parameterised runner for gpurun calls:  tools/run_gpu.sh <tag> <step> [<step> ...]
# steps: pytest | bench | bench63 | ref | twin31 | twin63 | sanitize | launches | ncu | cli | big
set -u
TAG="${1:-r02}"; shift; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for step in "$@"; do
  case "$step" in
    pynew)    timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_cli_dropin.py -x -q > "$OUT/pytest_new.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_new.log"; tail -15 "$OUT/pytest_new.log" ;;
    pycount)  timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "reference_runs or tiny_smem or synthetic or histo2d or heavy or multi_rank or forced or regrown or pass_loop" > "$OUT/pytest_count.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_count.log"; tail -4 "$OUT/pytest_count.log" ;;
    pyfast)   timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "device_planner or packed_bin or fine_histogram or record_sub or msd_multi or solid_set_ordering or multi_rank or multi_finish or synthetic_vs_oracle or tiny_smem or heavy or forced or regrown or pass_loop or auto_cutoff" > "$OUT/pytest_fast.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_fast.log"; tail -25 "$OUT/pytest_fast.log" ;;
    pytest)   timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -5 "$OUT/pytest_gpu.log" ;;
    bench)    timeout 600 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 3000 "$OUT/bench_n1.json"; tail -3 "$OUT/bench_n1.err" ;;
    benchq)   timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; tail -c 3000 "$OUT/bench_n1.json"; tail -3 "$OUT/bench_n1.err" ;;
    bench63)  timeout 600 python bench.py --kmer-size 63 --no-cpu-baseline > "$OUT/bench_k63.json" 2> "$OUT/bench_k63.err"; tail -c 2000 "$OUT/bench_k63.json" ;;
    ref)      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference_arm.json" 2> "$OUT/bench_reference_arm.err"; cat "$OUT/bench_reference_arm.json" ;;
    twin31)   timeout 900 python tools/twin_check.py --kmer-size 31 > "$OUT/twin_k31.json" 2> "$OUT/twin_k31.err"; cat "$OUT/twin_k31.json"; tail -3 "$OUT/twin_k31.err" ;;
    twin63)   timeout 900 python tools/twin_check.py --kmer-size 63 > "$OUT/twin_k63.json" 2> "$OUT/twin_k63.err"; cat "$OUT/twin_k63.json"; tail -3 "$OUT/twin_k63.err" ;;
    sanitize) for tool in memcheck racecheck; do
                timeout 900 compute-sanitizer --tool $tool --log-file "$OUT/sanitizer_${tool}_smoke.log" python __graft_entry__.py smoke > "$OUT/sanitizer_${tool}_smoke.out" 2>&1
                echo "$tool smoke exit $?"; tail -3 "$OUT/sanitizer_${tool}_smoke.log"
                timeout 1200 compute-sanitizer --tool $tool --log-file "$OUT/sanitizer_${tool}_split.log" python -m pytest tests/test_gpu_parity.py -x -q -k "tiny_smem_table_overflow_splits and (c1_k31 or c1_k63 or histo2d_k31) and 64" > "$OUT/sanitizer_${tool}_split.out" 2>&1
                echo "$tool split exit $?"; tail -3 "$OUT/sanitizer_${tool}_split.log"; tail -2 "$OUT/sanitizer_${tool}_split.out"
              done ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/launches_bench.log" 2>&1; echo "launches exit $?" ;;
    ncu)      timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNELS:-k_}" -s "${NCU_SKIP:-60}" -c "${NCU_COUNT:-24}" -f -o "$OUT/ncu_full" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/ncu_bench.log" 2>&1; echo "ncu exit $?"; ls -la "$OUT" ;;
    big)      timeout 900 python bench.py --genome 375000000 --coverage 30 --device-synth --minimizer-size ${BIG_M:-14} --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > "$OUT/bench_big.json" 2> "$OUT/bench_big.err"; cut -c1-2500 "$OUT/bench_big.json"; tail -3 "$OUT/bench_big.err" ;;
    bigmsd)   DSKGPU_MSD_MIN_PARTS=1 timeout 900 python bench.py --genome 375000000 --coverage 30 --device-synth --minimizer-size ${BIG_M:-14} --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > "$OUT/bench_bigmsd.json" 2> "$OUT/bench_bigmsd.err"; cut -c1-2500 "$OUT/bench_bigmsd.json"; tail -3 "$OUT/bench_bigmsd.err" ;;
    biglaunches) timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches_big.csv" python bench.py --genome 375000000 --coverage 30 --device-synth --minimizer-size ${BIG_M:-14} --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > "$OUT/launches_big.log" 2>&1; echo "biglaunches exit $?" ;;
    ncuk)     # one full capture per named kernel (small report): NCU_LIST="k_scan_emit k_superkmers ..."
              for kn in ${NCU_LIST:-k_scan_tables k_scan_emit k_superkmers k_rs_onesweep k_msd_pass k_count_smem}; do
                case "$kn" in k_count_smem) sk=1 ;; k_msd_pass) sk=2 ;; k_rs_onesweep) sk=3 ;; k_rs_fix) sk=1 ;; *) sk=6 ;; esac   # a launch of the second (timed) step (C2, device-resident input: 5 pieces of 128 MiB, 2 scatter passes, 3 ordering passes per step)
                timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kn" -s "$sk" -c 1 -f -o "$OUT/ncu_$kn" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/ncu_$kn.log" 2>&1; echo "ncu $kn exit $?"
              done; ls -la "$OUT" ;;
    cli)      timeout 900 bash tools/run_cli_check.sh "$OUT" ;;
    proxy)    # one GPU standing in for one of the 8 ranks of configs[2]/[3]: same k-mers per rank, 2^19 fine bins = the k-mers per bin
              # of 2^22 bins at 8 GPUs.  PROXY_RUNS="k:m:fine ..."
              for r in ${PROXY_RUNS:-31:14:19 63:14:19 63:13:19 63:15:22}; do
                IFS=: read -r pk pm pf <<< "$r"
                DSKGPU_FINE_LOG2=$pf timeout 600 python bench.py --kmer-size $pk --genome 375000000 --coverage 30 --device-synth --minimizer-size $pm --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/proxy_k${pk}_m${pm}_f${pf}.json" 2> "$OUT/proxy_k${pk}_m${pm}_f${pf}.err"
                python - "$OUT/proxy_k${pk}_m${pm}_f${pf}.json" <<'PYEOF'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d = json.loads(line)
        print(sys.argv[1], '%.1f G/s %.1f ms' % (d['value'], d['ms_per_step']), {k: round(v, 1) for k, v in d['stage_ms'].items()},
              {k: d['engine'][k] for k in ('log2_bins', 'partitions', 'smem_splits', 'hash_groups', 'sampled_density')}, d['checks'])
PYEOF
                tail -2 "$OUT/proxy_k${pk}_m${pm}_f${pf}.err"
              done ;;
    pywide)   timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_cli_dropin.py -q -k "wide or unhandled" > "$OUT/pytest_wide.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_wide.log"; tail -30 "$OUT/pytest_wide.log" | cut -c1-400 ;;
    pyplug)   timeout 900 python -m pytest tests/test_cli_dropin.py -q -k "plugin" > "$OUT/pytest_plugin.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_plugin.log"; tail -40 "$OUT/pytest_plugin.log" | cut -c1-600 ;;
    clitime)  # wall-clock of the drop-in CLI on the C2 read set (file on tmpfs -> .h5 on tmpfs) with the host / device phase timeline
              python - <<'PYEOF'
from dsk_b200.synth import reads_fasta
buf, n, _ = reads_fasta(G=5_000_000, coverage=100, L=150, err=0.01, seed=42)
open("/dev/shm/c2.fa", "wb").write(buf[:n].tobytes())
PYEOF
              for i in 1 2 3; do s=$(date +%s.%N); DSKGPU_TRACE=1 host/_build/dsk_gpu -file /dev/shm/c2.fa -kmer-size 31 -abundance-min 2 -out /dev/shm/c2_gpu -histo 1 -verbose 0 > "$OUT/clitime_gpu_$i.log" 2>&1; e=$(date +%s.%N); echo "dsk_gpu run $i wall $(echo "$e - $s" | bc) s"; done
              s=$(date +%s.%N); oracle/_ref/bin/dsk -file /dev/shm/c2.fa -kmer-size 31 -abundance-min 2 -out /dev/shm/c2_ref -histo 1 -verbose 0 -out-tmp /dev/shm > "$OUT/clitime_ref.log" 2>&1; e=$(date +%s.%N); echo "reference dsk wall $(echo "$e - $s" | bc) s"
              cat "$OUT/clitime_gpu_3.log" | cut -c1-160
              rm -f /dev/shm/c2.fa /dev/shm/c2_gpu* /dev/shm/c2_ref* ;;
    pycli)    timeout 900 python -m pytest tests/test_cli_dropin.py tests/test_zz_cli_scanner_fallback.py -q > "$OUT/pytest_cli.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_cli.log"; tail -30 "$OUT/pytest_cli.log" | cut -c1-500 ;;
    sanitize2) # the code added after r03a: wide-span kernels (memcheck: shared-memory indexing of the 128-position halo) and the early
              # exit of a lost table pass (racecheck on the split tests)
              timeout 900 compute-sanitizer --tool memcheck --log-file "$OUT/sanitizer_memcheck_wide.log" python -m pytest tests/test_gpu_wide.py -x -q -k "c1_k64 or longreads250_k127-min1 or histo2d_k95 or scatter_paths and 100 or multi_rank and 80" > "$OUT/sanitizer_memcheck_wide.out" 2>&1
              echo "memcheck wide exit $?"; tail -3 "$OUT/sanitizer_memcheck_wide.log"; tail -2 "$OUT/sanitizer_memcheck_wide.out"
              timeout 900 compute-sanitizer --tool racecheck --log-file "$OUT/sanitizer_racecheck_split.log" python -m pytest tests/test_gpu_parity.py -x -q -k "tiny_smem_table_overflow_splits and (c1_k31 or c1_k63) and 64" > "$OUT/sanitizer_racecheck_split.out" 2>&1
              echo "racecheck split exit $?"; tail -3 "$OUT/sanitizer_racecheck_split.log"; tail -2 "$OUT/sanitizer_racecheck_split.out" ;;
    pyseq)    timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_cli_dropin.py -q -k "sequence_statistics or bank_statistics or c123 or c1234 or two_contexts or larger_than_one" > "$OUT/pytest_seq.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_seq.log"; tail -40 "$OUT/pytest_seq.log" | cut -c1-700 ;;
    benchwide) for kk in ${WIDE_KS:-95 127}; do timeout 600 python bench.py --kmer-size $kk --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/bench_k$kk.json" 2> "$OUT/bench_k$kk.err"; python - "$OUT/bench_k$kk.json" <<'PYEOF'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d = json.loads(line)
        print(sys.argv[1], '%.2f G/s %.1f ms' % (d['value'], d['ms_per_step']), {k: round(v, 1) for k, v in d['stage_ms'].items()}, d['engine'], d['checks'])
PYEOF
              tail -2 "$OUT/bench_k$kk.err"; done ;;
    pykff)    timeout 600 python -m pytest tests/test_cli_dropin.py -q -k "kff" > "$OUT/pytest_kff.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_kff.log"; tail -40 "$OUT/pytest_kff.log" | cut -c1-700 ;;
    smoke)    timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -2 "$OUT/smoke.log" ;;
    pymin)    timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "minimizer_sizes or tiny_smem or record_sub or fine_histogram or heavy" > "$OUT/pytest_min.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_min.log"; tail -6 "$OUT/pytest_min.log" ;;
    *)        echo "unknown step $step" ;;
  esac
done
