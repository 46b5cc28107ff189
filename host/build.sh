#!/usr/bin/env bash
# Builds host/_build/dsk_gpu: the reference `dsk` command line with the counting class replaced by the device path.
# Needs the reference tree (headers) and the out-of-tree reference build that oracle/build_ref.sh makes
# (libgatbcore.a, libhdf5.a, generated config headers).  The binary is static against those and dynamic against
# dsk_b200/libdskgpu.so (rpath $ORIGIN/../../dsk_b200), so it travels to the GPU box with the snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
OUT="$HERE/_build"
if [ ! -d "$REF" ]; then echo "no reference tree at $REF (GPU box uses the prebuilt host/_build/dsk_gpu)"; exit 0; fi
# gatb-core to link against: the build with the reference's default KSIZE_LIST "32 64 96 128" (oracle/build_ref_wide.sh; k <= 127
# through GpuSortingCount<96> / <128>) when it can be had, else the "32 64" build of oracle/build_ref.sh (k <= 63)
BW="${TMPDIR:-/tmp}/dsk_ref_build_wide"
BN="${TMPDIR:-/tmp}/dsk_ref_build"
if [ -n "${REF_BUILD:-}" ]; then B="$REF_BUILD"
else
  if [ ! -f "$BW/ext/gatb-core/lib/Release/libgatbcore.a" ] && [ "${DSKGPU_HOST_NARROW:-0}" != "1" ]; then
    rm -rf "$ROOT/oracle/_ref/wide/bin/dsk"; "$ROOT/oracle/build_ref_wide.sh" "$REF" || echo "wide reference build failed: falling back to KSIZE_LIST '32 64'"
  fi
  if [ -f "$BW/ext/gatb-core/lib/Release/libgatbcore.a" ] && [ "${DSKGPU_HOST_NARROW:-0}" != "1" ]; then B="$BW"; else B="$BN"; fi
fi
if [ ! -f "$B/ext/gatb-core/lib/Release/libgatbcore.a" ]; then
  rm -rf "$ROOT/oracle/_ref/bin/dsk"; "$ROOT/oracle/build_ref.sh" "$REF"
fi
G="$REF/thirdparty/gatb-core/gatb-core"
mkdir -p "$OUT"
if [ -x "$OUT/dsk_gpu" ] && [ -x "$OUT/plugin_check" ] && [ "$OUT/dsk_gpu" -nt "$HERE/GpuSortingCount.hpp" ] && [ "$OUT/dsk_gpu" -nt "$HERE/dsk_gpu_main.cpp" ] \
   && [ "$OUT/plugin_check" -nt "$HERE/GpuSortingCount.hpp" ] && [ "$OUT/plugin_check" -nt "$HERE/plugin_check.cpp" ] \
   && [ "$OUT/dsk_gpu" -nt "$ROOT/include/dskgpu.h" ] && [ -f "$OUT/.built_against" ] && [ "$(cat "$OUT/.built_against")" = "$B" ]; then echo "host/_build/dsk_gpu up to date"; exit 0; fi
# dsk_gpu = the `dsk` command line over the device path; plugin_check = the 5-argument constructor + ICountProcessor surface
# driven the way Graph.cpp drives SortingCountAlgorithm (GPU tests)
for prog in dsk_gpu:dsk_gpu_main.cpp plugin_check:plugin_check.cpp; do
  exe="${prog%%:*}"; src="${prog##*:}"
  g++ -std=c++11 -O2 -DNDEBUG -D_FILE_OFFSET_BITS=64 -D_GNU_SOURCE -D_LARGEFILE64_SOURCE -D_LARGEFILE_SOURCE -DINT128_FOUND \
      -include cstdint -Wno-invalid-offsetof -Wno-format -Wno-unknown-pragmas \
      -I"$B/ext/gatb-core/include" -I"$B/ext/gatb-core/include/Release" -I"$G/src" -I"$G/thirdparty" \
      -I"$B/ext/gatb-core/thirdparty/hdf5/src" -I"$G/thirdparty/hdf5/src" \
      "$HERE/$src" -o "$OUT/$exe" \
      -L"$B/ext/gatb-core/lib/Release" -lgatbcore -lhdf5 -L"$ROOT/dsk_b200" -ldskgpu \
      -Wl,-rpath,'$ORIGIN/../../dsk_b200' -ldl -lpthread -lz &
done
wait
[ -x "$OUT/dsk_gpu" ] && [ -x "$OUT/plugin_check" ] || { echo "host build failed"; exit 1; }
echo "$B" > "$OUT/.built_against"
echo "built $OUT/dsk_gpu and $OUT/plugin_check (against $B)"
