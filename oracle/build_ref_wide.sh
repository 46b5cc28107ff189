#!/usr/bin/env bash
# Builds the UNMODIFIED reference with KSIZE_LIST "32 64 96 128" (k up to 127) into oracle/_ref/wide/bin.
# Test infrastructure only: generates the golden vectors that pin the wide-span oracle (oracle/dsk_oracle.c built with
# -DORC_WIDE) for SURVEY.md 8(f)-4.  Separate from oracle/build_ref.sh so that the "32 64" build the host adapter links
# against stays untouched.  Reference sources are never copied; outputs land only under oracle/_ref/.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref/wide"
if [ -x "$OUT/bin/dsk" ] && [ -x "$OUT/bin/dsk2ascii" ]; then echo "oracle/_ref/wide already built"; exit 0; fi
if [ ! -d "$REF" ]; then echo "no reference tree at $REF"; exit 0; fi
B="${TMPDIR:-/tmp}/dsk_ref_build_wide"
mkdir -p "$B" "$OUT/bin"
cd "$B"
cmake -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release -DKSIZE_LIST="32 64 96 128" \
      -DCMAKE_CXX_FLAGS="-include cstdint" "$REF" > cmake.log 2>&1
make -j"$(nproc)" dsk dsk2ascii > make.log 2>&1
cp bin/dsk bin/dsk2ascii "$OUT/bin/"
strip "$OUT/bin/"* || true
echo "wide reference built into $OUT/bin"
