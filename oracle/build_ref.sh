#!/usr/bin/env bash
# Builds the UNMODIFIED reference (dsk + dsk2ascii + gatb-h5dump) out of tree from /root/reference
# into oracle/_ref/bin.  Test infrastructure only: the binaries are the parity checker and the
# CPU baseline ("cpu_baseline.kind": "reference"); nothing in the product links or runs them.
#
# Why cmake and not a short hand-written Makefile: the counting path of gatb-core cannot be
# compiled from "a few source files" -- it is welded to the vendored HDF5 1.10.5, whose build
# needs configure-generated headers (H5pubconf.h, H5Tinit.c via H5detect).  The reference's own
# CMake project is therefore driven as-is, with the two workarounds SURVEY.md 8(c) lists.
# Reference sources are never copied into this repo; outputs land only under oracle/_ref/.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref"
if [ -x "$OUT/bin/dsk" ] && [ -x "$OUT/bin/dsk2ascii" ]; then echo "oracle/_ref already built"; exit 0; fi
if [ ! -d "$REF" ]; then echo "no reference tree at $REF (GPU box uses prebuilt oracle/_ref)"; exit 0; fi
B="${TMPDIR:-/tmp}/dsk_ref_build"
mkdir -p "$B" "$OUT/bin"
cd "$B"
cmake -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release -DKSIZE_LIST="32 64" \
      -DCMAKE_CXX_FLAGS="-include cstdint" "$REF" > cmake.log 2>&1
make -j"$(nproc)" dsk dsk2ascii gatb-h5dump > make.log 2>&1
cp bin/dsk bin/dsk2ascii ext/gatb-core/bin/Release/gatb-h5dump "$OUT/bin/"
strip "$OUT/bin/"* || true
echo "reference built into $OUT/bin"
