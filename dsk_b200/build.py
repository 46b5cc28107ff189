"""Builds dsk_b200/libdskgpu.so (hand-written CUDA for sm_100a behind the C ABI of include/dskgpu.h)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SO = os.path.join(PKG, "libdskgpu.so")
SRC = os.path.join(PKG, "csrc", "dskgpu.cu")
DEPS = [os.path.join(PKG, "csrc", f) for f in ("dskgpu.cu", "kmer_bits.cuh", "scan.cuh", "superk.cuh", "count.cuh", "count_smem.cuh", "radix.cuh", "kmer_wide.cuh", "plan.cuh", "seqstats.cuh")]
DEPS.append(os.path.join(ROOT, "include", "dskgpu.h"))

NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("DSKGPU_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdskgpu.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
