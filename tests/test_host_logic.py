"""CPU-side checks of the product library (no GPU): the C-ABI exports, and the host copies of the device bit
logic (record scanner tables, minimizer values, super-k-mer pack/expand) against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from dsk_b200 import _lib
from dsk_b200.build import build
from util import read_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["read50x_ref10K_e001.fasta.gz", "longread.fasta", "readN.fasta", "shortread.fasta", "IUPAC.fasta",
         "multiline.fasta", "reads.fastq", "reads_plusname.fastq", "varlen.fasta", "lowcomplexity.fasta", "weird.fasta",
         "assembly.fasta"]


@pytest.fixture(scope="module")
def L():
    build()
    return _lib.lib()


def test_abi_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "dskgpu.h")).read()
    declared = set(re.findall(r"\b(dskgpu_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dskgpu_ctx", "dskgpu_config", "dskgpu_stats"}
    assert declared == set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.dskgpu_abi_version() == 1


def test_struct_sizes_match(L):
    # the ctypes mirrors must have the C layout: config_default must not scribble outside the struct
    cfg = _lib.Config()
    L.dskgpu_config_default(C.byref(cfg))
    assert cfg.kmer_size == 31 and cfg.minimizer_size == 10 and cfg.abundance_min[0] == 2
    assert cfg.abundance_max == 2**31 - 1 and cfg.world_size == 1 and cfg.solid_vec[15] == 1


def test_no_device_fails_loudly(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dsk_b200 import GpuCounter, DskGpuError
    with pytest.raises(DskGpuError) as e:
        GpuCounter()
    assert e.value.code == _lib.ERR_NODEVICE


def test_unhandled_kmer_size(L):
    cfg = _lib.Config()
    L.dskgpu_config_default(C.byref(cfg))
    cfg.kmer_size = 64
    h = C.c_void_p()
    assert L.dskgpu_create(C.byref(cfg), C.byref(h)) == -1
    assert b"unhandled kmer size 64" in L.dskgpu_last_error(None)


def encode(seq):
    a = np.frombuffer(seq, dtype=np.uint8)
    u = a & 0xDF
    ok = (u == ord("A")) | (u == ord("C")) | (u == ord("G")) | (u == ord("T"))
    return (((a >> 1) & 3) | np.where(ok, 0, 4)).astype(np.uint8)


def scan(L, data, fmt=0):
    arr = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(arr.size + 16, np.uint8)
    n = L.dskgpu_selftest_scan(arr.ctypes.data, arr.size, fmt, out.ctypes.data, out.size)
    assert n >= 0, "scanner self-check failed with %d" % n
    return out[:n]


def split_records(codes):
    """list of code arrays, one per record (separator = 8)"""
    idx = np.nonzero(codes == 8)[0]
    return idx, codes


@pytest.mark.parametrize("name", FILES)
def test_scanner_matches_reference_parser(L, name):
    data = read_input(name)
    nrec, nt, concat = oracle.parse_stats(data)
    codes = scan(L, data)
    seps = np.nonzero(codes == 8)[0]
    assert len(seps) == nrec
    assert len(codes) - len(seps) == nt
    # sequences in order: FASTA emits the separator before a record, FASTQ after it
    ref_seqs = concat.split(b"\n")[:-1]
    body = codes.tobytes().split(b"\x08")
    got = body[1:] if data.lstrip()[:1] != b"@" or name.endswith(".fasta") else body[:-1]
    if name.endswith(".fastq"):
        got = body[:-1]
    assert len(got) == len(ref_seqs)
    for g, r in zip(got, ref_seqs):
        assert g == encode(r).tobytes()


def test_scanner_lines_format(L):
    data = b"ACGT\nNNAC\n\nTTGA"
    codes = scan(L, data, fmt=3)
    assert codes.tobytes() == encode(b"ACGT").tobytes() + b"\x08" + encode(b"NNAC").tobytes() + b"\x08\x08" + encode(b"TTGA").tobytes()


def test_scanner_flags_fastq_in_fasta(L):
    arr = np.frombuffer(b">a\nACGT\n+\nIIII\n", dtype=np.uint8)
    out = np.zeros(64, np.uint8)
    assert L.dskgpu_selftest_scan(arr.ctypes.data, arr.size, 1, out.ctypes.data, out.size) < 0


@pytest.mark.parametrize("k,m", [(31, 10), (63, 10), (15, 7), (12, 5), (32, 10), (21, 8), (5, 4), (31, 12), (31, 14), (63, 14), (21, 13), (47, 11)])
def test_minimizer_function_matches_oracle(L, k, m):
    rng = np.random.default_rng(k * 100 + m)
    seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 600).tobytes())
    seq = seq[:200] + b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA" + seq[200:300] + b"N" + seq[300:]
    codes = encode(seq)
    n = len(seq) - k + 1
    mn = np.zeros(n, np.uint32); valid = np.zeros(n, np.uint8)
    assert L.dskgpu_selftest_minimizers(codes.ctypes.data, codes.size, k, m, mn.ctypes.data, valid.ctypes.data) == n
    _, _, ovalid, omn, _ = oracle.kmers_of(seq, k, m=m)
    assert (valid.astype(bool) == ovalid).all()
    assert (mn[ovalid] == omn[ovalid]).all()


@pytest.mark.parametrize("k", [11, 21, 31, 32, 33, 47, 63])
def test_superkmer_pack_expand_roundtrip(L, k):
    m = min(10, k - 1)
    rng = np.random.default_rng(k)
    seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 3000).tobytes())
    seq = seq[:1000] + b"A" * 300 + b"N" + seq[1000:2000] + b"ACGT" * 100 + seq[2000:]
    codes = encode(seq)
    words = 1 if k < 32 else 2
    cap = len(seq)
    out = np.zeros(cap * words, np.uint64)
    nrec = C.c_uint64()
    n = L.dskgpu_selftest_superkmers(codes.ctypes.data, codes.size, k, m, out.ctypes.data, cap, C.byref(nrec))
    lo, hi, valid, _, _ = oracle.kmers_of(seq, k)
    assert n == int(valid.sum())
    got = out[: n * words].reshape(n, words)
    if words == 1:
        assert (got[:, 0] == lo[valid]).all()
    else:
        assert (got[:, 0] == lo[valid]).all() and (got[:, 1] == hi[valid]).all()
    assert 0 < nrec.value <= n


# ---------------------------------------------------------------- -abundance-min auto: host-side cutoff heuristic
def test_compute_threshold_matches_reference_cutoffs():
    """the product's restatement of Histogram::compute_threshold on the histograms of real reference runs"""
    import json
    from dsk_b200.histogram import compute_threshold, auto_thresholds
    runs = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_runs_auto.json")))["runs"]
    seen = set()
    for t in runs:
        if t["solidity_kind"] not in ("sum", "min", "max") or len(t["files"]) > 1 and t["solidity_kind"] != "sum":
            continue                                  # the committed histogram is the histogram of the sum
        h = np.zeros(10001, np.uint64)
        for i, v in t["hist"].items():
            h[int(i)] = v
        assert compute_threshold(h, 3)[0] == t["cutoffs"][0], t["name"]
        seen.add(t["cutoffs"][0])
    assert len(seen) >= 4                             # 3 (the floor) and real valleys
    assert auto_thresholds([-1, 4, -1], [7, 9]) == [7, 4, 4]          # CountProcessorSolidity.hpp:45-66
    assert auto_thresholds([-1], [5]) == [5]
    with pytest.raises(ValueError):
        auto_thresholds([-1], [3, 3])


def test_compute_threshold_agrees_with_oracle_restatement_on_random_histograms():
    import oracle
    from dsk_b200.histogram import compute_threshold
    rng = np.random.default_rng(3)
    for trial in range(200):
        h = np.zeros(10001, np.uint64)
        n = int(rng.integers(3, 200))
        peak = int(rng.integers(2, n))
        x = np.arange(1, n + 1)
        shape = rng.integers(1, 10**6) * np.exp(-x / rng.uniform(0.5, 3)) + rng.integers(0, 10**5) * np.exp(-((x - peak) ** 2) / (2 * rng.uniform(1, 30) ** 2))
        h[1:n + 1] = (shape * rng.uniform(0.7, 1.3, n)).astype(np.uint64)
        a = compute_threshold(h, 3)
        b = oracle.histogram_threshold(h, 3)
        assert (a[0], a[1]) == b, trial


def test_suggested_minimizer_size_grows_with_the_job(L):
    # the role of ConfigurationAlgorithm's volume-driven sizing: 10 (reference default) for small jobs, longer for big ones
    f = L.dskgpu_suggest_minimizer_size
    assert f(0, 31) == 10 and f(400_000_000, 31) == 10 and f(800_000_000, 31) == 12 and f(3_000_000_000, 31) == 12
    assert f(72_000_000_000, 31) == 14 and f(72_000_000_000, 63) == 14
    assert f(72_000_000_000, 11) == 10 and f(10, 5) == 4                      # clipped to k-1 (ConfigurationAlgorithm.cpp:249-251)


@pytest.mark.parametrize("k", [5, 31, 32, 33, 63, 64, 65, 71, 95, 96, 97, 111, 127])
def test_wide_kmer_logic_matches_the_wide_oracle(L, k):
    # dsk_b200/csrc/kmer_wide.cuh (N-word roll / revcomp / extraction from a packed record), host-compiled, against the
    # oracle (wide build for k >= 64, pinned against the reference built with KSIZE_LIST "32 64 96 128")
    rng = np.random.default_rng(1000 + k)
    seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 700).tobytes())
    seq = seq[:300] + b"N" + seq[301:500] + b"GGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG" + seq[500:]
    codes = encode(seq)
    n = len(seq) - k + 1
    words = np.zeros((n, 4), np.uint64); valid = np.zeros(n, np.uint8)
    got = L.dskgpu_selftest_wide_kmers(codes.ctypes.data, codes.size, k, words.ctypes.data, valid.ctypes.data)
    assert got == n, "rolling and record extraction disagree at window %d" % (-1 - got)
    ow, ovalid, _, _ = oracle.kmers_of_words(seq, k)
    assert (valid.astype(bool) == ovalid).all()
    assert (words[ovalid] == ow[ovalid]).all()
