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
    assert L.dskgpu_abi_version() == 4


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
    cfg.kmer_size = 128                                     # KSIZE_LIST "32 64 96 128": k <= 127
    h = C.c_void_p()
    assert L.dskgpu_create(C.byref(cfg), C.byref(h)) == -1
    assert b"unhandled kmer size 128" in L.dskgpu_last_error(None)


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


@pytest.mark.parametrize("k,m", [(31, 10), (63, 10), (15, 7), (12, 5), (32, 10), (21, 8), (5, 4), (31, 12), (31, 14), (63, 14), (21, 13), (47, 11), (63, 15), (31, 15), (16, 15)])
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
    # windows with an allowed m-mer: the reference's minimizer; windows without one: the reference's default (4^m - 1), which
    # the device replaces by (1 << 2m | smallest banned m-mer) to spread those k-mers over bins (unobservable in the results)
    banned = (mn >> np.uint32(2 * m)) != 0
    ref_like = np.where(banned, np.uint32((1 << (2 * m)) - 1), mn)
    assert (ref_like[ovalid] == omn[ovalid]).all()


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
    assert f(0, 31) == 10 and f(100_000_000, 31) == 10 and f(400_000_000, 31) == 11 and f(3_000_000_000, 31) == 12
    assert f(30_000_000, 63) == 10 and f(100_000_000, 63) == 12 and f(293_000_000, 63) == 14
    assert f(72_000_000_000, 31) == 14 and f(6_600_000_000, 63) == 14 and f(52_800_000_000, 63) == 15
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


def _plan(L, level, gh, lh, world, nb_counts=1, slots=15360, density=0.26, mode=0, forced=0):
    nb = 1 << level
    b2p = np.zeros(nb, np.uint32); pk = np.zeros(nb + world, np.uint64); pr = np.zeros(nb + world, np.uint64)
    P = L.dskgpu_selftest_plan(level, gh.ctypes.data, lh.ctypes.data, world, nb_counts, slots, density, mode, forced,
                               b2p.ctypes.data, pk.ctypes.data, pr.ctypes.data, pk.size)
    assert P > 0
    return int(P), b2p, pk[:P], pr[:P]


@pytest.mark.parametrize("world,level", [(1, 16), (2, 17), (4, 18), (8, 16), (5, 16)])
def test_partition_planner_invariants_and_rank_agreement(L, world, level):
    # what every rank derives from the all-reduced minimizer-bin histogram (plan_host, the sequential mirror of the device
    # planner in dsk_b200/csrc/plan.cuh): skewed bins (a few hot minimizers), ranks holding different shares of every bin
    rng = np.random.default_rng(level * 10 + world)
    nb = 1 << level
    km = (rng.pareto(1.3, nb) * 3000).astype(np.uint64) + rng.integers(0, 2000, nb).astype(np.uint64)
    km[rng.integers(0, nb, 5)] += np.uint64(2_000_000)                 # hot bins far beyond one shared-memory table
    km[rng.integers(0, nb, 50)] = 0                                    # and empty ones
    rec = (km // np.uint64(11)) + (km > 0).astype(np.uint64)
    share = rng.dirichlet(np.ones(world), nb)                          # every rank's share of every bin
    lrec = np.floor(share * rec[:, None]).astype(np.uint64)
    lrec[:, 0] += rec - lrec.sum(1)
    lkm = np.floor(share * km[:, None]).astype(np.uint64)
    lkm[:, 0] += km - lkm.sum(1)
    gh = np.concatenate([rec, km]).astype(np.uint64)
    slots, density = 15360, 0.26
    T = int(slots * 0.52 / density * 0.85)                             # the cut: 85 % of the target (make_plan_params)
    plans = []
    for r in range(world):
        lh = np.concatenate([lrec[:, r], lkm[:, r]]).astype(np.uint64)
        plans.append(_plan(L, level, gh, lh, world, slots=slots, density=density))
    P, b2p, pk, _ = plans[0]
    for Pr, b2, pkr, _ in plans[1:]:                                   # same plan on every rank
        assert Pr == P and (b2 == b2p).all() and (pkr == pk).all()
    assert b2p.max() < P and int(pk.sum()) == int(km.sum())
    assert (np.bincount(b2p, weights=km.astype(np.float64), minlength=P).astype(np.uint64) == pk).all()
    # the ranks' local records of a partition add up to the whole-job records of its bins
    tot_local = np.sum([p[3] for p in plans], axis=0)
    assert (tot_local == np.bincount(b2p, weights=rec.astype(np.float64), minlength=P).astype(np.uint64)).all()
    # the cut rule: a partition is the run of bins that START inside one window of T k-mers, so it holds less than T plus its
    # last bin (a bin cannot be split: the minimizer decides the partition), and no partition is empty of bins
    nbins_of = np.bincount(b2p, minlength=P)
    assert (nbins_of > 0).all()
    last_bin = np.zeros(P, np.uint64)
    last_bin[b2p] = km                                                 # bins are visited in increasing order: the last one wins
    assert (pk < np.uint64(T) + last_bin + np.uint64(1)).all()
    # bins of a partition are consecutive
    first = np.full(P, nb, np.int64); last = np.zeros(P, np.int64)
    np.minimum.at(first, b2p, np.arange(nb)); np.maximum.at(last, b2p, np.arange(nb))
    assert (last - first + 1 == nbins_of).all()
    # partitions beyond the reach of the shared-memory path (AUTO mode) are numbered after all the others, in bin order
    lim = int(slots * 0.75 / density * 16)                             # fit x 2^max_split0 (default 4: 16 record sub-passes)
    heavy = np.nonzero(pk > lim)[0]
    assert len(heavy) >= 5
    light = np.nonzero(pk <= lim)[0]
    assert heavy.min() > light.max()
    assert (np.diff(first[heavy]) > 0).all() and (np.diff(first[light]) > 0).all()


def test_partition_planner_forced_partition_count(L):
    nb = 1 << 16
    km = np.full(nb, 1000, np.uint64); rec = np.full(nb, 90, np.uint64)
    gh = np.concatenate([rec, km])
    P, b2p, pk, pr = _plan(L, 16, gh, gh, 1, mode=2, forced=64)      # hash mode, -nb-partitions style override
    assert P == 64 and (pk == 1024 * 1000).all() and (np.diff(b2p.astype(np.int64)) >= 0).all() and int(pr.sum()) == int(rec.sum())


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    # the counting path is CUDA-only: without libdskgpu.so the binding raises, it never falls back to anything
    monkeypatch.setattr(_lib, "SO", str(tmp_path / "libdskgpu.so"))
    monkeypatch.setattr(_lib, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_product_package_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under dsk_b200/ (nor the C++ host, nor the C ABI) may reference it
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for sub in ("dsk_b200", "host", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            if "_build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".sh")):
                    continue
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r'(^\s*(import|from)\s+oracle\b)|(#include\s+[<"][^>"]*oracle)|(-loracle)|((dlopen|CDLL)\([^)]*oracle)', txt, re.M):
                    offenders.append(os.path.join(sub, f))
    assert offenders == []
