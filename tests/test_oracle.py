"""Pins the CPU oracle (oracle/dsk_oracle.c) against every golden vector the reference holds for the
counting path and against outputs of the real reference binary (ref_runs.json)."""
import numpy as np
import pytest

import oracle
from util import (load_json, read_input, digest, sparse_hist, sparse_hist2d, kmer_to_str, str_to_kmer)

SHELL = load_json("ref_shell_tests.json")["tests"]
UNIT = load_json("ref_unit_vectors.json")
RUNS = load_json("ref_runs.json")["runs"]


def run_oracle(files, k, **kw):
    banks = [read_input(f) for f in files]
    return oracle.count_files(banks, k, **kw)


# ---------------------------------------------------------------- R/scripts/simple_test.sh
@pytest.mark.parametrize("t", SHELL, ids=[t["name"] for t in SHELL])
def test_shell_goldens(t):
    r = run_oracle(t["files"], t["k"], abundance_min=t["abundance_min"])
    if "hist" in t:
        assert sparse_hist(r.hist) == t["hist"]
    if "dsk2ascii" in t:
        lo, hi, cnt = r.solid_kmers()
        txt = "".join("%s %d\n" % (kmer_to_str(a, b, t["k"]), c) for a, b, c in zip(lo, hi, cnt))
        assert txt == t["dsk2ascii"]


# ---------------------------------------------------------------- TestDSK.cpp literals
def test_dsk_check1():
    d = UNIT["DSK_check1"]
    assert len(d["checks"]) == 33
    for c in d["checks"]:
        o = oracle.Oracle(c["k"])
        for s in d["seqsets"][c["seqs"]]:
            o.add_sequence(s)
        r = o.finish(abundance_min=c["nks"])
        assert r.nb_solid == c["nb_solid"], c


def test_dsk_check2_values_and_checksum():
    d = UNIT["DSK_check2"]
    o = oracle.Oracle(d["k"])
    o.add_sequence(d["seq"])
    r = o.finish(abundance_min=1)
    lits = {int(x, 16) for x in d["hex_literals"]}
    checksum = 0x8b0c176c3b43d207
    assert checksum in lits
    vals = {int(v) for v in r.keys_lo}
    assert vals == lits - {checksum}
    assert sum(vals) % (1 << 64) == checksum
    assert (r.keys_hi == 0).all() and (r.sums == 1).all()


@pytest.mark.parametrize("name", ["DSK_perBank1", "DSK_perBank2"])
def test_dsk_perbank(name):
    d = UNIT[name]
    assert len(d["checks"]) in (9, 45)
    for c in d["checks"]:
        o = oracle.Oracle(c["k"], nbanks=len(d["banks"]))
        for b, s in enumerate(d["banks"]):
            o.add_sequence(s, bank=b)
        r = o.finish(abundance_min=c["min"], abundance_max=c["max"], kind=c["kind"])
        assert r.nb_solid == c["nb_solid"], c


# ---------------------------------------------------------------- TestKmer.cpp literals
def test_kmer3_direct_and_canonical():
    d = UNIT["kmer3"]
    lo, _, valid, _, _ = oracle.kmers_of(d["seq"], d["k"], forward=True)
    assert list(map(int, lo)) == d["direct"] and valid.all()
    lo, _, _, _, _ = oracle.kmers_of(d["seq"], d["k"])
    assert list(map(int, lo)) == d["canonical"]


def test_kmer5_canonical():
    d = UNIT["kmer5"]
    lo, _, _, _, _ = oracle.kmers_of(d["seq"], d["k"])
    assert list(map(int, lo)) == d["canonical"]


def test_minimizer_canonical_table():
    d = UNIT["minimizer3"]
    lo, _, valid, mn, _ = oracle.kmers_of(d["seq"], d["k"], m=d["m"])
    assert len(d["rows"]) == len(lo) == 18
    for i, row in enumerate(d["rows"]):
        assert kmer_to_str(lo[i], 0, d["k"]) == row["kmer"]
        assert kmer_to_str(mn[i], 0, d["m"]) == row["minimizer"]
        changed = i == 0 or mn[i] != mn[i - 1]
        assert changed == row["changed"]


def test_minimizer_bruteforce_rule():
    # TestKmer.cpp:286-304 : min over m-mers whose (m-1)-suffix has no "AA", default 4^m-1
    rng = np.random.default_rng(7)
    for k, m in ((12, 5), (31, 10), (63, 10), (27, 8), (15, 7)):
        seq = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
        lo, hi, _, mn, _ = oracle.kmers_of(seq, k, m=m, forward=True)
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        for i in range(len(lo)):
            s = seq[i:i + k]
            best = 4 ** m - 1
            for j in range(k - m + 1):
                mm = s[j:j + m]
                rc = "".join(comp[c] for c in reversed(mm))
                cand = min(str_to_kmer(mm), str_to_kmer(rc))
                cs = kmer_to_str(cand, 0, m)
                if "AA" in cs[1:]:
                    continue
                best = min(best, cand)
            assert best == int(mn[i]), (k, m, i)


def test_badchar_validity():
    d = UNIT["badchar"]
    lo, _, valid, _, _ = oracle.kmers_of(d["seq"], d["k"], forward=True)
    assert len(d["rows"]) == len(lo) == 32
    for i, row in enumerate(d["rows"]):
        assert bool(valid[i]) == row["valid"]
        # N encodes as G (TestKmer.cpp:533-536)
        assert kmer_to_str(lo[i], 0, d["k"]) == row["kmer"].replace("N", "G")


# ---------------------------------------------------------------- real reference runs
@pytest.mark.parametrize("t", RUNS, ids=[t["name"] for t in RUNS])
def test_reference_runs(t):
    kw = dict(abundance_min=t["abundance_min"], histo2d=t["histo2d"])
    if "abundance_max" in t:
        kw["abundance_max"] = t["abundance_max"]
    if "solidity_kind" in t:
        kw["kind"] = t["solidity_kind"]
    r = run_oracle(t["files"], t["k"], **kw)
    assert r.kmers_nb_valid == t["kmers_nb_valid"]
    assert r.nb_distinct == t["kmers_nb_distinct"]
    assert r.nb_solid == t["nb_solid"] == t["kmers_nb_solid"]
    assert sparse_hist(r.hist) == t["hist"]
    lo, hi, cnt = r.solid_kmers()
    dg, pairs = digest(lo, hi, cnt, t["k"])
    assert [list(p) for p in pairs[:3]] == t["first_kmers"]
    assert dg == t["kmers_sha256"]
    if t["histo2d"]:
        assert sparse_hist2d(r.hist2d) == t["hist2d"]


def test_survey_md5_anchor():
    """md5 of `dsk2ascii | LC_ALL=C sort` for config C1 at k=31 and k=63."""
    import hashlib
    for k, md5 in ((31, "9905e88981187c2c47ac784674297fa7"), (63, "5a3e08caab679e2bdb407c0d478fbf73")):
        r = run_oracle(["read50x_ref10K_e001.fasta.gz"], k, abundance_min=2)
        lo, hi, cnt = r.solid_kmers()
        _, pairs = digest(lo, hi, cnt, k)
        txt = "".join("%s %d\n" % p for p in pairs)
        assert hashlib.md5(txt.encode()).hexdigest() == md5


# ---------------------------------------------------------------- -abundance-min auto (two-pass cutoff chain)
AUTO = load_json("ref_runs_auto.json")["runs"]


def _amin_list(s):
    return [-1 if x == "auto" else int(x) for x in str(s).split(",")]


@pytest.mark.parametrize("t", AUTO, ids=[t["name"] for t in AUTO])
def test_reference_auto_cutoff_runs(t):
    r = run_oracle(t["files"], t["k"], abundance_min=_amin_list(t["abundance_min"]), kind=t["solidity_kind"])
    assert r.cutoffs == t["cutoffs"]
    assert r.nb_solid == t["nb_solid"] == t["kmers_nb_solid"]
    assert sparse_hist(r.hist) == t["hist"]
    lo, hi, cnt = r.solid_kmers()
    dg, _ = digest(lo, hi, cnt, t["k"])
    assert dg == t["kmers_sha256"]


BANKSTATS = load_json("ref_bankstats.json")["runs"]


@pytest.mark.parametrize("t", BANKSTATS, ids=["+".join(t["files"]) for t in BANKSTATS])
def test_oracle_bank_statistics_against_the_reference(t):
    """BankStats (K/BankKmers.hpp:166-215) + kmersNbInvalid (K/Sequence2SuperKmer.hpp:95-108) as `dsk -verbose 1` prints them
    (tests/golden/ref_bankstats.json, from the unmodified reference): pins the oracle's record parser on sequence boundaries
    (empty records, multi-line FASTA, FASTQ, several banks) and its validity window on the invalid-window count"""
    res = oracle.count_files([read_input(f) for f in t["files"]], t["k"], abundance_min=2)
    got = res.bank_stats_strings()
    for key, want in t.items():
        if key in got:
            assert got[key] == want, (key, got[key], want)


HISTOMAX = load_json("ref_runs_histomax.json")["runs"]


@pytest.mark.parametrize("t", HISTOMAX, ids=[t["name"] for t in HISTOMAX])
def test_histo_max_rule_against_the_reference(t):
    """-histo-max N: the <out>.histo of the reference is the first N bins of the 10000-bin histogram with bin N empty (counts >= N
    are clamped into a bin that is never merged: Histogram.hpp:92,221); the 2-D clamp column collects everything >= N
    (Histogram.hpp:97).  This is the rule host/GpuSortingCount.hpp applies to the device bins."""
    amin = -1 if t["abundance_min"] == "auto" else t["abundance_min"]
    res = oracle.count_files([read_input(f) for f in t["files"]], t["k"], abundance_min=amin, histo2d=bool(t.get("histo2d")))
    N = t["histo_max"]
    want = "".join("%d\t%d\n" % (i, int(res.hist[i]) if i < N else 0) for i in range(1, N + 1))
    assert want == t["histo_text"]
    assert res.nb_solid == t["nb_solid"]
    if t.get("histo2d"):
        rows = t["histo2d_text"].splitlines()
        assert len(rows) == N + 1
        for i, row in enumerate(rows):
            vals = [int(x) for x in row.split(":")[1].split()]
            for j in range(11):
                exp = int(res.hist2d[j, i]) if i < N else int(res.hist2d[j, N:].sum())
                assert vals[j] == exp, (i, j)
