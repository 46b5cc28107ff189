"""Parity tests (-m gpu) of the wide spans on the device (KSIZE_LIST 96 / 128, 64 <= k <= 127: 192- and
256-bit keys, super-k-mer records of 6 / 8 words, counted through the sort path).  Bit-exact against
  * the outputs of the unmodified reference built with KSIZE_LIST "32 64 96 128" (tests/golden/ref_runs_wide.json), and
  * the wide build of the oracle (oracle/liboracle_wide.so, itself pinned against those runs by tests/test_oracle_wide.py)
on seeded synthetic reads: same multiset of (canonical k-mer, abundance), same histogram(s), ascending order as delivered
(most significant word first, LargeInt.hpp:502-509)."""
import numpy as np
import pytest

import oracle
from dsk_b200 import SortingCountAlgorithm, BankBytes, BankAlbum, GpuCounter
from dsk_b200.counter import multi_finish
from dsk_b200.synth import reads_fasta
from util import load_json, read_input, digest_words, sparse_hist, sparse_hist2d

pytestmark = pytest.mark.gpu

WIDE = load_json("ref_runs_wide.json")["runs"]


def words_of(k):
    return 1 if k < 32 else 2 if k < 64 else 3 if k < 96 else 4


def as_int(keys):
    v = keys[:, 0].astype(object)
    for j in range(1, keys.shape[1]):
        v = v | (keys[:, j].astype(object) << (64 * j))
    return v


def run_gpu(files, k, abundance_min=2, histo2d=False, kind=None, **engine):
    banks = BankAlbum([BankBytes(read_input(f)) for f in files]) if len(files) > 1 else BankBytes(read_input(files[0]))
    props = {"-kmer-size": k, "-abundance-min": str(abundance_min), "-histo2D": int(histo2d)}
    if kind:
        props["-solidity-kind"] = kind
    return SortingCountAlgorithm(banks, props, **engine).execute()


def check_against_run(sc, t):
    info = sc.getInfo()
    assert info["kmers_nb_valid"] == t["kmers_nb_valid"]
    assert info["kmers_nb_distinct"] == t["kmers_nb_distinct"]
    assert info["kmers_nb_solid"] == t["nb_solid"]
    keys, cnt = sc.getSolidCounts()
    assert keys.shape[1] == words_of(t["k"])
    h1, h2 = sc.getHistogram()
    assert sparse_hist(h1) == t["hist"]
    dg, pairs = digest_words(keys, cnt, t["k"])
    assert [list(p) for p in pairs[:3]] == t["first_kmers"]
    assert dg == t["kmers_sha256"]
    assert int(cnt.astype(np.int64).sum()) == t["sum_counts"]
    if t["histo2d"]:
        assert sparse_hist2d(h2) == t["hist2d"]
    v = as_int(keys)
    assert all(v[i] < v[i + 1] for i in range(len(v) - 1))          # ascending as delivered


@pytest.mark.parametrize("t", WIDE, ids=[t["name"] for t in WIDE])
def test_wide_reference_runs(t):
    sc = run_gpu(t["files"], t["k"], t["abundance_min"], t["histo2d"], t.get("solidity_kind"))
    check_against_run(sc, t)
    st = sc.getInfo()["engine"]
    assert st["nb_parts_smem"] == 0 and st["nb_groups_hash"] == 0       # the sort path, whatever count_mode says
    if t["kmers_nb_valid"]:
        assert st["nb_groups_sort"] > 0


def assert_equals_wide_oracle(keys, cnts, hist, ref, k):
    order = np.lexsort(tuple(keys[:, j] for j in range(keys.shape[1])))       # last key = most significant word
    keys, cnts = keys[order], cnts[order]
    words, sums = ref.solid_kmer_words()
    W = words_of(k)
    assert keys.shape[1] == W and len(cnts) == len(sums)
    assert (keys == words[:, :W]).all() and not words[:, W:].any()
    assert (cnts.astype(np.int64) == sums).all()
    assert (hist == ref.hist).all()


@pytest.mark.parametrize("k", [64, 71, 95, 96, 110, 127])
def test_wide_synthetic_vs_oracle(k):
    buf, n, _ = reads_fasta(G=200_000, coverage=30, L=250, err=0.005, seed=640 + k)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": k, "-abundance-min": "2"}).execute()
    keys, cnt = sc.getSolidCounts()
    assert sc.getInfo()["kmers_nb_valid"] == ref.kmers_nb_valid and sc.getInfo()["kmers_nb_distinct"] == ref.nb_distinct
    assert_equals_wide_oracle(keys, cnt, sc.getHistogram()[0], ref, k)
    v = as_int(keys)
    assert all(v[i] < v[i + 1] for i in range(len(v) - 1))


@pytest.mark.parametrize("k,m", [(80, 8), (80, 15), (127, 12)])
def test_wide_minimizer_sizes_and_chunked_push(k, m):
    """4 KiB push granularity: every chunk boundary cuts a read, the carry holds k - 1 (up to 126) codes; invalid bases (N)
    inside reads break the 128-flag validity window in both of its halves"""
    buf, n, _ = reads_fasta(G=100_000, coverage=20, L=300, err=0.01, seed=700 + k + m)
    data = bytearray(buf[:n].tobytes())
    rng = np.random.default_rng(k)
    for pos in rng.integers(0, len(data), 400):                             # sprinkle Ns over sequence lines
        if data[pos] in b"ACGT":
            data[pos] = ord("N")
    data = bytes(data)
    ref = oracle.count_files([data], k, abundance_min=2)
    sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": k, "-abundance-min": "2", "-minimizer-size": m}, push_chunk_bytes=4096).execute()
    keys, cnt = sc.getSolidCounts()
    assert sc.getInfo()["kmers_nb_valid"] == ref.kmers_nb_valid and sc.getInfo()["kmers_nb_distinct"] == ref.nb_distinct
    assert_equals_wide_oracle(keys, cnt, sc.getHistogram()[0], ref, k)


@pytest.mark.parametrize("k,min_parts,nparts", [(80, 1, 0), (100, 1, 3000), (127, 10**9, 500)])
def test_wide_partition_scatter_paths(k, min_parts, nparts, monkeypatch):
    """records of 6 / 8 words through both scatter kernels: the MSD multi-split passes (forced) and the per-record cursor"""
    monkeypatch.setenv("DSKGPU_MSD_MIN_PARTS", str(min_parts))
    buf, n, _ = reads_fasta(G=200_000, coverage=25, L=250, err=0.01, seed=800 + k)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    with GpuCounter(kmer_size=k, abundance_min=2, nb_partitions=nparts) as eng:
        eng.push_bytes(data)
        eng.finish()
        st = eng.stats()
        if nparts:
            assert st["nb_partitions"] >= nparts // 2
        kk, cc = eng.solid()
        assert st["kmers_nb_distinct"] == ref.nb_distinct
        assert_equals_wide_oracle(kk, cc, eng.histogram()[0], ref, k)


def split_records(data, parts):
    cuts = [0]
    for i in range(1, parts):
        j = data.find(b"\n>", len(data) * i // parts)
        cuts.append(len(data) if j < 0 else j + 1)
    cuts.append(len(data))
    return [data[cuts[i]:cuts[i + 1]] for i in range(parts)]


@pytest.mark.parametrize("W,k", [(2, 80), (3, 127)])
def test_wide_multi_rank_exchange_in_process(W, k):
    """several contexts on one GPU through dskgpu_multi_finish: wide records cross the (peer) exchange, every rank counts the
    partitions it owns by the sort path; the union of the ranks equals the oracle"""
    buf, n, _ = reads_fasta(G=200_000, coverage=30, L=250, err=0.01, seed=900 + k)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    engines = [GpuCounter(kmer_size=k, abundance_min=2, rank=r, world_size=W, device=0) for r in range(W)]
    try:
        for e, piece in zip(engines, split_records(data, W)):
            e.push_bytes(piece)
        multi_finish(engines)
        keys, cnts, hist, valid, distinct = [], [], np.zeros(10001, np.uint64), 0, 0
        for e in engines:
            kk, cc = e.solid()
            keys.append(kk); cnts.append(cc)
            hist += e.histogram()[0]
            st = e.stats()
            valid += st["kmers_nb_valid"]; distinct += st["kmers_nb_distinct"]
        assert valid == ref.kmers_nb_valid and distinct == ref.nb_distinct
        assert_equals_wide_oracle(np.concatenate(keys), np.concatenate(cnts), hist, ref, k)
    finally:
        for e in engines:
            e.close()


def test_wide_pass_loop_and_auto_cutoff():
    """nb_passes = 3 at k = 90 (union of the passes), and -abundance-min auto at k = 100 (histogram pass, then a recount from HBM)"""
    buf, n, _ = reads_fasta(G=150_000, coverage=30, L=250, err=0.01, seed=77)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], 90, abundance_min=2)
    sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": 90, "-abundance-min": "2"}, nb_passes=3).execute()
    keys, cnt = sc.getSolidCounts()
    assert sc.getInfo()["kmers_nb_distinct"] == ref.nb_distinct
    assert_equals_wide_oracle(keys, cnt, sc.getHistogram()[0], ref, 90)
    ref = oracle.count_files([data], 100, abundance_min=-1)
    sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": 100, "-abundance-min": "auto"}).execute()
    assert sc.getInfo()["cutoffs_auto"] == ref.cutoffs
    keys, cnt = sc.getSolidCounts()
    assert_equals_wide_oracle(keys, cnt, sc.getHistogram()[0], ref, 100)


@pytest.mark.parametrize("k,bits", [(80, 12), (127, 20), (100, 64)])
def test_wide_hash_ordering_collisions_fall_back_to_the_full_sort(k, bits, monkeypatch):
    """the wide spans bring equal k-mers together by sorting a 64-bit hash of the key; when two different k-mers share a hash
    (forced here by narrowing the hash) the verification pass must notice and the group must be redone by the exact
    full-width sort; with the full hash the fast path runs.  Same results either way, and with DSKGPU_WIDE_FULLSORT=1."""
    buf, n, _ = reads_fasta(G=150_000, coverage=25, L=250, err=0.01, seed=1200 + k)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    monkeypatch.setenv("DSKGPU_TEST_HASH_BITS", str(bits))
    with GpuCounter(kmer_size=k, abundance_min=2) as eng:
        eng.push_bytes(data)
        eng.finish()
        st = eng.stats()
        assert (st["sort_fallbacks"] > 0) == (bits < 64)
        kk, cc = eng.solid()
        assert st["kmers_nb_distinct"] == ref.nb_distinct
        assert_equals_wide_oracle(kk, cc, eng.histogram()[0], ref, k)
    monkeypatch.delenv("DSKGPU_TEST_HASH_BITS")
    monkeypatch.setenv("DSKGPU_WIDE_FULLSORT", "1")
    with GpuCounter(kmer_size=k, abundance_min=2) as eng:
        eng.push_bytes(data)
        eng.finish()
        kk, cc = eng.solid()
        assert eng.stats()["sort_fallbacks"] == 0
        assert_equals_wide_oracle(kk, cc, eng.histogram()[0], ref, k)
