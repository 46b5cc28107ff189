"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against
  * the committed outputs of the real reference binary (tests/golden/ref_runs.json, ref_shell_tests.json),
  * the reference's unit-test literals (ref_unit_vectors.json),
  * the CPU oracle on seeded synthetic inputs,
  * size-independent properties at BASELINE.json's full single-GPU size.
Bit-exact everywhere (integer work): same multiset of (canonical k-mer, abundance), same histogram."""
import numpy as np
import pytest

import oracle
from dsk_b200 import SortingCountAlgorithm, BankBytes, BankStrings, BankAlbum, GpuCounter
from dsk_b200.synth import reads_fasta, genome_codes, assembly_fasta
from util import load_json, read_input, digest, sparse_hist, sparse_hist2d, kmer_to_str

pytestmark = pytest.mark.gpu

SHELL = load_json("ref_shell_tests.json")["tests"]
UNIT = load_json("ref_unit_vectors.json")
RUNS = load_json("ref_runs.json")["runs"]


def run_gpu(files, k, abundance_min=2, histo2d=False, kind=None, abundance_max=None, **engine):
    banks = BankAlbum([BankBytes(read_input(f)) for f in files]) if len(files) > 1 else BankBytes(read_input(files[0]))
    props = {"-kmer-size": k, "-abundance-min": str(abundance_min), "-histo2D": int(histo2d)}
    if kind:
        props["-solidity-kind"] = kind
    if abundance_max is not None:
        props["-abundance-max"] = abundance_max
    return SortingCountAlgorithm(banks, props, **engine).execute()


def check_against_run(sc, t):
    info = sc.getInfo()
    assert info["kmers_nb_valid"] == t["kmers_nb_valid"]
    assert info["kmers_nb_distinct"] == t["kmers_nb_distinct"]
    assert info["kmers_nb_solid"] == t["nb_solid"]
    keys, cnt = sc.getSolidCounts()
    h1, h2 = sc.getHistogram()
    assert sparse_hist(h1) == t["hist"]
    hi = keys[:, 1] if keys.shape[1] == 2 else np.zeros(len(keys), np.uint64)
    dg, pairs = digest(keys[:, 0], hi, cnt, t["k"])
    assert [list(p) for p in pairs[:3]] == t["first_kmers"]
    assert dg == t["kmers_sha256"]
    if t["histo2d"]:
        assert sparse_hist2d(h2) == t["hist2d"]
    # ascending order inside the partition, like the reference's dump
    v = keys[:, 0].astype(object) if keys.shape[1] == 1 else (keys[:, 1].astype(object) << 64) | keys[:, 0].astype(object)
    assert all(v[i] < v[i + 1] for i in range(len(v) - 1))


@pytest.mark.parametrize("t", RUNS, ids=[t["name"] for t in RUNS])
def test_reference_runs(t):
    sc = run_gpu(t["files"], t["k"], t["abundance_min"], t["histo2d"], t.get("solidity_kind"), t.get("abundance_max"))
    check_against_run(sc, t)


SUBSET = [t for t in RUNS if t["name"] in ("c1_k31", "c1_k63", "lowcomplexity.fasta_k31", "lowcomplexity.fasta_k63", "reads.fastq_k31",
                                           "multiline.fasta_k63", "histo2d_k31", "c123_k31_all", "c1_k21", "c1_k32")]


@pytest.mark.parametrize("mode", ["sort", "hash", "smem"])
@pytest.mark.parametrize("t", SUBSET, ids=[t["name"] for t in SUBSET])
def test_forced_count_modes(t, mode):
    sc = run_gpu(t["files"], t["k"], t["abundance_min"], t["histo2d"], t.get("solidity_kind"), t.get("abundance_max"), count_mode=mode)
    check_against_run(sc, t)
    st = sc.getInfo()["engine"]
    if mode == "smem":
        # shared-memory tables everywhere except the partitions too big for a few split passes (lowcomplexity: one hot minimizer);
        # per-bank counts (-histo2D, solidity kinds over several banks) use the multi-count variant of the same kernel
        assert st["nb_parts_smem"] > 0 and st["nb_groups_sort"] == 0
    else:
        assert (st["nb_groups_sort"] > 0) == (mode == "sort") and (st["nb_groups_hash"] > 0) == (mode == "hash") and st["nb_parts_smem"] == 0


@pytest.mark.parametrize("t", SUBSET, ids=[t["name"] for t in SUBSET])
def test_small_table_many_partitions_and_chunked_push(t):
    # tiny global hash table => many partitions/groups and the occupancy picker sending big partitions to the sort path;
    # 4 KiB push granularity => records straddle every chunk boundary
    sc = run_gpu(t["files"], t["k"], t["abundance_min"], t["histo2d"], t.get("solidity_kind"), t.get("abundance_max"),
                 count_mode="hash", hash_log2_slots=10, push_chunk_bytes=4096)
    check_against_run(sc, t)
    assert sc.getInfo()["engine"]["nb_partitions"] > 1


@pytest.mark.parametrize("slots", [64, 256, 2048])
@pytest.mark.parametrize("t", SUBSET, ids=[t["name"] for t in SUBSET])
def test_tiny_smem_table_overflow_splits(t, slots):
    # a shared-memory table far too small for its partitions: the kernel must answer every overflow by splitting the pass
    # (recursively) and still produce the reference's counts; partitions beyond 16 x slots k-mers go to the global paths
    sc = run_gpu(t["files"], t["k"], t["abundance_min"], t["histo2d"], t.get("solidity_kind"), t.get("abundance_max"),
                 count_mode="smem", smem_table_slots=slots, nb_partitions=1000 if slots == 64 else 0)
    check_against_run(sc, t)
    st = sc.getInfo()["engine"]
    assert st["smem_table_slots"] == slots
    if t["kmers_nb_distinct"] > 20000:                       # (lowcomplexity: a few hot minimizers, all beyond 16 x slots)
        assert st["nb_parts_smem"] > 0
        if slots == 64:
            assert st["nb_smem_splits"] > 0


@pytest.mark.parametrize("t", SHELL, ids=[t["name"] for t in SHELL])
def test_shell_goldens(t):
    sc = run_gpu(t["files"], t["k"], t["abundance_min"])
    if "hist" in t:
        assert sparse_hist(sc.getHistogram()[0]) == t["hist"]
    if "dsk2ascii" in t:
        assert "".join("%s %d\n" % p for p in sc.solidKmerStrings()) == t["dsk2ascii"]


def test_dsk_check1_literals():
    d = UNIT["DSK_check1"]
    for c in d["checks"]:
        sc = SortingCountAlgorithm(BankStrings(d["seqsets"][c["seqs"]]), {"-kmer-size": c["k"], "-abundance-min": str(c["nks"])}).execute()
        assert sc.getInfo()["kmers_nb_solid"] == c["nb_solid"], c


def test_dsk_check2_values():
    d = UNIT["DSK_check2"]
    sc = SortingCountAlgorithm(BankStrings(d["seq"]), {"-kmer-size": 31, "-abundance-min": "1"}).execute()
    keys, cnt = sc.getSolidCounts()
    lits = {int(x, 16) for x in d["hex_literals"]}
    vals = {int(v) for v in keys[:, 0]}
    assert vals == lits - {0x8b0c176c3b43d207}
    assert sum(vals) % (1 << 64) == 0x8b0c176c3b43d207 and (cnt == 1).all()


@pytest.mark.parametrize("name", ["DSK_perBank1", "DSK_perBank2"])
def test_dsk_perbank_literals(name):
    d = UNIT[name]
    for c in d["checks"]:
        album = BankAlbum([BankStrings(s) for s in d["banks"]])
        props = {"-kmer-size": c["k"], "-abundance-min": str(c["min"]), "-abundance-max": c["max"], "-solidity-kind": c["kind"]}
        sc = SortingCountAlgorithm(album, props).execute()
        assert len(sc.getSolidCounts()[1]) == c["nb_solid"], c


def test_unhandled_kmer_size():
    with pytest.raises(RuntimeError, match="unhandled kmer size 128"):
        SortingCountAlgorithm(BankStrings("ACGT" * 40), {"-kmer-size": 128}).execute()


def test_empty_and_short_inputs():
    for data in (b"", b">only header\n", b">a\nACGT\n", b"no header at all\n"):
        sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": 31, "-abundance-min": "1"}).execute()
        assert sc.getInfo()["kmers_nb_valid"] == 0 and len(sc.getSolidCounts()[1]) == 0
        assert sc.getHistogram()[0].sum() == 0


def test_format_error_is_reported():
    from dsk_b200 import DskGpuError
    with pytest.raises(DskGpuError) as e:
        SortingCountAlgorithm(BankBytes(b">a\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"),
                              {"-kmer-size": 11}).execute()
    assert e.value.code == -4


def compare_with_oracle(data_banks, k, engine=None, **kw):
    ref = oracle.count_files(data_banks, k, abundance_min=kw.get("abundance_min", 2), histo2d=kw.get("histo2d", False),
                             kind=kw.get("kind", "sum"))
    props = {"-kmer-size": k, "-abundance-min": str(kw.get("abundance_min", 2)), "-histo2D": int(kw.get("histo2d", False)),
             "-solidity-kind": kw.get("kind", "sum")}
    bank = BankAlbum([BankBytes(d) for d in data_banks]) if len(data_banks) > 1 else BankBytes(data_banks[0])
    sc = SortingCountAlgorithm(bank, props, **(engine or {})).execute()
    keys, cnt = sc.getSolidCounts()
    lo, hi, rc = ref.solid_kmers()
    assert sc.getInfo()["kmers_nb_valid"] == ref.kmers_nb_valid
    assert sc.getInfo()["kmers_nb_distinct"] == ref.nb_distinct
    assert len(cnt) == len(rc)
    assert (keys[:, 0] == lo).all() and (cnt.astype(np.int64) == rc).all()
    if keys.shape[1] == 2:
        assert (keys[:, 1] == hi).all()
    h1, h2 = sc.getHistogram()
    assert (h1 == ref.hist).all()
    if kw.get("histo2d"):
        assert (h2 == ref.hist2d).all()
    return sc


@pytest.mark.parametrize("k", [31, 63, 21, 47])
def test_synthetic_vs_oracle(k):
    buf, n, _ = reads_fasta(G=300_000, coverage=40, L=150, err=0.01, seed=100 + k)
    compare_with_oracle([buf[:n].tobytes()], k)


@pytest.mark.parametrize("k,m", [(31, 8), (31, 12), (31, 14), (63, 12), (63, 14), (21, 13), (63, 15), (31, 15), (33, 16)])
def test_minimizer_sizes(k, m):
    # -minimizer-size only moves k-mers between partitions: same counts whatever m (big jobs run with m = 12 / 14 / 15; 16 is clipped to 15)
    buf, n, _ = reads_fasta(G=300_000, coverage=30, L=150, err=0.01, seed=300 + m)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": k, "-abundance-min": "2", "-minimizer-size": m}).execute()
    keys, cnt = sc.getSolidCounts()
    lo, hi, rc = ref.solid_kmers()
    assert sc.getInfo()["kmers_nb_valid"] == ref.kmers_nb_valid and sc.getInfo()["kmers_nb_distinct"] == ref.nb_distinct
    assert len(cnt) == len(rc) and (keys[:, 0] == lo).all() and (cnt.astype(np.int64) == rc).all()
    assert (sc.getHistogram()[0] == ref.hist).all()


@pytest.mark.parametrize("mode", ["auto", "sort", "hash", "smem"])
def test_synthetic_medium_modes(mode):
    buf, n, _ = reads_fasta(G=1_000_000, coverage=30, L=150, err=0.01, seed=5)
    sc = compare_with_oracle([buf[:n].tobytes()], 31, engine=dict(count_mode=mode, hash_log2_slots=18))
    assert sc.getInfo()["engine"]["nb_partitions"] > 4


@pytest.mark.parametrize("heavy", ["bucket", "table"])
@pytest.mark.parametrize("k", [31, 63])
def test_auto_mixed_shared_memory_and_heavy_partitions(k, heavy, monkeypatch):
    # a shared-memory table far smaller than the heavy minimizer bins, in AUTO mode: the partitions beyond two sub-passes
    # are renumbered to the end and expanded into hash buckets of flat keys counted in shared memory (default) or counted
    # by the global table (DSKGPU_HEAVY_PATH=table), the others by the shared-memory path, in the same job
    monkeypatch.setenv("DSKGPU_HEAVY_PATH", heavy)
    monkeypatch.setenv("DSKGPU_SMEM_MAX_SPLIT0", "1")                     # (default 4: up to 16 record sub-passes stay in shared memory)
    buf, n, _ = reads_fasta(G=1_000_000, coverage=30, L=150, err=0.01, seed=11)
    sc = compare_with_oracle([buf[:n].tobytes()], k, engine=dict(count_mode="auto", smem_table_slots=256, hash_log2_slots=18))
    st = sc.getInfo()["engine"]
    assert st["nb_parts_smem"] > 0 and st["smem_table_slots"] == 256
    assert (st["nb_groups_bucket"] > 0) == (heavy == "bucket") and (st["nb_groups_hash"] > 0) == (heavy == "table")


def test_heavy_bucket_overflow_falls_back_to_the_table(monkeypatch):
    # one k-mer with a huge multiplicity (poly-A reads) outgrows any hash-bucket slab: that group must fall back to the
    # global table, never lose counts
    monkeypatch.setenv("DSKGPU_HEAVY_PATH", "bucket")
    monkeypatch.setenv("DSKGPU_SMEM_MAX_SPLIT0", "1")
    buf, n, _ = reads_fasta(G=200_000, coverage=20, L=150, err=0.01, seed=21)
    data = buf[:n].tobytes() + b"".join(b">p%d\n%s\n" % (i, b"A" * 150) for i in range(3000))
    sc = compare_with_oracle([data], 31, engine=dict(count_mode="auto", smem_table_slots=256, hash_log2_slots=16))
    st = sc.getInfo()["engine"]
    assert st["nb_groups_hash"] > 0


def test_heavy_buckets_with_per_bank_counts(monkeypatch):
    # -histo2D + heavy partitions: the bank id rides in the two spare top bits of the flat keys
    monkeypatch.setenv("DSKGPU_SMEM_MAX_SPLIT0", "1")
    g = genome_codes(300_000, seed=19)
    asm = assembly_fasta(g)
    buf, n, _ = reads_fasta(coverage=25, L=150, err=0.01, seed=19, genome=g)
    for k in (31, 63, 21):
        sc = compare_with_oracle([asm, buf[:n].tobytes()], k, histo2d=True, engine=dict(count_mode="auto", smem_table_slots=256, hash_log2_slots=16))
        st = sc.getInfo()["engine"]
        assert st["nb_groups_bucket"] > 0 and st["nb_parts_smem"] > 0


def test_histo2d_assembly_vs_reads():
    g = genome_codes(200_000, seed=9)
    asm = assembly_fasta(g)
    buf, n, _ = reads_fasta(coverage=20, L=150, err=0.01, seed=9, genome=g)
    compare_with_oracle([asm, buf[:n].tobytes()], 31, histo2d=True)
    compare_with_oracle([asm, buf[:n].tobytes()], 63, histo2d=True, engine=dict(count_mode="sort"))
    # per-bank counts in the shared-memory tables (two counts per slot), with and without overflow splits, 64- and 128-bit keys
    for k, eng in ((31, dict(count_mode="smem")), (63, dict(count_mode="smem")), (31, dict(count_mode="smem", smem_table_slots=256)),
                   (63, dict(count_mode="auto", smem_table_slots=512, hash_log2_slots=16)), (31, dict(count_mode="hash"))):
        sc = compare_with_oracle([asm, buf[:n].tobytes()], k, histo2d=True, engine=eng)
        st = sc.getInfo()["engine"]
        assert (st["nb_parts_smem"] > 0) == (eng["count_mode"] != "hash")


def test_push_reads_equals_push_bytes():
    buf, n, _ = reads_fasta(G=50_000, coverage=10, L=100, err=0.02, seed=3)
    data = buf[:n].tobytes()
    seqs = [ln for ln in data.split(b"\n") if ln and not ln.startswith(b">")]
    a = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": 27}).execute()
    b = SortingCountAlgorithm(BankStrings(seqs), {"-kmer-size": 27}).execute()
    assert (a.getSolidCounts()[0] == b.getSolidCounts()[0]).all() and (a.getSolidCounts()[1] == b.getSolidCounts()[1]).all()


def test_device_resident_input_and_reset():
    import torch
    buf, n, _ = reads_fasta(G=200_000, coverage=30, L=150, err=0.01, seed=11)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], 31, abundance_min=2)
    d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    with GpuCounter(kmer_size=31, abundance_min=2, stream=torch.cuda.current_stream().cuda_stream) as eng:
        for _ in range(3):                                   # reset() must give identical results every step
            eng.reset()
            eng.push_device_bytes(d.data_ptr(), n)
            eng.finish()
            keys, cnt = eng.solid()
            lo, hi, rc = ref.solid_kmers()
            assert (keys[:, 0] == lo).all() and (cnt == rc.astype(np.uint32)).all()
            assert (eng.histogram()[0] == ref.hist).all()
        # unaligned device pointer
        eng.reset()
        d2 = torch.empty(n + 7, dtype=torch.uint8, device="cuda")
        d2[3:3 + n] = d
        eng.push_device_bytes(d2.data_ptr() + 3, n)
        eng.finish()
        assert (eng.solid()[1] == rc.astype(np.uint32)).all()


@pytest.mark.parametrize("k", [31, 63])
def test_full_size_properties(k):
    """BASELINE.json configs[1]: 5 Mbp genome, 100x, 150 bp, 1 % error -- size-independent invariants."""
    buf, n, nreads = reads_fasta(G=5_000_000, coverage=100, L=150, err=0.01, seed=42)
    with GpuCounter(kmer_size=k, abundance_min=2) as eng:
        eng.push_bytes(buf[:n])
        eng.finish()
        st = eng.stats()
        keys, cnt = eng.solid()
        h1, _ = eng.histogram()
    assert st["kmers_nb_valid"] == nreads * (150 - k + 1)               # no N in this set
    assert st["nb_sequences"] == nreads and st["nb_nucleotides"] == nreads * 150
    assert int((h1 * np.arange(10001, dtype=np.uint64)).sum()) == st["kmers_nb_valid"]   # checksum of counts (no count >= 10000 here)
    assert int(h1.sum()) == st["kmers_nb_distinct"]
    assert int(h1[2:].sum()) == st["kmers_nb_solid"] == len(cnt)
    assert int(cnt.astype(np.uint64).sum()) == st["kmers_nb_valid"] - int(h1[1])
    hc = np.bincount(np.minimum(cnt, 10000), minlength=10001).astype(np.uint64)
    assert (hc[2:10000] == h1[2:10000]).all()
    if keys.shape[1] == 1:
        assert (keys[1:, 0] > keys[:-1, 0]).all()
    else:
        assert ((keys[1:, 1] > keys[:-1, 1]) | ((keys[1:, 1] == keys[:-1, 1]) & (keys[1:, 0] > keys[:-1, 0]))).all()
    if k == 31:   # SURVEY.md 6.2 [measured with the reference]: 103 530 226 distinct / 12 960 855 solid for seed 42 -- different generator, same shape
        assert 0.9e8 < st["kmers_nb_distinct"] < 1.2e8


def split_records(data, parts):
    """cut a FASTA image into `parts` slices at record boundaries"""
    cuts = [0]
    for i in range(1, parts):
        j = data.find(b"\n>", len(data) * i // parts)
        cuts.append(len(data) if j < 0 else j + 1)
    cuts.append(len(data))
    return [data[cuts[i]:cuts[i + 1]] for i in range(parts)]


@pytest.mark.parametrize("W,k,mode", [(2, 31, "auto"), (3, 31, "hash"), (2, 63, "auto"), (4, 31, "sort"), (3, 63, "smem"), (3, 31, "auto-tiny"), (5, 31, "auto-tiny"),
                                      (8, 31, "auto")])
def test_multi_rank_exchange_in_process(W, k, mode, monkeypatch):
    """N ranks as N contexts on one GPU: every rank parses a slice and plans the same partitions on the device, scatters its
    records into owner-major order, one contiguous copy per (sender, owner) pair moves them (sender-major receive layout), every
    rank counts only what it owns, each partition as W segments.  Union of the ranks' outputs == oracle."""
    from dsk_b200.counter import multi_finish
    buf, n, _ = reads_fasta(G=400_000, coverage=30, L=150, err=0.01, seed=77)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    extra = {}
    if mode == "auto-tiny":               # heavy partitions (global table) and light ones (shared memory) on every rank
        mode, extra = "auto", dict(smem_table_slots=256)
        monkeypatch.setenv("DSKGPU_SMEM_MAX_SPLIT0", "1")
    engines = [GpuCounter(kmer_size=k, abundance_min=2, rank=r, world_size=W, count_mode=mode, hash_log2_slots=16, **extra) for r in range(W)]
    try:
        for e, piece in zip(engines, split_records(data, W)):
            e.push_bytes(piece)
        multi_finish(engines)
        P = engines[0].stats()["nb_partitions"]
        # every rank derived the same plan on its device, and the ranks' local records of a partition add up to the job's
        plans = [e.debug_plan() for e in engines]
        for pl in plans[1:]:
            assert pl[0] == plans[0][0] and (pl[1] == plans[0][1]).all() and (pl[2] == plans[0][2]).all() and (pl[3] == plans[0][3]).all()
        assert (np.sum([pl[4] for pl in plans], axis=0) == plans[0][3]).all()
        assert int(plans[0][3].sum()) == sum(e.stats()["nb_superkmers"] for e in engines)
        keys, cnts, hist, valid, distinct = [], [], np.zeros(10001, np.uint64), 0, 0
        for r, e in enumerate(engines):
            kk, cc = e.solid()
            if kk.shape[1] == 1:
                assert (kk[1:, 0] > kk[:-1, 0]).all()
            keys.append(kk); cnts.append(cc)
            hist += e.histogram()[0]
            st = e.stats()
            valid += st["kmers_nb_valid"]; distinct += st["kmers_nb_distinct"]
            assert st["nb_partitions"] == P
            if extra and W <= 3:
                assert st["nb_parts_smem"] > 0 and st["nb_groups_hash"] + st["nb_groups_bucket"] > 0
        keys = np.concatenate(keys); cnts = np.concatenate(cnts)
        order = np.lexsort((keys[:, 0], keys[:, -1])) if keys.shape[1] == 2 else np.argsort(keys[:, 0], kind="stable")
        keys, cnts = keys[order], cnts[order]
        lo, hi, rc = ref.solid_kmers()
        assert valid == ref.kmers_nb_valid and distinct == ref.nb_distinct
        assert len(cnts) == len(rc) and (keys[:, 0] == lo).all() and (cnts.astype(np.int64) == rc).all()
        if keys.shape[1] == 2:
            assert (keys[:, 1] == hi).all()
        assert (hist == ref.hist).all()
    finally:
        for e in engines:
            e.close()


# ---------------------------------------------------------------- -abundance-min auto
AUTO = load_json("ref_runs_auto.json")["runs"]


@pytest.mark.parametrize("t", AUTO, ids=[t["name"] for t in AUTO])
def test_reference_auto_cutoff_runs(t):
    """two count passes over the partitions held in HBM: cutoff histogram(s) -> thresholds -> recount"""
    sc = run_gpu(t["files"], t["k"], t["abundance_min"], kind=t["solidity_kind"])
    assert sc.getInfo()["cutoffs_auto"] == t["cutoffs"]
    assert sc.getInfo()["kmers_nb_solid"] == t["nb_solid"]
    keys, cnt = sc.getSolidCounts()
    h1, _ = sc.getHistogram()
    assert sparse_hist(h1) == t["hist"]
    hi = keys[:, 1] if keys.shape[1] == 2 else np.zeros(len(keys), np.uint64)
    dg, _ = digest(keys[:, 0], hi, cnt, t["k"])
    assert dg == t["kmers_sha256"]


def test_auto_cutoff_synthetic_vs_oracle():
    buf, n, _ = reads_fasta(G=200_000, coverage=40, L=150, err=0.01, seed=5)
    data = buf[:n].tobytes()
    for k in (21, 31):
        ref = oracle.count_files([data], k, abundance_min=-1)
        sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": k, "-abundance-min": "auto"}).execute()
        assert sc.getInfo()["cutoffs_auto"] == ref.cutoffs
        keys, cnt = sc.getSolidCounts()
        lo, _, rc = ref.solid_kmers()
        assert len(cnt) == len(rc) and (keys[:, 0] == lo).all() and (cnt.astype(np.int64) == rc).all()
