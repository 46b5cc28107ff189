"""Parity tests (-m gpu) of what round 2 added to the path: the pass loop (jobs whose records exceed HBM), the solid-set buffers
sized from an estimate with an exact retry, the global-table groups that outgrow an estimated size, several contexts in one
process (dskgpu_multi_finish: what the multi-GPU `dsk_gpu` CLI runs), and -- when the box has at least two GPUs -- the real
multi-process exchange under torchrun.  Bit-exact against the oracle everywhere."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from dsk_b200 import SortingCountAlgorithm, BankBytes, BankAlbum, GpuCounter
from dsk_b200.counter import multi_finish
from dsk_b200.synth import reads_fasta, genome_codes, assembly_fasta

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sorted_pairs(keys, cnts):
    order = np.lexsort((keys[:, 0], keys[:, -1])) if keys.shape[1] == 2 else np.argsort(keys[:, 0], kind="stable")
    return keys[order], cnts[order]


def assert_equals_oracle(keys, cnts, hist, ref, hist2d=None):
    keys, cnts = sorted_pairs(keys, cnts)
    lo, hi, rc = ref.solid_kmers()
    assert len(cnts) == len(rc)
    assert (keys[:, 0] == lo).all() and (cnts.astype(np.int64) == rc).all()
    if keys.shape[1] == 2:
        assert (keys[:, 1] == hi).all()
    assert (hist == ref.hist).all()
    if hist2d is not None:
        assert (hist2d == ref.hist2d).all()


def split_records(data, parts):
    cuts = [0]
    for i in range(1, parts):
        j = data.find(b"\n>", len(data) * i // parts)
        cuts.append(len(data) if j < 0 else j + 1)
    cuts.append(len(data))
    return [data[cuts[i]:cuts[i + 1]] for i in range(parts)]


# ---------------------------------------------------------------- pass loop (K/SortingCountAlgorithm.cpp:678-689, :1086)
@pytest.mark.parametrize("k,nb_passes", [(31, 2), (31, 3), (63, 3), (21, 5)])
def test_pass_loop_union_equals_oracle(k, nb_passes):
    buf, n, _ = reads_fasta(G=300_000, coverage=30, L=150, err=0.01, seed=900 + k)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": k, "-abundance-min": "2"}, nb_passes=nb_passes).execute()
    info = sc.getInfo()
    assert info["kmers_nb_valid"] == ref.kmers_nb_valid                 # every pass sees the whole bank
    assert info["kmers_nb_distinct"] == ref.nb_distinct                 # ... and counts a disjoint share of it
    assert info["engine"]["kmers_in_pass"] == ref.kmers_nb_valid        # summed over the passes
    keys, cnt = sc.getSolidCounts()
    assert_equals_oracle(keys, cnt, sc.getHistogram()[0], ref)


def test_pass_loop_with_histo2d_and_auto_cutoff():
    g = genome_codes(150_000, seed=29)
    asm = assembly_fasta(g)
    buf, n, _ = reads_fasta(coverage=25, L=150, err=0.01, seed=29, genome=g)
    reads = buf[:n].tobytes()
    ref = oracle.count_files([asm, reads], 31, abundance_min=2, histo2d=True)
    sc = SortingCountAlgorithm(BankAlbum([BankBytes(asm), BankBytes(reads)]), {"-kmer-size": 31, "-abundance-min": "2", "-histo2D": 1},
                               nb_passes=3).execute()
    keys, cnt = sc.getSolidCounts()
    h1, h2 = sc.getHistogram()
    assert_equals_oracle(keys, cnt, h1, ref, h2)
    # -abundance-min auto: the cutoff comes from the histogram of ALL passes
    ref = oracle.count_files([reads], 31, abundance_min=-1)
    sc = SortingCountAlgorithm(BankBytes(reads), {"-kmer-size": 31, "-abundance-min": "auto"}, nb_passes=2).execute()
    assert sc.getInfo()["cutoffs_auto"] == ref.cutoffs
    keys, cnt = sc.getSolidCounts()
    assert_equals_oracle(keys, cnt, sc.getHistogram()[0], ref)


def test_set_pass_needs_an_empty_context():
    from dsk_b200 import DskGpuError
    buf, n, _ = reads_fasta(G=20_000, coverage=5, L=100, err=0.0, seed=1)
    with GpuCounter(kmer_size=21) as e:
        e.set_pass(1, 2)
        e.push_bytes(buf[:n].tobytes())
        with pytest.raises(DskGpuError) as ex:
            e.set_pass(0, 2)
        assert ex.value.code == -5
        with pytest.raises(DskGpuError):
            e.reset(); e.set_pass(2, 2)
    lib = __import__("dsk_b200")._lib.lib()
    assert lib.dskgpu_suggest_nb_passes(10**9, 31, 1, 0, 0) == 1
    assert lib.dskgpu_suggest_nb_passes(72 * 10**9, 31, 1, 180 << 30, 0) >= 2          # BASELINE configs[2] on ONE GPU: passes
    assert lib.dskgpu_suggest_nb_passes(72 * 10**9, 31, 8, 180 << 30, 0) == 1          # ... on eight: fits


# ---------------------------------------------------------------- estimates that turn out wrong are answered, never fatal
@pytest.mark.parametrize("k,mode", [(31, "auto"), (63, "auto"), (31, "hash"), (31, "sort")])
def test_solid_buffers_too_small_are_regrown_to_the_exact_size(k, mode, monkeypatch):
    monkeypatch.setenv("DSKGPU_TEST_SOLID_CAP", "1000")
    buf, n, _ = reads_fasta(G=200_000, coverage=30, L=150, err=0.01, seed=41)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": k, "-abundance-min": "2"}, count_mode=mode).execute()
    assert sc.getInfo()["engine"]["nb_solid_regrows"] == 1
    assert sc.getInfo()["kmers_nb_distinct"] == ref.nb_distinct
    keys, cnt = sc.getSolidCounts()
    assert_equals_oracle(keys, cnt, sc.getHistogram()[0], ref)


@pytest.mark.parametrize("k", [31, 63])
def test_global_table_group_outgrowing_its_estimate_is_regrouped(k, monkeypatch):
    # groups sized as if 2 % of the k-mers were distinct (they are ~35 %): every group overflows its table and must be redone
    # in sub-groups sized for distinct = total -- the reference never fails here, neither does this path
    monkeypatch.setenv("DSKGPU_TEST_HASH_RATIO", "0.02")
    buf, n, _ = reads_fasta(G=300_000, coverage=30, L=150, err=0.01, seed=43)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    sc = SortingCountAlgorithm(BankBytes(data), {"-kmer-size": k, "-abundance-min": "2"}, count_mode="hash", hash_log2_slots=15).execute()
    st = sc.getInfo()["engine"]
    assert st["nb_hash_regroups"] > 0
    assert sc.getInfo()["kmers_nb_distinct"] == ref.nb_distinct
    keys, cnt = sc.getSolidCounts()
    assert_equals_oracle(keys, cnt, sc.getHistogram()[0], ref)


# ---------------------------------------------------------------- the device planner against its host mirror
@pytest.mark.parametrize("k,slots,mode", [(31, 4096, "auto"), (63, 2048, "auto"), (31, 1024, "auto"), (31, 256, "auto")])
def test_device_planner_equals_host_mirror(k, slots, mode, monkeypatch):
    """dsk_b200/csrc/plan.cuh: prefix sums + cut flags + renumbering on the device give, bit for bit, the plan of the
    sequential restatement (plan_host via dskgpu_selftest_plan) on the histogram of a real job"""
    import torch
    from dsk_b200 import _lib
    monkeypatch.setenv("DSKGPU_SMEM_MAX_SPLIT0", "1")                     # heavy partitions exist at every table size of the test
    buf, n, _ = reads_fasta(G=2_000_000, coverage=20, L=150, err=0.01, seed=91)
    data = bytearray(buf[:n].tobytes())
    data[1000:1000] = b">low\n" + b"ACAC" * 40000 + b"\n"                       # one hot minimizer: heavy partitions
    eng = GpuCounter(kmer_size=k, abundance_min=2, smem_table_slots=slots, count_mode=mode, hash_log2_slots=18)
    try:
        eng.push_bytes(bytes(data))
        g4 = eng.xchg_prepare()
        level = eng.xchg_set_global(g4)
        G = torch.empty(2 << level, dtype=torch.int64, device="cuda")
        eng.xchg_hist(G.data_ptr())
        eng.xchg_sync()
        P, PW, need = eng.xchg_plan(G.data_ptr())
        lvl, b2p, pk, pr, pl = eng.debug_plan()
        assert lvl == level and len(pk) == P and PW == P
        gh = G.cpu().numpy().astype(np.uint64)
        density = min(1.0, max(0.01, float(g4[3]) / float(g4[2]))) if g4[2] >= 4096 else 1.0
        L = _lib.lib()
        nb = 1 << level
        hb2p = np.zeros(nb, np.uint32); hpk = np.zeros(nb + 1, np.uint64); hpl = np.zeros(nb + 1, np.uint64)
        hP = L.dskgpu_selftest_plan(level, gh.ctypes.data, gh.ctypes.data, 1, 1, slots, density, 0, 0,
                                    hb2p.ctypes.data, hpk.ctypes.data, hpl.ctypes.data, hpk.size)
        assert hP == P, (hP, P)
        assert (hb2p == b2p).all() and (hpk[:P] == pk).all() and (hpl[:P] == pl).all() and (pr == pl).all()
        assert int(need[0]) == int(pr.sum()) == int(g4[1])
        eng.finish()                                                            # and the job still counts right from that plan
        ref = oracle.count_files([bytes(data)], k, abundance_min=2)
        kk, cc = eng.solid()
        assert_equals_oracle(kk, cc, eng.histogram()[0], ref)
        if mode == "auto":
            st = eng.stats()
            assert st["nb_parts_smem"] > 0 and st["nb_groups_hash"] > 0
    finally:
        eng.close()


@pytest.mark.parametrize("k", [31, 63])
def test_packed_bin_histogram_rebuilt_when_it_could_wrap(k, monkeypatch):
    """the fine-bin histogram packs (records, k-mers) in one word; a bin heavy enough for the record field to wrap is flagged
    and the histogram is rebuilt exactly from the record meta -- forced here by lowering the limit; same plan, same counts"""
    buf, n, _ = reads_fasta(G=300_000, coverage=30, L=150, err=0.01, seed=93)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    plans = []
    for limit in (None, "50"):
        if limit:
            monkeypatch.setenv("DSKGPU_TEST_HIST_LIMIT", limit)
        with GpuCounter(kmer_size=k, abundance_min=2) as eng:
            eng.push_bytes(data)
            eng.finish()
            st = eng.stats()
            assert st["hist_rebuilt"] == (1 if limit else 0)
            kk, cc = eng.solid()
            assert_equals_oracle(kk, cc, eng.histogram()[0], ref)
            plans.append(eng.debug_plan())
    assert all((a == b).all() if hasattr(a, "all") else a == b for a, b in zip(plans[0], plans[1]))


@pytest.mark.parametrize("k,slots", [(31, 512), (63, 512), (31, 256), (21, 1024)])
def test_record_sub_passes_keep_big_partitions_in_shared_memory(k, slots):
    """a partition of up to 16 tables is counted in shared memory as record sub-passes over the sub-bins stored in the records
    (every record expanded in exactly one sub-pass); with a tiny table most partitions of this job are such partitions"""
    g = genome_codes(600_000, seed=23)
    buf, n, _ = reads_fasta(coverage=30, L=150, err=0.01, seed=23, genome=g)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    with GpuCounter(kmer_size=k, abundance_min=2, smem_table_slots=slots, nb_partitions=64) as eng:      # 64 partitions of ~270 K k-mers: far beyond one table
        eng.push_bytes(data)
        eng.finish()
        st = eng.stats()
        kk, cc = eng.solid()
        assert st["kmers_nb_distinct"] == ref.nb_distinct
        assert_equals_oracle(kk, cc, eng.histogram()[0], ref)
    with GpuCounter(kmer_size=k, abundance_min=2, smem_table_slots=slots) as eng:                         # planned sizes: partitions of up to 16 tables stay in shared memory
        eng.push_bytes(data)
        eng.finish()
        st = eng.stats()
        assert st["nb_parts_smem"] == st["nb_partitions"] and st["nb_groups_hash"] == 0 and st["nb_groups_sort"] == 0
        kk, cc = eng.solid()
        assert_equals_oracle(kk, cc, eng.histogram()[0], ref)
    # two banks, per-bank counts: the bank nibble shares the byte with the sub-bin
    asm = assembly_fasta(g)
    sc = SortingCountAlgorithm(BankAlbum([BankBytes(asm), BankBytes(data)]), {"-kmer-size": k, "-abundance-min": "2", "-histo2D": 1}, smem_table_slots=slots).execute()
    ref2 = oracle.count_files([asm, data], k, abundance_min=2, histo2d=True)
    keys, cnt = sc.getSolidCounts()
    h1, h2 = sc.getHistogram()
    assert_equals_oracle(keys, cnt, h1, ref2, h2)


@pytest.mark.parametrize("k,fine,limit", [(31, 24, None), (63, 24, None), (31, 18, None), (31, 24, "40")])
def test_fine_histogram_levels(k, fine, limit, monkeypatch):
    """the fine minimizer-bin histogram has 2^22 bins, or 2^24 for the jobs that run with 14-letter minimizers (forced here on a
    small job, also together with the exact-rebuild path); coarser and finer levels give the same results"""
    monkeypatch.setenv("DSKGPU_FINE_LOG2", str(fine))
    if limit:
        monkeypatch.setenv("DSKGPU_TEST_HIST_LIMIT", limit)
    buf, n, _ = reads_fasta(G=300_000, coverage=30, L=150, err=0.01, seed=95)
    data = buf[:n].tobytes()
    ref = oracle.count_files([data], k, abundance_min=2)
    with GpuCounter(kmer_size=k, abundance_min=2, minimizer_size=14 if fine == 24 else 10, smem_table_slots=1024) as eng:
        eng.push_bytes(data)
        eng.finish()
        st = eng.stats()
        assert st["hist_rebuilt"] == (1 if limit else 0)
        kk, cc = eng.solid()
        assert st["kmers_nb_distinct"] == ref.nb_distinct
        assert_equals_oracle(kk, cc, eng.histogram()[0], ref)
    run_multi([0, 0, 0], k, data, minimizer_size=14 if fine == 24 else 10)


@pytest.mark.parametrize("k,W,min_parts", [(31, 1, 1), (63, 1, 1), (31, 3, 1), (31, 1, 300), (31, 2, 5000)])
def test_msd_multi_split_scatter(k, W, min_parts, monkeypatch):
    """the partition scatter of jobs with millions of partitions (MSD multi-split passes with block-level binning in shared
    memory), forced on a small job: 1, 2 or 3 passes depending on the partition count, same results"""
    monkeypatch.setenv("DSKGPU_MSD_MIN_PARTS", str(min_parts))
    buf, n, _ = reads_fasta(G=500_000, coverage=30, L=150, err=0.01, seed=97)
    data = bytearray(buf[:n].tobytes())
    data[500:500] = b">low\n" + b"AC" * 30000 + b"\n"
    slots = {1: 0, 300: 1024, 5000: 64}[min_parts]                      # 0: ~1 K partitions (2 passes); 1024: ~15 K (2 passes); 64: ~250 K (3 passes)
    if W == 1:
        ref = oracle.count_files([bytes(data)], k, abundance_min=2)
        with GpuCounter(kmer_size=k, abundance_min=2, smem_table_slots=slots) as eng:
            eng.push_bytes(bytes(data))
            eng.finish()
            st = eng.stats()
            assert st["scatter_passes"] >= (1 if st["nb_partitions"] <= 256 else 2 if st["nb_partitions"] <= 65536 else 3)
            kk, cc = eng.solid()
            assert st["kmers_nb_distinct"] == ref.nb_distinct
            assert_equals_oracle(kk, cc, eng.histogram()[0], ref)
    else:
        run_multi([0] * W, k, bytes(data), smem_table_slots=slots)


@pytest.mark.parametrize("k", [31, 63])
def test_solid_set_ordering_fixup_and_fallback(k, monkeypatch):
    """ordering of the solid set: top ceil(log2 n) bits by one-sweep passes + neighbourhood fix-up; a set with long groups of
    equal prefixes (here 300 solid k-mers that start with the same 10 bases) makes the fix-up give up and the full-width sort
    run; both must give the ascending order of the plain sort"""
    rng = np.random.default_rng(5 + k)
    buf, n, _ = reads_fasta(G=150_000, coverage=20, L=150, err=0.01, seed=11)
    recs = [buf[:n].tobytes()]
    for i in range(300):
        tail = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), k - 10).tobytes())
        recs.append((b">p%d\n" % i + b"A" * 10 + tail + b"\n") * 3)
    data = b"".join(recs)
    ref = oracle.count_files([data], k, abundance_min=2)
    outs = []
    for full in (False, True):
        if full:
            monkeypatch.setenv("DSKGPU_SORT_FULL", "1")
        with GpuCounter(kmer_size=k, abundance_min=2) as eng:
            eng.push_bytes(data)
            eng.finish()
            kk, cc = eng.solid()
            st = eng.stats()
            assert st["sort_fallbacks"] == (0 if full else 1)
            assert_equals_oracle(kk, cc, eng.histogram()[0], ref)
            key = kk[:, 0].astype(object) if kk.shape[1] == 1 else (kk[:, 1].astype(object) << 64) | kk[:, 0].astype(object)
            assert all(key[i] < key[i + 1] for i in range(len(key) - 1))            # ascending as delivered, not only as a set
            outs.append((kk.copy(), cc.copy()))
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()
    monkeypatch.delenv("DSKGPU_SORT_FULL")
    # the plain case: no fallback, ascending
    buf, n, _ = reads_fasta(G=400_000, coverage=30, L=150, err=0.01, seed=12)
    with GpuCounter(kmer_size=k, abundance_min=2) as eng:
        eng.push_bytes(buf[:n].tobytes())
        eng.finish()
        kk, _ = eng.solid()
        assert eng.stats()["sort_fallbacks"] == 0
        key = kk[:, 0].astype(object) if kk.shape[1] == 1 else (kk[:, 1].astype(object) << 64) | kk[:, 0].astype(object)
        assert all(key[i] < key[i + 1] for i in range(len(key) - 1))


# ---------------------------------------------------------------- several ranks in one process (dskgpu_multi_finish)
def run_multi(devices, k, data, **kw):
    W = len(devices)
    ref = oracle.count_files([data], k, abundance_min=2)
    engines = [GpuCounter(kmer_size=k, abundance_min=2, rank=r, world_size=W, device=devices[r], **kw) for r in range(W)]
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
        assert_equals_oracle(np.concatenate(keys), np.concatenate(cnts), hist, ref)
    finally:
        for e in engines:
            e.close()


@pytest.mark.parametrize("W,k,mode", [(2, 31, "auto"), (3, 63, "auto"), (4, 31, "hash")])
def test_multi_finish_contexts_sharing_one_gpu(W, k, mode):
    buf, n, _ = reads_fasta(G=300_000, coverage=30, L=150, err=0.01, seed=71)
    run_multi([0] * W, k, buf[:n].tobytes(), count_mode=mode, hash_log2_slots=16)


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("k", [31, 63])
def test_multi_finish_across_real_gpus(k):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs at least two GPUs in the box")
    buf, nb, _ = reads_fasta(G=1_000_000, coverage=30, L=150, err=0.01, seed=73)
    run_multi(list(range(min(n, 4))), k, buf[:nb].tobytes())


def test_multi_process_exchange_under_torchrun():
    """tools/mgpu_check.py (one process per GPU, NCCL metadata + CUDA-IPC peer stores; union of the ranks against the oracle)"""
    n = _ngpus()
    if n < 2:
        pytest.skip("needs at least two GPUs in the box")
    w = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(w), "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tools", "mgpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, MGPU_SHORT="1"))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "MISMATCH" not in p.stdout


# ---------------------------------------------------------------- per-sequence statistics (K/BankKmers.hpp:166-215)
BANKSTATS = __import__("util").load_json("ref_bankstats.json")["runs"]


@pytest.mark.parametrize("chunk", [0, 4096, 1 << 16])
@pytest.mark.parametrize("t", BANKSTATS, ids=["+".join(t["files"]) for t in BANKSTATS])
def test_sequence_statistics_match_the_reference(t, chunk):
    """seq_number / seq_size_min / max / mean / deviation and kmers_nb_invalid as `dsk -verbose 1` prints them (goldens from the
    unmodified reference), computed on the device from the record separators of the code stream (seqstats.cuh): FASTA
    (separator in front of a record: the last sequence is closed by the end of the stream), FASTQ (separator behind), empty
    records, sequences longer than a tile, several banks; push granularities that cut sequences at chunk boundaries"""
    from util import read_input
    import math
    with GpuCounter(kmer_size=t["k"], abundance_min=2, nb_banks=len(t["files"]), sequence_stats=True, push_chunk_bytes=chunk) as eng:
        for b, f in enumerate(t["files"]):
            eng.push_bytes(read_input(f), bank=b)
        eng.finish()
        st = eng.stats()
    n = st["seq_stats_sequences"]
    assert n == st["nb_sequences"] == int(t["seq_number"])
    assert st["seq_len_sum"] == st["nb_nucleotides"] == int(t["bank_total_nt"])
    assert st["seq_len_min"] == int(t["seq_size_min"]) and st["seq_len_max"] == int(t["seq_size_max"])
    mean = st["seq_len_sum"] / n
    assert "%.1f" % mean == t["seq_size_mean"]
    assert "%.1f" % math.sqrt(max(0.0, st["seq_len_sumsq"] / n - mean * mean)) == t["seq_size_deviation"]
    assert st["kmers_nb_valid"] == int(t["kmers_nb_valid"]) and st["kmers_nb_invalid"] == int(t["kmers_nb_invalid"])


def test_sequence_statistics_of_pushed_reads_and_default_off():
    # one-sequence-per-line pushes (dskgpu_push_reads: what the adapter feeds from IBank::iterator), a last line without newline
    seqs = ["ACGTACGTAC" * 7, "A" * 40, "", "ACGTNACGT" * 9, "C" * 5]
    k = 21
    with GpuCounter(kmer_size=k, abundance_min=1, sequence_stats=True) as eng:
        eng.push_reads([s.encode() for s in seqs if s])
        eng.finish()
        st = eng.stats()
    lens = [len(s) for s in seqs if s]
    assert st["seq_stats_sequences"] == len(lens) and st["seq_len_min"] == min(lens) and st["seq_len_max"] == max(lens)
    assert st["seq_len_sum"] == sum(lens) and st["seq_len_sumsq"] == sum(x * x for x in lens)
    assert st["kmers_nb_invalid"] == sum(max(0, x - k + 1) for x in lens) - st["kmers_nb_valid"]
    with GpuCounter(kmer_size=k, abundance_min=1) as eng:                    # not asked for: nothing gathered, nothing launched for it
        eng.push_reads([s.encode() for s in seqs if s])
        eng.finish()
        st = eng.stats()
    assert st["seq_stats_sequences"] == 0 and st["kmers_nb_invalid"] == 0
