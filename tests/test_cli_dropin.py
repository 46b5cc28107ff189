"""Drop-in test of the C++ host side (-m gpu): `host/_build/dsk_gpu` is the reference `dsk` command line with the counting
class replaced by the device path (host/GpuSortingCount.hpp over include/dskgpu.h).  Its HDF5 output is read back by the
UNMODIFIED reference readers `dsk2ascii` and `gatb-h5dump` (oracle/_ref/bin, the checker) and compared with
  * the committed outputs of the real reference (tests/golden/ref_runs.json), and
  * the reference `dsk` binary run side by side on the same input, when it is available on the box.
This is the shape of the reference's own end-to-end tests (R/scripts/simple_test.sh:35-135)."""
import hashlib
import os
import subprocess

import pytest

from util import load_json, INPUTS, GOLDEN

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DSK_GPU = os.path.join(ROOT, "host", "_build", "dsk_gpu")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")
RUNS = load_json("ref_runs.json")["runs"]
SHELL = load_json("ref_shell_tests.json")["tests"]

need_bins = pytest.mark.skipif(not (os.path.exists(DSK_GPU) and os.path.exists(os.path.join(REFBIN, "dsk2ascii"))),
                               reason="host/_build/dsk_gpu or oracle/_ref/bin readers not built (python __graft_entry__.py)")


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    assert p.returncode == 0, "%s\n%s\n%s" % (" ".join(cmd), p.stdout[-2000:], p.stderr[-2000:])
    return p.stdout


def dsk_args(t, out):
    a = ["-file", ",".join(os.path.join(INPUTS, f) for f in t["files"]), "-kmer-size", str(t["k"]), "-abundance-min", str(t["abundance_min"]),
         "-out", out, "-histo", "1", "-verbose", "0"]
    if t.get("histo2d"):
        a += ["-histo2D", "1"]
    if t.get("solidity_kind"):
        a += ["-solidity-kind", t["solidity_kind"]]
    if t.get("abundance_max") is not None:
        a += ["-abundance-max", str(t["abundance_max"])]
    return a


def read_back(h5, tmp):
    """what the reference's readers see in an output file"""
    txt = os.path.join(tmp, os.path.basename(h5) + ".txt")
    run([os.path.join(REFBIN, "dsk2ascii"), "-file", h5, "-out", txt, "-verbose", "0"], tmp)
    lines = sorted(open(txt, "rb").read().splitlines())
    histo = run([os.path.join(REFBIN, "gatb-h5dump"), "-y", "-d", "histogram/histogram", h5], tmp)
    histo = "\n".join(ln for ln in histo.splitlines() if not ln.startswith("HDF5 "))      # first line names the file
    header = run([os.path.join(REFBIN, "gatb-h5dump"), "-H", h5], tmp)
    return lines, histo, header


CLI_RUNS = [t for t in RUNS if t["name"] in ("c1_k31", "c1_k63", "c1_k31_min3_max20", "c1234_k31", "longread_k63", "reads.fastq_k31", "multiline.fasta_k31",
                                            "weird.fasta_k31", "histo2d_k31", "histo2d_c123_k31", "c123_k31_min", "c123_k31_one", "c123_k31_all")]


@need_bins
@pytest.mark.parametrize("t", CLI_RUNS, ids=[t["name"] for t in CLI_RUNS])
def test_cli_against_committed_reference_outputs(t, tmp_path):
    tmp = str(tmp_path)
    out = os.path.join(tmp, "gpu_out")
    run([DSK_GPU] + dsk_args(t, out), tmp)
    lines, histo, header = read_back(out + ".h5", tmp)
    assert len(lines) == t["nb_solid"]
    m = hashlib.sha256()
    for ln in lines:
        m.update(ln + b"\n")
    assert m.hexdigest() == t["kmers_sha256"]
    # <out>.histo text: 10000 lines "i\tcount" (CountProcessorHistogram.hpp:111-142)
    rows = open(out + ".histo").read().splitlines()
    assert len(rows) == 10000
    got = {r.split("\t")[0]: int(r.split("\t")[1]) for r in rows if int(r.split("\t")[1])}
    assert got == t["hist"]
    # layout names the readers rely on
    for name in ('GROUP "dsk"', 'GROUP "solid"', 'DATASET "0"', 'GROUP "histogram"', 'DATASET "histogram"', 'DATASET "cutoff"',
                 'DATASET "nbsolidsforcutoff"', 'GROUP "minimizers"', 'DATASET "minimRepart"', 'GROUP "configuration"',
                 'ATTRIBUTE "kmer_size"', 'ATTRIBUTE "nb_partitions"', 'ATTRIBUTE "xml"'):
        assert name in header, name
    if t.get("histo2d"):
        assert len(open(out + ".histo2D").read().splitlines()) == 10001


@need_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(REFBIN, "dsk")), reason="reference dsk binary not on this box")
@pytest.mark.parametrize("name", ["c1_k31", "c1_k63", "histo2d_c123_k31", "reads.fastq_k31"])
def test_cli_side_by_side_with_reference_binary(name, tmp_path):
    t = [r for r in RUNS if r["name"] == name][0]
    tmp = str(tmp_path)
    a, b = os.path.join(tmp, "gpu_out"), os.path.join(tmp, "ref_out")
    run([DSK_GPU] + dsk_args(t, a), tmp)
    run([os.path.join(REFBIN, "dsk")] + dsk_args(t, b) + ["-out-tmp", tmp], tmp)
    la, ha, hda = read_back(a + ".h5", tmp)
    lb, hb, hdb = read_back(b + ".h5", tmp)
    assert la == lb                                          # dsk2ascii | sort
    assert ha == hb                                          # gatb-h5dump -y -d histogram/histogram
    assert open(a + ".histo", "rb").read() == open(b + ".histo", "rb").read()
    if t.get("histo2d"):
        assert open(a + ".histo2D", "rb").read() == open(b + ".histo2D", "rb").read()
    # same dataset types: the compound {value, abundance} of dsk/solid/0 and {index, abundance} of the histogram

    def types(h):
        return sorted(set(ln.strip() for ln in h.splitlines() if "H5T_" in ln))
    assert types(hda) == types(hdb)


@need_bins
def test_cli_shell_goldens(tmp_path):
    """R/scripts/simple_test.sh: histogram dumps and dsk2ascii text"""
    tmp = str(tmp_path)
    for t in SHELL:
        out = os.path.join(tmp, "o_" + t["name"])
        run([DSK_GPU] + dsk_args(t, out), tmp)
        if "dsk2ascii" in t:
            txt = out + ".txt"
            run([os.path.join(REFBIN, "dsk2ascii"), "-file", out + ".h5", "-out", txt, "-verbose", "0"], tmp)
            assert open(txt).read() == t["dsk2ascii"]
        if "hist" in t:
            rows = open(out + ".histo").read().splitlines()
            assert {r.split("\t")[0]: int(r.split("\t")[1]) for r in rows if int(r.split("\t")[1])} == t["hist"]


PLUGIN = os.path.join(ROOT, "host", "_build", "plugin_check")
AUTO = load_json("ref_runs_auto.json")["runs"]
PLUGIN_RUNS = [t for t in RUNS if t["name"] in ("c1_k31", "c1_k63", "c1_k31_min3_max20", "longread_k63", "reads.fastq_k31",
                                               "c1234_k31", "c123_k31_min", "c123_k31_one", "c123_k31_all", "histo2d_k31", "histo2d_c123_k31")] + \
              [t for t in AUTO if t["name"] in ("auto_c1_k31", "auto_asmreads_k21", "auto_c123_k31_all", "auto_c123_k15_mixed_all", "auto_asm_reads_k21_one")]


@need_bins
@pytest.mark.skipif(not os.path.exists(PLUGIN), reason="host/_build/plugin_check not built (python __graft_entry__.py)")
@pytest.mark.parametrize("t", PLUGIN_RUNS, ids=[t["name"] for t in PLUGIN_RUNS])
def test_plugin_surface_with_the_reference_processors(t, tmp_path):
    """the ICountProcessor plug-in surface (5-argument constructor, G/src/gatb/debruijn/impl/Graph.cpp:399-407): the REFERENCE'S
    OWN processor chain (getDefaultProcessorVector: histogram -> solidity -> dump; cutoff processor first for 'auto') is fed
    by the device path with every distinct k-mer; the .h5 its dump processor writes must hold what the reference dsk wrote,
    and our audit processor must have seen every distinct k-mer once, ascending, with one count per bank (several banks: one
    device job per bank, merged on the host -- the solidity kinds min / one / all and -histo2D read the per-bank counts)"""
    tmp = str(tmp_path)
    out = os.path.join(tmp, "plug")
    a = ["-file", ",".join(os.path.join(INPUTS, f) for f in t["files"]), "-kmer-size", str(t["k"]), "-abundance-min", str(t["abundance_min"]), "-out", out, "-histo", "1"]
    if t.get("abundance_max") is not None:
        a += ["-abundance-max", str(t["abundance_max"])]
    if t.get("solidity_kind"):
        a += ["-solidity-kind", t["solidity_kind"]]
    if t.get("histo2d"):
        a += ["-histo2D", "1"]
    stdout = run([PLUGIN] + a, tmp)
    audit = [ln for ln in stdout.splitlines() if ln.startswith("audit ")][-1].split()
    got = dict(zip(audit[1::2], (int(x) for x in audit[2::2])))
    if "kmers_nb_distinct" in t:                                                  # (the 'auto' goldens only carry the solid side)
        assert got["distinct"] == t["kmers_nb_distinct"] and got["occurrences"] == t["kmers_nb_valid"]
    else:
        assert got["distinct"] >= t["nb_solid"] and got["occurrences"] >= t["sum_counts"]
    assert got["unordered"] == 0 and got["bad_vectors"] == 0 and got["parts"] >= 1
    assert got["processors"] == (3 if "auto" in str(t["abundance_min"]) else 2)     # (cutoff,) default chain, audit
    lines, histo, header = read_back(out + ".h5", tmp)
    assert len(lines) == t["nb_solid"]
    m = hashlib.sha256()
    for ln in lines:
        m.update(ln + b"\n")
    assert m.hexdigest() == t["kmers_sha256"]
    rows = open(out + ".histo").read().splitlines()
    assert {r.split("\t")[0]: int(r.split("\t")[1]) for r in rows if int(r.split("\t")[1])} == t["hist"]


BANKSTATS = load_json("ref_bankstats.json")["runs"]


BANKSTATS_CLI = [t for t in BANKSTATS if t["files"][0] in ("weird.fasta", "read50x_ref10K_e001.fasta.gz", "reads.fastq", "multiline.fasta", "c1.fasta.gz", "longread.fasta", "readN.fasta")]


@need_bins
@pytest.mark.parametrize("t", BANKSTATS_CLI, ids=["+".join(t["files"]) for t in BANKSTATS_CLI])    # (every input: tests/test_gpu_round2.py, through the C ABI)
def test_cli_bank_statistics_as_the_reference_prints_them(t, tmp_path):
    """bank / sequences / kmers keys of `dsk -verbose 1` (K/SortingCountAlgorithm.cpp:728-742): same strings from `dsk_gpu`"""
    import re
    tmp = str(tmp_path)
    a = ["-file", ",".join(os.path.join(INPUTS, f) for f in t["files"]), "-kmer-size", str(t["k"]), "-abundance-min", "2", "-out", os.path.join(tmp, "o"), "-verbose", "1"]
    out = run([DSK_GPU] + a, tmp)
    for key in ("bank_total_nt", "seq_number", "seq_size_min", "seq_size_max", "seq_size_mean", "seq_size_deviation", "kmers_nb_valid", "kmers_nb_invalid"):
        m = re.search(r"^\s*%s\s*:\s*(\S+)\s*$" % re.escape(key), out, re.M)
        assert m, key
        assert m.group(1) == t[key], (key, m.group(1), t[key])


def kff_records(path):
    """(k, data_size, sorted list of (2-bit sequence bytes, data bytes)) of a KFF file whose raw sections hold one k-mer per
    block (max = 1), as CountProcessorDumpKff writes them: header, then sections 'v' (variables), 'r' (raw blocks), 'i' (index)"""
    d = open(path, "rb").read()
    assert d[:3] == b"KFF" and d[-3:] == b"KFF"
    pos = 3 + 2 + 1 + 1 + 1                                   # signature, version, encoding, uniqueness, canonicity
    free = int.from_bytes(d[pos:pos + 4], "big"); pos += 4 + free
    var, recs = {}, []
    while pos < len(d) - 3:
        t = d[pos:pos + 1]; pos += 1
        if t == b"v":
            n = int.from_bytes(d[pos:pos + 8], "big"); pos += 8
            for _ in range(n):
                e = d.index(b"\0", pos); name = d[pos:e].decode(); pos = e + 1
                var[name] = int.from_bytes(d[pos:pos + 8], "big"); pos += 8
        elif t == b"r":
            assert var["max"] == 1
            n = int.from_bytes(d[pos:pos + 8], "big"); pos += 8
            sb, db = (var["k"] * 2 + 7) // 8, var["data_size"]
            for _ in range(n):
                recs.append((d[pos:pos + sb], d[pos + sb:pos + sb + db])); pos += sb + db
        elif t == b"i":
            n = int.from_bytes(d[pos:pos + 8], "big"); pos += 8 + n * 9 + 8
        else:
            raise AssertionError("unknown KFF section %r at %d" % (t, pos - 1))
    return var["k"], var["data_size"], sorted(recs)


@need_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(REFBIN, "dsk")), reason="reference dsk binary not on this box")
@pytest.mark.parametrize("name", ["c1_k31", "c1_k63"])
def test_cli_kff_output_through_the_reference_dump_processor(name, tmp_path):
    """-kff: the job runs in plug-in mode with the reference's default processor chain (CountProcessorDumpKff included); the
    k-mers and counts in the .kff are the ones the reference dsk writes (section layout differs with the partition count),
    and the .h5 / .histo stay identical"""
    t = [r for r in RUNS if r["name"] == name][0]
    tmp = str(tmp_path)
    a, b = os.path.join(tmp, "gpu_out"), os.path.join(tmp, "ref_out")
    run([DSK_GPU] + dsk_args(t, a) + ["-kff"], tmp)
    run([os.path.join(REFBIN, "dsk")] + dsk_args(t, b) + ["-kff", "-out-tmp", tmp], tmp)
    ka = [f for f in os.listdir(tmp) if f.startswith("gpu_out") and f.endswith(".kff")]
    kb = [f for f in os.listdir(tmp) if f.startswith("ref_out") and f.endswith(".kff")]
    assert len(ka) == 1 and len(kb) == 1, (ka, kb)
    ra, rb = kff_records(os.path.join(tmp, ka[0])), kff_records(os.path.join(tmp, kb[0]))
    assert ra[0] == rb[0] == t["k"] and ra[1] == rb[1]
    assert len(ra[2]) == t["nb_solid"] and ra[2] == rb[2]
    la, ha, _ = read_back(a + ".h5", tmp)
    lb, hb, _ = read_back(b + ".h5", tmp)
    assert la == lb and ha == hb
    assert open(a + ".histo", "rb").read() == open(b + ".histo", "rb").read()


@need_bins
@pytest.mark.parametrize("given,used", [(None, "10"), ("10", "10"), ("12", "12")])
def test_cli_minimizer_size_used_is_reported(given, used, tmp_path):
    """-minimizer-size on the command line is honoured as given (IOptionsParser::saw tells an explicit 10 from the default, which
    big jobs replace by dskgpu_suggest_minimizer_size); the value in use is part of the statistics"""
    import re
    tmp = str(tmp_path)
    a = ["-file", os.path.join(INPUTS, "read50x_ref10K_e001.fasta.gz"), "-kmer-size", "31", "-out", os.path.join(tmp, "o"), "-verbose", "1"]
    if given:
        a += ["-minimizer-size", given]
    out = run([DSK_GPU] + a, tmp)
    m = re.search(r"^\s*minimizer_size_used\s*:\s*(\S+)\s*$", out, re.M)
    assert m and m.group(1) == used


HISTOMAX = load_json("ref_runs_histomax.json")["runs"]


@need_bins
@pytest.mark.parametrize("t", HISTOMAX, ids=[t["name"] for t in HISTOMAX])
def test_cli_histo_max_and_partition_balance_options(t, tmp_path):
    """-histo-max N (N < 10000): <out>.histo / <out>.histo2D byte-identical to the reference's (N lines, clamp bin empty in 1-D,
    populated in 2-D), 'auto' cutoff computed on the shortened histogram; -minimizer-type 1 / -repartition-type 1 ride along:
    they only change partition balance in the reference and nothing in the results"""
    tmp = str(tmp_path)
    out = os.path.join(tmp, "gpu_out")
    run([DSK_GPU] + dsk_args(t, out) + ["-histo-max", str(t["histo_max"]), "-minimizer-type", "1", "-repartition-type", "1"], tmp)
    assert open(out + ".histo").read() == t["histo_text"]
    if t.get("histo2d"):
        assert open(out + ".histo2D").read() == t["histo2d_text"]
    lines, _, _ = read_back(out + ".h5", tmp)
    assert len(lines) == t["nb_solid"]
    m = hashlib.sha256()
    for ln in lines:
        m.update(ln + b"\n")
    assert m.hexdigest() == t["kmers_sha256"]


WIDEBIN = os.path.join(ROOT, "oracle", "_ref", "wide", "bin")
WIDE_RUNS = [t for t in load_json("ref_runs_wide.json")["runs"] if t["name"] in ("c1_k64", "c1_k95", "longreads250_k96", "longreads250_k127", "c123_k71_all", "histo2d_k95")]


@need_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(WIDEBIN, "dsk2ascii")), reason="reference readers built with KSIZE_LIST '32 64 96 128' not on this box (oracle/build_ref_wide.sh)")
@pytest.mark.parametrize("t", WIDE_RUNS, ids=[t["name"] for t in WIDE_RUNS])
def test_cli_wide_spans_against_committed_reference_outputs(t, tmp_path):
    """spans 96 / 128 (64 <= k <= 127) through the C++ host side: GpuSortingCount<96> / <128>, dsk/solid items of 192 / 256
    bits, read back by the unmodified reference dsk2ascii built with the reference's default KSIZE_LIST"""
    tmp = str(tmp_path)
    out = os.path.join(tmp, "gpu_out")
    run([DSK_GPU] + dsk_args(t, out), tmp)
    txt = out + ".txt"
    run([os.path.join(WIDEBIN, "dsk2ascii"), "-file", out + ".h5", "-out", txt, "-verbose", "0"], tmp)
    lines = sorted(open(txt, "rb").read().splitlines())
    assert len(lines) == t["nb_solid"]
    m = hashlib.sha256()
    for ln in lines:
        m.update(ln + b"\n")
    assert m.hexdigest() == t["kmers_sha256"]
    rows = open(out + ".histo").read().splitlines()
    assert {r.split("\t")[0]: int(r.split("\t")[1]) for r in rows if int(r.split("\t")[1])} == t["hist"]


@need_bins
def test_cli_unhandled_kmer_size(tmp_path):
    p = subprocess.run([DSK_GPU, "-file", os.path.join(INPUTS, "shortread.fasta"), "-kmer-size", "128", "-out", str(tmp_path / "x")],
                       cwd=str(tmp_path), capture_output=True, text=True)
    assert p.returncode != 0 and "unhandled kmer size 128" in (p.stdout + p.stderr)


AUTO = load_json("ref_runs_auto.json")["runs"]


@need_bins
@pytest.mark.parametrize("t", AUTO, ids=[t["name"] for t in AUTO])
def test_cli_abundance_min_auto(t, tmp_path):
    """-abundance-min auto through the command line: cutoffs of the first pass and the solid set they select"""
    tmp = str(tmp_path)
    out = os.path.join(tmp, "gpu_out")
    a = dsk_args(t, out)
    a[a.index("-verbose") + 1] = "1"
    stats = run([DSK_GPU] + a, tmp).replace("\r", "\n").splitlines()
    i = [j for j, ln in enumerate(stats) if "cutoffs_auto" in ln][0]
    assert [int(x) for x in stats[i + 1].split(":", 1)[1].split()] == t["cutoffs"]
    lines, _, _ = read_back(out + ".h5", tmp)
    assert len(lines) == t["nb_solid"]
    m = hashlib.sha256()
    for ln in lines:
        m.update(ln + b"\n")
    assert m.hexdigest() == t["kmers_sha256"]
    rows = open(out + ".histo").read().splitlines()
    assert {r.split("\t")[0]: int(r.split("\t")[1]) for r in rows if int(r.split("\t")[1])} == t["hist"]


# ---------------------------------------------------------------- several devices / several passes behind the same command line
def _cli_check(t, tmp, env, extra_names=()):
    out = os.path.join(tmp, "gpu_out")
    a = dsk_args(t, out)
    a[a.index("-verbose") + 1] = "1"
    p = subprocess.run([DSK_GPU] + a, cwd=tmp, capture_output=True, text=True, env=dict(os.environ, **env))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lines, histo, header = read_back(out + ".h5", tmp)
    assert len(lines) == t["nb_solid"]
    m = hashlib.sha256()
    for ln in lines:
        m.update(ln + b"\n")
    assert m.hexdigest() == t["kmers_sha256"]
    rows = open(out + ".histo").read().splitlines()
    assert {r.split("\t")[0]: int(r.split("\t")[1]) for r in rows if int(r.split("\t")[1])} == t["hist"]
    for name in extra_names:
        assert name in header, name
    return p.stdout.replace("\r", "\n")


MULTI = [t for t in RUNS if t["name"] in ("c1_k31", "c1_k63", "reads.fastq_k31", "multiline.fasta_k31", "histo2d_c123_k31", "c123_k31_all", "longread_k63")]


@need_bins
@pytest.mark.parametrize("t", MULTI, ids=[t["name"] for t in MULTI])
def test_cli_two_contexts_byte_range_slices(t, tmp_path):
    """DSKGPU_DEVICES with two entries: one process, two contexts (here on the same GPU), each parsing a record-aligned byte
    range of every plain input file; dskgpu_multi_finish routes the super-k-mers; two output collections dsk/solid/{0,1}"""
    stats = _cli_check(t, str(tmp_path), {"DSKGPU_DEVICES": "0,0", "DSKGPU_SPLIT_MIN_BYTES": "1000"}, ('DATASET "0"', 'DATASET "1"'))
    assert "nb_devices" in stats


@need_bins
def test_cli_all_real_devices(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs in the box")
    for name in ("c1_k31", "c1_k63", "histo2d_c123_k31"):
        t = [r for r in RUNS if r["name"] == name][0]
        d = tmp_path / name
        d.mkdir()
        _cli_check(t, str(d), {"DSKGPU_DEVICES": "all", "DSKGPU_SPLIT_MIN_BYTES": "1000"}, ('DATASET "1"',))


@need_bins
@pytest.mark.parametrize("name,env", [("c1_k31", {"DSKGPU_NB_PASSES": "3"}), ("c1_k63", {"DSKGPU_NB_PASSES": "2"}),
                                      ("histo2d_c123_k31", {"DSKGPU_NB_PASSES": "2"}),
                                      ("c1_k31", {"DSKGPU_NB_PASSES": "2", "DSKGPU_DEVICES": "0,0", "DSKGPU_SPLIT_MIN_BYTES": "1000"})])
def test_cli_pass_loop(name, env, tmp_path):
    """several passes over the input (a job whose records exceed HBM): dsk/solid/<pass * devices + device>, same k-mers"""
    t = [r for r in RUNS if r["name"] == name][0]
    stats = _cli_check(t, str(tmp_path), env, ('DATASET "1"',))
    assert "nb_passes" in stats


@need_bins
def test_cli_abundance_min_auto_with_passes_and_devices(tmp_path):
    for i, env in enumerate(({"DSKGPU_NB_PASSES": "2"}, {"DSKGPU_DEVICES": "0,0", "DSKGPU_SPLIT_MIN_BYTES": "1000"})):
        t = AUTO[0]
        d = tmp_path / str(i)
        d.mkdir()
        stats = _cli_check(t, str(d), env).splitlines()
        j = [x for x, ln in enumerate(stats) if "cutoffs_auto" in ln][0]
        assert [int(x) for x in stats[j + 1].split(":", 1)[1].split()] == t["cutoffs"]


@need_bins
def test_cli_file_larger_than_one_push_block(tmp_path):
    """a 150 MB FASTA: the reader's double-buffered 64 MiB blocks are reused (dskgpu_push_sync before every refill); the
    reference binary, when it is on the box, must produce the same dump"""
    import numpy as np
    import sys
    sys.path.insert(0, ROOT)
    from dsk_b200.synth import reads_fasta
    buf, n, _ = reads_fasta(G=1_000_000, coverage=140, L=150, err=0.01, seed=5)
    fa = str(tmp_path / "big.fa")
    buf[:n].tofile(fa)
    assert n > (140 << 20)
    tmp = str(tmp_path)
    a, b = os.path.join(tmp, "gpu_out"), os.path.join(tmp, "ref_out")
    args = ["-file", fa, "-kmer-size", "31", "-abundance-min", "3", "-histo", "1", "-verbose", "0"]
    run([DSK_GPU] + args + ["-out", a], tmp)
    la, ha, _ = read_back(a + ".h5", tmp)
    if os.path.exists(os.path.join(REFBIN, "dsk")):
        run([os.path.join(REFBIN, "dsk")] + args + ["-out", b, "-out-tmp", tmp], tmp)
        lb, hb, _ = read_back(b + ".h5", tmp)
        assert la == lb and ha == hb
        assert open(a + ".histo", "rb").read() == open(b + ".histo", "rb").read()
    # ... and the same file through two contexts with real byte-range slices
    c = os.path.join(tmp, "gpu2_out")
    p = subprocess.run([DSK_GPU] + args + ["-out", c], cwd=tmp, capture_output=True, text=True, env=dict(os.environ, DSKGPU_DEVICES="0,0"))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lc, hc, _ = read_back(c + ".h5", tmp)
    assert la == lc and ha == hc
