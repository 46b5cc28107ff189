#!/usr/bin/env python
"""Regenerates tests/golden/* .  Runs ONLY in the authoring container (needs /root/reference and the
reference binaries built by oracle/build_ref.sh).  The GPU box and the test-suite only read the
committed outputs.

What it writes
  inputs/*                 test inputs: the reference's own fixtures (R/test/*, gz re-compressed so the
                           payload is byte-identical but the container is ours) + synthetic edge cases
  ref_shell_tests.json     expectations of R/scripts/simple_test.sh:35-135 (sparse histograms, dsk2ascii text)
  ref_unit_vectors.json    literal known answers transcribed *programmatically* from
                           G/test/unit/src/kmer/TestDSK.cpp and TestKmer.cpp
  ref_runs.json            outputs of the real reference `dsk` on every input (solid k-mer digest,
                           counts, sparse histogram, stats) -- the "outputs of the reference run here" pin
  ref_runs_auto.json       same for `-abundance-min auto` (cutoffs of the first pass + the solid set they select)
"""
import gzip
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.pyoracle import run_reference, stat_value, ref_available  # noqa: E402

R = "/root/reference/"
G = R + "thirdparty/gatb-core/gatb-core/"
INP = os.path.join(HERE, "inputs")
os.makedirs(INP, exist_ok=True)


def sparse(h):
    return {str(i): int(v) for i, v in enumerate(h) if v}


def read_histo(path):
    h = np.zeros(10001, np.uint64)
    for line in open(path):
        a, b = line.split()
        h[int(a)] = int(b)
    return h


def regz(src, dst):
    data = gzip.open(src, "rb").read()
    with open(dst, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as g:
            g.write(data)


def kmer_digest(kmers):
    """kmers: sorted list of (str, count). digest = sha256 of 'KMER count\n' lines (LC_ALL=C order)."""
    m = hashlib.sha256()
    for s, c in kmers:
        m.update(("%s %d\n" % (s, c)).encode())
    return m.hexdigest()


# ------------------------------------------------------------------ 1. fixtures of the shell tests
def shell_tests():
    for name in ("read50x_ref10K_e001.fasta.gz", "c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz", "c4.fasta.gz"):
        regz(R + "test/" + name, os.path.join(INP, name))
    for name in ("longread.fasta", "readN.fasta", "shortread.fasta", "IUPAC.fasta"):
        open(os.path.join(INP, name), "wb").write(open(R + "test/" + name, "rb").read())
    out = {
        "source": "R/scripts/simple_test.sh:35-135",
        "tests": [
            {"name": "k27", "files": ["read50x_ref10K_e001.fasta.gz"], "k": 27, "abundance_min": 2,
             "hist": sparse(read_histo(R + "test/k27.histo"))},
            {"name": "k27_multifile", "files": ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz", "c4.fasta.gz"], "k": 27,
             "abundance_min": 2, "hist": sparse(read_histo(R + "test/k27.histo"))},
            {"name": "longread", "files": ["longread.fasta"], "k": 27, "abundance_min": 2,
             "hist": sparse(read_histo(R + "test/rlong.histo"))},
            {"name": "shortread", "files": ["shortread.fasta"], "k": 15, "abundance_min": 1,
             "dsk2ascii": open(R + "test/short.parse_results").read()},
            {"name": "k_gt_readlen", "files": ["shortread.fasta"], "k": 16, "abundance_min": 1, "dsk2ascii": ""},
            {"name": "readN", "files": ["readN.fasta"], "k": 20, "abundance_min": 2,
             "hist": sparse(read_histo(R + "test/readN.histo"))},
        ],
    }
    json.dump(out, open(os.path.join(HERE, "ref_shell_tests.json"), "w"), indent=1)


# ------------------------------------------------------------------ 2. unit-test literals
def c_strings(block):
    """concatenate adjacent C string literals separated by commas -> list of strings"""
    items, cur = [], None
    for tok in re.finditer(r'"([^"]*)"|(,)', block):
        if tok.group(1) is not None:
            cur = (cur or "") + tok.group(1)
        else:
            if cur is not None:
                items.append(cur)
            cur = None
    if cur is not None:
        items.append(cur)
    return items


def unit_vectors():
    dsk = open(G + "test/unit/src/kmer/TestDSK.cpp").read()
    kmer = open(G + "test/unit/src/kmer/TestKmer.cpp").read()
    out = {"source": "G/test/unit/src/kmer/TestDSK.cpp, TestKmer.cpp"}

    # --- DSK_check1 (TestDSK.cpp:147-241)
    body = dsk[dsk.index("void DSK_check1 ()"):dsk.index("void DSK_check2_aux ()")]
    s1 = re.search(r'const char\* s1 = "([ACGT]+)"', body).group(1)
    seqsets = {"seqs1": [s1], "seqs2": [s1, s1], "seqs3": [s1, s1, s1]}
    blk = body[body.index("const char* seqs4[] = {"):body.index("} ;")]
    seqsets["seqs4"] = c_strings(blk[blk.index("{") + 1:])
    checks = []
    for m in re.finditer(r"DSK_check1_aux \((seqs\d), ARRAY_SIZE\(seqs\d\), (\d+), (\d+), (\d+)\);", body):
        checks.append({"seqs": m.group(1), "k": int(m.group(2)), "nks": int(m.group(3)), "nb_solid": int(m.group(4))})
    out["DSK_check1"] = {"seqsets": seqsets, "checks": checks}

    # --- DSK_check2 (TestDSK.cpp:244-341)
    body = dsk[dsk.index("void DSK_check2_aux ()"):dsk.index("void DSK_check2 ()")]
    seq = re.search(r'const char\* s1 = "([ACGT]+)"', body).group(1)
    vals = sorted(set(re.findall(r"0x[0-9a-fA-F]{15,16}", body)))
    out["DSK_check2"] = {"seq": seq, "k": 31, "nks": 1, "hex_literals": vals}

    # --- DSK_perBank1/2 (TestDSK.cpp:482-612)
    for name, nxt in (("DSK_perBank1", "void DSK_perBank2 ()"), ("DSK_perBank2", "void DSK_perBankKmer_aux")):
        body = dsk[dsk.index("void %s ()" % name):dsk.index(nxt)]
        blk = body[body.index("const char* seqs[] = {"):body.index("};")]
        blk = re.sub(r"//.*", "", blk)
        seqs = c_strings(blk[blk.index("{") + 1:])
        checks = []
        for m in re.finditer(r"DSK_perBank_aux<KSIZE_1> \(album, (\d+), (\d+), (\w+), KMER_SOLIDITY_(\w+), (\d+)\);", body):
            mx = m.group(3)
            checks.append({"k": int(m.group(1)), "min": int(m.group(2)), "max": (1 << 30) if mx == "nksMax" else int(mx),
                           "kind": m.group(4).lower(), "nb_solid": int(m.group(5))})
        out[name] = {"banks": seqs, "checks": checks}

    # --- TestKmer: direct / canonical 3-mers (TestKmer.cpp:141-190)
    body = kmer[kmer.index("void kmer_checkInfo ()") if "void kmer_checkInfo ()" in kmer else 0:]
    m = re.search(r'const char\* seq = "(CATTGATAGTGG)"', kmer)
    out["kmer3"] = {"seq": m.group(1), "k": 3}
    for key, pat in (("direct", r"long checkDirect \[\]\s*=\s*\{([^}]*)\}"), ("canonical", r"long checkBoth \[\]\s*=\s*\{([^}]*)\}")):
        mm = re.search(pat, kmer)
        out["kmer3"][key] = [int(x) for x in re.findall(r"\d+", mm.group(1))]
    # --- canonical 5-mers (TestKmer.cpp:233-261)
    i0 = kmer.index("void kmer_build ()")
    body = kmer[i0:i0 + 3000]
    seq = re.search(r'"(ACTACGATCGATGTA)"', body).group(1)
    vals = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", re.search(r"check\[\] = \{([^}]*)\}", body).group(1))]
    out["kmer5"] = {"seq": seq, "k": 5, "canonical": vals}
    # --- minimizer canonical table (TestKmer.cpp:437-469)
    i0 = kmer.index("void kmer_minimizer3 ()")
    body = kmer[i0:i0 + 4000]
    seq = re.search(r'const char\* seq = "([ACGT]+)"', body).group(1)
    rows = re.findall(r'\{"([ACGT]+)",\s*"([ACGT]+)",\s*(\d+),\s*(true|false)\s*\}', body)
    out["minimizer3"] = {"seq": seq, "k": 15, "m": 7,
                         "rows": [{"kmer": a, "minimizer": b, "position": int(c), "changed": d == "true"} for a, b, c, d in rows]}
    # --- bad char table (TestKmer.cpp:509-569)
    i0 = kmer.index("void kmer_badchar (void)")
    body = kmer[i0:i0 + 4000]
    seq = re.search(r'const char\* seq = "([ACGTN]+)"', body).group(1)
    rows = re.findall(r'\{"([ACGTN]+)",\s*(true|false)\s*\}', body)
    out["badchar"] = {"seq": seq, "k": 11, "rows": [{"kmer": a, "valid": b == "true"} for a, b in rows]}
    json.dump(out, open(os.path.join(HERE, "ref_unit_vectors.json"), "w"), indent=1)


# ------------------------------------------------------------------ 3. synthetic edge-case inputs + reference runs
def synth_inputs():
    rng = np.random.default_rng(1234)
    ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return ACGT[rng.integers(0, 4, n)].tobytes()

    genome = rnd(3000)
    comp = bytes.maketrans(b"ACGT", b"TGCA")

    def reads(n, L, err=0.02):
        out = []
        for _ in range(n):
            s = int(rng.integers(0, len(genome) - L))
            r = bytearray(genome[s:s + L])
            for j in np.nonzero(rng.random(L) < err)[0]:
                r[j] = ACGT[(int(np.where(ACGT == r[j])[0][0]) + int(rng.integers(1, 4))) % 4]
            r = bytes(r)
            if rng.random() < 0.5:
                r = r.translate(comp)[::-1]
            out.append(r)
        return out

    files = {}
    rs = reads(400, 100)
    # multi-line FASTA, 60 columns, some lowercase, some N, an empty line, CRLF on a few records
    buf = bytearray()
    for i, r in enumerate(rs):
        if i % 7 == 0:
            r = r.lower()
        if i % 11 == 0:
            r = r[:40] + b"N" + r[41:70] + b"nn" + r[72:]
        eol = b"\r\n" if i % 13 == 0 else b"\n"
        buf += b">read%d some comment ACGT" % i + eol
        for j in range(0, len(r), 60):
            buf += r[j:j + 60] + eol
        if i % 17 == 0:
            buf += b"\n"
    files["multiline.fasta"] = bytes(buf)
    # FASTQ, 4 lines, quality lines starting with '@' and '+' and '>' on purpose
    buf = bytearray()
    for i, r in enumerate(reads(400, 90)):
        q = bytearray(rng.integers(33, 74, len(r)).astype(np.uint8).tobytes())
        if i % 3 == 0:
            q[0] = ord("@")
        if i % 5 == 0:
            q[0] = ord("+")
        if i % 7 == 0:
            q[0] = ord(">")
        buf += b"@fq%d/1\n" % i + r + b"\n+\n" + bytes(q) + b"\n"
    files["reads.fastq"] = bytes(buf)
    # FASTQ with header repeated on '+' line and a final record without trailing newline
    buf = bytearray()
    for i, r in enumerate(reads(100, 75)):
        q = bytes(rng.integers(40, 74, len(r)).astype(np.uint8).tobytes())
        buf += b"@x%d\n" % i + r + b"\n+x%d\n" % i + q + b"\n"
    files["reads_plusname.fastq"] = bytes(buf[:-1])
    # reads shorter than k mixed with long ones, variable length, no trailing newline
    buf = bytearray()
    for i in range(300):
        L = int(rng.integers(1, 140))
        buf += b">v%d\n" % i + rnd(L) + b"\n"
    files["varlen.fasta"] = bytes(buf[:-1])
    # low complexity: poly-A / dinucleotide repeats (skewed minimizers, large counts, histogram wrap)
    buf = b">polyA\n" + b"A" * 70030 + b"\n>acgt\n" + b"ACGT" * 3000 + b"\n>polyA2\n" + b"A" * 20030 + b"\n"
    files["lowcomplexity.fasta"] = buf
    # garbage before first header, IUPAC letters, tabs/spaces inside sequence lines
    buf = b"garbage line\nACGT>hdr1 x\nACGTACGTAC GTACGTACGTACGTRYACGTAGCTAGCTAGCATCGATCGATCGATCGATCAGCTAGCTAGCTAGCATCG\n" \
          b">hdr2\nACGATCGATCGACTAGCTAGCTAGCTAG\tCTAGCTAGCTAGCTAGCATGCATGCATGCATGCATGCATGCATGCAT\n>empty\n>last\n" + rnd(200) + b"\n"
    files["weird.fasta"] = buf
    # 2 banks for histo2D: assembly (genome, multi-line) + reads
    asm = bytearray(b">contig1\n")
    for j in range(0, len(genome), 70):
        asm += genome[j:j + 70] + b"\n"
    files["assembly.fasta"] = bytes(asm)
    buf = bytearray()
    for i, r in enumerate(reads(1500, 100, 0.01)):
        buf += b">r%d\n" % i + r + b"\n"
    files["asm_reads.fasta"] = bytes(buf)
    for name, data in files.items():
        open(os.path.join(INP, name), "wb").write(data)
    return sorted(files)


def ref_runs(synth):
    assert ref_available(), "build the reference first: oracle/build_ref.sh"
    runs = []

    def one(name, files, k, amin=2, histo2d=False, **kw):
        paths = [os.path.join(INP, f) for f in files]
        r = run_reference(paths, k, abundance_min=amin, histo=True, histo2d=histo2d, nb_cores=2, **kw)
        ent = {"name": name, "files": files, "k": k, "abundance_min": amin, "histo2d": histo2d,
               "nb_solid": len(r["kmers"]), "kmers_sha256": kmer_digest(r["kmers"]),
               "sum_counts": int(sum(c for _, c in r["kmers"])),
               "first_kmers": r["kmers"][:3], "hist": sparse(r["hist"]),
               "kmers_nb_valid": int(stat_value(r["stats"], "kmers_nb_valid") or 0),
               "kmers_nb_distinct": int(stat_value(r["stats"], "kmers_nb_distinct") or 0),
               "kmers_nb_solid": int(stat_value(r["stats"], "kmers_nb_solid") or 0)}
        ent.update({k_: v for k_, v in kw.items() if k_ in ("solidity_kind", "abundance_max")})
        if histo2d:
            h2 = r["hist2d"]
            ent["hist2d"] = {"%d,%d" % (j, i): int(h2[j, i]) for j in range(h2.shape[0]) for i in range(h2.shape[1]) if h2[j, i]}
        runs.append(ent)
        print(name, ent["nb_solid"], ent["kmers_nb_valid"], flush=True)

    c1 = ["read50x_ref10K_e001.fasta.gz"]
    for k in (31, 63, 21, 32, 33, 15, 11, 47):
        one("c1_k%d" % k, c1, k)
    one("c1_k31_min1", c1, 31, amin=1)
    one("c1_k31_min3_max20", c1, 31, amin=3, abundance_max=20)
    one("c1234_k31", ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz", "c4.fasta.gz"], 31)
    one("longread_k31", ["longread.fasta"], 31)
    one("longread_k63", ["longread.fasta"], 63)
    one("iupac_k11", ["IUPAC.fasta"], 11, amin=1)
    for f in synth:
        if f in ("assembly.fasta", "asm_reads.fasta"):
            continue
        for k in (31, 63) if f != "varlen.fasta" else (21, 31, 63):
            one("%s_k%d" % (f, k), [f], k, amin=1 if "weird" in f or "varlen" in f else 2)
    one("histo2d_k31", ["assembly.fasta", "asm_reads.fasta"], 31, histo2d=True)
    one("histo2d_c123_k31", ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 31, histo2d=True)
    for kind in ("min", "max", "one", "all"):
        one("c123_k31_%s" % kind, ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 31, solidity_kind=kind)
    json.dump({"source": "oracle/_ref/bin/dsk (unmodified reference, built by oracle/build_ref.sh)", "runs": runs},
              open(os.path.join(HERE, "ref_runs.json"), "w"), indent=1)


def auto_cutoffs_of(stats):
    """`cutoffs_auto / values` of the -verbose 1 stats block (CountProcessorCutoff::getProperties)"""
    lines = stats.replace("\r", "\n").splitlines()
    for i, ln in enumerate(lines):
        if "cutoffs_auto" in ln:
            return [int(x) for x in lines[i + 1].split(":", 1)[1].split()]
    return None


def ref_runs_auto():
    """-abundance-min auto: the reference's two-pass cutoff chain on committed inputs"""
    assert ref_available(), "build the reference first: oracle/build_ref.sh"
    c1 = "read50x_ref10K_e001.fasta.gz"
    cases = [("auto_c1_k31", [c1], 31, "auto", None), ("auto_c1_k63", [c1], 63, "auto", None),
             ("auto_c1_k15", [c1], 15, "auto", None), ("auto_c1_k11", [c1], 11, "auto", None),
             ("auto_asmreads_k21", ["asm_reads.fasta"], 21, "auto", None),
             ("auto_lowcomplexity_k31", ["lowcomplexity.fasta"], 31, "auto", None),
             ("auto_c123_k31_sum", ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 31, "auto", None),
             ("auto_c123_k31_all", ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 31, "auto", "all"),
             ("auto_c123_k15_one", ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 15, "auto", "one"),
             ("auto_c123_k15_mixed_all", ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 15, "2,auto", "all"),
             ("auto_c123_k21_max", ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 21, "auto", "max"),
             ("auto_asm_reads_k21_one", ["assembly.fasta", "asm_reads.fasta"], 21, "auto,auto", "one")]
    runs = []
    for name, files, k, amin, kind in cases:
        r = run_reference([os.path.join(INP, f) for f in files], k, abundance_min=amin, histo=True, nb_cores=2, solidity_kind=kind)
        ent = {"name": name, "files": files, "k": k, "abundance_min": amin, "solidity_kind": kind or "sum",
               "cutoffs": auto_cutoffs_of(r["stats"]), "nb_solid": len(r["kmers"]), "kmers_sha256": kmer_digest(r["kmers"]),
               "sum_counts": int(sum(c for _, c in r["kmers"])), "hist": sparse(r["hist"]),
               "kmers_nb_solid": int(stat_value(r["stats"], "kmers_nb_solid") or 0)}
        runs.append(ent)
        print(name, ent["cutoffs"], ent["nb_solid"], flush=True)
    json.dump({"source": "oracle/_ref/bin/dsk -abundance-min auto (unmodified reference)", "runs": runs},
              open(os.path.join(HERE, "ref_runs_auto.json"), "w"), indent=1)


if __name__ == "__main__":
    if "--auto-only" in sys.argv:
        ref_runs_auto()
        sys.exit(0)
    shell_tests()
    unit_vectors()
    if "--units-only" not in sys.argv:
        s = synth_inputs()
        ref_runs(s)
        ref_runs_auto()
