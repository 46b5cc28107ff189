#!/usr/bin/env python
"""Regenerates tests/golden/ref_runs_wide.json (+ inputs/longreads250.fasta): outputs of the UNMODIFIED reference built with
KSIZE_LIST "32 64 96 128" (oracle/build_ref_wide.sh -> oracle/_ref/wide/bin) for 64 <= k <= 127, the spans SURVEY.md
8(f)-4 lists.  Runs ONLY in the authoring container; the test-suite reads the committed outputs.  These vectors pin the
wide build of the oracle (oracle/liboracle_wide.so = dsk_oracle.c with -DORC_WIDE, 256-bit keys)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.pyoracle import run_reference, stat_value, ref_wide_available  # noqa: E402
from make_golden import sparse, kmer_digest, INP  # noqa: E402


def synth_inputs():
    """250-bp reads at 25x over a 20 kb random genome, 1 % substitutions, a few N / lower-case / CRLF / multi-line records"""
    rng = np.random.default_rng(20261017)
    genome = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 20000).tobytes()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    buf = bytearray()
    for i in range(2000):
        s = int(rng.integers(0, len(genome) - 250))
        r = bytearray(genome[s:s + 250])
        if rng.random() < 0.5:
            r = bytearray(bytes(r).translate(comp)[::-1])
        for j in np.nonzero(rng.random(250) < 0.01)[0]:
            r[j] = int(rng.choice(list(b"ACGT".replace(bytes([r[j]]), b""))))
        if i % 37 == 0:
            r[100] = ord("N")
        if i % 41 == 0:
            r = bytearray(bytes(r).lower())
        eol = b"\r\n" if i % 29 == 0 else b"\n"
        buf += b">lr%d" % i + eol
        if i % 11 == 0:
            for j in range(0, 250, 80):
                buf += r[j:j + 80] + eol
        else:
            buf += r + eol
    import gzip
    with open(os.path.join(INP, "longreads250.fasta.gz"), "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as g:
            g.write(bytes(buf))


def main():
    assert ref_wide_available(), "build the wide reference first: oracle/build_ref_wide.sh"
    synth_inputs()
    runs = []

    def one(name, files, k, amin=2, histo2d=False, **kw):
        paths = [os.path.join(INP, f) for f in files]
        r = run_reference(paths, k, abundance_min=amin, histo=True, histo2d=histo2d, nb_cores=2, wide=True, **kw)
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
    for k in (64, 65, 71, 95, 96, 97):
        one("c1_k%d" % k, c1, k)
    for k in (64, 95, 96, 111, 127):
        one("longreads250_k%d" % k, ["longreads250.fasta.gz"], k)
    one("longreads250_k127_min1", ["longreads250.fasta.gz"], 127, amin=1)
    for k in (95, 127):
        one("longread_k%d" % k, ["longread.fasta"], k)
        one("lowcomplexity_k%d" % k, ["lowcomplexity.fasta"], k)
        one("multiline_k%d" % k, ["multiline.fasta"], k)
    one("histo2d_k95", ["assembly.fasta", "asm_reads.fasta"], 95, histo2d=True)
    one("c123_k71_all", ["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 71, solidity_kind="all")
    json.dump({"source": "oracle/_ref/wide/bin/dsk (unmodified reference, KSIZE_LIST 32 64 96 128, built by oracle/build_ref_wide.sh)",
               "runs": runs}, open(os.path.join(HERE, "ref_runs_wide.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
