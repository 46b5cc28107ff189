#!/usr/bin/env python
"""Golden vectors for `-histo-max N` (N != 10000): the <out>.histo / <out>.histo2D text files and the solid-set digest of the
unmodified reference `dsk` (oracle/_ref/bin, built by oracle/build_ref.sh).  Runs ONLY in the authoring container; the test
suite reads the committed tests/golden/ref_runs_histomax.json.  Reference: Histogram.hpp:92-98,221 (clamp to `length`, the
clamp bin is never merged), CountProcessorHistogram.hpp:104-159, K/SortingCountAlgorithm.cpp:213."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
INP = os.path.join(HERE, "inputs")

CASES = [
    dict(name="c1_k31_histomax50", files=["read50x_ref10K_e001.fasta.gz"], k=31, abundance_min=2, histo_max=50),
    dict(name="c1_k31_histomax5", files=["read50x_ref10K_e001.fasta.gz"], k=31, abundance_min=2, histo_max=5),
    dict(name="lowcomplexity_k31_histomax1000", files=["lowcomplexity.fasta"], k=31, abundance_min=2, histo_max=1000),
    dict(name="c1_k63_auto_histomax100", files=["read50x_ref10K_e001.fasta.gz"], k=63, abundance_min="auto", histo_max=100),
    dict(name="histo2d_k31_histomax30", files=["assembly.fasta", "asm_reads.fasta"], k=31, abundance_min=2, histo_max=30, histo2d=True),
]

runs = []
for c in CASES:
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "o")
        cmd = [os.path.join(BIN, "dsk"), "-file", ",".join(os.path.join(INP, f) for f in c["files"]), "-kmer-size", str(c["k"]),
               "-abundance-min", str(c["abundance_min"]), "-histo-max", str(c["histo_max"]), "-out", out, "-histo", "1", "-verbose", "0",
               "-nb-cores", "2", "-out-tmp", tmp]
        if c.get("histo2d"):
            cmd += ["-histo2D", "1"]
        subprocess.run(cmd, check=True, cwd=tmp, capture_output=True)
        subprocess.run([os.path.join(BIN, "dsk2ascii"), "-file", out + ".h5", "-out", out + ".txt", "-verbose", "0"], check=True, cwd=tmp, capture_output=True)
        lines = sorted(open(out + ".txt", "rb").read().splitlines())
        m = hashlib.sha256()
        for ln in lines:
            m.update(ln + b"\n")
        r = dict(c)
        r["nb_solid"] = len(lines)
        r["kmers_sha256"] = m.hexdigest()
        r["histo_text"] = open(out + ".histo").read()
        if c.get("histo2d"):
            r["histo2d_text"] = open(out + ".histo2D").read()
        runs.append(r)
        print(c["name"], len(lines), len(r["histo_text"].splitlines()), "histo lines")
json.dump({"source": "oracle/_ref/bin/dsk (unmodified reference), tests/golden/make_golden_histomax.py", "runs": runs},
          open(os.path.join(HERE, "ref_runs_histomax.json"), "w"), indent=1)
