#!/usr/bin/env python
"""Golden vectors for the per-sequence statistics of SortingCountAlgorithm::getInfo() (K/SortingCountAlgorithm.cpp:733-742,
BankStats K/BankKmers.hpp:166-215): seq_number, seq_size_min/max/mean/deviation, kmers_nb_valid, kmers_nb_invalid as the
unmodified reference `dsk -verbose 1` prints them.  Runs ONLY in the authoring container (oracle/_ref/bin/dsk); the test
suite reads the committed tests/golden/ref_bankstats.json."""
import json
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
INP = os.path.join(HERE, "inputs")
KEYS = ["bank_total_nt", "seq_number", "seq_size_min", "seq_size_max", "seq_size_mean", "seq_size_deviation", "kmers_nb_valid", "kmers_nb_invalid"]
CASES = [(["weird.fasta"], 31), (["read50x_ref10K_e001.fasta.gz"], 31), (["reads.fastq"], 21), (["multiline.fasta"], 31), (["varlen.fasta"], 31),
         (["readN.fasta"], 27), (["IUPAC.fasta"], 11), (["longread.fasta"], 63), (["lowcomplexity.fasta"], 31), (["shortread.fasta"], 31),
         (["reads_plusname.fastq"], 31), (["c1.fasta.gz", "c2.fasta.gz", "c3.fasta.gz"], 31), (["assembly.fasta", "asm_reads.fasta"], 31),
         (["longreads250.fasta.gz"], 31)]
runs = []
for files, k in CASES:
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [os.path.join(BIN, "dsk"), "-file", ",".join(os.path.join(INP, f) for f in files), "-kmer-size", str(k), "-abundance-min", "2",
               "-out", os.path.join(tmp, "o"), "-verbose", "1", "-nb-cores", "2", "-out-tmp", tmp]
        out = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True).stdout
        r = {"files": files, "k": k}
        for key in KEYS:
            m = re.search(r"^\s*%s\s*:\s*(\S+)\s*$" % re.escape(key), out, re.M)
            if m:
                r[key] = m.group(1)
        runs.append(r)
        print(files, k, {x: r.get(x) for x in KEYS})
json.dump({"source": "oracle/_ref/bin/dsk -verbose 1 (unmodified reference), tests/golden/make_golden_bankstats.py", "runs": runs},
          open(os.path.join(HERE, "ref_bankstats.json"), "w"), indent=1)
