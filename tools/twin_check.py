#!/usr/bin/env python
"""Scaled-twin parity check for the 3 Gbp configurations: a G = 100 Mbp twin of BASELINE configs[2]/[3]
(same 30x coverage, 150 bp reads, 1 % error) is counted by the CUDA path and by the UNMODIFIED reference `dsk` binary on
the same FASTA, and the two are compared on everything the reference reports without a multi-GB text dump: the abundance
histogram (10 000 bins), kmers_nb_valid, kmers_nb_distinct, kmers_nb_solid.

  python tools/twin_check.py [--genome 100000000] [--coverage 30] [--kmer-size 31|63]

Run on the GPU box (needs oracle/_ref/bin/dsk, which travels with the snapshot).  Test infrastructure: the reference is the
checker here, never on the product path.  Written at the end of round 1; first on the list for round 2's GPU budget."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--coverage", type=int, default=30)
    ap.add_argument("--kmer-size", type=int, default=31)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    import numpy as np
    import torch
    from dsk_b200 import GpuCounter, _lib
    from dsk_b200.synth import reads_fasta_device, genome_device
    from oracle.pyoracle import _ref_bin, stat_value

    g = genome_device(args.genome, seed=args.seed, device="cuda")
    dev, nreads = reads_fasta_device(args.genome, args.coverage, 150, 0.01, seed=args.seed + 1, device="cuda", genome=g)
    del g
    n = dev.numel()
    job_kmers = nreads * (150 - args.kmer_size + 1)
    m = _lib.lib().dskgpu_suggest_minimizer_size(job_kmers, args.kmer_size)
    t0 = time.time()
    with GpuCounter(kmer_size=args.kmer_size, abundance_min=2, minimizer_size=m, keep_results_on_device=True) as eng:
        eng.push_device_bytes(dev.data_ptr(), n, fmt="fasta")
        eng.finish()
        st = eng.stats()
        hist = eng.histogram()[0]
    t_gpu = time.time() - t0

    tmp = tempfile.mkdtemp(prefix="dsktwin_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fa = os.path.join(tmp, "twin.fa")
    try:
        host = torch.empty(n, dtype=torch.uint8)
        host.copy_(dev)
        host.numpy().tofile(fa)
        del host
        out = os.path.join(tmp, "ref")
        cmd = [_ref_bin("dsk"), "-file", fa, "-kmer-size", str(args.kmer_size), "-abundance-min", "2", "-histo", "1",
               "-out", out, "-out-tmp", tmp, "-out-dir", tmp, "-verbose", "1", "-nb-cores", str(os.cpu_count() or 1)]
        t0 = time.time()
        p = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
        t_ref = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError("reference dsk failed: " + p.stderr[-1000:])
        rh = np.zeros(10001, np.uint64)
        for line in open(out + ".histo"):
            a, b = line.split()
            rh[int(a)] = int(b)
        ref = {k: int(stat_value(p.stdout, k) or -1) for k in ("kmers_nb_valid", "kmers_nb_distinct", "kmers_nb_solid")}
    finally:
        subprocess.run(["rm", "-rf", tmp])
    ours = {k: int(st[k]) for k in ref}
    ok = ours == ref and bool((hist == rh).all())
    print(json.dumps({"twin": "G=%d, %dx, 150 bp, 1%% error, k=%d" % (args.genome, args.coverage, args.kmer_size), "minimizer_size": m,
                      "ours": ours, "reference": ref, "histogram_identical": bool((hist == rh).all()),
                      "histogram_bins_differing": int((hist != rh).sum()), "seconds_gpu_path": round(t_gpu, 2),
                      "seconds_reference_dsk": round(t_ref, 1), "parity": "OK" if ok else "MISMATCH"}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
