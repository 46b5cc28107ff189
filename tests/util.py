"""Shared helpers for the test-suite (golden loading, k-mer <-> string, digests)."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
INPUTS = os.path.join(GOLDEN, "inputs")

NT = "ACTG"  # code 0..3  (G/src/gatb/tools/misc/api/Data.hpp:185)


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def kmer_to_str(lo, hi, k):
    v = (int(hi) << 64) | int(lo)
    return "".join(NT[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def str_to_kmer(s):
    v = 0
    for ch in s:
        v = (v << 2) | NT.index(ch)
    return v


def kmers_to_strings(lo, hi, k):
    """vectorised conversion of arrays of k-mer values to a list of python strings"""
    n = len(lo)
    if n == 0:
        return []
    out = np.empty((n, k), dtype=np.uint8)
    lut = np.frombuffer(NT.encode(), dtype=np.uint8)
    lo = np.asarray(lo, dtype=np.uint64)
    hi = np.asarray(hi, dtype=np.uint64)
    for i in range(k):
        sh = 2 * (k - 1 - i)
        if sh >= 64:
            c = (hi >> np.uint64(sh - 64)) & np.uint64(3)
        else:
            c = (lo >> np.uint64(sh)) & np.uint64(3)
        out[:, i] = lut[c.astype(np.int64)]
    return [row.tobytes().decode() for row in out]


def digest(lo, hi, counts, k):
    """sha256 over 'KMER count\\n' lines sorted in C locale (same as `dsk2ascii | LC_ALL=C sort`)."""
    strs = kmers_to_strings(lo, hi, k)
    pairs = sorted(zip(strs, (int(c) for c in counts)))
    m = hashlib.sha256()
    for s, c in pairs:
        m.update(("%s %d\n" % (s, c)).encode())
    return m.hexdigest(), pairs


def sparse_hist(h):
    return {str(i): int(v) for i, v in enumerate(h) if v}


def sparse_hist2d(h2):
    return {"%d,%d" % (j, i): int(h2[j, i]) for j, i in zip(*np.nonzero(h2))}


def read_input(name):
    import gzip
    p = os.path.join(INPUTS, name)
    with open(p, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":
        return gzip.open(p, "rb").read()
    return open(p, "rb").read()


def kmer_words_to_strings(words, k):
    """uint64[n, W] value words (least significant first, W up to 4) -> list of k-mer strings"""
    words = np.asarray(words, dtype=np.uint64)
    n = len(words)
    if n == 0:
        return []
    out = np.empty((n, k), dtype=np.uint8)
    lut = np.frombuffer(NT.encode(), dtype=np.uint8)
    for i in range(k):
        sh = 2 * (k - 1 - i)
        c = (words[:, sh // 64] >> np.uint64(sh % 64)) & np.uint64(3)
        out[:, i] = lut[c.astype(np.int64)]
    return [row.tobytes().decode() for row in out]


def digest_words(words, counts, k):
    """sha256 over 'KMER count\\n' lines sorted in C locale, for keys given as 64-bit words (any span)"""
    strs = kmer_words_to_strings(words, k)
    pairs = sorted(zip(strs, (int(c) for c in counts)))
    m = hashlib.sha256()
    for s, c in pairs:
        m.update(("%s %d\n" % (s, c)).encode())
    return m.hexdigest(), pairs
