"""ctypes wrapper over oracle/liboracle.so (dsk_oracle.c) + runner for the reference binaries
in oracle/_ref/bin (when present).  Test infrastructure only."""
import ctypes as C
import gzip
import os
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_WIDE = None

KINDS = {"sum": 0, "min": 1, "max": 2, "one": 3, "all": 4, "custom": 5}


def _lib(k=31):
    """liboracle.so (128-bit keys, k <= 63: the pinned restatement of KSIZE_LIST "32 64") or, for k >= 64,
    liboracle_wide.so (same source built with -DORC_WIDE: 256-bit keys, spans 96 and 128)."""
    global _LIB, _LIB_WIDE
    wide = k >= 64
    if not wide and _LIB is not None:
        return _LIB
    if wide and _LIB_WIDE is not None:
        return _LIB_WIDE
    name = "liboracle_wide.so" if wide else "liboracle.so"
    so = os.path.join(ORACLE_DIR, name)
    src = os.path.join(ORACLE_DIR, "dsk_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", name])
    L = C.CDLL(so)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_add_sequence.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
    L.orc_add_file.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.orc_finish.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int64, C.c_int, C.c_char_p, C.c_int]
    for name in ("nb_valid", "nb_invalid", "nb_seq", "nb_nt", "nb_superkmers", "nb_distinct", "nb_solid", "seq_min", "seq_max", "seq_sumsq"):
        f = getattr(L, "orc_" + name)
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p]
    for name in ("keys_lo", "keys_hi", "keys_w2", "keys_w3", "counts", "sums", "solid", "hist", "hist2d"):
        f = getattr(L, "orc_" + name)
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p]
    L.orc_kmers_of.restype = C.c_int
    L.orc_kmers_of.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
    L.orc_kmers_of4.restype = C.c_int
    L.orc_kmers_of4.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
    L.orc_max_k.restype = C.c_int
    L.orc_parse_stats.restype = C.c_uint64
    L.orc_parse_stats.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_void_p, C.c_size_t]
    L.orc_mmer_lut.restype = C.c_uint32
    L.orc_mmer_lut.argtypes = [C.c_uint32, C.c_int]
    if not wide:
        L.orc_parse_dump.restype = C.c_int64
        L.orc_parse_dump.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    if wide:
        _LIB_WIDE = L
    else:
        _LIB = L
    return L


def read_maybe_gz(path):
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":
        with gzip.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


@dataclass
class OracleResult:
    k: int
    nbanks: int
    kmers_nb_valid: int
    kmers_nb_invalid: int
    nb_seq: int
    nb_nt: int
    nb_superkmers: int
    keys_lo: np.ndarray      # distinct canonical k-mers, ascending (low 64 bits)
    keys_hi: np.ndarray      # high 64 bits (all zero for k<=32)
    keys_w2: np.ndarray      # bits 128..191 (k >= 65, wide build; zeros otherwise)
    keys_w3: np.ndarray      # bits 192..255 (k >= 97)
    counts: np.ndarray       # int32 [ndistinct, nbanks]
    sums: np.ndarray         # int32 [ndistinct]
    solid: np.ndarray        # bool  [ndistinct]
    hist: np.ndarray         # uint64[10001]  (index = abundance; bins 0 and 10000 are always 0)
    hist2d: np.ndarray       # uint64[11, 10001]  ([dim2, dim1])
    cutoffs: object = None   # -abundance-min auto: the cutoffs of the first pass
    seq_min: int = 0         # BankStats (K/BankKmers.hpp:166-215): shortest / longest sequence, sum of squared lengths
    seq_max: int = 0
    seq_sumsq: int = 0

    def bank_stats_strings(self):
        """the bank / sequences / kmers keys as SortingCountAlgorithm::getInfo() formats them (K/SortingCountAlgorithm.cpp:733-742)"""
        import math
        n = self.nb_seq
        mean = self.nb_nt / n if n else 0.0
        dev = math.sqrt(max(0.0, self.seq_sumsq / n - mean * mean)) if n else 0.0
        return {"bank_total_nt": str(self.nb_nt), "seq_number": str(n), "seq_size_min": str(self.seq_min), "seq_size_max": str(self.seq_max),
                "seq_size_mean": "%.1f" % mean, "seq_size_deviation": "%.1f" % dev,
                "kmers_nb_valid": str(self.kmers_nb_valid), "kmers_nb_invalid": str(self.kmers_nb_invalid)}

    @property
    def nb_distinct(self):
        return len(self.keys_lo)

    @property
    def nb_solid(self):
        return int(self.solid.sum())

    def solid_kmers(self):
        s = self.solid
        return self.keys_lo[s], self.keys_hi[s], self.sums[s]

    def solid_kmer_words(self):
        """(uint64[n, 4] value words, least significant first; int32[n] abundance) of the solid k-mers, ascending"""
        s = self.solid
        return np.stack([self.keys_lo[s], self.keys_hi[s], self.keys_w2[s], self.keys_w3[s]], axis=1), self.sums[s]


def _np_from(ptr, dtype, n):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class Oracle:
    """Accumulate sequences/files per bank, then finish() -> OracleResult."""

    def __init__(self, k, nbanks=1, m=0):
        self.L = _lib(k)
        if k > self.L.orc_max_k():
            raise ValueError("oracle built for k <= %d" % self.L.orc_max_k())
        self.k, self.nbanks = k, nbanks
        self.h = self.L.orc_create(k, nbanks, m)

    def add_sequence(self, seq, bank=0):
        if isinstance(seq, str):
            seq = seq.encode()
        self.L.orc_add_sequence(self.h, bank, seq, len(seq))

    def add_file_bytes(self, data, bank=0):
        arr = np.frombuffer(data, dtype=np.uint8)
        self.L.orc_add_file(self.h, bank, arr.ctypes.data, arr.size)

    def finish(self, abundance_min=2, abundance_max=2**31 - 1, kind="sum", solid_vec=None, histo2d=False):
        nb = self.nbanks
        if isinstance(abundance_min, int):
            abundance_min = [abundance_min] * nb
        amin = (C.c_int64 * nb)(*abundance_min)
        sv = bytes(solid_vec if solid_vec is not None else [1] * nb)
        self.L.orc_finish(self.h, amin, abundance_max, KINDS[kind], sv, int(histo2d))
        L, h = self.L, self.h
        nd = L.orc_nb_distinct(h)
        res = OracleResult(
            k=self.k, nbanks=nb,
            kmers_nb_valid=L.orc_nb_valid(h), kmers_nb_invalid=L.orc_nb_invalid(h),
            nb_seq=L.orc_nb_seq(h), nb_nt=L.orc_nb_nt(h), nb_superkmers=L.orc_nb_superkmers(h),
            keys_lo=_np_from(L.orc_keys_lo(h), np.uint64, nd),
            keys_hi=_np_from(L.orc_keys_hi(h), np.uint64, nd),
            keys_w2=_np_from(L.orc_keys_w2(h), np.uint64, nd),
            keys_w3=_np_from(L.orc_keys_w3(h), np.uint64, nd),
            counts=_np_from(L.orc_counts(h), np.int32, nd * nb).reshape(nd, nb),
            sums=_np_from(L.orc_sums(h), np.int32, nd),
            solid=_np_from(L.orc_solid(h), np.uint8, nd).astype(bool),
            hist=_np_from(L.orc_hist(h), np.uint64, 10001),
            hist2d=_np_from(L.orc_hist2d(h), np.uint64, 10001 * 11).reshape(11, 10001),
            seq_min=L.orc_seq_min(h), seq_max=L.orc_seq_max(h), seq_sumsq=L.orc_seq_sumsq(h),
        )
        L.orc_destroy(h)
        self.h = None
        return res

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)


def histogram_threshold(hist, min_auto_threshold=3, length=10000):
    """Histogram::compute_threshold (G/src/gatb/tools/misc/impl/Histogram.cpp:61-190) restated: smoothed histogram
    (u64 casts of 0.6/0.4 and 0.2/0.6/0.2 double mixes), first increase, highest smoothed bin after it, lowest smoothed
    bin between the two = cutoff, capped at the first abundance whose cumulated volume reaches 25 %, floored at
    `min_auto_threshold`.  Returns (cutoff, nbsolids)."""
    H = [int(x) for x in hist[:length + 1]]
    S = [0] * (length + 1)
    S[1] = int(0.6 * float(H[1]) + 0.4 * float(H[2]))                       # :68
    total = H[1] + H[length] * length                                      # :69,:98
    inc_at = peak_at = -1
    peak = 0
    for i in range(2, length):                                             # :78-96
        total += H[i] * i
        S[i] = int(0.2 * float(H[i - 1]) + 0.6 * float(H[i]) + 0.2 * float(H[i + 1]))
        if inc_at == -1 and S[i - 1] < S[i]:
            inc_at = i - 1
        if inc_at > 0 and S[i] > peak:
            peak, peak_at = S[i], i
    if inc_at == -1:                                                       # :101-105
        return min_auto_threshold, 0
    cut, low = 0, 10000000000
    for i in range(inc_at, peak_at + 1):                                   # :115-122
        if S[i] < low:
            low, cut = S[i], i
    cap, gone = 0, 0
    for i in range(length + 1):                                            # :130-142
        gone += H[i] * i
        if float(gone) / float(total) >= 0.25:
            cap = i + 1
            break
    cut = max(min(cut, cap), min_auto_threshold)                           # :144-148
    return cut, int(sum(H[cut:length + 1]))


def _bank_histogram(counts_b):
    """CountProcessorHistogram on one bank's counts: u16 truncation, clamp, bins 1..9999 survive (Histogram.hpp:92,221)."""
    idx = (counts_b.astype(np.int64) & 0xFFFF)
    idx = np.minimum(idx, 10000)
    h = np.bincount(idx, minlength=10001).astype(np.uint64)
    h[0] = 0
    h[10000] = 0
    return h


def count_files(banks, k, m=0, **kw):
    """banks: list of byte strings (one file image per bank).  abundance_min may hold -1 entries ("auto"): the cutoffs
    are then computed first, as the reference's cutoff pass does (SortingCountAlgorithm.cpp:455-514,
    CountProcessorCutoff.hpp:86-124, CountProcessorSolidity.hpp:45-66); the result carries them in `.cutoffs`."""
    def run(**kw2):
        o = Oracle(k, len(banks), m)
        for b, data in enumerate(banks):
            o.add_file_bytes(data, b)
        return o.finish(**kw2)
    amin = kw.get("abundance_min", 2)
    if isinstance(amin, int):
        amin = [amin] * len(banks)
    amin = list(amin) + [amin[-1]] * (len(banks) - len(amin))
    if -1 not in amin:
        res = run(**kw)
        res.cutoffs = None
        return res
    kind = kw.get("kind", "sum") if len(banks) > 1 else "sum"
    first = run(**dict(kw, abundance_min=[2**31 - 1] * len(banks)))
    if kind in ("sum", "min", "max"):
        hs = [first.hist]                                                  # histogram of the sum over all banks
    else:
        hs = [_bank_histogram(first.counts[:, b]) for b in range(len(banks))]
    cutoffs = [histogram_threshold(h, 3)[0] for h in hs]
    new = [c if a == -1 else a for a, c in zip(amin, cutoffs)]
    new = new + [new[-1]] * (len(amin) - len(new))
    res = run(**dict(kw, abundance_min=new))
    res.cutoffs = cutoffs
    return res


def kmers_of(seq, k, m=0, forward=False):
    if isinstance(seq, str):
        seq = seq.encode()
    n = max(0, len(seq) - k + 1)
    lo = np.zeros(n, np.uint64); hi = np.zeros(n, np.uint64)
    valid = np.zeros(n, np.uint8); mn = np.zeros(n, np.uint32); mp = np.zeros(n, np.int32)
    got = _lib(k).orc_kmers_of(seq, len(seq), k, m, int(forward), lo.ctypes.data, hi.ctypes.data,
                               valid.ctypes.data, mn.ctypes.data, mp.ctypes.data)
    assert got == n
    return lo, hi, valid.astype(bool), mn, mp


def kmers_of_words(seq, k, m=0, forward=False):
    """k-mer values as uint64[n, 4] (least significant word first), validity, minimizers (k up to 127)"""
    if isinstance(seq, str):
        seq = seq.encode()
    n = max(0, len(seq) - k + 1)
    w = [np.zeros(n, np.uint64) for _ in range(4)]
    valid = np.zeros(n, np.uint8); mn = np.zeros(n, np.uint32); mp = np.zeros(n, np.int32)
    got = _lib(k).orc_kmers_of4(seq, len(seq), k, m, int(forward), w[0].ctypes.data, w[1].ctypes.data, w[2].ctypes.data,
                                w[3].ctypes.data, valid.ctypes.data, mn.ctypes.data, mp.ctypes.data)
    assert got == n
    return np.stack(w, axis=1), valid.astype(bool), mn, mp


def parse_stats(data):
    arr = np.frombuffer(data, dtype=np.uint8)
    nt = C.c_uint64(0)
    cap = arr.size + 16
    out = np.zeros(cap, np.uint8)
    nrec = _lib().orc_parse_stats(arr.ctypes.data, arr.size, C.byref(nt), out.ctypes.data, cap)
    concat = out[: nt.value + nrec].tobytes()
    return nrec, nt.value, concat


def mmer_lut(x, m):
    return _lib().orc_mmer_lut(x, m)


# ------------------------------------------------------------------------------------------------
# the real reference (oracle/_ref/bin/{dsk,dsk2ascii,gatb-h5dump}), when it was built
# ------------------------------------------------------------------------------------------------
def _ref_bin(name, wide=False):
    """oracle/_ref/bin (KSIZE_LIST "32 64") or oracle/_ref/wide/bin (KSIZE_LIST "32 64 96 128", golden generation only)"""
    return os.path.join(ORACLE_DIR, "_ref", "wide", "bin", name) if wide else os.path.join(ORACLE_DIR, "_ref", "bin", name)


def ref_wide_available():
    return all(os.access(_ref_bin(n, True), os.X_OK) for n in ("dsk", "dsk2ascii"))


def ref_available():
    return all(os.access(_ref_bin(n), os.X_OK) for n in ("dsk", "dsk2ascii", "gatb-h5dump"))


def run_reference(files, k, abundance_min=2, histo=True, histo2d=False, nb_cores=0, extra=(), workdir=None,
                  want_kmers=True, solidity_kind=None, abundance_max=None, wide=False):
    """Runs the reference `dsk` on `files` (list of paths; comma-joined like the CLI) and returns
    dict(kmers=[(str,count)...] sorted, hist=np.uint64[10001], hist2d=..., stats=str, seconds=float)."""
    import time
    tmp = workdir or tempfile.mkdtemp(prefix="dskref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    out = os.path.join(tmp, "ref")
    cmd = [_ref_bin("dsk", wide), "-file", ",".join(files), "-kmer-size", str(k), "-abundance-min", str(abundance_min),
           "-out", out, "-out-tmp", tmp, "-out-dir", tmp, "-verbose", "1", "-nb-cores", str(nb_cores)]
    if histo:
        cmd += ["-histo", "1"]
    if histo2d:
        cmd += ["-histo2D", "1"]
    if solidity_kind:
        cmd += ["-solidity-kind", solidity_kind]
    if abundance_max is not None:
        cmd += ["-abundance-max", str(abundance_max)]
    cmd += list(extra)
    t0 = time.time()
    p = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
    dt = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("reference dsk failed: %s\n%s" % (p.stdout[-2000:], p.stderr[-2000:]))
    res = {"seconds": dt, "stats": p.stdout, "h5": out + ".h5", "tmp": tmp}
    if histo and os.path.exists(out + ".histo"):
        h = np.zeros(10001, np.uint64)
        for line in open(out + ".histo"):
            a, b = line.split()
            h[int(a)] = int(b)
        res["hist"] = h
    if histo2d and os.path.exists(out + ".histo2D"):
        rows = []
        for line in open(out + ".histo2D"):
            parts = line.replace(":", " ").split()
            rows.append([int(x) for x in parts[1:]])
        res["hist2d"] = np.array(rows, dtype=np.uint64).T   # -> [dim2(11), dim1(10001)]
    if want_kmers:
        txt = os.path.join(tmp, "ref.txt")
        q = subprocess.run([_ref_bin("dsk2ascii", wide), "-file", out + ".h5", "-out", txt], cwd=tmp, capture_output=True, text=True)
        if q.returncode != 0:
            raise RuntimeError("dsk2ascii failed: " + q.stderr[-2000:])
        km = []
        if os.path.exists(txt):
            for line in open(txt):
                a, b = line.split()
                km.append((a, int(b)))
        km.sort()
        res["kmers"] = km
    return res


def stat_value(stats_text, key):
    """Pull `key : value` out of the -verbose 1 stats block."""
    for line in stats_text.splitlines():
        s = line.strip()
        if s.startswith(key + " ") or s.startswith(key + ":"):
            return s.split(":", 1)[1].strip()
    return None


def parse_dump(path, k):
    """(lo u64[n], hi u64[n], count u32[n]) of a `dsk2ascii` text dump, sorted ascending by k-mer value (k <= 63)."""
    text = np.fromfile(path, dtype=np.uint8)
    cap = int(text.size // (k + 3)) + 16
    lo = np.zeros(cap, np.uint64); hi = np.zeros(cap, np.uint64); cnt = np.zeros(cap, np.uint32)
    n = _lib(k).orc_parse_dump(text.ctypes.data, text.size, k, lo.ctypes.data, hi.ctypes.data, cnt.ctypes.data, cap)
    if n < 0 or n > cap:
        raise RuntimeError("malformed dsk2ascii dump %s (%d)" % (path, n))
    lo, hi, cnt = lo[:n], hi[:n], cnt[:n]
    order = np.lexsort((lo, hi)) if k > 32 else np.argsort(lo, kind="stable")
    return lo[order], hi[order], cnt[order]
