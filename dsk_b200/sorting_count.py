"""Host-side mirror of the reference operator for the counting path.

`SortingCountAlgorithm(bank, props)` follows gatb-core's class of the same name
(G/src/gatb/kmer/impl/SortingCountAlgorithm.hpp:66-192; options :202-236 of the .cpp) closely enough that
the parity tests read like the reference's own (G/test/unit/src/kmer/TestDSK.cpp): same option names, same
defaults, same error for an unhandled k, `execute()`, `getInfo()`, `getSolidCounts()`.  All counting is done
by libdskgpu.so through the C ABI; this file only parses options, reads bank files and hands bytes over.
"""
import gzip
import os

import numpy as np

from . import _lib
from .counter import GpuCounter, DskGpuError
from .histogram import compute_threshold, auto_thresholds, MIN_AUTO_THRESHOLD

# G/src/gatb/tools/misc/impl/StringsRepository.hpp:81-127
STR_KMER_SIZE = "-kmer-size"
STR_KMER_ABUNDANCE_MIN = "-abundance-min"
STR_KMER_ABUNDANCE_MAX = "-abundance-max"
STR_SOLIDITY_KIND = "-solidity-kind"
STR_SOLIDITY_CUSTOM = "-solidity-custom"
STR_HISTO2D = "-histo2D"
STR_HISTO = "-histo"
STR_MINIMIZER_SIZE = "-minimizer-size"
STR_URI_FILE = "-file"
STR_URI_OUTPUT = "-out"

NT = "ACTG"


def getDefaultProperties():
    """SortingCountAlgorithm::getOptionsParser defaults (SortingCountAlgorithm.cpp:208-231)."""
    return {STR_KMER_SIZE: 31, STR_KMER_ABUNDANCE_MIN: "2", STR_KMER_ABUNDANCE_MAX: 2147483647,
            STR_SOLIDITY_KIND: "sum", STR_HISTO2D: 0, STR_HISTO: 0, STR_MINIMIZER_SIZE: 10}


class BankStrings:
    """In-memory bank of sequences (role of gatb::core::bank::BankStrings in the reference's unit tests)."""

    def __init__(self, *seqs):
        if len(seqs) == 1 and isinstance(seqs[0], (list, tuple)):
            seqs = seqs[0]
        self.seqs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]


class BankFile:
    """A FASTA/FASTQ file, optionally gzipped (Bank::open on a single uri)."""

    def __init__(self, path):
        self.path = path

    def read(self):
        with open(self.path, "rb") as f:
            head = f.read(2)
        if head == b"\x1f\x8b":
            with gzip.open(self.path, "rb") as f:
                return f.read()
        with open(self.path, "rb") as f:
            return f.read()


class BankBytes:
    """File image already in memory."""

    def __init__(self, data):
        self.data = data

    def read(self):
        return self.data


class BankAlbum:
    """List of banks (comma separated -file list / BankAlbum)."""

    def __init__(self, banks=None):
        self.banks = list(banks or [])

    def addBank(self, b):
        self.banks.append(b)


def open_bank(uri):
    """Bank::open(uri) (G/src/gatb/bank/impl/Bank.cpp:143-160): comma separated list -> album."""
    parts = [p for p in uri.split(",") if p]
    if len(parts) == 1:
        return BankFile(parts[0])
    return BankAlbum([BankFile(p) for p in parts])


class SortingCountAlgorithm:
    def __init__(self, bank, props=None, device=0, **engine_kw):
        self.props = dict(getDefaultProperties())
        if props:
            self.props.update(props)
        if isinstance(bank, str):
            bank = open_bank(bank)
        self.banks = bank.banks if isinstance(bank, BankAlbum) else [bank]
        self.device = device
        self.engine_kw = engine_kw
        self.info = {}
        self._solid = None
        self._hist = None

    # -- option parsing (ConfigurationAlgorithm.cpp:195-264, :478-498) -----------------------------------
    def _configure(self):
        p = self.props
        k = int(p[STR_KMER_SIZE])
        nb = len(self.banks)
        kind = str(p[STR_SOLIDITY_KIND])
        histo2d = int(p.get(STR_HISTO2D, 0)) != 0
        amin_s = str(p[STR_KMER_ABUNDANCE_MIN])
        amin = [-1 if x == "auto" else int(x) for x in amin_s.split(",")]      # ConfigurationAlgorithm.cpp:478-498
        if len(amin) > nb:
            raise ValueError("Kmer solidity has more thresholds than banks")
        amin = amin + [amin[-1]] * (nb - len(amin))
        if nb == 1:
            kind = "sum"                                     # ConfigurationAlgorithm.cpp:261-264
        per_bank = nb > 1 and (kind != "sum" or histo2d)     # SortingCountAlgorithm.cpp:604-620
        if histo2d and nb < 2:
            raise ValueError("histo2D requires at least two banks")
        solid_vec = None
        if kind == "custom":
            sv = str(p.get(STR_SOLIDITY_CUSTOM, ""))
            solid_vec = [int(c) for c in sv] + [0] * (nb - len(sv))
        # "auto" anywhere: a first count pass builds the histogram(s) the cutoffs come from -- one histogram of the sum
        # for sum/min/max, one per bank for one/all/custom (SortingCountAlgorithm.cpp:455-514)
        self._auto = -1 in amin
        self._auto_per_bank = self._auto and nb > 1 and kind in ("one", "all", "custom")
        return dict(kmer_size=k, abundance_min=amin, abundance_max=int(p[STR_KMER_ABUNDANCE_MAX]), nb_banks=nb,
                    per_bank_counts=per_bank, solidity_kind=kind, solid_vec=solid_vec, histo2d=histo2d,
                    minimizer_size=int(p[STR_MINIMIZER_SIZE]), bank_histograms=self._auto_per_bank)

    def execute(self):
        cfg = self._configure()
        if cfg["minimizer_size"] == 10:
            # reference default: sized from the estimated volume (~0.7 k-mers per input byte), as host/GpuSortingCount.hpp does
            nbytes = 0
            for bank in self.banks:
                if isinstance(bank, BankStrings):
                    nbytes += sum(len(x) for x in bank.seqs)
                elif isinstance(bank, BankBytes):
                    nbytes += len(bank.data)
                elif os.path.exists(bank.path):
                    nbytes += os.path.getsize(bank.path) * (4 if bank.path.endswith(".gz") else 1)
            cfg["minimizer_size"] = _lib.lib().dskgpu_suggest_minimizer_size(int(0.7 * nbytes), cfg["kmer_size"])
        cfg.update(self.engine_kw)
        user_amin = list(cfg["abundance_min"])
        if self._auto:
            # pass 1 (the cutoff processor) dumps nothing: no abundance reaches this threshold
            cfg["abundance_min"] = [2**31 - 1] * len(user_amin)
        try:
            eng = GpuCounter(device=self.device, **cfg)
        except DskGpuError as e:
            if e.code == -1 and "unhandled kmer size" in str(e):
                raise RuntimeError("Failure because of unhandled kmer size %d" % cfg["kmer_size"]) from e
            raise
        # pass loop of SortingCountAlgorithm::execute (SortingCountAlgorithm.cpp:678-689): with nb_passes > 1 the banks are
        # pushed once per pass and every pass keeps its own share of the minimizers; the results of the passes are disjoint
        nb_passes = max(1, int(cfg.get("nb_passes", 1) or 1))

        def feed(eng):
            for b, bank in enumerate(self.banks):
                if isinstance(bank, BankStrings):
                    eng.push_reads(bank.seqs, bank=b)
                else:
                    eng.push_bytes(bank.read(), bank=b, last=True)

        with eng:
            self.cutoffs = None
            solids, st = [], None
            h1 = np.zeros(_lib.HISTO_LEN, np.uint64); h2 = np.zeros((_lib.HISTO2D_DIM2, _lib.HISTO_LEN), np.uint64)
            new_amin = None
            if self._auto and nb_passes > 1:
                # the cutoffs come from the histogram(s) of the WHOLE job: one round of passes for them, one for the dump
                hs = None
                for ps in range(nb_passes):
                    eng.reset(); eng.set_pass(ps, nb_passes); feed(eng); eng.finish()
                    h = eng.bank_histograms() if self._auto_per_bank else eng.histogram()[0][None, :]
                    hs = h.copy() if hs is None else hs + h
                self.cutoffs = [compute_threshold(h, MIN_AUTO_THRESHOLD)[0] for h in hs]
                new_amin = auto_thresholds(user_amin, self.cutoffs)
            for ps in range(nb_passes):
                if nb_passes > 1:
                    eng.reset(); eng.set_pass(ps, nb_passes)
                feed(eng)
                eng.finish()
                if self._auto and nb_passes == 1:
                    hs = eng.bank_histograms() if self._auto_per_bank else [eng.histogram()[0]]
                    self.cutoffs = [compute_threshold(h, MIN_AUTO_THRESHOLD)[0] for h in hs]
                    eng.recount(auto_thresholds(user_amin, self.cutoffs))          # pass 2: the dsk processor chain
                elif new_amin is not None:
                    eng.recount(new_amin)
                solids.append(eng.solid())
                a, b2 = eng.histogram(); h1 += a; h2 += b2
                s1 = eng.stats()
                if st is None:
                    st = s1
                else:
                    for key in ("kmers_nb_distinct", "kmers_nb_solid", "nb_superkmers", "nb_partitions", "gpu_launches", "kmers_in_pass",
                                "nb_parts_smem", "nb_smem_splits", "nb_groups_hash", "nb_groups_sort", "nb_groups_bucket"):
                        st[key] += s1[key]
            self._solid = (np.concatenate([x[0] for x in solids]), np.concatenate([x[1] for x in solids])) if len(solids) > 1 else solids[0]
            self._hist = (h1, h2)
        self.k = cfg["kmer_size"]
        self.info = {
            "kmers_nb_valid": st["kmers_nb_valid"], "kmers_nb_distinct": st["kmers_nb_distinct"],
            "kmers_nb_solid": st["kmers_nb_solid"], "kmers_nb_weak": st["kmers_nb_distinct"] - st["kmers_nb_solid"],
            "nb_superkmers": st["nb_superkmers"], "nb_partitions": st["nb_partitions"], "seq_number": st["nb_sequences"],
            "bank_total_nt": st["nb_nucleotides"], "solidity_kind": cfg["solidity_kind"], "engine": st,
            "cutoffs_auto": self.cutoffs,
        }
        return self

    # -- results -------------------------------------------------------------------------------------------
    def getInfo(self):
        return self.info

    def getSolidCounts(self):
        """(keys uint64[n, words], abundance uint32[n]) -- the content of Partition<Count> "dsk/solid"."""
        return self._solid

    def getHistogram(self):
        return self._hist

    def solidKmerStrings(self):
        keys, cnt = self._solid
        out = []
        for row, c in zip(keys, cnt):
            v = 0
            for w in reversed(row):
                v = (v << 64) | int(w)
            out.append(("".join(NT[(v >> (2 * (self.k - 1 - i))) & 3] for i in range(self.k)), int(c)))
        return out

    def writeHisto(self, path):
        """`<out>.histo` text (CountProcessorHistogram.hpp:111-142): 10000 lines "i\\tcount"."""
        h1, _ = self._hist
        with open(path, "w") as f:
            for i in range(1, 10001):
                f.write("%d\t%d\n" % (i, int(h1[i])))

    def writeHisto2D(self, path):
        """`<out>.histo2D` text: 10001 rows "%5i:\\t" + 11 x "\\t%6lli" (CountProcessorHistogram.hpp:127-142)."""
        _, h2 = self._hist
        with open(path, "w") as f:
            for i in range(10001):
                f.write("%5d:\t" % i + "".join("\t%6d" % int(h2[j, i]) for j in range(11)) + "\n")
