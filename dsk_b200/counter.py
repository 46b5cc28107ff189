"""Thin object wrapper over the C ABI (include/dskgpu.h).  All compute happens in libdskgpu.so."""
import ctypes as C

import numpy as np

from . import _lib

FMT = {"auto": 0, "fasta": 1, "fastq": 2, "lines": 3}
SOLIDITY = {"sum": 0, "min": 1, "max": 2, "one": 3, "all": 4, "custom": 5}
COUNT_MODE = {"auto": 0, "sort": 1, "vector": 1, "hash": 2, "smem": 3}


class DskGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dskgpu error %d: %s" % (code, msg))
        self.code = code


class GpuCounter:
    """One counting context on one GPU: push bank bytes, finish, read the solid k-mers + histogram."""

    def __init__(self, kmer_size=31, abundance_min=2, abundance_max=2**31 - 1, nb_banks=1, per_bank_counts=False,
                 solidity_kind="sum", solid_vec=None, histo2d=False, minimizer_size=10, device=0, count_mode="auto",
                 hash_log2_slots=0, nb_partitions=0, keep_results_on_device=False, stream=None, rank=0, world_size=1, push_chunk_bytes=0,
                 smem_table_slots=0, bank_histograms=False, nb_passes=1, pass_id=0, sequence_stats=False):
        self.L = _lib.lib()
        cfg = _lib.Config()
        self.L.dskgpu_config_default(C.byref(cfg))
        cfg.kmer_size = kmer_size
        cfg.minimizer_size = minimizer_size
        cfg.nb_banks = nb_banks
        cfg.per_bank_counts = int(per_bank_counts)
        cfg.solidity_kind = SOLIDITY[solidity_kind] if isinstance(solidity_kind, str) else solidity_kind
        amin = [abundance_min] * _lib.MAX_BANKS if isinstance(abundance_min, int) else list(abundance_min)
        amin = amin + [amin[-1]] * (_lib.MAX_BANKS - len(amin))
        for i in range(_lib.MAX_BANKS):
            cfg.abundance_min[i] = amin[i]
        cfg.abundance_max = abundance_max
        if solid_vec is not None:
            for i, v in enumerate(solid_vec):
                cfg.solid_vec[i] = int(v)
        cfg.histo2d = int(histo2d)
        cfg.device = device
        cfg.count_mode = COUNT_MODE[count_mode] if isinstance(count_mode, str) else count_mode
        cfg.hash_log2_slots = hash_log2_slots
        cfg.nb_partitions = nb_partitions
        cfg.keep_results_on_device = int(keep_results_on_device)
        cfg.stream = stream
        self.stream = stream
        cfg.rank, cfg.world_size = rank, world_size
        cfg.push_chunk_bytes = push_chunk_bytes
        cfg.smem_table_slots = smem_table_slots
        cfg.bank_histograms = int(bank_histograms)
        cfg.nb_passes, cfg.pass_id = nb_passes, pass_id
        cfg.sequence_stats = int(sequence_stats)
        self.cfg = cfg
        self.k = kmer_size
        self.h = C.c_void_p()
        rc = self.L.dskgpu_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            msg = self.L.dskgpu_last_error(None).decode() or self.L.dskgpu_strerror(rc).decode()
            self.h = None
            raise DskGpuError(rc, msg)

    # -- helpers ------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise DskGpuError(rc, (self.L.dskgpu_last_error(self.h) or b"").decode() or self.L.dskgpu_strerror(rc).decode())

    # -- input --------------------------------------------------------------------------------------
    def push_bytes(self, data, bank=0, fmt="auto", last=True):
        """data: bytes / bytearray / numpy uint8 array / (ptr, n) of pinned host memory."""
        if isinstance(data, tuple):
            ptr, n = data
        else:
            arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
            self._keep = arr
            ptr, n = arr.ctypes.data, arr.size
        self._check(self.L.dskgpu_push_bytes(self.h, bank, ptr, n, FMT[fmt], 1 if last else 0))

    def push_device_bytes(self, dev_ptr, n, bank=0, fmt="auto", last=True):
        self._check(self.L.dskgpu_push_device_bytes(self.h, bank, dev_ptr, n, FMT[fmt], 1 if last else 0))

    def push_reads(self, reads, bank=0):
        """reads: list of bytes/str sequences (the IBank::iterator() flavour of the input)."""
        bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
        offs = np.zeros(len(bs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
        blob = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8)
        self._check(self.L.dskgpu_push_reads(self.h, bank, blob.ctypes.data, offs.ctypes.data, len(bs)))

    # -- compute ------------------------------------------------------------------------------------
    def finish(self):
        self._check(self.L.dskgpu_finish(self.h))

    def reset(self):
        self._check(self.L.dskgpu_reset(self.h))

    def set_pass(self, pass_id, nb_passes):
        """the next pushes keep the super-k-mers of pass `pass_id` of `nb_passes` (call on a fresh / reset context)"""
        self._check(self.L.dskgpu_set_pass(self.h, pass_id, nb_passes))

    def push_sync(self):
        """every host buffer handed to push_bytes has been copied to the device"""
        self._check(self.L.dskgpu_push_sync(self.h))

    def recount(self, abundance_min):
        """second pass of -abundance-min auto: same partitions (still in HBM), new thresholds"""
        a = list(abundance_min) + [abundance_min[-1]] * (_lib.MAX_BANKS - len(abundance_min))
        arr = (C.c_int64 * _lib.MAX_BANKS)(*a)
        self._check(self.L.dskgpu_recount(self.h, arr))

    def bank_histograms(self):
        h = np.zeros((max(1, self.cfg.nb_banks if self.cfg.per_bank_counts else 1), _lib.HISTO_LEN), np.uint64)
        self._check(self.L.dskgpu_bank_histograms(self.h, h.ctypes.data))
        return h

    # -- results ------------------------------------------------------------------------------------
    def solid(self):
        """(keys uint64[n, words] low word first, counts uint32[n]) of every partition, concatenated."""
        ks, cs = [], []
        words = 1
        for p in range(self.L.dskgpu_num_partitions(self.h)):
            kp, cp, n, w = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_int()
            self._check(self.L.dskgpu_partition(self.h, p, C.byref(kp), C.byref(cp), C.byref(n), C.byref(w)))
            words = w.value
            if n.value:
                kb = (C.c_uint64 * (n.value * words)).from_address(kp.value)
                cb = (C.c_uint32 * n.value).from_address(cp.value)
                ks.append(np.frombuffer(kb, dtype=np.uint64).reshape(n.value, words).copy())
                cs.append(np.frombuffer(cb, dtype=np.uint32).copy())
        if not ks:
            return np.zeros((0, words), np.uint64), np.zeros(0, np.uint32)
        return np.concatenate(ks), np.concatenate(cs)

    def solid_device(self, p=0):
        kp, cp, n, w = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_int()
        self._check(self.L.dskgpu_partition_device(self.h, p, C.byref(kp), C.byref(cp), C.byref(n), C.byref(w)))
        return kp.value, cp.value, n.value, w.value

    def histogram(self):
        h1 = np.zeros(_lib.HISTO_LEN, np.uint64)
        h2 = np.zeros((_lib.HISTO2D_DIM2, _lib.HISTO_LEN), np.uint64)
        self._check(self.L.dskgpu_histogram(self.h, h1.ctypes.data, h2.ctypes.data))
        return h1, h2

    def stats(self):
        st = _lib.Stats()
        self._check(self.L.dskgpu_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def close(self):
        if getattr(self, "h", None):
            self.L.dskgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def multi_finish(engines):
    """dskgpu_multi_finish: exchange + counting for several contexts (ranks 0..n-1) living in this process"""
    L = _lib.lib()
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    rc = L.dskgpu_multi_finish(arr, len(engines))
    if rc != 0:
        msgs = [(L.dskgpu_last_error(e.h) or b"").decode() for e in engines]
        raise DskGpuError(rc, "; ".join(m for m in msgs if m) or L.dskgpu_strerror(rc).decode())


# ---- multi-GPU exchange (include/dskgpu.h "multi-GPU exchange") ----------------------------------------------
def _xchg_methods():
    def xchg_local_totals(self):
        km, nr = C.c_uint64(), C.c_uint64()
        self._check(self.L.dskgpu_xchg_local_totals(self.h, C.byref(km), C.byref(nr)))
        return km.value, nr.value

    def xchg_prepare(self):
        v = np.zeros(4, dtype=np.uint64)
        self._check(self.L.dskgpu_xchg_prepare(self.h, v.ctypes.data))
        return v

    def xchg_set_global(self, global4):
        g = np.ascontiguousarray(global4, dtype=np.uint64)
        lvl = C.c_int()
        self._check(self.L.dskgpu_xchg_set_global(self.h, g.ctypes.data, C.byref(lvl)))
        self._bin_level = lvl.value
        return lvl.value

    def xchg_sketch(self):
        """HyperLogLog registers of the distinct k-mers of this rank's density sample (merge over ranks with max)"""
        v = np.zeros(4096, dtype=np.uint32)
        self._check(self.L.dskgpu_xchg_sketch(self.h, v.ctypes.data))
        return v

    def xchg_set_sketch(self, merged):
        m = np.ascontiguousarray(merged, dtype=np.uint32)
        assert m.size == 4096
        self._check(self.L.dskgpu_xchg_set_sketch(self.h, m.ctypes.data))

    def xchg_hist(self, d_out):
        """this rank's bin histogram [2 << level] u64 -> device buffer of the caller (all-reduce it in place)"""
        self._check(self.L.dskgpu_xchg_hist(self.h, C.c_void_p(d_out)))

    def xchg_plan(self, d_global_hist):
        """device planner; returns (P, PW, need_records[W])"""
        P, PW = C.c_uint32(), C.c_uint32()
        need = np.zeros(self.cfg.world_size, dtype=np.uint64)
        self._check(self.L.dskgpu_xchg_plan(self.h, C.c_void_p(d_global_hist), C.byref(P), C.byref(PW), need.ctypes.data))
        return P.value, PW.value, need

    def xchg_counts(self, d_out):
        """[W][PW] records per partition grouped by owner, then [W] chunk sizes -> device buffer of the caller"""
        self._check(self.L.dskgpu_xchg_counts(self.h, C.c_void_p(d_out)))

    def xchg_ensure_recv(self, capacity_records):
        self._check(self.L.dskgpu_xchg_ensure_recv(self.h, int(capacity_records)))

    def xchg_recv_buffer(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.L.dskgpu_xchg_recv_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def xchg_ipc_handle(self):
        buf = (C.c_ubyte * 64)()
        self._check(self.L.dskgpu_xchg_ipc_handle(self.h, buf))
        return bytes(buf)

    def xchg_open_peer(self, handle):
        p = C.c_void_p()
        hb = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self.L.dskgpu_xchg_open_peer(self.h, hb, C.byref(p)))
        return p.value

    def xchg_close_peer(self, ptr):
        self._check(self.L.dskgpu_xchg_close_peer(self.h, C.c_void_p(ptr)))

    def xchg_set_peers(self, ptrs):
        arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(p) for p in ptrs])
        self._check(self.L.dskgpu_xchg_set_peers(self.h, arr))

    def xchg_scatter(self, d_recv_counts, d_send_matrix):
        self._check(self.L.dskgpu_xchg_scatter(self.h, C.c_void_p(d_recv_counts), C.c_void_p(d_send_matrix)))

    def xchg_sync(self):
        self._check(self.L.dskgpu_xchg_sync(self.h))

    def debug_plan(self):
        """the plan the DEVICE derived: (level, bin2part, part_kmers, part_recs, part_local_recs)"""
        lvl = C.c_int()
        cap = (1 << 24) + 16
        pk = np.zeros(cap, np.uint64); pr = np.zeros(cap, np.uint64); pl = np.zeros(cap, np.uint64)
        b2p = np.zeros(1 << 24, np.uint32)
        P = self.L.dskgpu_debug_plan(self.h, C.byref(lvl), b2p.ctypes.data, pk.ctypes.data, pr.ctypes.data, pl.ctypes.data, cap)
        if P < 0:
            self._check(int(P))
        return lvl.value, b2p[:1 << lvl.value].copy(), pk[:P].copy(), pr[:P].copy(), pl[:P].copy()

    for f in (xchg_local_totals, xchg_prepare, xchg_set_global, xchg_sketch, xchg_set_sketch, xchg_hist, xchg_plan, xchg_counts, xchg_ensure_recv, xchg_recv_buffer, xchg_ipc_handle,
              xchg_open_peer, xchg_close_peer, xchg_set_peers, xchg_scatter, xchg_sync, debug_plan):
        setattr(GpuCounter, f.__name__, f)


_xchg_methods()
