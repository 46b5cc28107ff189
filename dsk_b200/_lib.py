"""ctypes binding of include/dskgpu.h.  Fails loudly when the shared library is missing: the counting
path has no CPU fallback."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("DSKGPU_LIB") or os.path.join(PKG, "libdskgpu.so")      # DSKGPU_LIB: tuning builds under variants/

MAX_BANKS = 16
HISTO_LEN = 10001
HISTO2D_DIM2 = 11
NBINS = 65536
NBINS_MAX = 1 << 22

ERR_NODEVICE = -6


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("kmer_size", C.c_int32), ("minimizer_size", C.c_int32), ("nb_banks", C.c_int32),
        ("per_bank_counts", C.c_int32), ("solidity_kind", C.c_int32),
        ("abundance_min", C.c_int64 * MAX_BANKS), ("abundance_max", C.c_int64),
        ("solid_vec", C.c_uint8 * MAX_BANKS),
        ("histo2d", C.c_int32), ("device", C.c_int32), ("count_mode", C.c_int32), ("hash_log2_slots", C.c_int32),
        ("nb_partitions", C.c_int32), ("keep_results_on_device", C.c_int32),
        ("stream", C.c_void_p), ("rank", C.c_int32), ("world_size", C.c_int32), ("push_chunk_bytes", C.c_int32), ("smem_table_slots", C.c_int32), ("bank_histograms", C.c_int32),
        ("nb_passes", C.c_int32), ("pass_id", C.c_int32), ("sequence_stats", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("nb_sequences", C.c_uint64), ("nb_nucleotides", C.c_uint64), ("kmers_nb_valid", C.c_uint64),
        ("nb_superkmers", C.c_uint64), ("kmers_nb_distinct", C.c_uint64), ("kmers_nb_solid", C.c_uint64),
        ("nb_partitions", C.c_uint64), ("nb_groups_hash", C.c_uint64), ("nb_groups_sort", C.c_uint64),
        ("superkmer_bytes", C.c_uint64), ("gpu_launches", C.c_uint64),
        ("ms_parse", C.c_float), ("ms_superk", C.c_float), ("ms_partition", C.c_float), ("ms_count", C.c_float),
        ("ms_sort", C.c_float), ("ms_total", C.c_float), ("ms_dominant_kernel", C.c_float),
        ("dominant_kernel_launches", C.c_uint32), ("nb_parts_smem", C.c_uint32), ("nb_smem_splits", C.c_uint32),
        ("smem_table_slots", C.c_uint32), ("density_ppm", C.c_uint32), ("log2_bins", C.c_uint32), ("nb_groups_bucket", C.c_uint32), ("nb_hash_regroups", C.c_uint32),
        ("exchange_bytes_out", C.c_uint64), ("ms_exchange", C.c_float), ("nb_solid_regrows", C.c_uint32), ("kmers_in_pass", C.c_uint64),
        ("ms_plan", C.c_float), ("ms_push_wall", C.c_float), ("hist_rebuilt", C.c_uint32), ("scatter_passes", C.c_uint32), ("ms_count_heavy", C.c_float), ("sort_fallbacks", C.c_uint32),
        ("seq_stats_sequences", C.c_uint64), ("seq_len_min", C.c_uint64), ("seq_len_max", C.c_uint64), ("seq_len_sum", C.c_uint64),
        ("seq_len_sumsq", C.c_uint64), ("kmers_nb_invalid", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("reserved")}


# every symbol include/dskgpu.h declares (tests check the library exports all of them)
SYMBOLS = [
    "dskgpu_config_default", "dskgpu_create", "dskgpu_push_bytes", "dskgpu_push_device_bytes", "dskgpu_push_reads",
    "dskgpu_finish", "dskgpu_num_partitions", "dskgpu_partition", "dskgpu_partition_device", "dskgpu_histogram",
    "dskgpu_recount", "dskgpu_bank_histograms",
    "dskgpu_get_stats", "dskgpu_reset", "dskgpu_destroy", "dskgpu_host_alloc", "dskgpu_host_free", "dskgpu_strerror",
    "dskgpu_last_error", "dskgpu_device_count", "dskgpu_abi_version",
    "dskgpu_xchg_local_totals", "dskgpu_xchg_prepare", "dskgpu_xchg_set_global", "dskgpu_xchg_sketch", "dskgpu_xchg_set_sketch", "dskgpu_xchg_hist", "dskgpu_xchg_counts", "dskgpu_xchg_ensure_recv", "dskgpu_xchg_plan", "dskgpu_xchg_recv_buffer", "dskgpu_xchg_ipc_handle",
    "dskgpu_xchg_open_peer", "dskgpu_xchg_set_peers", "dskgpu_xchg_scatter", "dskgpu_xchg_sync", "dskgpu_xchg_layout", "dskgpu_record_bytes",
    "dskgpu_set_pass", "dskgpu_push_sync", "dskgpu_suggest_nb_passes", "dskgpu_xchg_close_peer", "dskgpu_multi_finish",
    "dskgpu_debug_plan",
    "dskgpu_selftest_scan", "dskgpu_selftest_minimizers", "dskgpu_selftest_superkmers", "dskgpu_suggest_minimizer_size", "dskgpu_selftest_wide_kmers", "dskgpu_selftest_plan", "dskgpu_selftest_wide_superkmers",
]

_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO):
        raise RuntimeError(
            "dsk_b200/libdskgpu.so is missing: build it with `python -m dsk_b200.build` "
            "(the counting path is CUDA-only, there is no CPU fallback)")
    L = C.CDLL(SO)
    P = C.POINTER
    L.dskgpu_config_default.argtypes = [P(Config)]
    L.dskgpu_config_default.restype = None
    L.dskgpu_create.argtypes = [P(Config), P(C.c_void_p)]
    L.dskgpu_push_bytes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    L.dskgpu_push_device_bytes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    L.dskgpu_push_reads.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    L.dskgpu_finish.argtypes = [C.c_void_p]
    L.dskgpu_set_pass.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.dskgpu_push_sync.argtypes = [C.c_void_p]
    L.dskgpu_suggest_nb_passes.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int]
    L.dskgpu_xchg_close_peer.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_multi_finish.argtypes = [C.c_void_p, C.c_int]
    L.dskgpu_num_partitions.argtypes = [C.c_void_p]
    L.dskgpu_partition.argtypes = [C.c_void_p, C.c_int, P(C.c_void_p), P(C.c_void_p), P(C.c_uint64), P(C.c_int)]
    L.dskgpu_partition_device.argtypes = [C.c_void_p, C.c_int, P(C.c_void_p), P(C.c_void_p), P(C.c_uint64), P(C.c_int)]
    L.dskgpu_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.dskgpu_get_stats.argtypes = [C.c_void_p, P(Stats)]
    L.dskgpu_recount.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_bank_histograms.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_reset.argtypes = [C.c_void_p]
    L.dskgpu_destroy.argtypes = [C.c_void_p]
    L.dskgpu_destroy.restype = None
    L.dskgpu_host_alloc.argtypes = [C.c_size_t]
    L.dskgpu_host_alloc.restype = C.c_void_p
    L.dskgpu_host_free.argtypes = [C.c_void_p]
    L.dskgpu_host_free.restype = None
    L.dskgpu_strerror.argtypes = [C.c_int]
    L.dskgpu_strerror.restype = C.c_char_p
    L.dskgpu_last_error.argtypes = [C.c_void_p]
    L.dskgpu_last_error.restype = C.c_char_p
    L.dskgpu_xchg_local_totals.argtypes = [C.c_void_p, P(C.c_uint64), P(C.c_uint64)]
    L.dskgpu_xchg_prepare.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_xchg_set_global.argtypes = [C.c_void_p, C.c_void_p, P(C.c_int)]
    L.dskgpu_xchg_hist.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_xchg_sketch.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_xchg_set_sketch.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_xchg_plan.argtypes = [C.c_void_p, C.c_void_p, P(C.c_uint32), P(C.c_uint32), C.c_void_p]
    L.dskgpu_xchg_counts.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_xchg_ensure_recv.argtypes = [C.c_void_p, C.c_uint64]
    L.dskgpu_debug_plan.argtypes = [C.c_void_p, P(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.dskgpu_debug_plan.restype = C.c_int64
    L.dskgpu_xchg_recv_buffer.argtypes = [C.c_void_p, P(C.c_void_p), P(C.c_size_t)]
    L.dskgpu_xchg_ipc_handle.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_xchg_open_peer.argtypes = [C.c_void_p, C.c_void_p, P(C.c_void_p)]
    L.dskgpu_xchg_set_peers.argtypes = [C.c_void_p, C.c_void_p]
    L.dskgpu_xchg_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.dskgpu_xchg_sync.argtypes = [C.c_void_p]
    L.dskgpu_xchg_layout.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.dskgpu_record_bytes.argtypes = [C.c_void_p]
    L.dskgpu_selftest_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
    L.dskgpu_selftest_scan.restype = C.c_int64
    L.dskgpu_selftest_minimizers.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.dskgpu_selftest_superkmers.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, P(C.c_uint64)]
    L.dskgpu_selftest_superkmers.restype = C.c_int64
    L.dskgpu_suggest_minimizer_size.argtypes = [C.c_uint64, C.c_int]
    L.dskgpu_selftest_wide_kmers.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    L.dskgpu_selftest_wide_kmers.restype = C.c_int64
    L.dskgpu_selftest_plan.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_double, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.dskgpu_selftest_plan.restype = C.c_int64
    L.dskgpu_selftest_wide_superkmers.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t, P(C.c_uint64)]
    L.dskgpu_selftest_wide_superkmers.restype = C.c_int64
    _LIB = L
    return L
