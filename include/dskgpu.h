/*
 * dskgpu.h -- C ABI of the B200-native DSK counting hot path (libdskgpu.so).
 *
 * This is the drop-in boundary for gatb-core's SortingCountAlgorithm<span>::execute()
 * (G/src/gatb/kmer/impl/SortingCountAlgorithm.cpp:636-781; G/ = thirdparty/gatb-core/gatb-core/):
 * everything between "bank bytes in" and "(canonical k-mer, count) per partition + abundance
 * histogram out".  Nothing like it exists in the reference (pure C++/pthreads, no FFI); each
 * entry point names the reference code it replaces.  Plain pointers and sizes only, no C++/torch
 * types, never throws; every call returns 0 or a negative DSKGPU_ERR_* code.
 *
 * One context drives one GPU (one process per GPU); multi-GPU runs create one context per rank
 * and route super-k-mers between ranks with the dskgpu_xchg_* calls.
 */
#ifndef DSKGPU_H
#define DSKGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSKGPU_ABI_VERSION   4
#define DSKGPU_MAX_BANKS     16
#define DSKGPU_HISTO_LEN     10001          /* bins 0..10000 (Histogram.hpp:92, length 10000) */
#define DSKGPU_HISTO2D_DIM2  11             /* bins 0..10   (CountProcessorHistogram.hpp:173-184) */
#define DSKGPU_MAX_MINIMIZER 15             /* -minimizer-size is clipped to this (32-bit m-mer arithmetic) */
#define DSKGPU_MAX_KMER      127            /* KSIZE_LIST "32 64 96 128": k<32 -> 64-bit keys, k<64 -> 128-bit, k<96 -> 192-bit, k<128 -> 256-bit */
#define DSKGPU_NBINS         65536          /* minimizer bins packed into partitions at finish: the coarsest level ... */
#define DSKGPU_NBINS_MAX     (1u << 24)     /* ... and the finest (jobs of tens of G k-mers); the level is picked from the job size */

/* error codes */
enum {
    DSKGPU_OK            =  0,
    DSKGPU_ERR_ARG       = -1,   /* bad argument / unsupported k (Integer.hpp:479 "unhandled kmer size") */
    DSKGPU_ERR_CUDA      = -2,   /* CUDA runtime failure (see dskgpu_last_error) */
    DSKGPU_ERR_NOMEM     = -3,
    DSKGPU_ERR_FORMAT    = -4,   /* input is not a FASTA / 4-line FASTQ stream the device scanner accepts */
    DSKGPU_ERR_STATE     = -5,   /* call order violated (push after finish, results before finish, ...) */
    DSKGPU_ERR_NODEVICE  = -6,   /* no CUDA device: there is NO CPU fallback on the counting path */
    DSKGPU_ERR_OVERFLOW  = -7    /* internal capacity exceeded */
};

/* -solidity-kind (G/src/gatb/kmer/impl/CountProcessorSolidity.hpp:186-300) */
enum { DSKGPU_SOLIDITY_SUM = 0, DSKGPU_SOLIDITY_MIN = 1, DSKGPU_SOLIDITY_MAX = 2,
       DSKGPU_SOLIDITY_ONE = 3, DSKGPU_SOLIDITY_ALL = 4, DSKGPU_SOLIDITY_CUSTOM = 5 };

/* how a partition is counted (SortingCountAlgorithm.cpp:1489-1497 picks vector vs hash by partition occupancy).
 * AUTO: partitions are sized for the shared-memory hash table of one SM (DSKGPU_COUNT_SMEM); partitions too big for it
 *       go to the L2-resident global hash table, and those too big for that to the radix-sort path.
 * SORT / HASH / SMEM force one path (SMEM still hands oversize partitions to HASH). */
enum { DSKGPU_COUNT_AUTO = 0, DSKGPU_COUNT_SORT = 1 /* PartitionsByVectorCommand */, DSKGPU_COUNT_HASH = 2 /* PartitionsByHashCommand */,
       DSKGPU_COUNT_SMEM = 3 };

/* input stream formats understood by the device record scanner (BankFasta.cpp:485-572) */
enum { DSKGPU_FMT_AUTO = 0, DSKGPU_FMT_FASTA = 1, DSKGPU_FMT_FASTQ = 2, DSKGPU_FMT_LINES = 3 /* one sequence per line */ };

/* push flags */
enum { DSKGPU_PUSH_LAST = 1 /* last chunk of this bank (file) */ };

typedef struct dskgpu_ctx dskgpu_ctx;

/* mirrors the options SortingCountAlgorithm reads (SortingCountAlgorithm.cpp:202-236) */
typedef struct dskgpu_config {
    int32_t  abi_version;                        /* DSKGPU_ABI_VERSION */
    int32_t  kmer_size;                          /* -kmer-size        (default 31) */
    int32_t  minimizer_size;                     /* -minimizer-size   (default 10; clipped to k-1, ConfigurationAlgorithm.cpp:249-251) */
    int32_t  nb_banks;                           /* number of input banks (files of a comma list), 1..DSKGPU_MAX_BANKS */
    int32_t  per_bank_counts;                    /* 0: banks are summed while counting (default "sum" path);
                                                    1: keep one count per bank (-histo2D, -solidity-kind != sum) */
    int32_t  solidity_kind;                      /* DSKGPU_SOLIDITY_* */
    int64_t  abundance_min[DSKGPU_MAX_BANKS];    /* -abundance-min (one value or one per bank) */
    int64_t  abundance_max;                      /* -abundance-max   (default 2^31-1) */
    uint8_t  solid_vec[DSKGPU_MAX_BANKS];        /* -solidity-custom */
    int32_t  histo2d;                            /* -histo2D */
    int32_t  device;                             /* CUDA device ordinal */
    int32_t  count_mode;                         /* DSKGPU_COUNT_* */
    int32_t  hash_log2_slots;                    /* 0 = auto; size of the L2-resident hash table */
    int32_t  nb_partitions;                      /* 0 = auto */
    int32_t  keep_results_on_device;             /* 1: dskgpu_finish does not copy the solid set to the host */
    void*    stream;                             /* cudaStream_t to run on, NULL = library-owned stream */
    int32_t  rank, world_size;                   /* multi-GPU: this context owns partitions p with p % world_size == rank */
    int32_t  push_chunk_bytes;                   /* 0 = default (64 MiB): granularity of the streamed H2D copy + scan */
    int32_t  smem_table_slots;                   /* 0 = auto (all the shared memory of an SM); tests shrink it to force splits */
    int32_t  bank_histograms;                    /* 1: also keep one abundance histogram per bank (dskgpu_bank_histograms; the first
                                                    pass of -abundance-min auto with -solidity-kind one/all/custom).  Global atomics
                                                    per distinct k-mer: leave 0 unless those cutoffs are needed */
    int32_t  nb_passes;                          /* 0/1 = one pass.  >1: this context keeps only the super-k-mers whose minimizer bin falls in */
    int32_t  pass_id;                            /*   pass `pass_id` (the reference's `minimizer % nbPass == pass`, SortingCountAlgorithm.cpp:1086); the
                                                    caller pushes the whole input once per pass (dskgpu_set_pass + dskgpu_reset in between) and the
                                                    results of the passes are disjoint: what makes a job whose records exceed HBM fit */
    int32_t  sequence_stats;                     /* 1: also gather the per-sequence statistics of BankStats (K/BankKmers.hpp:166-215): seq_size_min/max,
                                                  *    sum of squares (deviation) and the k-mer windows of every sequence (kmers_nb_invalid = windows - valid).
                                                  *    One more read of the code stream per chunk; the reference gets them for free while parsing */
    int32_t  reserved[2];
} dskgpu_config;

/* stats block: the keys of SortingCountAlgorithm::getInfo() (SortingCountAlgorithm.cpp:728-780) */
typedef struct dskgpu_stats {
    uint64_t nb_sequences;          /* bank/sequences/seq_number */
    uint64_t nb_nucleotides;        /* bank/bank_total_nt */
    uint64_t kmers_nb_valid;        /* bank/kmers/kmers_nb_valid */
    uint64_t nb_superkmers;         /* stats/temp_files/nb_superkmers */
    uint64_t kmers_nb_distinct;     /* stats/kmers/kmers_nb_distinct */
    uint64_t kmers_nb_solid;        /* stats/kmers/kmers_nb_solid */
    uint64_t nb_partitions;         /* stats/partitions/nb_partitions */
    uint64_t nb_groups_hash;        /* stats/partitions/kind/hash   */
    uint64_t nb_groups_sort;        /* stats/partitions/kind/vector */
    uint64_t superkmer_bytes;       /* stats/temp_files/total_size */
    uint64_t gpu_launches;          /* kernels launched by this context since create/reset */
    /* device time per stage, milliseconds (CUDA events on the context stream) */
    float ms_parse, ms_superk, ms_partition, ms_count, ms_sort, ms_total;
    float ms_dominant_kernel;       /* summed duration of the dominant counting kernel (smem count, hash insert or radix passes) */
    uint32_t dominant_kernel_launches;
    uint32_t nb_parts_smem;         /* partitions counted in shared memory */
    uint32_t nb_smem_splits;        /* table overflows answered by splitting a pass */
    uint32_t smem_table_slots;      /* capacity of the shared-memory table */
    uint32_t density_ppm;           /* sampled distinct / total k-mers x 1e6 (0 = sample too small) */
    uint32_t log2_bins;             /* minimizer-bin level the partitions were packed from (16..20) */
    uint32_t nb_groups_bucket;      /* heavy partitions: groups expanded into hash buckets of flat keys, counted in shared memory */
    uint32_t nb_hash_regroups;      /* global-table groups that outgrew the table (sized from an estimate) and were redone at the worst-case size */
    /* multi-GPU exchange (this rank): bytes stored into other ranks' HBM over NVLink, duration of the segment-copy kernel */
    uint64_t exchange_bytes_out;
    float    ms_exchange;
    uint32_t nb_solid_regrows;      /* counting stage redone because the solid-set buffers (sized from an estimate) were too small */
    uint64_t kmers_in_pass;         /* valid k-mers whose minimizer belongs to this context's pass (== kmers_nb_valid with one pass) */
    float    ms_plan;               /* density sample + device planner (scans, renumbering, layout tables) */
    float    ms_push_wall;          /* first to last kernel of the push phase on the stream (parse + super-k-mers + any gaps between chunks) */
    uint32_t hist_rebuilt;          /* 1: the packed minimizer-bin histogram could have wrapped and was rebuilt exactly from the records */
    uint32_t scatter_passes;        /* 0: single-pass partition scatter; n: MSD multi-split passes (jobs with millions of partitions) */
    float    ms_count_heavy;        /* part of ms_count spent on the heavy partitions (gather + global table / sort / bucket paths) */
    uint32_t sort_fallbacks;        /* ordering of the solid set: neighbourhood fix-up gave up (long groups of equal prefixes), full-width sort ran */
    /* per-sequence statistics (cfg.sequence_stats = 1; else all zero): bank/sequences/seq_size_{min,max,mean,deviation}, bank/kmers/kmers_nb_invalid */
    uint64_t seq_stats_sequences;   /* sequences measured (== nb_sequences) */
    uint64_t seq_len_min, seq_len_max;
    uint64_t seq_len_sum;           /* == nb_nucleotides */
    uint64_t seq_len_sumsq;         /* BankStats::sequencesTotalLengthSquare */
    uint64_t kmers_nb_invalid;      /* k-mer windows of all sequences that hold a non-ACGT letter (K/Sequence2SuperKmer.hpp:95-108) */
} dskgpu_stats;

/* fills *cfg with the reference defaults (SortingCountAlgorithm.cpp:208-231) */
void dskgpu_config_default(dskgpu_config* cfg);

/* replaces: the part of ConfigurationAlgorithm::execute (ConfigurationAlgorithm.cpp:245-467) that sizes the partitioning
 * from the estimated volume.  Returns the minimizer length to put in cfg->minimizer_size for a job of `expected_kmers`
 * k-mers over ALL ranks (k < 32: 10 up to 150 M k-mers, 11 up to 1.5 G, 12 up to 12 G, 14 beyond; k >= 32, whose 128-bit
 * tables hold fewer k-mers: 10 up to 40 M, 12 up to 150 M, 14 up to 16 G, 15 beyond; clipped to kmer_size-1): a partition cannot be lighter
 * than its heaviest minimizer bin, so the bins must shrink as the job grows to stay inside a shared-memory table.
 * Which partition a k-mer lands in is unobservable in the results.  Host-only, no device needed. */
int dskgpu_suggest_minimizer_size(uint64_t expected_kmers, int kmer_size);

/* replaces: the pass sizing of ConfigurationAlgorithm::execute (ConfigurationAlgorithm.cpp:245-467: nb_passes from the estimated
 * volume against -max-disk / -max-memory).  Here the bound is HBM: returns how many passes a job of `expected_kmers` k-mers over
 * `world_size` GPUs needs so that one pass's super-k-mer records (two copies: input order + partition order), the receive
 * buffer and the solid-set buffers fit in `hbm_bytes` per GPU (0 = ask the device: free memory of `device`).  >= 1. */
int dskgpu_suggest_nb_passes(uint64_t expected_kmers, int kmer_size, int world_size, uint64_t hbm_bytes, int device);

/* replaces: SortingCountAlgorithm ctor + configure() (SortingCountAlgorithm.cpp:525-625) */
int dskgpu_create(const dskgpu_config* cfg, dskgpu_ctx** out);

/* replaces: the pass loop of SortingCountAlgorithm::execute (SortingCountAlgorithm.cpp:678-689).  Valid on a context that holds
 * no pushed data (after create or reset): the next pushes keep the super-k-mers of pass `pass_id` of `nb_passes`. */
int dskgpu_set_pass(dskgpu_ctx* ctx, int pass_id, int nb_passes);

/* replaces: fillPartitions() for one chunk of one bank (SortingCountAlgorithm.cpp:1216-1349).
 * `bytes` are raw FASTA/FASTQ file bytes (already gunzipped), any chunking; records may straddle chunks.
 * The library scans records, 2-bit encodes, builds minimizers and super-k-mers on the device.
 * Pass DSKGPU_PUSH_LAST with the final chunk of the bank.  Host memory may be pageable or pinned
 * (dskgpu_host_alloc); pinned memory is copied without an intermediate staging copy. */
int dskgpu_push_bytes(dskgpu_ctx* ctx, int bank_id, const char* bytes, size_t n, int format, int flags);

/* Blocks until every host buffer handed to dskgpu_push_bytes so far has been copied to the device.  push_bytes returns as soon
 * as the H2D copy is QUEUED: a caller that refills a buffer (double-buffered file reader) must call this first, or not touch
 * a buffer until two further pushes have returned. */
int dskgpu_push_sync(dskgpu_ctx* ctx);

/* same, for bytes that already live in device memory (HBM-resident benchmark leg) */
int dskgpu_push_device_bytes(dskgpu_ctx* ctx, int bank_id, const void* dev_bytes, size_t n, int format, int flags);

/* replaces: IBank::iterator() -> Sequence2SuperKmer (Sequence2SuperKmer.hpp:138-159) for callers that
 * already hold parsed sequences: `bases` = concatenated sequences, offsets[nreads+1] delimit them. */
int dskgpu_push_reads(dskgpu_ctx* ctx, int bank_id, const char* bases, const uint64_t* offsets, size_t nreads);

/* replaces: fillSolidKmers() (SortingCountAlgorithm.cpp:1414-1607) + CountProcessor chain
 * (histogram -> solidity -> dump; CountProcessorChain.hpp:128-134): partitions the super-k-mers,
 * counts every partition (hash or sort), filters, histograms, sorts the solid set. */
int dskgpu_finish(dskgpu_ctx* ctx);

/* replaces: Partition<Count>& getSolidCounts() (SortingCountAlgorithm.hpp:66-192) */
int dskgpu_num_partitions(dskgpu_ctx* ctx);
/* partition p: *kmers -> n values of `words` little-endian uint64 each (words = 1 for k<32, 2 for k<64, 3 for k<96,
 * 4 for k<128 -- the spans of KSIZE_LIST "32 64 96 128", Integer.hpp:453-471; low word first), ascending as
 * LargeInt compares (most significant word first, LargeInt.hpp:502-509); *counts -> n uint32 abundances (Count::abundance, Abundance.hpp:108-125).
 * Host pointers owned by the context until destroy/reset. */
int dskgpu_partition(dskgpu_ctx* ctx, int p, const uint64_t** kmers, const uint32_t** counts, uint64_t* n, int* words);
/* same data, device pointers (valid until destroy/reset) */
int dskgpu_partition_device(dskgpu_ctx* ctx, int p, const void** d_kmers, const void** d_counts, uint64_t* n, int* words);

/* replaces: Histogram::save / .histo / .histo2D (Histogram.cpp:43-51; CountProcessorHistogram.hpp:104-159).
 * hist1d[i] = number of distinct k-mers of abundance i, with the reference's quirks (bins 0 and 10000
 * always 0, uint16 wrap).  hist2d may be NULL; layout hist2d[dim2 * 10001 + dim1]. */
int dskgpu_histogram(dskgpu_ctx* ctx, uint64_t* hist1d /*[10001]*/, uint64_t* hist2d /*[11*10001] or NULL*/);

/* replaces: the second fillSolidKmers pass of "-abundance-min auto" (SortingCountAlgorithm.cpp:1393-1402 run once per
 * processor of getDefaultProcessorVector, :454-514; thresholds updated by CountProcessorCustomProxy::endPass, :419-444).
 * Call after dskgpu_finish: counts the partitions again -- the super-k-mer records are still in HBM -- with new
 * abundance_min[nb_banks]; results, histograms and stats are replaced. */
int dskgpu_recount(dskgpu_ctx* ctx, const int64_t* abundance_min);
/* replaces: the per-bank histograms of CountProcessorCutoff (CountProcessorCutoff.hpp:101-116), from which the reference
 * derives one cutoff per bank for -solidity-kind one/all/custom.  hist[b * 10001 + i] = distinct k-mers whose count in
 * bank b is i (same quirks as dskgpu_histogram); needs cfg.bank_histograms = 1.  With summed banks (per_bank_counts = 0)
 * this is the 1-D histogram. */
int dskgpu_bank_histograms(dskgpu_ctx* ctx, uint64_t* hist /*[nb_banks * 10001]*/);

int dskgpu_get_stats(dskgpu_ctx* ctx, dskgpu_stats* out);

/* forget all pushed data and results, keep device buffers (next benchmark step / next pass) */
int dskgpu_reset(dskgpu_ctx* ctx);

void dskgpu_destroy(dskgpu_ctx* ctx);

/* pinned host memory helpers */
void* dskgpu_host_alloc(size_t n);
void  dskgpu_host_free(void* p);

const char* dskgpu_strerror(int code);
const char* dskgpu_last_error(dskgpu_ctx* ctx);     /* detail text of the last failure (ctx may be NULL) */
int  dskgpu_device_count(void);
int  dskgpu_abi_version(void);

/* ---- multi-GPU exchange (replaces the SuperKmerBinFiles temp tier, Storage.cpp:310-589) --------------
 * One context per rank (one process per GPU).  Partition p is owned by rank p % world_size (world_size <= 16).  All the
 * metadata stays on the device: the caller only moves DEVICE buffers between ranks (NCCL through torch.distributed, or peer
 * copies inside one process) and the host reads one small header.  After all pushes:
 *   0. xchg_prepare      -> this rank's {k-mers, records, density-sample k-mers, density-sample distinct}; all-reduce (sum) the
 *                           four numbers, hand the sums to xchg_set_global: every rank then agrees on the bin level
 *                           (2^16 .. 2^22 bins) and on the partition size.  The distinct / total ratio of the JOB needs the
 *                           distinct k-mers of the UNION of the ranks' samples: xchg_sketch returns HyperLogLog registers of
 *                           this rank's sample, all-reduce them with MAX and hand them to xchg_set_sketch before xchg_set_global
 *   1. xchg_hist         copies this rank's (records, k-mers) per minimizer bin, [2 * B] u64, into a DEVICE buffer of the
 *                           caller, who all-reduces it in place
 *   2. xchg_plan         plans the partitions ON THE DEVICE from the all-reduced histogram (dsk_b200/csrc/plan.cuh); every
 *                           rank derives the same plan.  Returns P, the partitions per rank PW = ceil(P / W) and
 *                           need_records[r] = records rank r receives (the same numbers on every rank, so every rank knows
 *                           when a peer has to grow its receive buffer and the IPC handles must travel again)
 *   3. xchg_counts       copies [W][PW] u64 = this rank's records of every partition, grouped by owner (row o goes to rank
 *                           o: an all-to-all of rows), followed by [W] u64 = records this rank holds for each rank (all-gather
 *                           them into the [W][W] chunk matrix), into a DEVICE buffer of the caller ((W * PW + W) u64)
 *   4. xchg_ensure_recv  sizes the receive buffer (capacity >= own need), then xchg_ipc_handle / open_peer / set_peers
 *   5. xchg_scatter      d_recv_counts = [W][PW] rows received in step 3 (row s = rank s's records of MY partitions),
 *                           d_send_matrix = [W][W] from step 3.  Scatters the local records into owner-major partition order
 *                           and stores this rank's chunk for every other rank into its receive buffer with ONE contiguous
 *                           copy per peer (16-byte vector stores through the peer pointers: NVLink P2P, no NCCL on the data
 *                           path, sender-major receive layout).  A rank's own chunk never moves.
 *   6. a stream-ordered barrier of the caller (e.g. a one-element all-reduce on the same stream), then dskgpu_finish counts the
 *      owned partitions: the counting kernel reads every partition as W segments, one per sender. */
int dskgpu_xchg_local_totals(dskgpu_ctx* ctx, uint64_t* kmers, uint64_t* records);
int dskgpu_xchg_prepare(dskgpu_ctx* ctx, uint64_t* local4 /*[4]*/);
#define DSKGPU_SKETCH_LEN 4096
int dskgpu_xchg_sketch(dskgpu_ctx* ctx, uint32_t* sketch /*[DSKGPU_SKETCH_LEN]*/);
int dskgpu_xchg_set_sketch(dskgpu_ctx* ctx, const uint32_t* merged /*[DSKGPU_SKETCH_LEN], element-wise max over ranks*/);
int dskgpu_xchg_set_global(dskgpu_ctx* ctx, const uint64_t* global4 /*[4] sums over ranks*/, int* log2_bins /*out: B = 1 << *log2_bins*/);
int dskgpu_xchg_hist(dskgpu_ctx* ctx, void* d_hist_out /*device, [2*B] u64*/);
int dskgpu_xchg_plan(dskgpu_ctx* ctx, const void* d_global_hist /*device, [2*B] u64*/, uint32_t* nparts, uint32_t* parts_per_rank,
                     uint64_t* need_records /*[world_size] or NULL*/);
int dskgpu_xchg_counts(dskgpu_ctx* ctx, void* d_out /*device, [W*PW + W] u64*/);
int dskgpu_xchg_ensure_recv(dskgpu_ctx* ctx, uint64_t capacity_records);
int dskgpu_xchg_recv_buffer(dskgpu_ctx* ctx, void** d_recv, size_t* bytes);
int dskgpu_xchg_ipc_handle(dskgpu_ctx* ctx, void* handle64 /*cudaIpcMemHandle_t of the receive buffer*/);
int dskgpu_xchg_open_peer(dskgpu_ctx* ctx, const void* handle64, void** d_ptr);
int dskgpu_xchg_close_peer(dskgpu_ctx* ctx, void* d_ptr /*from xchg_open_peer: unmap a peer buffer that was re-allocated*/);
int dskgpu_xchg_set_peers(dskgpu_ctx* ctx, void* const* d_peer_recv /*[world_size]; own entry ignored*/);
int dskgpu_xchg_scatter(dskgpu_ctx* ctx, const void* d_recv_counts /*device, [W][PW] u64*/, const void* d_send_matrix /*device, [W][W] u64*/);
int dskgpu_xchg_sync(dskgpu_ctx* ctx);
/* ---- several GPUs inside ONE process (what the `dsk_gpu` CLI does with DSKGPU_DEVICES): ctxs[r] is the context of rank r
 * (cfg.rank = r, cfg.world_size = n, any device each -- peer access is enabled between distinct devices; contexts may also
 * share a device).  After all pushes (one host thread per context is fine), this runs the whole exchange above with peer
 * reads / copies in place of the collectives -- totals, bin histogram, device plan, count rows, receive buffers, scatter + one
 * contiguous peer copy per rank pair -- and then counts the owned partitions of every context concurrently (one host thread
 * each).  Results are read per context, as after dskgpu_finish. */
int dskgpu_multi_finish(dskgpu_ctx* const* ctxs, int n);

/* host-only mirror of rank `owner`'s receive layout (CPU test-suite): counts[s * W * PW + q] = records rank s holds for the
 * partition at position q = (p % W) * PW + p / W.  region_base[s] = first record of sender s's chunk in the receive buffer
 * (senders != owner in rank order; region_base[W] = records received); seg_off[s * PW + j] = first record of (sender s, owned
 * job j), inside the receive buffer for s != owner, inside the owner's own chunk for s == owner. */
int dskgpu_xchg_layout(int world_size, uint32_t parts_per_rank, const uint64_t* counts, int owner, uint64_t* region_base /*[W+1]*/, uint64_t* seg_off /*[W*PW]*/);
int dskgpu_record_bytes(dskgpu_ctx* ctx);

/* ---- host-side self checks of the device bit logic (no GPU needed; used by the CPU test-suite) ------- */
/* runs the record-scanner state machine sequentially on the host; out gets one byte per emitted code
 * (0..3 base, 4|code invalid base, 8 separator).  Returns number of codes or <0. */
int64_t dskgpu_selftest_scan(const char* bytes, size_t n, int format, uint8_t* out, size_t out_cap);
/* minimizer of every k-mer window of a code string (host copy of the device function).  Values below 4^m are the reference's
 * minimizer (K/Model.hpp:1254-1287); a value with bit 2m set means "no allowed m-mer in this window": the reference uses its
 * default minimizer (4^m - 1) there, the device keeps the smallest banned m-mer instead so that such k-mers spread over bins */
int dskgpu_selftest_minimizers(const uint8_t* codes, size_t n, int k, int m, uint32_t* out_min, uint8_t* out_valid);
/* super-k-mer packing round trip: codes -> records -> canonical k-mers (host copy of pack + expand) */
int64_t dskgpu_selftest_superkmers(const uint8_t* codes, size_t n, int k, int m, uint64_t* out_kmers /*[n*words]*/, size_t cap, uint64_t* n_records);
/* the partition planner on the host alone (what every rank derives from the all-reduced minimizer-bin histogram):
 * hist arrays are [2 << level] = records per bin, then k-mers per bin; returns the number of partitions */
int64_t dskgpu_selftest_plan(int level, const uint64_t* global_hist, const uint64_t* local_hist, int world_size, int nb_counts,
                             uint32_t smem_slots, double density, int count_mode, int forced_nb_partitions,
                             uint32_t* bin2part, uint64_t* part_kmers, uint64_t* part_local_recs, size_t max_parts);
/* the plan a context derived on the DEVICE (after dskgpu_finish or dskgpu_xchg_plan), for the tests that compare it with the
 * host mirror above: same outputs (+ whole-job records per partition); returns the number of partitions */
int64_t dskgpu_debug_plan(dskgpu_ctx* ctx, int* level, uint32_t* bin2part, uint64_t* part_kmers, uint64_t* part_recs, uint64_t* part_local_recs, size_t max_parts);
/* wide spans groundwork (k <= 127, dsk_b200/csrc/kmer_wide.cuh; the counting path itself still rejects k >= 64): canonical
 * k-mers of a code stream as 4 words each, computed by rolling and by extraction from a packed record (must agree) */
int64_t dskgpu_selftest_wide_kmers(const uint8_t* codes, size_t n, int k, uint64_t* out_words /*[n-k+1][4]*/, uint8_t* out_valid);

/* super-k-mer record round trip with the N-word logic (k <= 127, records of 2, 4, 6 or 8 words); 4 words per k-mer out */
int64_t dskgpu_selftest_wide_superkmers(const uint8_t* codes, size_t n, int k, int m, uint64_t* out_kmers /*[cap][4]*/, size_t cap, uint64_t* n_records);

#ifdef __cplusplus
}
#endif
#endif /* DSKGPU_H */
