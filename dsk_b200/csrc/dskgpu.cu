// dskgpu.cu -- C ABI (include/dskgpu.h) and host orchestration of the B200 counting path.
//
// Stage map (reference -> here; K/ = thirdparty/gatb-core/gatb-core/src/gatb/kmer/impl/):
//   fillPartitions  (K/SortingCountAlgorithm.cpp:1216-1349)  -> push_*: K1 scan.cuh, K2 superk.cuh (per chunk, streaming)
//   SuperKmerBinFiles temp tier (Storage.cpp:310-589)        -> records stay in HBM; K3 partition scatter at finish
//   fillSolidKmers  (K/SortingCountAlgorithm.cpp:1414-1607)  -> finish: groups of partitions counted by the hash path
//                                                               (count.cuh) or the sort path (radix.cuh + k_rle_emit)
//   CountProcessor chain                                     -> fused into k_hash_scan / k_rle_emit
// There is no CPU fallback: without a CUDA device dskgpu_create fails with DSKGPU_ERR_NODEVICE.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dskgpu.h"
#include "kmer_bits.cuh"
#include "scan.cuh"
#include "seqstats.cuh"
#include "superk.cuh"
#include "count.cuh"
#include "plan.cuh"
#include "count_smem.cuh"
#include "kmer_wide.cuh"
#include "radix.cuh"

using namespace dsk;

static thread_local std::string g_last_error;

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct dskgpu_ctx {
    dskgpu_config cfg;
    int k = 0, m = 0, KW = 1, RW = 2, NB = 1;         // NB = counts kept per k-mer
    int nb_passes = 1, pass_id = 0;                  // this context keeps the super-k-mers whose minimizer bin % nb_passes == pass_id
    cudaStream_t stream = nullptr; bool own_stream = false;
    cudaStream_t copy_stream = nullptr;
    int state = 0;                                   // 0 accepting pushes, 1 finished
    std::string err;
    // device state
    DevBuf ss, ctr, hist, hist2d, bank_hist, raw[2], codes, tabs, tin;
    DevBuf seqst, seqtab; SeqStats* h_seqst = nullptr;   // per-sequence statistics (cfg.sequence_stats): accumulators, per-tile separator table
    DevBuf recs, meta;                               // staging records (input order)
    DevBuf precs;                                    // partitioned records
    DevBuf cursor, bin_hist, bin_fold, bin2part, work_ctr;
    DevBuf sample_recs, stab_keys, stab_counts, hll; // density sample: selected records, small hash table, HyperLogLog registers of its distinct k-mers
    u32* h_hll = nullptr; double sketch_distinct = -1.0;   // pinned copy; distinct k-mers of the union of all ranks' samples (merged sketch), < 0 = not given
    DevBuf tkeys, tcounts;                           // hash table
    DevBuf skeys[2], svals[2];                       // solid (k-mer, abundance) ping-pong
    DevBuf keys[2], banks[2];                        // sort path ping-pong
    DevBuf whash[2], widx[2];                        // wide spans: (64-bit hash, index) pairs of the keys, ping-pong
    DevBuf rs_hist, rs_status, rs_tilectr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_probe[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_push0 = nullptr, ev_push1 = nullptr; bool push_timed = false;   // first / last push kernel of the job (wall time of the push phase)
    // host pinned mirrors
    Counters* h_ctr = nullptr; StreamState* h_ss = nullptr;
    unsigned long long* h_nrec_probe = nullptr;
    u64* h_skeys = nullptr; u32* h_svals = nullptr; size_t h_solid_cap = 0;
    unsigned long long* h_hist = nullptr;            // [10001 + 11*10001]
    // stream bookkeeping
    int cur_bank = -1, cur_fmt = 0; bool stream_open = false; int pending_cr = 0;
    u64 rec_cap = 0; u64 nrec_known = 0; int chunk_parity = 0;
    // record-count probes of the chunks in flight (the host stays two chunks ahead of the device instead of draining it per chunk)
    struct Probe { int slot; u64 bytes; };
    std::vector<Probe> probes; int probe_slot = 0;
    size_t push_chunk = (size_t)64 << 20;
    // results
    u64 n_solid = 0; int solid_buf = 0; bool results_on_host = false; u64 solid_cap = 0;
    int sort_src = 0; bool sort_fixup = false;       // ordering of the solid set: input buffer of the fix-up pass, whether it ran
    u32 nparts = 0;
    u32 smem_cap = 0; int num_sms = 148;
    // device-side plan (plan.cuh): scratch of the planner, q-ordered per-partition tables, exchange table
    DevBuf pl_ex, pl_E, pl_H, pl_bsum, pl_pk, pl_pr, pl_pl, pl_newid, pl_nvals, pl_hdr;
    DevBuf gk_q, gr_q, lcnt_q, loff, xX, xtab, xS, hoff, hrecs;
    PlanHdr* h_hdr = nullptr; XchgTab* h_xtab = nullptr;             // pinned mirrors
    bool planned = false, hist_suspect = false;
    u32 nl_me = 0, np_me = 0;                        // owned jobs: [0, nl_me) counted in shared memory, [nl_me, np_me) heavy
    std::vector<u64> heavy_recs, heavy_kmers;        // whole-job records / k-mers of the owned heavy partitions (increasing id)
    u64* h_heavy = nullptr; size_t h_heavy_cap = 0;  // pinned staging for them
    const void* rcnt_dev = nullptr;                  // [W][PW] records of my partitions held by every rank (W = 1: lcnt_q)
    // multi-GPU
    std::vector<void*> peer_recv; bool xchg_planned = false; bool xchg_scattered = false;
    std::vector<void*> ipc_opened;
    bool totals_done = false; u64 local_nrec = 0, local_nkm = 0;     // records / k-mers this context holds (this pass)
    u64 bank_nkm = 0;                                // valid k-mers of everything pushed (all passes)
    u64 sample_solid = 0;
    u64 bytes_pushed = 0;                            // raw input bytes so far (sizes the density sample before the totals are known)
    bool sample_queued = false; u64 sample_nkm = 0, sample_distinct = 0; double sample_wmult = 0.0;   // wmult: occurrence-weighted multiplicity (this rank's sample)
    bool global_set = false; u64 g_total_kmers = 0, g_total_recs = 0; double density = 1.0; bool density_known = false;
    int bin_level = NBINS_LOG2; bool hist_fetched = false;
    int fine_log2 = 22;                              // fine histogram bins (2^22 = 32 MB: stays L2-resident under the record stream)
    // records in q order (owner-major partition order): what the counting kernels read on one GPU, what crosses NVLink as
    // W - 1 contiguous chunks on several (precs is then the receive buffer)
    DevBuf lrecs, xpeers, bcur, ghist, mkeys;
    void* qrecs = nullptr;                           // the q-ordered records after the scatter (lrecs, or recs after an even number of MSD passes)
    u64 xchg_bytes_out = 0;
    dskgpu_stats st;
    // timing
    std::vector<cudaEvent_t> evpool; size_t ev_used = 0;
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    char b_[512]; snprintf(b_, sizeof b_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    if (ctx) ctx->err = b_; g_last_error = b_; return DSKGPU_ERR_CUDA; } } while (0)
#define FAIL(code, ...) do { char b_[512]; snprintf(b_, sizeof b_, __VA_ARGS__); if (ctx) ctx->err = b_; g_last_error = b_; return (code); } while (0)
#define LAUNCHED() do { ctx->st.gpu_launches++; } while (0)
// key width dispatch: KW = 64-bit words per k-mer (1: k < 32, 2: k < 64, 3: k < 96, 4: k < 128 -- Integer.hpp:463, KSIZE_LIST "32 64 96 128")
#define KW_DISPATCH(ctx_, fn, ...) ((ctx_)->KW == 1 ? fn<1>(__VA_ARGS__) : (ctx_)->KW == 2 ? fn<2>(__VA_ARGS__) : (ctx_)->KW == 3 ? fn<3>(__VA_ARGS__) : fn<4>(__VA_ARGS__))
static inline int kw_of(int k) { return k < 32 ? 1 : k < 64 ? 2 : k < 96 ? 3 : 4; }
// every ABI entry point runs on the context's device, whatever the caller's current device is (several contexts, one per
// GPU, can live in one process: host/GpuSortingCount.hpp with DSKGPU_DEVICES, dskgpu_multi_finish)
static inline void use_device(const dskgpu_ctx* ctx) { int d = -1; if (cudaGetDevice(&d) != cudaSuccess || d != ctx->cfg.device) cudaSetDevice(ctx->cfg.device); }

// DSKGPU_TRACE=1: host wall-clock of the stages of push / finish on stderr (tuning aid)
#include <chrono>
#include <unistd.h>
static bool g_trace = getenv("DSKGPU_TRACE") != nullptr;
static std::chrono::steady_clock::time_point g_t0;
static void trace(const char* what)
{
    if (!g_trace) return;
    auto now = std::chrono::steady_clock::now();
    if (!what) { g_t0 = now; return; }
    fprintf(stderr, "[dskgpu pid %d] %-48s +%8.3f ms\n", (int)getpid(), what, std::chrono::duration<double, std::milli>(now - g_t0).count());
}

static int ensure(dskgpu_ctx* ctx, DevBuf& b, size_t bytes, bool keep = false, size_t keep_bytes = 0)
{
    if (bytes <= b.cap) return 0;
    size_t ncap = std::max(bytes, b.cap + b.cap / 2);
    ncap = (ncap + 255) & ~(size_t)255;
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) { (void)cudaGetLastError(); FAIL(DSKGPU_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", ncap, cudaGetErrorString(e)); }
    if (keep && b.p && keep_bytes) { CK(cudaMemcpyAsync(np, b.p, keep_bytes, cudaMemcpyDeviceToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
    if (b.p) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(b.p); }
    b.p = np; b.cap = ncap;
    return 0;
}

enum { SPAN_PARSE = 0, SPAN_SUPERK = 1, SPAN_PART = 2, SPAN_COUNT = 3, SPAN_SORT = 4, SPAN_DOM = 5, SPAN_TOTAL = 6, SPAN_SORTPASS = 7, SPAN_XCHG = 8, SPAN_PLAN = 9, SPAN_HEAVY = 10 };

static cudaEvent_t get_event(dskgpu_ctx* ctx)
{
    if (ctx->ev_used == ctx->evpool.size()) { cudaEvent_t e; cudaEventCreate(&e); ctx->evpool.push_back(e); }
    return ctx->evpool[ctx->ev_used++];
}
struct SpanGuard {
    dskgpu_ctx* c; cudaEvent_t a; int kind;
    SpanGuard(dskgpu_ctx* ctx, int kind_) : c(ctx), kind(kind_) { a = get_event(ctx); cudaEventRecord(a, ctx->stream); }
    ~SpanGuard() { cudaEvent_t b = get_event(c); cudaEventRecord(b, c->stream); c->spans.push_back({a, b, kind}); }
};

extern "C" {

void dskgpu_config_default(dskgpu_config* c)
{
    memset(c, 0, sizeof(*c));
    c->abi_version = DSKGPU_ABI_VERSION;
    c->kmer_size = 31; c->minimizer_size = 10; c->nb_banks = 1; c->per_bank_counts = 0;
    c->solidity_kind = DSKGPU_SOLIDITY_SUM;
    for (int i = 0; i < DSKGPU_MAX_BANKS; i++) { c->abundance_min[i] = 2; c->solid_vec[i] = 1; }
    c->abundance_max = 2147483647LL;
    c->count_mode = DSKGPU_COUNT_AUTO; c->world_size = 1;
}

const char* dskgpu_strerror(int code)
{
    switch (code) {
    case DSKGPU_OK: return "ok";
    case DSKGPU_ERR_ARG: return "bad argument / unhandled kmer size";
    case DSKGPU_ERR_CUDA: return "CUDA runtime error";
    case DSKGPU_ERR_NOMEM: return "out of device memory";
    case DSKGPU_ERR_FORMAT: return "input format not accepted by the device record scanner";
    case DSKGPU_ERR_STATE: return "call order violated";
    case DSKGPU_ERR_NODEVICE: return "no CUDA device (the counting path has no CPU fallback)";
    case DSKGPU_ERR_OVERFLOW: return "internal capacity exceeded";
    }
    return "unknown error";
}
const char* dskgpu_last_error(dskgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }
int dskgpu_abi_version(void) { return DSKGPU_ABI_VERSION; }
int dskgpu_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; } return n; }
void* dskgpu_host_alloc(size_t n) { void* p = nullptr; if (cudaMallocHost(&p, n) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; } return p; }
void dskgpu_host_free(void* p) { if (p) cudaFreeHost(p); }

int dskgpu_create(const dskgpu_config* cfg, dskgpu_ctx** out)
{
    dskgpu_ctx* ctx = nullptr;
    if (!cfg || !out) FAIL(DSKGPU_ERR_ARG, "null argument");
    if (cfg->abi_version != DSKGPU_ABI_VERSION) FAIL(DSKGPU_ERR_ARG, "abi version mismatch");
    if (cfg->kmer_size < 2 || cfg->kmer_size > DSKGPU_MAX_KMER) FAIL(DSKGPU_ERR_ARG, "Failure because of unhandled kmer size %d", cfg->kmer_size);
    if (cfg->nb_banks < 1 || cfg->nb_banks > DSKGPU_MAX_BANKS) FAIL(DSKGPU_ERR_ARG, "nb_banks %d out of range", cfg->nb_banks);
    if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) FAIL(DSKGPU_ERR_ARG, "bad rank/world_size");
    if (dskgpu_device_count() <= 0) FAIL(DSKGPU_ERR_NODEVICE, "no CUDA device visible");
    ctx = new dskgpu_ctx();
    ctx->cfg = *cfg;
    ctx->k = cfg->kmer_size;
    int m = cfg->minimizer_size > 0 ? cfg->minimizer_size : 10;
    if (m > ctx->k - 1) m = ctx->k - 1;              // ConfigurationAlgorithm.cpp:249-251
    if (m > DSKGPU_MAX_MINIMIZER) m = DSKGPU_MAX_MINIMIZER;   // m-mer values are 32-bit (kmer_bits.cuh: mmer_value)
    if (m < 2) m = 2;
    if (m > ctx->k) m = ctx->k;
    ctx->m = m;
    ctx->KW = kw_of(ctx->k);                         // Integer.hpp:463 : span = first K in KSIZE_LIST with k < K
    // spans 96 / 128: 192/256-bit keys have no single-instruction claim (the widest CAS is 128 bits), so
    // their partitions are counted by the sort path -- what the reference itself does for every partition that does not
    // take its hash path (PartitionsByVectorCommand, K/PartitionsCommand.cpp:1600-1805)
    if (ctx->KW > 2) ctx->cfg.count_mode = DSKGPU_COUNT_SORT;
    ctx->RW = 2 * ctx->KW;
    ctx->NB = (cfg->per_bank_counts && cfg->nb_banks > 1) ? cfg->nb_banks : 1;
    if (cfg->push_chunk_bytes > 0) ctx->push_chunk = (size_t)std::max(cfg->push_chunk_bytes, 64);
    ctx->nb_passes = cfg->nb_passes > 1 ? cfg->nb_passes : 1;
    ctx->pass_id = cfg->nb_passes > 1 ? cfg->pass_id : 0;
    if (ctx->pass_id < 0 || ctx->pass_id >= ctx->nb_passes) { delete ctx; ctx = nullptr; FAIL(DSKGPU_ERR_ARG, "pass_id %d outside [0, nb_passes = %d)", cfg->pass_id, cfg->nb_passes); }
    memset(&ctx->st, 0, sizeof ctx->st);
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { delete ctx; ctx = nullptr; FAIL(DSKGPU_ERR_CUDA, "cudaSetDevice(%d): %s", cfg->device, cudaGetErrorString(e)); }
    if (cfg->stream) ctx->stream = (cudaStream_t)cfg->stream;
    else { CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) { CK(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming)); }
    for (int i = 0; i < 4; i++) CK(cudaEventCreateWithFlags(&ctx->ev_probe[i], cudaEventDisableTiming));
    CK(cudaEventCreate(&ctx->ev_push0)); CK(cudaEventCreate(&ctx->ev_push1));
    CK(cudaMallocHost((void**)&ctx->h_ctr, sizeof(Counters)));
    CK(cudaMallocHost((void**)&ctx->h_ss, sizeof(StreamState)));
    CK(cudaMallocHost((void**)&ctx->h_nrec_probe, 64));
    CK(cudaMallocHost((void**)&ctx->h_hist, sizeof(unsigned long long) * (DSKGPU_HISTO_LEN * (1 + DSKGPU_HISTO2D_DIM2))));
    CK(cudaMallocHost((void**)&ctx->h_hdr, sizeof(PlanHdr)));
    CK(cudaMallocHost((void**)&ctx->h_hll, sizeof(u32) * HLL_M));
    CK(cudaMallocHost((void**)&ctx->h_xtab, sizeof(XchgTab)));
    CK(cudaMallocHost((void**)&ctx->h_seqst, sizeof(SeqStats)));
    memset(ctx->h_seqst, 0, sizeof(SeqStats));
    int rc;
    if ((rc = ensure(ctx, ctx->seqst, sizeof(SeqStats)))) return rc;
    // 2^22 bins (32 MB) stay L2-resident while the records stream by; 2^24 bins (128 MB > L2) tripled k_superkmers on the 9 G
    // k-mer-per-GPU job (69 -> 190 ms, profiles/r03f against r03b).  Partitions finer than a bin come from the record sub-bins.
    ctx->fine_log2 = 22;
    if (const char* e = getenv("DSKGPU_FINE_LOG2")) ctx->fine_log2 = std::min(NBINS_FINE_LOG2_MAX, std::max(NBINS_LOG2, atoi(e)));
    if ((rc = ensure(ctx, ctx->bin_hist, sizeof(unsigned long long) << ctx->fine_log2))) return rc;
    if ((rc = ensure(ctx, ctx->bin_fold, (sizeof(unsigned long long) * 2) << ctx->fine_log2))) return rc;
    if ((rc = ensure(ctx, ctx->work_ctr, 64))) return rc;
    if ((rc = ensure(ctx, ctx->hll, sizeof(u32) * HLL_M))) return rc;
    if ((rc = ensure(ctx, ctx->ss, sizeof(StreamState)))) return rc;
    if ((rc = ensure(ctx, ctx->ctr, sizeof(Counters)))) return rc;
    if ((rc = ensure(ctx, ctx->hist, sizeof(unsigned long long) * DSKGPU_HISTO_LEN))) return rc;
    if ((rc = ensure(ctx, ctx->hist2d, sizeof(unsigned long long) * DSKGPU_HISTO_LEN * DSKGPU_HISTO2D_DIM2))) return rc;
    if (ctx->NB > 1 && cfg->bank_histograms && (rc = ensure(ctx, ctx->bank_hist, sizeof(unsigned long long) * DSKGPU_HISTO_LEN * (size_t)ctx->NB))) return rc;
    // dynamic shared memory opt-in for the one-sweep kernels
    const int smem1 = RsCfg<1>::TILE * 8 + RsCfg<1>::TILE * 4, smem2 = RsCfg<2>::TILE * 16 + RsCfg<2>::TILE * 4;
    const int smem3 = RsCfg<3>::TILE * 24 + RsCfg<3>::TILE * 4, smem4 = RsCfg<4>::TILE * 32 + RsCfg<4>::TILE * 4;
    CK(cudaFuncSetAttribute(k_rs_onesweep<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    CK(cudaFuncSetAttribute(k_rs_onesweep<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    CK(cudaFuncSetAttribute(k_rs_onesweep<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    CK(cudaFuncSetAttribute(k_rs_onesweep<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    CK(cudaFuncSetAttribute(k_rs_onesweep<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    CK(cudaFuncSetAttribute(k_rs_onesweep<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    CK(cudaFuncSetAttribute(k_rs_onesweep<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
    CK(cudaFuncSetAttribute(k_rs_onesweep<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
    // shared-memory counting path: CS_CTAS_PER_SM CTAs share the SM's shared memory; each table takes what is left of its share
    {
        int nsm = 0;
        CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, cfg->device));
        ctx->num_sms = nsm > 0 ? nsm : 148;
    }
    if (ctx->KW <= 2) {
        int max_optin = 0, per_sm = 0, reserved = 0;
        CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
        CK(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, cfg->device));
        CK(cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, cfg->device));
        // the kernel variant this context will launch: 64/128-bit keys x (one summed count | one count per bank)
        const bool mb = ctx->NB > 1;
        const void* fn = ctx->KW == 1 ? (mb ? (const void*)k_count_smem<1, true> : (const void*)k_count_smem<1, false>)
                                      : (mb ? (const void*)k_count_smem<2, true> : (const void*)k_count_smem<2, false>);
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, fn));
        size_t avail = std::min<size_t>((size_t)max_optin, (size_t)per_sm / CS_CTAS_PER_SM - (size_t)reserved) - fa.sharedSizeBytes;
        const size_t fixed = ctx->KW == 1 ? cs_smem_bytes<1>(0) : cs_smem_bytes<2>(0);
        u32 cap = avail > fixed ? (u32)((avail - fixed) / (size_t)(8 * ctx->KW + 4 * ctx->NB)) : 0;
        cap = cap / 256 * 256;
        if (cap > 16384u) cap = 16384u;                            // the sweep keeps one solid bit per slot of a thread in 32 bits
        if (cfg->smem_table_slots > 0) cap = std::min<u32>(cap, std::max<u32>(64u, (u32)cfg->smem_table_slots / 4 * 4));
        if (ctx->NB > CS_MAX_BANKS) cap = 0;                       // many banks: the global-table path
        ctx->smem_cap = cap;
        const size_t dyn = ctx->KW == 1 ? cs_smem_bytes<1>(cap, ctx->NB) : cs_smem_bytes<2>(cap, ctx->NB);
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        // the flat-key variant (heavy partitions) shares the table geometry
        const void* fk = ctx->KW == 1 ? (mb ? (const void*)k_count_smem<1, true, true> : (const void*)k_count_smem<1, false, true>)
                                      : (mb ? (const void*)k_count_smem<2, true, true> : (const void*)k_count_smem<2, false, true>);
        cudaFuncAttributes fb;
        CK(cudaFuncGetAttributes(&fb, fk));
        if (fb.sharedSizeBytes > fa.sharedSizeBytes) FAIL(DSKGPU_ERR_CUDA, "internal: key variant of k_count_smem needs more static shared memory");
        CK(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        CK(cudaFuncSetAttribute(fk, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    *out = ctx;
    int r = dskgpu_reset(ctx);
    if (r) { *out = nullptr; return r; }
    return DSKGPU_OK;
}

// The partition of a k-mer is a pure function of its minimizer, so a partition can never be lighter than its heaviest
// minimizer bin: ~9e-5 of the job at m = 10 (k = 31), about 16 times less for every two more letters.  A shared-memory
// table takes ~30 K k-mers, so the minimizer length has to grow with the job to keep the bins inside its reach -- the
// role ConfigurationAlgorithm (K/ConfigurationAlgorithm.cpp:245-467) gives to nb_partitions, which it derives from the
// estimated volume.  The choice only moves k-mers between partitions: unobservable in the results.
int dskgpu_suggest_minimizer_size(uint64_t expected_kmers, int kmer_size)
{
    // measured on B200 (profiles/r02z*): C2 (400 M k-mers, k = 31): m = 10 -> 10.77 ms/step (21 table splits), 11 -> 10.51,
    // 12 -> 10.65 (more records); the same reads at k = 63 (293 M k-mers; 128-bit keys: tables of 8 K slots, density 0.48):
    // m = 10 -> 19.6 ms (15 % of the partitions too heavy for shared memory), 12 -> 13.0, 14 -> 12.7
    const double n = (double)expected_kmers;
    int m = 10;                                        // the reference's default (-minimizer-size)
    if (kmer_size < 32) {
        if (n > 150e6) m = 11;
        if (n > 1.5e9) m = 12;
        if (n > 12e9) m = 14;
    } else {
        if (n > 40e6) m = 12;
        if (n > 150e6) m = 14;
        // 128-bit keys: a table takes ~12 K k-mers, and ONE minimizer cannot be split by records.  The hottest 14-letter minimizers
        // hold w / (4^14 / 2) = 3.7e-7 of the k-mers of a random genome (w = k - m + 1 windows): 19.7 K k-mers of BASELINE
        // configs[3] (52.8 G k-mers) -- 37 % of its k-mers sat in minimizers heavier than a table (profiles/r03f: 229 787
        // overflow splits per rank).  One more letter: 4.8 K.
        if (n > 16e9) m = 15;
    }
    if (m > kmer_size - 1) m = kmer_size - 1;
    if (m < 2) m = 2;
    return m;
}

int dskgpu_reset(dskgpu_ctx* ctx)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemsetAsync(ctx->ss.p, 0, sizeof(StreamState), ctx->stream));
    CK(cudaMemsetAsync(ctx->seqst.p, 0, sizeof(SeqStats), ctx->stream));
    CK(cudaMemsetAsync(ctx->ctr.p, 0, sizeof(Counters), ctx->stream));
    CK(cudaMemsetAsync(ctx->hist.p, 0, sizeof(unsigned long long) * DSKGPU_HISTO_LEN, ctx->stream));
    CK(cudaMemsetAsync(ctx->hist2d.p, 0, sizeof(unsigned long long) * DSKGPU_HISTO_LEN * DSKGPU_HISTO2D_DIM2, ctx->stream));
    if (ctx->bank_hist.p) CK(cudaMemsetAsync(ctx->bank_hist.p, 0, sizeof(unsigned long long) * DSKGPU_HISTO_LEN * (size_t)ctx->NB, ctx->stream));
    CK(cudaMemsetAsync(ctx->bin_hist.p, 0, sizeof(unsigned long long) << ctx->fine_log2, ctx->stream));
    CK(cudaMemsetAsync(ctx->hll.p, 0, sizeof(u32) * HLL_M, ctx->stream));
    ctx->sketch_distinct = -1.0;
    ctx->state = 0; ctx->cur_bank = -1; ctx->stream_open = false; ctx->pending_cr = 0;
    ctx->nrec_known = 0; ctx->probes.clear(); ctx->chunk_parity = 0; ctx->push_timed = false;
    ctx->n_solid = 0; ctx->results_on_host = false; ctx->nparts = 0; ctx->planned = false; ctx->nl_me = ctx->np_me = 0; ctx->heavy_recs.clear(); ctx->heavy_kmers.clear(); ctx->rcnt_dev = nullptr;
    ctx->xchg_planned = false; ctx->xchg_scattered = false; ctx->xchg_bytes_out = 0; ctx->totals_done = false; ctx->local_nrec = ctx->local_nkm = 0;
    ctx->bytes_pushed = 0; ctx->sample_queued = false; ctx->sample_nkm = ctx->sample_distinct = 0; ctx->sample_wmult = 0.0;
    ctx->global_set = false; ctx->g_total_kmers = 0; ctx->density = 1.0; ctx->density_known = false; ctx->bin_level = NBINS_LOG2; ctx->hist_fetched = false;
    ctx->peer_recv.clear();
    ctx->ev_used = 0; ctx->spans.clear();
    u64 launches = 0;
    memset(&ctx->st, 0, sizeof ctx->st); ctx->st.gpu_launches = launches;
    return DSKGPU_OK;
}

void dskgpu_destroy(dskgpu_ctx* ctx)
{
    if (!ctx) return;
    cudaStreamSynchronize(ctx->stream);
    DevBuf* all[] = {&ctx->ss, &ctx->ctr, &ctx->hist, &ctx->hist2d, &ctx->bank_hist, &ctx->raw[0], &ctx->raw[1], &ctx->codes, &ctx->tabs, &ctx->tin,
                     &ctx->seqst, &ctx->seqtab, &ctx->recs, &ctx->meta, &ctx->precs, &ctx->cursor, &ctx->bin_hist, &ctx->bin_fold, &ctx->sample_recs, &ctx->stab_keys, &ctx->stab_counts, &ctx->hll, &ctx->bin2part, &ctx->work_ctr,
                     &ctx->tkeys, &ctx->tcounts, &ctx->skeys[0], &ctx->skeys[1], &ctx->svals[0], &ctx->svals[1], &ctx->whash[0], &ctx->whash[1], &ctx->widx[0], &ctx->widx[1], &ctx->keys[0],
                     &ctx->keys[1], &ctx->banks[0], &ctx->banks[1], &ctx->rs_hist, &ctx->rs_status, &ctx->rs_tilectr,
                     &ctx->lrecs, &ctx->xpeers, &ctx->bcur, &ctx->ghist, &ctx->mkeys, &ctx->pl_ex, &ctx->pl_E, &ctx->pl_H, &ctx->pl_bsum, &ctx->pl_pk, &ctx->pl_pr, &ctx->pl_pl,
                     &ctx->pl_newid, &ctx->pl_nvals, &ctx->pl_hdr, &ctx->gk_q, &ctx->gr_q, &ctx->lcnt_q, &ctx->loff, &ctx->xX, &ctx->xtab, &ctx->xS, &ctx->hoff, &ctx->hrecs};
    for (DevBuf* b : all) b->release();
    for (void* q : ctx->ipc_opened) cudaIpcCloseMemHandle(q);
    for (cudaEvent_t e : ctx->evpool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) { if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]); if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]); }
    for (int i = 0; i < 4; i++) if (ctx->ev_probe[i]) cudaEventDestroy(ctx->ev_probe[i]);
    if (ctx->ev_push0) cudaEventDestroy(ctx->ev_push0);
    if (ctx->ev_push1) cudaEventDestroy(ctx->ev_push1);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    if (ctx->h_ss) cudaFreeHost(ctx->h_ss);
    if (ctx->h_nrec_probe) cudaFreeHost(ctx->h_nrec_probe);
    if (ctx->h_hist) cudaFreeHost(ctx->h_hist);
    if (ctx->h_hdr) cudaFreeHost(ctx->h_hdr);
    if (ctx->h_hll) cudaFreeHost(ctx->h_hll);
    if (ctx->h_xtab) cudaFreeHost(ctx->h_xtab);
    if (ctx->h_seqst) cudaFreeHost(ctx->h_seqst);
    if (ctx->h_heavy) cudaFreeHost(ctx->h_heavy);
    if (ctx->h_skeys) cudaFreeHost(ctx->h_skeys);
    if (ctx->h_svals) cudaFreeHost(ctx->h_svals);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// push path
// ---------------------------------------------------------------------------------------------------------------
static int open_stream(dskgpu_ctx* ctx, int bank, int fmt)
{
    if (ctx->stream_open && ctx->cur_bank == bank) return 0;
    if (bank < 0 || bank >= ctx->cfg.nb_banks) FAIL(DSKGPU_ERR_ARG, "bank %d out of range", bank);
    k_scan_reset_stream<<<1, 1, 0, ctx->stream>>>((StreamState*)ctx->ss.p, fmt); LAUNCHED();
    ctx->stream_open = true; ctx->cur_bank = bank; ctx->cur_fmt = fmt; ctx->pending_cr = 0;
    return 0;
}
static void close_stream(dskgpu_ctx* ctx)
{
    // the run of codes still open at the end of a bank is a sequence (seqstats.cuh)
    if (ctx->cfg.sequence_stats && ctx->stream_open) { k_seqstat_close<<<1, 1, 0, ctx->stream>>>(ctx->k, ctx->cur_fmt == FMT_FASTA ? 1 : 0, (SeqStats*)ctx->seqst.p); LAUNCHED(); }
    ctx->stream_open = false; ctx->cur_bank = -1; ctx->pending_cr = 0;
}

// scans bytes [lo, hi) of the 16-byte aligned device buffer `raw` and appends super-k-mer records
static int process_chunk(dskgpu_ctx* ctx, const u8* raw, u64 lo, u64 hi, int next_after)
{
    if (hi <= lo) return 0;
    const int fmt = ctx->cur_fmt;
    const u64 n = hi - lo;
    ctx->bytes_pushed += n;
    const u64 tile_first = lo / SCAN_TILE;
    const u64 ntiles = (hi + SCAN_TILE - 1) / SCAN_TILE - tile_first;
    int rc;
    if ((rc = ensure(ctx, ctx->tabs, ntiles * sizeof(TileTab)))) return rc;
    if ((rc = ensure(ctx, ctx->tin, ntiles * sizeof(TileIn)))) return rc;
    if ((rc = ensure(ctx, ctx->codes, std::max<u64>(n, 2 * ctx->push_chunk) + 128 + SK_TP + 1024, true, 128))) return rc;   // (128 >= k - 1: the carry)
    // record capacity: worst case one record per position of this chunk on top of what is known to be used
    // The host may run two chunks ahead of the device: a chunk still in flight is charged at its worst case (one record per
    // byte) until its probe (the record counter copied out behind its kernels) has landed.
    auto harvest = [&](bool wait) -> int {
        while (!ctx->probes.empty()) {
            const int sl = ctx->probes.front().slot;
            if (wait) CK(cudaEventSynchronize(ctx->ev_probe[sl]));
            else { cudaError_t q = cudaEventQuery(ctx->ev_probe[sl]); if (q == cudaErrorNotReady) { (void)cudaGetLastError(); break; } CK(q); }
            ctx->nrec_known = ctx->h_nrec_probe[sl];
            ctx->probes.erase(ctx->probes.begin());
            wait = false;                                                 // one blocking wait per call, then whatever else is ready
        }
        return 0;
    };
    if ((rc = harvest(false))) return rc;
    while (ctx->probes.size() >= 2) if ((rc = harvest(true))) return rc;
    auto upper = [&]() { u64 u = ctx->nrec_known + n + 64; for (auto& pr : ctx->probes) u += pr.bytes; return u; };
    while (upper() > ctx->rec_cap && !ctx->probes.empty()) if ((rc = harvest(true))) return rc;
    const u64 need = upper();
    if (need > ctx->rec_cap) {
        u64 ncap = std::max(need, ctx->rec_cap + ctx->rec_cap / 2);
        if ((rc = ensure(ctx, ctx->recs, ncap * ctx->RW * 8, true, ctx->nrec_known * ctx->RW * 8))) return rc;
        if ((rc = ensure(ctx, ctx->meta, ncap * 4, true, ctx->nrec_known * 4))) return rc;
        ctx->rec_cap = std::min<u64>(ctx->recs.cap / (ctx->RW * 8), ctx->meta.cap / 4);
    }
    StreamState* ss = (StreamState*)ctx->ss.p;
    if (!ctx->push_timed) { CK(cudaEventRecord(ctx->ev_push0, ctx->stream)); ctx->push_timed = true; }
    {
        SpanGuard g(ctx, SPAN_PARSE);
        const unsigned gt = (unsigned)ntiles;
        switch (fmt) {
        case FMT_FASTA: k_scan_tables<FMT_FASTA><<<gt, SCAN_THREADS, 0, ctx->stream>>>(raw, lo, hi, tile_first, ss, next_after, (TileTab*)ctx->tabs.p); break;
        case FMT_FASTQ: k_scan_tables<FMT_FASTQ><<<gt, SCAN_THREADS, 0, ctx->stream>>>(raw, lo, hi, tile_first, ss, next_after, (TileTab*)ctx->tabs.p); break;
        default:        k_scan_tables<FMT_LINES><<<gt, SCAN_THREADS, 0, ctx->stream>>>(raw, lo, hi, tile_first, ss, next_after, (TileTab*)ctx->tabs.p); break;
        }
        LAUNCHED();
        k_scan_tiles<<<1, 1024, 0, ctx->stream>>>((const TileTab*)ctx->tabs.p, ntiles, (TileIn*)ctx->tin.p, ss, raw, lo, hi); LAUNCHED();
        switch (fmt) {
        case FMT_FASTA: k_scan_emit<FMT_FASTA><<<gt, SCAN_THREADS, 0, ctx->stream>>>(raw, lo, hi, tile_first, ss, ss, next_after, (const TileIn*)ctx->tin.p, (u8*)ctx->codes.p); break;
        case FMT_FASTQ: k_scan_emit<FMT_FASTQ><<<gt, SCAN_THREADS, 0, ctx->stream>>>(raw, lo, hi, tile_first, ss, ss, next_after, (const TileIn*)ctx->tin.p, (u8*)ctx->codes.p); break;
        default:        k_scan_emit<FMT_LINES><<<gt, SCAN_THREADS, 0, ctx->stream>>>(raw, lo, hi, tile_first, ss, ss, next_after, (const TileIn*)ctx->tin.p, (u8*)ctx->codes.p); break;
        }
        LAUNCHED();
    }
    if (ctx->cfg.sequence_stats) {
        // sequence lengths from the separators of the codes this chunk added ([carry, total): both still on the device)
        const u64 ntl = (n + 15 + SQ_TILE - 1) / SQ_TILE + 1;
        if ((rc = ensure(ctx, ctx->seqtab, ntl * 16))) return rc;
        k_seqstat_tiles<<<(unsigned)ntl, SQ_THREADS, 0, ctx->stream>>>((const u8*)ctx->codes.p, ss, ctx->k, (unsigned long long*)ctx->seqtab.p, (SeqStats*)ctx->seqst.p); LAUNCHED();
        k_seqstat_stitch<<<1, 1024, 0, ctx->stream>>>(ss, ctx->k, fmt == FMT_FASTA ? 1 : 0, ntl, (const unsigned long long*)ctx->seqtab.p, (SeqStats*)ctx->seqst.p); LAUNCHED();
    }
    {
        SpanGuard g(ctx, SPAN_SUPERK);
        const unsigned gk = (unsigned)((n + (ctx->KW <= 2 ? 64 : 128) + SK_TP - 1) / SK_TP);
        const int bank = ctx->NB > 1 ? ctx->cur_bank : 0;
#define SUPERK_LAUNCH(W_) k_superkmers<W_><<<gk, SK_THREADS, 0, ctx->stream>>>((const u8*)ctx->codes.p, ss, ctx->k, ctx->m, bank, (u64*)ctx->recs.p, (u32*)ctx->meta.p, ctx->rec_cap, (Counters*)ctx->ctr.p, (unsigned long long*)ctx->bin_hist.p, (u32)ctx->nb_passes, (u32)ctx->pass_id, META_BIN_BITS - ctx->fine_log2)
        if (ctx->KW == 1) SUPERK_LAUNCH(1); else if (ctx->KW == 2) SUPERK_LAUNCH(2); else if (ctx->KW == 3) SUPERK_LAUNCH(3); else SUPERK_LAUNCH(4);
#undef SUPERK_LAUNCH
        LAUNCHED();
        k_scan_carry<<<1, 128, 0, ctx->stream>>>((u8*)ctx->codes.p, ss, ctx->k); LAUNCHED();
    }
    {
        const int sl = ctx->probe_slot; ctx->probe_slot = (sl + 1) & 3;
        CK(cudaMemcpyAsync(ctx->h_nrec_probe + sl, &((Counters*)ctx->ctr.p)->nrec, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->ev_probe[sl], ctx->stream));
        ctx->probes.push_back({sl, n});
    }
    CK(cudaEventRecord(ctx->ev_push1, ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

static int detect_format(const char* b, size_t n, size_t* skip)
{
    // BankFasta.cpp:492-498 : everything before the first '>' or '@' is skipped
    for (size_t i = 0; i < n; i++) {
        if (b[i] == '>') { *skip = i; return FMT_FASTA; }
        if (b[i] == '@') { *skip = i; return FMT_FASTQ; }
    }
    *skip = n; return 0;
}

extern "C" {

int dskgpu_push_bytes(dskgpu_ctx* ctx, int bank_id, const char* bytes, size_t n, int format, int flags)
{
    if (!ctx || (!bytes && n)) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (ctx->state != 0) FAIL(DSKGPU_ERR_STATE, "push after finish");
    const bool last = flags & DSKGPU_PUSH_LAST;
    size_t skip = 0;
    if (!ctx->stream_open || ctx->cur_bank != bank_id) {
        int fmt = format;
        if (fmt == DSKGPU_FMT_AUTO || fmt == DSKGPU_FMT_FASTA || fmt == DSKGPU_FMT_FASTQ) {
            size_t s = 0; int det = detect_format(bytes, n, &s);
            if (det == 0) { if (last) return DSKGPU_OK; FAIL(DSKGPU_ERR_FORMAT, "no FASTA/FASTQ header in the first chunk of bank %d", bank_id); }
            if (fmt == DSKGPU_FMT_AUTO) fmt = det;
            skip = s;
        }
        int rc = open_stream(ctx, bank_id, fmt); if (rc) return rc;
    }
    size_t pos = skip;
    const size_t CH = ctx->push_chunk;
    for (;;) {
        const size_t remaining = n - pos;
        if (remaining == 0 && !(last && ctx->pending_cr)) break;
        const size_t len = std::min(CH, remaining);
        const bool final_piece = (pos + len == n);
        // hold back a trailing CR unless the stream ends here: its fate depends on the next byte (BankFasta.cpp:471)
        const int hold = (final_piece && !last && len > 0 && bytes[pos + len - 1] == '\r') ? 1 : 0;
        const size_t eff = len - hold;
        if (eff == 0 && !ctx->pending_cr) { ctx->pending_cr = hold; pos += len; continue; }
        const int par = ctx->chunk_parity; ctx->chunk_parity ^= 1;
        int rc = ensure(ctx, ctx->raw[par], CH + 64); if (rc) return rc;
        u8* d = (u8*)ctx->raw[par].p;
        // raw[par] was last read by the kernels of two chunks ago; order the copy after them
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[par], 0));
        u64 lo = 16;
        if (ctx->pending_cr) { static const char cr = '\r'; lo = 15; CK(cudaMemcpyAsync(d + 15, &cr, 1, cudaMemcpyHostToDevice, ctx->copy_stream)); ctx->pending_cr = 0; }
        if (eff) CK(cudaMemcpyAsync(d + 16, bytes + pos, eff, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_copy[par], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[par], 0));
        int next_after;
        if (!final_piece) next_after = (int)(unsigned char)bytes[pos + len];
        else if (hold) next_after = '\r';
        else next_after = last ? -1 : '\n';
        rc = process_chunk(ctx, d, lo, 16 + eff, next_after);
        if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_done[par], ctx->stream));
        ctx->pending_cr = hold;
        pos += len;
    }
    if (last) close_stream(ctx);
    return DSKGPU_OK;
}

int dskgpu_push_device_bytes(dskgpu_ctx* ctx, int bank_id, const void* dev_bytes, size_t n, int format, int flags)
{
    if (!ctx || (!dev_bytes && n)) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (ctx->state != 0) FAIL(DSKGPU_ERR_STATE, "push after finish");
    const bool last = flags & DSKGPU_PUSH_LAST;
    const u8* p = (const u8*)dev_bytes;
    size_t skip = 0;
    if (!ctx->stream_open || ctx->cur_bank != bank_id) {
        int fmt = format;
        if (fmt != DSKGPU_FMT_LINES) {
            // sniff the head of the stream on the host (one small D2H copy per bank)
            size_t hn = std::min<size_t>(n, 1 << 16);
            std::vector<char> head(hn);
            CK(cudaMemcpyAsync(head.data(), p, hn, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            size_t s = 0; int det = detect_format(head.data(), hn, &s);
            if (det == 0) { if (last && hn == n) return DSKGPU_OK; FAIL(DSKGPU_ERR_FORMAT, "no FASTA/FASTQ header in the first 64 KiB of bank %d", bank_id); }
            if (fmt == DSKGPU_FMT_AUTO) fmt = det;
            skip = s;
        }
        int rc = open_stream(ctx, bank_id, fmt); if (rc) return rc;
    }
    // process in place, in pieces that bound the worst-case record reservation
    const uintptr_t addr = (uintptr_t)p;
    const u8* base = (const u8*)(addr & ~(uintptr_t)15);
    const u64 off0 = (u64)(addr - (uintptr_t)base);
    const size_t CH = ctx->push_chunk * 2;
    size_t pos = skip;
    while (pos < n) {
        size_t len = std::min(CH, n - pos);
        // keep piece boundaries 16-byte aligned in absolute terms so vector loads stay aligned
        if (pos + len < n) { u64 endabs = off0 + pos + len; endabs &= ~(u64)(SCAN_TILE - 1); if (endabs > off0 + pos) len = (size_t)(endabs - off0 - pos); }
        const bool final_piece = (pos + len == n);
        // next byte: read on the device side would need a copy; non-final pieces pass -2 => kernels read raw[hi]
        int next_after = final_piece ? (last ? -1 : '\n') : -2;
        int rc = process_chunk(ctx, base, off0 + pos, off0 + pos + len, next_after); if (rc) return rc;
        pos += len;
    }
    if (last) close_stream(ctx);
    return DSKGPU_OK;
}

int dskgpu_push_reads(dskgpu_ctx* ctx, int bank_id, const char* bases, const uint64_t* offsets, size_t nreads)
{
    if (!ctx || !bases || !offsets) return DSKGPU_ERR_ARG;
    use_device(ctx);
    // one sequence per line: the separator is inserted on the host while staging
    size_t total = (size_t)(offsets[nreads] - offsets[0]) + nreads;
    char* tmp = (char*)dskgpu_host_alloc(total ? total : 1);
    if (!tmp) FAIL(DSKGPU_ERR_NOMEM, "pinned staging alloc failed");
    size_t w = 0;
    for (size_t i = 0; i < nreads; i++) {
        size_t len = (size_t)(offsets[i + 1] - offsets[i]);
        memcpy(tmp + w, bases + offsets[i], len); w += len; tmp[w++] = '\n';
    }
    int rc = dskgpu_push_bytes(ctx, bank_id, tmp, w, DSKGPU_FMT_LINES, DSKGPU_PUSH_LAST);
    if (rc == 0) { cudaStreamSynchronize(ctx->copy_stream); }
    dskgpu_host_free(tmp);
    return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// finish path
// ---------------------------------------------------------------------------------------------------------------
template <int KW, bool HAS_VAL>
static int radix_sort(dskgpu_ctx* ctx, u64* keys[2], u32* vals[2], u64 n, int npass, int* result_buf, int first_pass = 0, int start_buf = 0)
{
    *result_buf = start_buf;
    if (n == 0) return 0;
    if (n >= ((u64)1 << 30)) FAIL(DSKGPU_ERR_OVERFLOW, "radix sort of %llu keys exceeds the 2^30 limit of the tile status words", (unsigned long long)n);
    constexpr int TILE = RsCfg<KW>::TILE;
    const u64 ntiles = (n + TILE - 1) / TILE;
    int rc;
    if ((rc = ensure(ctx, ctx->rs_hist, (size_t)npass * 256 * 8))) return rc;
    if ((rc = ensure(ctx, ctx->rs_status, ntiles * 256 * 4))) return rc;
    if ((rc = ensure(ctx, ctx->rs_tilectr, 64 * 4))) return rc;
    CK(cudaMemsetAsync(ctx->rs_hist.p, 0, (size_t)npass * 256 * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->rs_tilectr.p, 0, 64 * 4, ctx->stream));
    const unsigned hb = (unsigned)std::min<u64>((n + RS_THREADS * 8 - 1) / (RS_THREADS * 8), 148 * 8);
    k_rs_hist<KW><<<hb, RS_THREADS, npass * 256 * 4, ctx->stream>>>(keys[start_buf], n, npass, (unsigned long long*)ctx->rs_hist.p, first_pass); LAUNCHED();
    k_rs_scan<<<npass, 256, 0, ctx->stream>>>((unsigned long long*)ctx->rs_hist.p); LAUNCHED();
    const int smem = TILE * KW * 8 + (HAS_VAL ? TILE * 4 : 16);
    int cur = start_buf;
    for (int p = first_pass; p < npass; p++) {
        CK(cudaMemsetAsync(ctx->rs_status.p, 0, ntiles * 256 * 4, ctx->stream));
        cudaEvent_t a = get_event(ctx), b = get_event(ctx);
        cudaEventRecord(a, ctx->stream);
        k_rs_onesweep<KW, HAS_VAL><<<(unsigned)ntiles, RS_THREADS, smem, ctx->stream>>>(
            keys[cur], keys[cur ^ 1], HAS_VAL ? vals[cur] : nullptr, HAS_VAL ? vals[cur ^ 1] : nullptr, n, p,
            (const unsigned long long*)ctx->rs_hist.p + (size_t)p * 256, (u32*)ctx->rs_status.p, (u32*)ctx->rs_tilectr.p + p);
        LAUNCHED();
        cudaEventRecord(b, ctx->stream);
        ctx->spans.push_back({a, b, HAS_VAL ? SPAN_SORTPASS : SPAN_DOM});
        cur ^= 1;
    }
    *result_buf = cur;
    CK(cudaGetLastError());
    return 0;
}

// ascending order of the solid set (as the reference emits inside a partition).  Top ceil(log2 n) bits by one-sweep LSD
// passes, the rest by the neighbourhood fix-up (radix.cuh); `full` (or keys too short to gain anything) = plain LSD sort.
// ctx->sort_src = the buffer the fix-up read (intact): where a flagged fix-up restarts from.
template <int KW>
static int sort_solid(dskgpu_ctx* ctx, bool full, int start_buf)
{
    u64* kk[2] = {(u64*)ctx->skeys[0].p, (u64*)ctx->skeys[1].p};
    u32* vv[2] = {(u32*)ctx->svals[0].p, (u32*)ctx->svals[1].p};
    const u64 n = ctx->n_solid;
    const int npass_full = (2 * ctx->k + 7) / 8;
    int lg = 1; while (((u64)1 << lg) < n) lg++;
    const int msd = (lg + 7) / 8;
    ctx->sort_fixup = false;
    if (KW > 2 || full || getenv("DSKGPU_SORT_FULL") || msd + 1 >= npass_full)
        return radix_sort<KW, true>(ctx, kk, vv, n, npass_full, &ctx->solid_buf, 0, start_buf);
    if constexpr (KW <= 2) {
    int r = start_buf, rc;
    if ((rc = radix_sort<KW, true>(ctx, kk, vv, n, npass_full, &r, npass_full - msd, start_buf))) return rc;
    Counters* ctr = (Counters*)ctx->ctr.p;
    cudaEvent_t a = get_event(ctx), b = get_event(ctx);
    cudaEventRecord(a, ctx->stream);
    k_rs_fix<KW><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(kk[r], vv[r], kk[r ^ 1], vv[r ^ 1], n, 8 * (npass_full - msd), &ctr->sort_fallback); LAUNCHED();
    cudaEventRecord(b, ctx->stream);
    ctx->spans.push_back({a, b, SPAN_SORTPASS});
    ctx->sort_src = r; ctx->solid_buf = r ^ 1; ctx->sort_fixup = true;
    CK(cudaMemcpyAsync(ctx->h_nrec_probe + 7, &ctr->sort_fallback, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaGetLastError());
    }
    return 0;
}

static SolidityParams make_sp(dskgpu_ctx* ctx)
{
    SolidityParams sp; memset(&sp, 0, sizeof sp);
    sp.kind = ctx->cfg.solidity_kind; sp.nbanks = ctx->NB; sp.histo2d = ctx->cfg.histo2d;
    for (int i = 0; i < MAXB; i++) { sp.amin[i] = ctx->cfg.abundance_min[i]; sp.solid_vec[i] = ctx->cfg.solid_vec[i]; }
    sp.amax = ctx->cfg.abundance_max;
    sp.bank_hist = (ctx->NB > 1 && ctx->bank_hist.p) ? (unsigned long long*)ctx->bank_hist.p : nullptr;
    if (sp.nbanks == 1) sp.kind = DSKGPU_SOLIDITY_SUM;          // ConfigurationAlgorithm.cpp:261-264
    return sp;
}

// count records [rb, re) (nk k-mers in total) through the sort path
template <int KW>
static int count_by_sort(dskgpu_ctx* ctx, const u64* recs, u64 rb, u64 re, u64 nk, u64 out_cap)
{
    if (nk == 0) return 0;
    int rc;
    for (int i = 0; i < 2; i++) {
        if ((rc = ensure(ctx, ctx->keys[i], nk * KW * 8 + 64))) return rc;
        if (ctx->NB > 1 && (rc = ensure(ctx, ctx->banks[i], nk * 4 + 64))) return rc;
    }
    Counters* ctr = (Counters*)ctx->ctr.p;
    CK(cudaMemsetAsync(&ctr->expand_cursor, 0, 8, ctx->stream));
    const unsigned gb = (unsigned)std::min<u64>((re - rb + 255) / 256, 148 * 16);
    k_expand_keys<KW><<<gb, 256, 0, ctx->stream>>>(recs, rb, re, ctx->k, (u64*)ctx->keys[0].p, (u32*)ctx->banks[0].p, ctx->NB, ctr); LAUNCHED();
    if constexpr (KW > 2) {
        // wide keys: equal k-mers are brought together by a 64-bit hash (8 passes over 12 bytes per key instead of 24 - 32 passes
        // over 24 - 32), the full keys are only compared (count.cuh).  DSKGPU_WIDE_FULLSORT=1 keeps the full-width sort;
        // DSKGPU_TEST_HASH_BITS narrows the hash so that the tests see collisions (-> verified, redone by the full sort)
        if (!getenv("DSKGPU_WIDE_FULLSORT") && nk < ((u64)1 << 31)) {
            u64 hmask = ~0ULL;
            if (const char* e = getenv("DSKGPU_TEST_HASH_BITS")) { const int b = std::min(64, std::max(1, atoi(e))); hmask = b >= 64 ? ~0ULL : (((u64)1 << b) - 1); }
            for (int i = 0; i < 2; i++) {
                if ((rc = ensure(ctx, ctx->whash[i], nk * 8 + 64))) return rc;
                if ((rc = ensure(ctx, ctx->widx[i], nk * 4 + 64))) return rc;
            }
            const unsigned gh = (unsigned)std::min<u64>((nk + 255) / 256, (u64)ctx->num_sms * 16);
            k_hash_keys<KW><<<gh, 256, 0, ctx->stream>>>((const u64*)ctx->keys[0].p, nk, (u64*)ctx->whash[0].p, (u32*)ctx->widx[0].p, hmask); LAUNCHED();
            u64* hk[2] = {(u64*)ctx->whash[0].p, (u64*)ctx->whash[1].p};
            u32* hv[2] = {(u32*)ctx->widx[0].p, (u32*)ctx->widx[1].p};
            int hres = 0;
            if ((rc = radix_sort<1, true>(ctx, hk, hv, nk, 8, &hres))) return rc;
            CK(cudaMemsetAsync(&ctr->sort_fallback, 0, sizeof(unsigned int), ctx->stream));
            k_verify_hashed<KW><<<gh, 256, 0, ctx->stream>>>(hk[hres], hv[hres], (const u64*)ctx->keys[0].p, nk, &ctr->sort_fallback); LAUNCHED();
            CK(cudaMemcpyAsync(ctx->h_nrec_probe + 6, &ctr->sort_fallback, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            const bool collided = *(volatile unsigned int*)(ctx->h_nrec_probe + 6) != 0;
            CK(cudaMemsetAsync(&ctr->sort_fallback, 0, sizeof(unsigned int), ctx->stream));
            if (!collided) {
                const unsigned gr = (unsigned)std::min<u64>((nk + 255) / 256, 148 * 16);
                k_rle_emit_hashed<KW><<<gr, 256, 0, ctx->stream>>>(hk[hres], hv[hres], (const u64*)ctx->keys[0].p, (const u32*)ctx->banks[0].p, nk, make_sp(ctx),
                                                                   (u64*)ctx->skeys[0].p, (u32*)ctx->svals[0].p, out_cap,
                                                                   (unsigned long long*)ctx->hist.p, (unsigned long long*)ctx->hist2d.p, ctr); LAUNCHED();
                ctx->st.nb_groups_sort++;
                CK(cudaGetLastError());
                return 0;
            }
            ctx->st.sort_fallbacks++;                               // two k-mers shared a hash: the exact full-width sort (keys[0] is intact)
        }
    }
    u64* kk[2] = {(u64*)ctx->keys[0].p, (u64*)ctx->keys[1].p};
    u32* vv[2] = {(u32*)ctx->banks[0].p, (u32*)ctx->banks[1].p};
    const int npass = (2 * ctx->k + 7) / 8;
    int res = 0;
    if (ctx->NB > 1) rc = radix_sort<KW, true>(ctx, kk, vv, nk, npass, &res);
    else rc = radix_sort<KW, false>(ctx, kk, vv, nk, npass, &res);
    if (rc) return rc;
    const unsigned gr = (unsigned)std::min<u64>((nk + 255) / 256, 148 * 16);
    k_rle_emit<KW><<<gr, 256, 0, ctx->stream>>>(kk[res], vv[res], nk, make_sp(ctx), (u64*)ctx->skeys[0].p, (u32*)ctx->svals[0].p, out_cap,
                                                (unsigned long long*)ctx->hist.p, (unsigned long long*)ctx->hist2d.p, ctr); LAUNCHED();
    ctx->st.nb_groups_sort++;
    CK(cudaGetLastError());
    return 0;
}

template <int KW>
static int count_all(dskgpu_ctx* ctx, const u64* recs, const std::vector<u64>& prec, const std::vector<u64>& pkm, u64 out_cap)
{
    // prec/pkm: records / k-mers of each partition, stored contiguously in `recs` in this order
    const size_t np = prec.size();
    std::vector<u64> off(np + 1, 0);
    for (size_t i = 0; i < np; i++) off[i + 1] = off[i] + prec[i];
    Counters* ctr = (Counters*)ctx->ctr.p;
    const SolidityParams sp = make_sp(ctx);
    int rc;
    const int mode = ctx->cfg.count_mode;
    const u64 sort_cap = (u64)1 << 28;                             // keys per sort-path group
    if (mode == DSKGPU_COUNT_SORT || KW > 2) {                     // (wide spans: always -- dskgpu_create forces the mode)
        size_t p = 0;
        while (p < np) {
            size_t q = p; u64 km = 0;
            while (q < np && (q == p || km + pkm[q] <= sort_cap)) { km += pkm[q]; q++; }
            if ((rc = count_by_sort<KW>(ctx, recs, off[p], off[q], km, out_cap))) return rc;
            p = q;
        }
        return 0;
    }
    if constexpr (KW <= 2) {
    const int log2s = ctx->cfg.hash_log2_slots > 0 ? std::max(10, ctx->cfg.hash_log2_slots) : 23;
    u64 nslots = (u64)1 << log2s;
    const double load_max = 0.6;

    auto init_table = [&](u64 slots) -> int {
        int rc2;
        if ((rc2 = ensure(ctx, ctx->tkeys, slots * KW * 8))) return rc2;
        if ((rc2 = ensure(ctx, ctx->tcounts, slots * 4 * (u64)ctx->NB))) return rc2;
        k_fill_u64<<<148 * 4, 256, 0, ctx->stream>>>((u64*)ctx->tkeys.p, slots * KW, ~0ULL); LAUNCHED();
        CK(cudaMemsetAsync(ctx->tcounts.p, 0, slots * 4 * (u64)ctx->NB, ctx->stream));
        return 0;
    };
    // one group of consecutive partitions [pb, pe) through the table.  `checked`: the group was sized from an ESTIMATE of
    // distinct / total, so the insert is followed by a look at the overflow flag before the sweep; *overflowed tells the
    // caller to redo the group at the worst-case ratio (the table has been cleared again)
    auto run_hash = [&](size_t pb, size_t pe, u64 slots, bool checked, bool* overflowed) -> int {
        const u64 rb = off[pb], re = off[pe];
        const unsigned gi_blocks = (unsigned)std::min<u64>((re - rb + 255) / 256, 148 * 8);
        cudaEvent_t a = get_event(ctx), b = get_event(ctx);
        cudaEventRecord(a, ctx->stream);
        k_hash_insert<KW><<<gi_blocks ? gi_blocks : 1, 256, 0, ctx->stream>>>(recs, rb, re, ctx->k, (u64*)ctx->tkeys.p, (u32*)ctx->tcounts.p,
                                                                              (u32)(slots - 1), ctx->NB, ctr); LAUNCHED();
        cudaEventRecord(b, ctx->stream);
        ctx->spans.push_back({a, b, SPAN_DOM});
        if (overflowed) *overflowed = false;
        if (checked) {
            CK(cudaMemcpyAsync(ctx->h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->h_ctr->hash_overflow) {
                CK(cudaMemsetAsync(&ctr->hash_overflow, 0, sizeof(unsigned int), ctx->stream));
                int rc2 = init_table(slots); if (rc2) return rc2;
                *overflowed = true;
                return 0;
            }
        }
        const unsigned gs = (unsigned)std::min<u64>((slots / 8 + 255) / 256, 148 * 8);
        if (ctx->NB == 1)
            k_hash_scan<KW, true><<<gs, 256, 0, ctx->stream>>>((u64*)ctx->tkeys.p, (u32*)ctx->tcounts.p, (u32)slots, sp, 0, (u64*)ctx->skeys[0].p,
                                                               (u32*)ctx->svals[0].p, out_cap, (unsigned long long*)ctx->hist.p,
                                                               (unsigned long long*)ctx->hist2d.p, ctr);
        else
            k_hash_scan<KW, false><<<gs, 256, 0, ctx->stream>>>((u64*)ctx->tkeys.p, (u32*)ctx->tcounts.p, (u32)slots, sp, 0, (u64*)ctx->skeys[0].p,
                                                                (u32*)ctx->svals[0].p, out_cap, (unsigned long long*)ctx->hist.p,
                                                                (unsigned long long*)ctx->hist2d.p, ctr);
        LAUNCHED();
        ctx->st.nb_groups_hash++;
        return 0;
    };

    // hash (forced) or auto.  The table is sized once; groups of consecutive partitions are sized so that the
    // estimated number of distinct k-mers stays under load_max * nslots.  The distinct/total ratio r comes from the
    // density sample (or is measured on a first group sized for the worst case r = 1).  It is an estimate: a group whose
    // k-mers are less repetitive than the job's average (hot low-complexity bins) can outgrow the table -- that group is
    // then redone in sub-groups sized for r = 1, which cannot overflow (never an error, as in the reference).
    u64 max_part = 0; for (size_t i = 0; i < np; i++) max_part = std::max(max_part, pkm[i]);
    if (mode == DSKGPU_COUNT_HASH) { while ((double)nslots * load_max < (double)max_part && nslots < ((u64)1 << 31)) nslots <<= 1; }
    if (ctx->cfg.hash_log2_slots <= 0) {
        // a call that covers little (a few heavy partitions next to the shared-memory path) gets a table to match:
        // initialising and sweeping 2^23 slots would cost more than the counting itself
        u64 tot = 0; for (size_t i = 0; i < np; i++) tot += pkm[i];
        while (nslots > ((u64)1 << 16) && (double)(nslots >> 1) * load_max >= (double)tot) nslots >>= 1;
    }
    if ((rc = init_table(nslots))) return rc;
    double r = 1.0; bool have_r = false;
    if (ctx->density_known) { r = std::min(1.0, std::max(0.02, ctx->density * 1.3 + 0.01)); have_r = true; }
    if (const char* e = getenv("DSKGPU_TEST_HASH_RATIO")) { r = std::min(1.0, std::max(0.001, atof(e))); have_r = true; }   // test hook: a wrong estimate
    // partitions [pb, pe) in groups sized for the ratio rr; a single partition beyond the table goes to the sort path
    std::function<int(size_t, size_t, double)> run_range = [&](size_t pb, size_t pe, double rr) -> int {
        size_t p = pb;
        while (p < pe) {
            const double capk = (double)nslots * load_max / rr;
            size_t q = p; u64 km = 0;
            while (q < pe && km + pkm[q] <= (u64)capk) { km += pkm[q]; q++; }
            if (q == p) {                                                  // occupancy picks the sort path
                int rc2 = count_by_sort<KW>(ctx, recs, off[p], off[p + 1], pkm[p], out_cap); if (rc2) return rc2;
                p++; continue;
            }
            if (km == 0) { p = q; continue; }
            const bool calibrate = !have_r;
            u64 d0 = 0;
            if (calibrate) { CK(cudaMemcpyAsync(ctx->h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); d0 = ctx->h_ctr->distinct_n; }
            bool ovf = false;
            int rc2 = run_hash(p, q, nslots, rr < 1.0, &ovf); if (rc2) return rc2;
            if (ovf) { ctx->st.nb_hash_regroups++; rc2 = run_range(p, q, 1.0); if (rc2) return rc2; p = q; continue; }
            if (calibrate) {
                CK(cudaMemcpyAsync(ctx->h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream));
                if (ctx->h_ctr->hash_overflow) FAIL(DSKGPU_ERR_OVERFLOW, "hash table overflow in a group sized for distinct = total (internal)");
                const double dr = (double)(ctx->h_ctr->distinct_n - d0) / (double)km;
                r = std::min(1.0, std::max(0.02, dr * 1.3 + 0.01)); have_r = true;
                rr = r;
            }
            p = q;
        }
        return 0;
    };
    if ((rc = run_range(0, np, r))) return rc;
    CK(cudaGetLastError());
    }
    return 0;
}

static float span_ms(dskgpu_ctx* ctx, int kind, u32* count = nullptr)
{
    float tot = 0; u32 c = 0;
    for (auto& s : ctx->spans) if (s.kind == kind) { float ms = 0; if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) tot += ms; c++; }
    if (count) *count = c;
    return tot;
}

// ---- density sample: distinct / total k-mers of the job, estimated on whole bins ---------------------------------------
// The records of the fine bins below a threshold (about SAMPLE_KMERS k-mers) are copied out, counted in a small hash
// table and only the number of occupied slots is kept.  Every occurrence of a k-mer lies in the same bin, so the ratio is
// unbiased for the job; it sizes the partitions (shared-memory table load) and the groups of the global hash path, the
// job the reference gives to its sampling pass (K/RepartitionAlgorithm.cpp:395-492, K/ConfigurationAlgorithm.cpp:245-467).
constexpr u64 SAMPLE_KMERS = 512 * 1024;
constexpr u64 SAMPLE_KM_CAP = 1u << 20;
constexpr u32 SAMPLE_SLOTS = 1u << 21;

template <int KW>
static int queue_sample(dskgpu_ctx* ctx)
{
    if (ctx->sample_queued) return 0;
    ctx->sample_queued = true;
    if (ctx->recs.p == nullptr || ctx->bytes_pushed == 0) return 0;
    int rc;
    if ((rc = ensure(ctx, ctx->sample_recs, SAMPLE_KM_CAP * (u64)ctx->RW * 8))) return rc;
    if (ctx->stab_keys.p == nullptr) {
        if ((rc = ensure(ctx, ctx->stab_keys, (size_t)SAMPLE_SLOTS * KW * 8))) return rc;
        if ((rc = ensure(ctx, ctx->stab_counts, (size_t)SAMPLE_SLOTS * 4))) return rc;
        k_fill_u64<<<148 * 4, 256, 0, ctx->stream>>>((u64*)ctx->stab_keys.p, (u64)SAMPLE_SLOTS * KW, ~0ULL); LAUNCHED();
        CK(cudaMemsetAsync(ctx->stab_counts.p, 0, (size_t)SAMPLE_SLOTS * 4, ctx->stream));
    }
    Counters* ctr = (Counters*)ctx->ctr.p;
    SpanGuard g(ctx, SPAN_PLAN);
    // the exact k-mer total is still on the device: size the sample from the bytes pushed (~0.7 k-mers per FASTA byte)
    const double est = std::max(1.0, (double)ctx->bytes_pushed * 0.7);
    double f = (double)SAMPLE_KMERS / est;
    u32 thresh = f >= 1.0 ? (1u << META_BIN_BITS) : (u32)std::max(1.0, f * (double)(1u << META_BIN_BITS) + 0.5);
    k_sample_select<KW><<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((const u64*)ctx->recs.p, (const u32*)ctx->meta.p, &ctr->nrec, thresh,
                                                                    (u64*)ctx->sample_recs.p, SAMPLE_KM_CAP, ctr); LAUNCHED();
    k_hash_insert<KW><<<ctx->num_sms * 2, 256, 0, ctx->stream>>>((const u64*)ctx->sample_recs.p, 0, SAMPLE_KM_CAP, ctx->k, (u64*)ctx->stab_keys.p,
                                                                  (u32*)ctx->stab_counts.p, SAMPLE_SLOTS - 1, 1, ctr, &ctr->sample_nrec); LAUNCHED();
    SolidityParams sp; memset(&sp, 0, sizeof sp); sp.nbanks = 1;
    // share of the sampled distinct k-mers whose summed count reaches the smallest threshold: sizes the solid-set buffers
    sp.amin[0] = ctx->cfg.abundance_min[0]; sp.amax = ctx->cfg.abundance_max;
    for (int b = 1; b < ctx->cfg.nb_banks; b++) sp.amin[0] = std::min<long long>(sp.amin[0], ctx->cfg.abundance_min[b]);
    if (ctx->cfg.solidity_kind == DSKGPU_SOLIDITY_CUSTOM) sp.amin[0] = 1;
    if (ctx->NB > 1) sp.amax = 0x7FFFFFFFFFFFFFFFLL;                    // per-bank ranges: the summed count only bounds from below
    k_hash_scan<KW, true><<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((u64*)ctx->stab_keys.p, (u32*)ctx->stab_counts.p, SAMPLE_SLOTS, sp, 2, nullptr, nullptr, 0,
                                                                      (unsigned long long*)ctx->hist.p, (unsigned long long*)ctx->hist2d.p, ctr, (u32*)ctx->hll.p); LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

static bool use_smem_path(const dskgpu_ctx* ctx);
static u64 plan_target_kmers(const dskgpu_ctx* ctx, u64 global_kmers);
constexpr int SMEM_MAX_SPLIT0_DEFAULT = 4;

// whole-job figures every rank plans from: k-mer total and density sample.  Picks the bin level.
static void set_global(dskgpu_ctx* ctx, u64 g_kmers, u64 g_recs, u64 g_sample_kmers, u64 g_sample_distinct)
{
    ctx->g_total_kmers = g_kmers; ctx->g_total_recs = g_recs;
    ctx->density_known = g_sample_kmers >= 4096;
    // several ranks: the distinct k-mers of the UNION of the ranks' samples come from the merged HyperLogLog sketch (a k-mer
    // sampled by two ranks is one distinct k-mer of the job; summing the ranks' own distinct counts would overstate the density
    // more and more as the ranks' shares of the coverage shrink)
    const double g_distinct = ctx->sketch_distinct >= 0.0 ? std::min(ctx->sketch_distinct, (double)g_sample_distinct) : (double)g_sample_distinct;
    ctx->density = ctx->density_known ? std::min(1.0, std::max(0.01, g_distinct / (double)g_sample_kmers)) : 1.0;
    const u64 T = plan_target_kmers(ctx, g_kmers);
    // bins of the chosen level should average an eighth of a partition: a partition overshoots the cut by part of its last bin
    // (plan.cuh), and the planner runs on the device, so a finer level costs microseconds
    int L = NBINS_LOG2;
    while (L < ctx->fine_log2 && ((u64)1 << L) * (T / 8 + 1) < g_kmers) L++;
    ctx->bin_level = L;
    ctx->global_set = true; ctx->hist_fetched = false;
}

// ---- stage 1: close the input, read the totals --------------------------------------------------------------------
static int stage_totals(dskgpu_ctx* ctx)
{
    if (ctx->totals_done) return 0;
    Counters* ctr = (Counters*)ctx->ctr.p;
    if (ctx->stream_open) close_stream(ctx);
    // (wide spans: no density sample -- the sort path needs no table sizing; the solid-set buffers start from the default share)
    if (ctx->KW <= 2) { int rc = ctx->KW == 1 ? queue_sample<1>(ctx) : queue_sample<2>(ctx); if (rc) return rc; }
    {
        // packed bin histogram: could a record field have wrapped?  (test hook: a lower limit exercises the exact rebuild)
        unsigned long long lim = 1ULL << 28;
        if (const char* e = getenv("DSKGPU_TEST_HIST_LIMIT")) lim = (unsigned long long)std::max(1LL, atoll(e));
        k_check_bins<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((const unsigned long long*)ctx->bin_hist.p, 1u << ctx->fine_log2, lim, ctr); LAUNCHED();
    }
    CK(cudaMemcpyAsync(ctx->h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_ss, ctx->ss.p, sizeof(StreamState), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_hll, ctx->hll.p, sizeof(u32) * HLL_M, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->cfg.sequence_stats) CK(cudaMemcpyAsync(ctx->h_seqst, ctx->seqst.p, sizeof(SeqStats), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->probes.clear();
    if (ctx->h_ss->err) FAIL(DSKGPU_ERR_FORMAT, "device record scanner rejected the input (flags 0x%x): not plain FASTA / 4-line FASTQ", ctx->h_ss->err);
    if (ctx->h_ctr->overflow) FAIL(DSKGPU_ERR_OVERFLOW, "super-k-mer record buffer overflow");
    if (ctx->h_ctr->kmers_pass != ctx->h_ctr->kmers_in_recs)
        FAIL(DSKGPU_ERR_OVERFLOW, "internal: %llu valid k-mers in this pass but %llu packed in records", ctx->h_ctr->kmers_pass, ctx->h_ctr->kmers_in_recs);
    ctx->local_nrec = ctx->h_ctr->nrec; ctx->local_nkm = ctx->h_ctr->kmers_pass; ctx->bank_nkm = ctx->h_ctr->kmers_valid;
    ctx->sample_solid = ctx->h_ctr->sample_solid;
    ctx->hist_suspect = ctx->h_ctr->hist_suspect != 0;
    ctx->sample_nkm = ctx->h_ctr->sample_nkm; ctx->sample_distinct = ctx->h_ctr->sample_distinct;
    ctx->sample_wmult = ctx->sample_nkm >= 4096 ? (double)ctx->h_ctr->sample_sumsq / (double)ctx->sample_nkm : 0.0;
    ctx->st.nb_sequences = ctx->h_ss->nsep; ctx->st.nb_nucleotides = ctx->h_ss->nbase;
    if (ctx->cfg.sequence_stats) {
        const SeqStats& q = *ctx->h_seqst;
        ctx->st.seq_stats_sequences = q.n; ctx->st.seq_len_sum = q.sum; ctx->st.seq_len_sumsq = q.sumsq;
        ctx->st.seq_len_max = q.max_len; ctx->st.seq_len_min = q.n ? ~q.min_inv : 0;
        ctx->st.kmers_nb_invalid = q.windows >= ctx->h_ctr->kmers_valid ? q.windows - ctx->h_ctr->kmers_valid : 0;
    }
    ctx->st.kmers_nb_valid = ctx->bank_nkm; ctx->st.kmers_in_pass = ctx->local_nkm; ctx->st.nb_superkmers = ctx->local_nrec;
    ctx->st.superkmer_bytes = ctx->local_nrec * (u64)ctx->RW * 8;
    ctx->totals_done = true;
    return 0;
}

// ---- stage 2: plan the partitions from the whole-job bin histogram (on the device, plan.cuh) ----------------------------
// this rank's bin histogram at the planning level, on the device: [2 << bin_level] = records per bin, then k-mers per bin
static int fold_local_hist(dskgpu_ctx* ctx, const void** d_hist)
{
    if (!ctx->global_set) set_global(ctx, ctx->local_nkm, ctx->local_nrec, ctx->sample_nkm, ctx->sample_distinct);
    const int shift = ctx->fine_log2 - ctx->bin_level;
    const u32 nb = 1u << ctx->bin_level;
    if (!ctx->hist_fetched) {
        if (!ctx->hist_suspect) {
            k_fold_bins<<<(nb + 255) / 256, 256, 0, ctx->stream>>>((const unsigned long long*)ctx->bin_hist.p, 1u << ctx->fine_log2, shift, (unsigned long long*)ctx->bin_fold.p); LAUNCHED();
        } else {
            // a packed record count may have wrapped (one minimizer with hundreds of millions of k-mers): exact rebuild from the meta
            CK(cudaMemsetAsync(ctx->bin_fold.p, 0, sizeof(unsigned long long) * 2 * nb, ctx->stream));
            if (ctx->local_nrec) {
                k_rebuild_hist<<<(unsigned)std::min<u64>((ctx->local_nrec + 255) / 256, (u64)ctx->num_sms * 16), 256, 0, ctx->stream>>>(
                    (const u32*)ctx->meta.p, ctx->local_nrec, META_BIN_BITS - ctx->bin_level, nb, (unsigned long long*)ctx->bin_fold.p); LAUNCHED();
            }
            ctx->st.hist_rebuilt = 1;
        }
        CK(cudaGetLastError());
    }
    *d_hist = ctx->bin_fold.p;
    ctx->hist_fetched = true;
    return 0;
}

static bool use_smem_path(const dskgpu_ctx* ctx)
{
    const int mode = ctx->cfg.count_mode;
    return ctx->KW <= 2 && ctx->NB <= CS_MAX_BANKS && (mode == DSKGPU_COUNT_AUTO || mode == DSKGPU_COUNT_SMEM) && ctx->smem_cap >= 64;
}

// Shared-memory path limits.  One pass over a partition takes `fit` k-mers (table filled to 75 % at the sampled density);
// a bigger partition starts pre-split into 2^split0 RECORD sub-passes over the sub-bins stored in the records (count_smem.cuh):
// every record is expanded in exactly one sub-pass, so a partition of up to 16 tables costs little more than 16 partitions of
// one table.  (Round 1 split by k-mer hash residue: every sub-pass re-extracted every k-mer, and beyond two sub-passes the
// L2-resident global table was the cheaper path -- 48 % of the k-mers of BASELINE configs[3] at 8 GPUs took it,
// profiles/r03b.)  The size of a partition is bounded below by its heaviest minimizer bin: jobs of tens of G k-mers are
// where this matters.
static int smem_max_split0(const dskgpu_ctx*)
{
    const char* e = getenv("DSKGPU_SMEM_MAX_SPLIT0");                     // (read per call: the tests lower it to exercise the heavy paths)
    const int x = e ? atoi(e) : SMEM_MAX_SPLIT0_DEFAULT;
    return std::min(std::max(x, 0), (int)CS_MAX_SPLIT0);
}
static double smem_fit_kmers(const dskgpu_ctx* ctx) { return std::max(64.0, (double)ctx->smem_cap * 0.75 / ctx->density); }
static u64 smem_max_kmers(const dskgpu_ctx* ctx)
{
    // forced SMEM mode (tests) starts everything in one pass and lets the kernel discover the splits
    if (ctx->cfg.count_mode == DSKGPU_COUNT_SMEM) return (u64)ctx->smem_cap * 16;
    return (u64)(smem_fit_kmers(ctx) * (double)(1 << smem_max_split0(ctx)));
}

// k-mers a partition should hold AT MOST in the common case.  Shared-memory path: what fills the table to ~52 % given the
// sampled density (distinct / total k-mers: 0.26 for 100x reads at k=31, 0.47 at k=63, 0.43 for 30x reads).
// Global-table path: a quarter of the table capacity, so that groups of partitions can be sized to the measured occupancy.
static u64 plan_target_kmers(const dskgpu_ctx* ctx, u64 global_kmers)
{
    if (ctx->cfg.nb_partitions > 0) return std::max<u64>(1, (global_kmers + ctx->cfg.nb_partitions - 1) / (u64)ctx->cfg.nb_partitions);
    if (use_smem_path(ctx)) {
        // table load after the last insert = T * density / slots: 52 % is the measured optimum (C2: density 0.26, T = 200 % of
        // the slots -> 4.03 ms; 125 % -> 4.36 ms; beyond 65 % overflows (= split passes) appear)
        const char* e = getenv("DSKGPU_SMEM_LOAD_PCT");
        const double load = (e ? (double)std::max(5, atoi(e)) : 52.0) / 100.0;
        double t = std::min((double)ctx->smem_cap * load / ctx->density, (double)ctx->smem_cap * 4.0);
        // ... and whose records fit the shared-memory job buffer in one slice (cs_bufrec): the whole job is then staged by one
        // bulk copy that runs under the previous job's sweep
        if (ctx->g_total_recs > 0 && ctx->cfg.count_mode == DSKGPU_COUNT_AUTO && ctx->cfg.smem_table_slots <= 0) {
            const double avg_nk = (double)ctx->g_total_kmers / (double)ctx->g_total_recs;
            t = std::min(t, (double)(ctx->KW == 1 ? cs_bufrec<1>() : cs_bufrec<2>()) * avg_nk * 0.93);   // (KW <= 2 here: use_smem_path)
        }
        return (u64)std::max(t, 64.0);
    }
    const int log2s = ctx->cfg.hash_log2_slots > 0 ? std::max(10, ctx->cfg.hash_log2_slots) : 23;
    return std::max<u64>(((u64)1 << log2s) * 6 / 10 / 4, 4096);
}

// The cut rule of plan.cuh gives partitions of T k-mers on average and T + (the last bin) at most; bins average an eighth of
// a partition (set_global), so cutting at 85 % of the target keeps the common case under it.  A forced partition count
// (-nb-partitions style) is honoured exactly.
static PlanParams make_plan_params(const dskgpu_ctx* ctx)
{
    PlanParams pp;
    const u64 T = plan_target_kmers(ctx, ctx->g_total_kmers);
    pp.T = ctx->cfg.nb_partitions > 0 ? T : std::max<u64>(1, (u64)((double)T * 0.85));
    pp.lim = smem_max_kmers(ctx);
    pp.nbins = 1u << ctx->bin_level;
    pp.W = (u32)ctx->cfg.world_size; pp.me = (u32)ctx->cfg.rank;
    pp.smem_ok = use_smem_path(ctx) ? 1u : 0u;
    return pp;
}

// d_ghist: whole-job histogram (device, [2 << bin_level]; on one GPU the local one).  Leaves on the device: bin2part (bin -> q),
// gk_q / gr_q / lcnt_q (per partition, q order), loff (exclusive prefix of lcnt_q); on the host: the header and the owned heavy
// partitions.  One stream sync (two when there are heavy partitions).
static int plan_device(dskgpu_ctx* ctx, const void* d_ghist)
{
    const void* d_lhist = nullptr;
    int rc;
    if ((rc = fold_local_hist(ctx, &d_lhist))) return rc;
    if (!d_ghist) d_ghist = d_lhist;
    const PlanParams pp = make_plan_params(ctx);
    const u32 nb = pp.nbins, W = pp.W, me = pp.me;
    const u64 qcap = (u64)nb + W;                                        // W * PW <= P + W <= nbins + W
    if ((rc = ensure(ctx, ctx->pl_ex, (nb + 1) * 8ull))) return rc;
    if ((rc = ensure(ctx, ctx->pl_E, (nb + 1) * 8ull))) return rc;
    if ((rc = ensure(ctx, ctx->pl_H, (nb + 1) * 8ull))) return rc;
    if ((rc = ensure(ctx, ctx->pl_bsum, ps_bsum_bytes(qcap)))) return rc;
    if ((rc = ensure(ctx, ctx->pl_pk, nb * 8ull * 3))) return rc;         // pk | pr | pl, cleared together
    if ((rc = ensure(ctx, ctx->pl_newid, nb * 4ull))) return rc;
    if ((rc = ensure(ctx, ctx->pl_nvals, 64))) return rc;
    if ((rc = ensure(ctx, ctx->pl_hdr, sizeof(PlanHdr)))) return rc;
    if ((rc = ensure(ctx, ctx->gk_q, qcap * 8 * 3))) return rc;            // gk_q | gr_q | lcnt_q, cleared together
    if ((rc = ensure(ctx, ctx->loff, (qcap + 1) * 8))) return rc;
    if ((rc = ensure(ctx, ctx->bin2part, nb * 4ull))) return rc;
    cudaStream_t st = ctx->stream;
    cudaEvent_t pa = get_event(ctx), pb = get_event(ctx);
    cudaEventRecord(pa, st);
    const u64* gh = (const u64*)d_ghist; const u64* lh = (const u64*)d_lhist;
    u64* ex = (u64*)ctx->pl_ex.p; u64* E = (u64*)ctx->pl_E.p; u64* H = (u64*)ctx->pl_H.p; u64* bsum = (u64*)ctx->pl_bsum.p;
    unsigned long long* pk = (unsigned long long*)ctx->pl_pk.p; unsigned long long* pr = pk + nb; unsigned long long* pl = pr + nb;
    u64* gk_q = (u64*)ctx->gk_q.p; u64* gr_q = gk_q + qcap; u64* lcnt_q = gr_q + qcap;
    u64* nvals = (u64*)ctx->pl_nvals.p; PlanHdr* hdr = (PlanHdr*)ctx->pl_hdr.p;
    const unsigned gb = (unsigned)std::min<u32>((nb + 255) / 256, (u32)ctx->num_sms * 8);
    CK(cudaMemsetAsync(pk, 0, nb * 8ull * 3, st));
    CK(cudaMemsetAsync(gk_q, 0, qcap * 8 * 3, st));
    CK(cudaMemsetAsync(hdr, 0, sizeof(PlanHdr), st));
    ctx->st.gpu_launches += ps_scan(st, PsLoad{gh + nb}, nullptr, nb, bsum, ex);                       // k-mer offset of every bin
    ctx->st.gpu_launches += ps_scan(st, PsCutFlag{ex, pp.T}, nullptr, nb, bsum, E);                    // cuts before every bin
    k_plan_sums<<<gb, 256, 0, st>>>(gh, lh, ex, E, pp, pk, pr, pl, nvals); LAUNCHED();
    ctx->st.gpu_launches += ps_scan(st, PsHeavyFlag{(const u64*)pk, (const u64*)pr, pp.lim, pp.smem_ok}, nvals, nb, bsum, H);
    k_plan_renumber<<<gb, 256, 0, st>>>((const u64*)pk, (const u64*)pr, (const u64*)pl, H, pp, nvals, (u32*)ctx->pl_newid.p, gk_q, gr_q, lcnt_q, hdr); LAUNCHED();
    k_plan_bin2q<<<gb, 256, 0, st>>>(ex, E, (const u32*)ctx->pl_newid.p, pp, nvals, (u32*)ctx->bin2part.p); LAUNCHED();
    ctx->st.gpu_launches += ps_scan(st, PsLoad{lcnt_q}, nvals + 1, qcap, bsum, (u64*)ctx->loff.p);     // first record of every partition in q order
    k_plan_hdr<<<W, 256, 0, st>>>(gk_q, gr_q, (const u64*)ctx->loff.p, hdr); LAUNCHED();
    CK(cudaMemcpyAsync(ctx->h_hdr, hdr, sizeof(PlanHdr), cudaMemcpyDeviceToHost, st));
    cudaEventRecord(pb, st);
    ctx->spans.push_back({pa, pb, SPAN_PLAN});
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const PlanHdr& h = *ctx->h_hdr;
    ctx->nparts = h.P; ctx->st.nb_partitions = h.P;
    ctx->np_me = me < h.P ? (h.P - me + W - 1) / W : 0;
    ctx->nl_me = me < h.nlight ? (h.nlight - me + W - 1) / W : 0;
    ctx->heavy_recs.clear(); ctx->heavy_kmers.clear();
    const u32 nh = ctx->np_me - ctx->nl_me;
    if (nh) {
        if ((size_t)nh * 16 > ctx->h_heavy_cap) {
            if (ctx->h_heavy) cudaFreeHost(ctx->h_heavy);
            ctx->h_heavy = nullptr; ctx->h_heavy_cap = 0;
            CK(cudaMallocHost((void**)&ctx->h_heavy, (size_t)nh * 16 + 4096));
            ctx->h_heavy_cap = (size_t)nh * 16 + 4096;
        }
        const u64 q0 = (u64)me * h.PW + ctx->nl_me;
        CK(cudaMemcpyAsync(ctx->h_heavy, gr_q + q0, (size_t)nh * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->h_heavy + nh, gk_q + q0, (size_t)nh * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->heavy_recs.assign(ctx->h_heavy, ctx->h_heavy + nh);
        ctx->heavy_kmers.assign(ctx->h_heavy + nh, ctx->h_heavy + 2 * (size_t)nh);
    }
    ctx->planned = true;
    return 0;
}

// ---- stage 3: scatter my records into q order: owner-major, so every owner's records are one contiguous chunk -------------
// Up to ~1 M partitions: one pass, one L2 atomic + one 16/32-byte store per record (k_part_scatter).  Beyond (multi-G k-mer
// jobs on several GPUs): MSD multi-split passes with block-level binning in shared memory (k_msd_pass), ping-ponging between
// the input-order buffer and lrecs; ctx->qrecs says where the q-ordered records ended up.
static u64 msd_min_parts(int KW)
{
    // measured (profiles/r02v-r02w): 16-byte records -- the passes win from ~10 K partitions on (C2: 0.72 against 0.80 ms;
    // 563 K partitions: 28.8 against 38.3 ms; 3.3 M partitions on 8 GPUs: 96 ms for the single pass); 32-byte records move
    // twice the bytes per pass and only win once the single pass has lost its L2 locality
    const char* e = getenv("DSKGPU_MSD_MIN_PARTS");                       // (read per call: the tests force the multi-pass path on small jobs)
    return e ? (u64)std::max(1LL, atoll(e)) : (KW == 1 ? (u64)8192 : (u64)1 << 18);
}

template <int KW>
static int stage_scatter(dskgpu_ctx* ctx)
{
    const PlanHdr& h = *ctx->h_hdr;
    int rc;
    if ((rc = ensure(ctx, ctx->lrecs, ctx->local_nrec * (u64)ctx->RW * 8 + 64))) return rc;
    ctx->qrecs = ctx->lrecs.p;
    if (ctx->local_nrec == 0) return 0;
    const u64 nq = (u64)ctx->cfg.world_size * h.PW;
    if ((rc = ensure(ctx, ctx->cursor, (nq + 1) * 8))) return rc;
    SpanGuard g(ctx, SPAN_PART);
    const int bin_shift = META_BIN_BITS - ctx->bin_level;
    if (nq < msd_min_parts(KW)) {
        CK(cudaMemcpyAsync(ctx->cursor.p, ctx->loff.p, nq * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        const unsigned sb = (unsigned)std::min<u64>((ctx->local_nrec + SC_THREADS - 1) / SC_THREADS, (u64)ctx->num_sms * 32);
        k_part_scatter<KW><<<sb, SC_THREADS, 0, ctx->stream>>>((const u64*)ctx->recs.p, (const u32*)ctx->meta.p, ctx->local_nrec,
                                                              (const u32*)ctx->bin2part.p, bin_shift, (u64*)ctx->lrecs.p,
                                                              (unsigned long long*)ctx->cursor.p); LAUNCHED();
        CK(cudaGetLastError());
        return 0;
    }
    int bits = 1; while (((u64)1 << bits) < nq) bits++;
    const int npass = (bits + 7) / 8, per = (bits + npass - 1) / npass;
    if ((rc = ensure(ctx, ctx->mkeys, ctx->local_nrec * 4 + 64))) return rc;
    const unsigned grid = (unsigned)std::min<u64>((ctx->local_nrec + MS_THREADS * (KW == 1 ? 8 : 4) - 1) / (MS_THREADS * (KW == 1 ? 8 : 4)), (u64)ctx->num_sms * 8);
    u64* rb[2] = {(u64*)ctx->recs.p, (u64*)ctx->lrecs.p};
    u32* kb[2] = {(u32*)ctx->meta.p, (u32*)ctx->mkeys.p};
    for (int p = 0; p < npass; p++) {
        const int shift = per * (npass - 1 - p), parent_shift = per * (npass - p);
        const bool first = p == 0, last = p == npass - 1;
        k_msd_cursor<<<(unsigned)std::min<u64>(((nq >> shift) + 256) / 256, (u64)ctx->num_sms * 8), 256, 0, ctx->stream>>>((const u64*)ctx->loff.p, nq, shift, (unsigned long long*)ctx->cursor.p); LAUNCHED();
        const u64* src = rb[p & 1]; u64* dst = rb[(p + 1) & 1];
        const u32* sk = kb[p & 1]; u32* dk = kb[(p + 1) & 1];
        unsigned long long* cur = (unsigned long long*)ctx->cursor.p;
        const u32* b2q = (const u32*)ctx->bin2part.p;
#define MSD_LAUNCH(F, L) do { CK(cudaFuncSetAttribute(k_msd_pass<KW, F, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ms_smem_bytes<KW>())); \
        k_msd_pass<KW, F, L><<<grid, MS_THREADS, ms_smem_bytes<KW>(), ctx->stream>>>(src, sk, ctx->local_nrec, b2q, bin_shift, shift, parent_shift, cur, dst, dk); } while (0)
        if (first && last) MSD_LAUNCH(true, true);
        else if (first) MSD_LAUNCH(true, false);
        else if (last) MSD_LAUNCH(false, true);
        else MSD_LAUNCH(false, false);
#undef MSD_LAUNCH
        LAUNCHED();
    }
    ctx->qrecs = rb[npass & 1];
    ctx->st.scatter_passes = (u32)npass;
    CK(cudaGetLastError());
    return 0;
}

// the segment table the counting kernels read (and, on several GPUs, where every chunk goes).  d_rcnt: [W][PW] records of my
// partitions held by every rank; d_S: [W][W] chunk sizes (records rank s holds for rank o); peers: receive buffers.
static int stage_bases(dskgpu_ctx* ctx, const void* d_rcnt, const void* d_S)
{
    const PlanHdr& h = *ctx->h_hdr;
    const u32 W = (u32)ctx->cfg.world_size, me = (u32)ctx->cfg.rank;
    const u64 nq = (u64)W * h.PW;
    int rc;
    if ((rc = ensure(ctx, ctx->xtab, sizeof(XchgTab)))) return rc;
    if ((rc = ensure(ctx, ctx->xpeers, (size_t)PLAN_MAXW * 8))) return rc;
    const u64* X = (const u64*)ctx->loff.p;
    if (W > 1) {
        if ((rc = ensure(ctx, ctx->xX, (nq + 1) * 8))) return rc;
        if ((rc = ensure(ctx, ctx->pl_bsum, ps_bsum_bytes(nq)))) return rc;
        ctx->st.gpu_launches += ps_scan(ctx->stream, PsLoad{(const u64*)d_rcnt}, nullptr, nq, (u64*)ctx->pl_bsum.p, (u64*)ctx->xX.p);
        X = (const u64*)ctx->xX.p;
        CK(cudaMemcpyAsync(ctx->xpeers.p, ctx->peer_recv.data(), (size_t)W * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->rcnt_dev = X;
    const u64* S = W > 1 ? (const u64*)d_S : ((const PlanHdr*)ctx->pl_hdr.p)->send_recs;
    k_xchg_bases<<<1, 32, 0, ctx->stream>>>((const PlanHdr*)ctx->pl_hdr.p, (const u64*)ctx->loff.p, X, S, (const u64*)ctx->xpeers.p,
                                            (u64)(uintptr_t)ctx->qrecs, W, me, (u32)ctx->RW * 8, (XchgTab*)ctx->xtab.p); LAUNCHED();
    CK(cudaMemcpyAsync(ctx->h_xtab, ctx->xtab.p, sizeof(XchgTab), cudaMemcpyDeviceToHost, ctx->stream));   // read at the next sync (bad flags)
    CK(cudaGetLastError());
    return 0;
}

// ---- heavy partitions: hash-bucketed flat keys, counted in shared memory ---------------------------------------------
// prec/pkm: records / k-mers of the partitions of one contiguous run starting at `recs`.  Groups of consecutive partitions
// (<= BUCKET_GROUP_KMERS k-mers) are expanded into S = k-mers / T hash buckets with fixed-size slabs (uniform hashing: the
// slab is the mean + 15 % + 6 sigma), then counted bucket by bucket by k_count_smem<KW, MB, true>.  A slab can only
// overflow when a single k-mer has a huge multiplicity (low-complexity input): that group falls back to the global table.
constexpr u64 BUCKET_GROUP_KMERS = (u64)1 << 27;

template <int KW>
static int count_all(dskgpu_ctx* ctx, const u64* recs, const std::vector<u64>& prec, const std::vector<u64>& pkm, u64 out_cap);

template <int KW>
static int count_by_buckets(dskgpu_ctx* ctx, const u64* recs, const std::vector<u64>& prec, const std::vector<u64>& pkm, u64 out_cap)
{
    Counters* ctr = (Counters*)ctx->ctr.p;
    const size_t np = prec.size();
    // k-mers per bucket: what fills the shared-memory table to ~52 % at the sampled density (same rule as the partitions)
    const u64 T = (u64)std::min(std::max((double)ctx->smem_cap * 0.52 / ctx->density, 64.0), (double)ctx->smem_cap * 4.0);
    int rc;
    size_t p = 0; u64 roff = 0;
    while (p < np) {
        size_t q = p; u64 km = 0, nr = 0;
        while (q < np && (q == p || km + pkm[q] <= BUCKET_GROUP_KMERS)) { km += pkm[q]; nr += prec[q]; q++; }
        if (km == 0) { p = q; roff += nr; continue; }
        const u64 S64 = std::max<u64>(1, (km + T - 1) / T);
        // slab = mean + 8 sigma.  Equal k-mers share a bucket, so a bucket's load is a compound Poisson: its variance is
        // mean x (occurrence-weighted multiplicity of a k-mer), measured on the density sample (x 1.5: hot minimizer bins
        // are more repetitive than the average bin); without a sample, a generous 256.
        const double mean = (double)km / (double)S64;
        const double wm = ctx->sample_wmult > 0.0 ? std::max(1.0, ctx->sample_wmult * 1.5) : 256.0;
        const u64 slab64 = (u64)std::max(mean * 1.25, mean + 8.0 * std::sqrt(mean * wm)) + 64;
        bool fallback = S64 > 0x7FFFFFFFull || slab64 > 0x7FFFFFFFull;
        if (!fallback) {
            const u32 S = (u32)S64, slab = (u32)slab64;
            if ((rc = ensure(ctx, ctx->keys[0], (u64)S * slab * KW * 8 + 64))) return rc;
            if ((rc = ensure(ctx, ctx->bcur, (size_t)S * 4))) return rc;
            CK(cudaMemsetAsync(ctx->bcur.p, 0, (size_t)S * 4, ctx->stream));
            const unsigned gb = (unsigned)std::min<u64>((nr + 255) / 256, (u64)ctx->num_sms * 8);
            k_expand_bucket<KW><<<gb ? gb : 1, 256, 0, ctx->stream>>>(recs + roff * (u64)ctx->RW, 0, nr, ctx->k, ctx->NB, S, slab, (u64*)ctx->keys[0].p,
                                                                     (u32*)ctx->bcur.p, ctr); LAUNCHED();
            CK(cudaMemcpyAsync(ctx->h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->h_ctr->bucket_overflow) { fallback = true; CK(cudaMemsetAsync(&ctr->bucket_overflow, 0, sizeof(unsigned int), ctx->stream)); }
            else {
                CK(cudaMemsetAsync(ctx->work_ctr.p, 0, 64, ctx->stream));
                const unsigned grid = (unsigned)std::min<u64>(S, (u64)ctx->num_sms * CS_CTAS_PER_SM);
                const size_t dyn = cs_smem_bytes<KW>(ctx->smem_cap, ctx->NB);
                cudaEvent_t a = get_event(ctx), b = get_event(ctx);
                cudaEventRecord(a, ctx->stream);
                if (ctx->NB == 1)
                    k_count_smem<KW, false, true><<<grid, CS_THREADS, dyn, ctx->stream>>>((const u64*)ctx->keys[0].p, CsSegs(), S, ctx->k, ctx->smem_cap,
                        (long long)ctx->cfg.abundance_min[0], (long long)ctx->cfg.abundance_max, (u64*)ctx->skeys[0].p, (u32*)ctx->svals[0].p, out_cap,
                        (unsigned long long*)ctx->hist.p, ctr, (u32*)ctx->work_ctr.p, 1, SolidityParams(), nullptr, (const u32*)ctx->bcur.p, slab);
                else
                    k_count_smem<KW, true, true><<<grid, CS_THREADS, dyn, ctx->stream>>>((const u64*)ctx->keys[0].p, CsSegs(), S, ctx->k, ctx->smem_cap,
                        (long long)ctx->cfg.abundance_min[0], (long long)ctx->cfg.abundance_max, (u64*)ctx->skeys[0].p, (u32*)ctx->svals[0].p, out_cap,
                        (unsigned long long*)ctx->hist.p, ctr, (u32*)ctx->work_ctr.p, ctx->NB, make_sp(ctx), (unsigned long long*)ctx->hist2d.p,
                        (const u32*)ctx->bcur.p, slab);
                LAUNCHED();
                cudaEventRecord(b, ctx->stream);
                ctx->spans.push_back({a, b, SPAN_DOM});
                ctx->st.nb_groups_bucket++;
                CK(cudaGetLastError());
            }
        }
        if (fallback) {
            std::vector<u64> rp(prec.begin() + p, prec.begin() + q), rk(pkm.begin() + p, pkm.begin() + q);
            if ((rc = count_all<KW>(ctx, recs + roff * (u64)ctx->RW, rp, rk, out_cap))) return rc;
        }
        roff += nr; p = q;
    }
    return 0;
}

// Which path takes the heavy partitions (measured on B200, profiles/r01u-r01v): with one count per k-mer the L2-resident
// global table wins (3 G k-mer job: count stage 65 ms against 116 ms -- k_expand_bucket's per-k-mer cursor atomics and
// scattered stores run at 25 G k-mers/s); with per-bank counts the key buckets win (-histo2D, 4.1 G k-mers: 169 ms against
// 267 ms).  DSKGPU_HEAVY_PATH=bucket|table overrides.  The first answer to heavy bins is a longer minimizer
// (dskgpu_suggest_minimizer_size): both paths are what is left for low-complexity input.
static bool heavy_by_buckets(const dskgpu_ctx* ctx)
{
    if (!(ctx->cfg.count_mode == DSKGPU_COUNT_AUTO && use_smem_path(ctx))) return false;
    const char* e = getenv("DSKGPU_HEAVY_PATH");
    if (e && strcmp(e, "table") == 0) return false;
    if (e && strcmp(e, "bucket") == 0) return true;
    return ctx->NB > 1;
}

// ---- stage 4: count the partitions this rank owns, order the solid set, copy results out --------------------------------
// Light partitions (jobs [0, nl_me) of my chunk of the q-ordered tables) go to the shared-memory kernel, which reads the
// planner's device tables itself; the owned heavy partitions (host vectors heavy_recs / heavy_kmers) to the host-driven paths.
template <int KW>
static int stage_count_once(dskgpu_ctx* ctx, u64 cap_request, u64* need_cap)
{
    Counters* ctr = (Counters*)ctx->ctr.p;
    *need_cap = 0;
    int rc;
    const PlanHdr& h = *ctx->h_hdr;
    const u32 W = (u32)ctx->cfg.world_size, me = (u32)ctx->cfg.rank;
    const u64 nrec = h.need_recs[me], nkm = h.need_kmers[me];
    ctx->st.smem_table_slots = ctx->smem_cap;
    ctx->st.density_ppm = ctx->density_known ? (u32)(ctx->density * 1e6) : 0u; ctx->st.log2_bins = (u32)ctx->bin_level;
    if (nrec) {
        // Capacity of the solid set.  Worst case: every solid k-mer holds at least min(abundance_min) occurrences -- 108 GB of
        // ping-pong buffers for a 9 G k-mer job whose solid set is 5 GB.  Big jobs therefore start from an ESTIMATE (share of
        // the sampled distinct k-mers that reach the threshold, with margin); the kernels keep counting past the capacity, so
        // an overflowing pass still reports the exact size and the counting stage is redone once at that size (stage_count).
        long long amin = ctx->cfg.abundance_min[0];
        for (int b = 1; b < ctx->NB; b++) amin = std::min<long long>(amin, ctx->cfg.abundance_min[b]);
        if (amin < 1 || ctx->cfg.solidity_kind == DSKGPU_SOLIDITY_CUSTOM) amin = 1;
        const u64 worst = nkm / (u64)amin + 1024;
        u64 out_cap = worst;
        if (cap_request) out_cap = std::min(worst, cap_request);
        else if (worst * (u64)(KW * 8 + 4) * 2 > ((u64)256 << 20)) {
            double share = 0.08;                                                   // of the k-mers counted here
            if (ctx->cfg.world_size == 1 && ctx->sample_nkm >= 4096) share = 1.5 * (double)ctx->sample_solid / (double)ctx->sample_nkm + 0.01;
            const u64 have = std::min<u64>(ctx->skeys[0].cap / (KW * 8), std::min<u64>(ctx->skeys[1].cap / (KW * 8), std::min<u64>(ctx->svals[0].cap / 4, ctx->svals[1].cap / 4)));
            out_cap = std::min(worst, std::max(have, (u64)((double)nkm * share) + ((u64)1 << 20)));
        }
        if (const char* e = getenv("DSKGPU_TEST_SOLID_CAP")) { if (!cap_request) out_cap = std::min(worst, (u64)std::max(1LL, atoll(e))); }   // test hook: a wrong estimate
        ctx->solid_cap = out_cap;
        for (int i = 0; i < 2; i++) {
            if ((rc = ensure(ctx, ctx->skeys[i], out_cap * KW * 8))) return rc;
            if ((rc = ensure(ctx, ctx->svals[i], out_cap * 4))) return rc;
        }
        SpanGuard g(ctx, SPAN_COUNT);
        if constexpr (KW > 2) { if (ctx->nl_me) FAIL(DSKGPU_ERR_STATE, "internal: shared-memory jobs planned for a wide span"); }
        else if (ctx->nl_me) {
            // occupancy picks the path of every partition (K/SortingCountAlgorithm.cpp:1489-1497): the planner already put
            // the partitions within reach of a few split passes first.  A partition expected to fill the table beyond 75 %
            // (k-mers x sampled density) starts as 2^split0 sub-passes over hash residues (computed by the kernel from gk_q);
            // forced SMEM mode (tests) starts everything in one pass and lets the kernel discover the splits.
            CsSegs sg;
            sg.tab = (const XchgTab*)ctx->xtab.p; sg.X = (const u64*)ctx->rcnt_dev; sg.gk_q = (const u64*)ctx->gk_q.p;
            sg.W = W; sg.PW = h.PW; sg.qbase = me * h.PW;
            sg.fit = (float)(smem_fit_kmers(ctx) * 0.85);                            // sub-bins are minimizers: sub-passes are not of equal size
            sg.max_split0 = ctx->cfg.count_mode != DSKGPU_COUNT_SMEM ? (u32)smem_max_split0(ctx) : 0u;
            CK(cudaMemsetAsync(ctx->work_ctr.p, 0, 64, ctx->stream));
            const unsigned grid = (unsigned)std::min<size_t>(ctx->nl_me, (size_t)ctx->num_sms * CS_CTAS_PER_SM);
            const size_t dyn = cs_smem_bytes<KW>(ctx->smem_cap, ctx->NB);
            cudaEvent_t a = get_event(ctx), b = get_event(ctx);
            cudaEventRecord(a, ctx->stream);
            if (ctx->NB == 1)
                k_count_smem<KW, false><<<grid, CS_THREADS, dyn, ctx->stream>>>(nullptr, sg, ctx->nl_me, ctx->k, ctx->smem_cap,
                                                                   (long long)ctx->cfg.abundance_min[0], (long long)ctx->cfg.abundance_max,
                                                                   (u64*)ctx->skeys[0].p, (u32*)ctx->svals[0].p, out_cap,
                                                                   (unsigned long long*)ctx->hist.p, ctr, (u32*)ctx->work_ctr.p, 1, SolidityParams(), nullptr);
            else
                k_count_smem<KW, true><<<grid, CS_THREADS, dyn, ctx->stream>>>(nullptr, sg, ctx->nl_me, ctx->k, ctx->smem_cap,
                                                                   (long long)ctx->cfg.abundance_min[0], (long long)ctx->cfg.abundance_max,
                                                                   (u64*)ctx->skeys[0].p, (u32*)ctx->svals[0].p, out_cap,
                                                                   (unsigned long long*)ctx->hist.p, ctr, (u32*)ctx->work_ctr.p, ctx->NB, make_sp(ctx),
                                                                   (unsigned long long*)ctx->hist2d.p);
            LAUNCHED();
            cudaEventRecord(b, ctx->stream);
            ctx->spans.push_back({a, b, SPAN_DOM});
            ctx->st.nb_parts_smem = ctx->nl_me;
            CK(cudaGetLastError());
            trace("count kernel launched");
        }
        // the owned heavy partitions: one contiguous run of records for the global hash / sort / bucket paths.  On one GPU
        // they already are the tail of lrecs (q order = id order, heavy ids last); on several their W segments are gathered.
        if (!ctx->heavy_recs.empty()) {
            SpanGuard gh(ctx, SPAN_HEAVY);
            u64 hrec = 0;
            for (u64 r : ctx->heavy_recs) hrec += r;
            const u64* hp = nullptr;
            if (W == 1) hp = (const u64*)ctx->qrecs + (nrec - hrec) * (u64)ctx->RW;
            else if (hrec) {
                const size_t nh = ctx->heavy_recs.size();
                if ((rc = ensure(ctx, ctx->hrecs, hrec * (u64)ctx->RW * 8 + 64))) return rc;
                if ((rc = ensure(ctx, ctx->hoff, (nh + 1) * 8))) return rc;
                u64* ho = ctx->h_heavy;                                           // pinned, >= 2 * nh words: reuse as staging of the prefix
                { u64 o = 0; for (size_t i = 0; i < nh; i++) { ho[i] = o; o += ctx->heavy_recs[i]; } }
                CK(cudaMemcpyAsync(ctx->hoff.p, ho, nh * 8, cudaMemcpyHostToDevice, ctx->stream));
                const u32 total_blocks = (u32)ctx->num_sms * 8;
                const u32 G = (u32)std::max<u64>(1, std::min<u64>(total_blocks, total_blocks / std::max<u64>(1, (u64)nh * W)));
                const unsigned gg = total_blocks / G * G;
                k_gather_heavy<<<gg, 256, 0, ctx->stream>>>((const XchgTab*)ctx->xtab.p, (const u64*)ctx->rcnt_dev, h.PW, W, ctx->nl_me, (u32)nh,
                                                           (const u64*)ctx->hoff.p, (ulonglong2*)ctx->hrecs.p, (u32)ctx->RW / 2, G); LAUNCHED();
                CK(cudaStreamSynchronize(ctx->stream));                           // h_heavy is rewritten by the next plan; cheap next to the heavy paths
                hp = (const u64*)ctx->hrecs.p;
            }
            if (hrec) {
                bool buckets = false;
                if constexpr (KW <= 2) { if (heavy_by_buckets(ctx)) { buckets = true; rc = count_by_buckets<KW>(ctx, hp, ctx->heavy_recs, ctx->heavy_kmers, out_cap); } }
                if (!buckets) rc = count_all<KW>(ctx, hp, ctx->heavy_recs, ctx->heavy_kmers, out_cap);
                if (rc) return rc;
            }
        }
    }
    CK(cudaMemcpyAsync(ctx->h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    trace("count done (sync)");
    if (ctx->h_xtab->bad) FAIL(DSKGPU_ERR_STATE, "exchange layout inconsistent on the device (flags 0x%x): the ranks disagree on the plan or on the record counts", ctx->h_xtab->bad);
    if (ctx->h_ctr->hash_overflow) FAIL(DSKGPU_ERR_OVERFLOW, "hash table overflow (distinct k-mer estimate too low)");
    if (ctx->h_ctr->smem_failed) FAIL(DSKGPU_ERR_OVERFLOW, "shared-memory table overflow at the deepest split (%u passes)", ctx->h_ctr->smem_failed);
    if (ctx->h_ctr->overflow) {                                    // the cursor kept counting: solid_n is the exact size needed
        if (ctx->h_ctr->solid_n <= ctx->solid_cap) FAIL(DSKGPU_ERR_OVERFLOW, "solid k-mer buffer overflow (internal)");
        *need_cap = ctx->h_ctr->solid_n + 1024;
        return DSKGPU_OK;
    }
    ctx->n_solid = ctx->h_ctr->solid_n;
    ctx->st.kmers_nb_distinct = ctx->h_ctr->distinct_n; ctx->st.kmers_nb_solid = ctx->n_solid;
    ctx->st.nb_smem_splits = ctx->h_ctr->smem_splits;

    // order the solid set (ascending k-mer value, as the reference emits within a partition), results to the host
    ctx->solid_buf = 0;
    auto order_and_copy = [&](bool full, int start_buf) -> int {
        if (ctx->n_solid) {
            SpanGuard g(ctx, SPAN_SORT);
            int rc2 = sort_solid<KW>(ctx, full, start_buf); if (rc2) return rc2;
        }
        CK(cudaMemcpyAsync(ctx->h_hist, ctx->hist.p, sizeof(unsigned long long) * DSKGPU_HISTO_LEN, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_hist + DSKGPU_HISTO_LEN, ctx->hist2d.p, sizeof(unsigned long long) * DSKGPU_HISTO_LEN * DSKGPU_HISTO2D_DIM2,
                           cudaMemcpyDeviceToHost, ctx->stream));
        trace("ordering queued");
        if (!ctx->cfg.keep_results_on_device && ctx->n_solid) {
            if (ctx->n_solid > ctx->h_solid_cap) {
                if (ctx->h_skeys) cudaFreeHost(ctx->h_skeys);
                if (ctx->h_svals) cudaFreeHost(ctx->h_svals);
                ctx->h_skeys = nullptr; ctx->h_svals = nullptr;
                ctx->h_solid_cap = ctx->n_solid + ctx->n_solid / 4;
                CK(cudaMallocHost((void**)&ctx->h_skeys, ctx->h_solid_cap * KW * 8));
                CK(cudaMallocHost((void**)&ctx->h_svals, ctx->h_solid_cap * 4));
            }
            trace("host result buffers ready");
            CK(cudaMemcpyAsync(ctx->h_skeys, ctx->skeys[ctx->solid_buf].p, ctx->n_solid * KW * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(ctx->h_svals, ctx->svals[ctx->solid_buf].p, ctx->n_solid * 4, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->results_on_host = true;
        }
        CK(cudaStreamSynchronize(ctx->stream));
        trace("results on the host");
        return 0;
    };
    *(volatile unsigned int*)(ctx->h_nrec_probe + 7) = 0;
    if ((rc = order_and_copy(false, 0))) return rc;
    if (ctx->sort_fixup && *(volatile unsigned int*)(ctx->h_nrec_probe + 7)) {
        // a prefix group too long for the neighbourhood fix-up (low-complexity set): plain full-width sort from its intact input
        CK(cudaMemsetAsync(&ctr->sort_fallback, 0, sizeof(unsigned int), ctx->stream));
        ctx->st.sort_fallbacks++;
        if ((rc = order_and_copy(true, ctx->sort_src))) return rc;
    }
    ctx->st.ms_parse = span_ms(ctx, SPAN_PARSE); ctx->st.ms_superk = span_ms(ctx, SPAN_SUPERK);
    ctx->st.ms_partition = span_ms(ctx, SPAN_PART); ctx->st.ms_count = span_ms(ctx, SPAN_COUNT);
    ctx->st.ms_sort = span_ms(ctx, SPAN_SORT);
    ctx->st.ms_dominant_kernel = span_ms(ctx, SPAN_DOM, &ctx->st.dominant_kernel_launches);
    ctx->st.ms_exchange = span_ms(ctx, SPAN_XCHG); ctx->st.exchange_bytes_out = ctx->xchg_bytes_out;
    ctx->st.ms_plan = span_ms(ctx, SPAN_PLAN); ctx->st.ms_count_heavy = span_ms(ctx, SPAN_HEAVY);
    ctx->st.ms_push_wall = 0;
    if (ctx->push_timed) { float ms = 0; if (cudaEventElapsedTime(&ms, ctx->ev_push0, ctx->ev_push1) == cudaSuccess) ctx->st.ms_push_wall = ms; }
    ctx->st.ms_total = ctx->st.ms_parse + ctx->st.ms_superk + ctx->st.ms_partition + ctx->st.ms_count + ctx->st.ms_sort;
    ctx->state = 1;
    return DSKGPU_OK;
}

// clears what a counting pass accumulates (solid cursor, distinct counter, histograms, split/overflow flags, the spans of
// the counting stage) so that the stage can run again from the partitioned records still in HBM
static int reset_count_state(dskgpu_ctx* ctx)
{
    Counters* ctr = (Counters*)ctx->ctr.p;
    CK(cudaMemsetAsync(&ctr->solid_n, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(&ctr->distinct_n, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(&ctr->smem_splits, 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(&ctr->overflow, 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(&ctr->sort_fallback, 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(ctx->hist.p, 0, sizeof(unsigned long long) * DSKGPU_HISTO_LEN, ctx->stream));
    CK(cudaMemsetAsync(ctx->hist2d.p, 0, sizeof(unsigned long long) * DSKGPU_HISTO_LEN * DSKGPU_HISTO2D_DIM2, ctx->stream));
    if (ctx->bank_hist.p) CK(cudaMemsetAsync(ctx->bank_hist.p, 0, sizeof(unsigned long long) * DSKGPU_HISTO_LEN * (size_t)ctx->NB, ctx->stream));
    ctx->n_solid = 0; ctx->results_on_host = false;
    ctx->st.nb_groups_hash = ctx->st.nb_groups_sort = 0; ctx->st.nb_groups_bucket = 0; ctx->st.nb_hash_regroups = 0;
    // timing spans of the counting stage belong to the pass that produced the results
    std::vector<dskgpu_ctx::Span> keep;
    for (auto& sp : ctx->spans) if (sp.kind != SPAN_COUNT && sp.kind != SPAN_SORT && sp.kind != SPAN_DOM && sp.kind != SPAN_SORTPASS && sp.kind != SPAN_HEAVY) keep.push_back(sp);
    ctx->spans.swap(keep);
    return 0;
}

template <int KW>
static int stage_count(dskgpu_ctx* ctx, u64 cap_request = 0)
{
    u64 need = 0;
    int rc = stage_count_once<KW>(ctx, cap_request, &need);
    if (rc || !need) return rc;
    // the estimate was too small: once more at the exact size (the records are still in HBM)
    if ((rc = reset_count_state(ctx))) return rc;
    ctx->st.nb_solid_regrows++;
    rc = stage_count_once<KW>(ctx, need, &need);
    if (rc) return rc;
    if (need) FAIL(DSKGPU_ERR_OVERFLOW, "solid k-mer buffer overflow after the exact-size pass (internal)");
    return DSKGPU_OK;
}

template <int KW>
static int finish_single(dskgpu_ctx* ctx)
{
    int rc;
    trace(nullptr);
    if ((rc = stage_totals(ctx))) return rc;
    trace("totals (push kernels done)");
    if (g_trace) { const void* dh = nullptr; if ((rc = fold_local_hist(ctx, &dh))) return rc; CK(cudaStreamSynchronize(ctx->stream)); trace("bin histogram folded (sync: tracing only)"); }
    if ((rc = plan_device(ctx, nullptr))) return rc;
    trace("plan (device) + header");
    if ((rc = stage_scatter<KW>(ctx))) return rc;
    if ((rc = stage_bases(ctx, nullptr, nullptr))) return rc;
    trace("scatter launched");
    rc = stage_count<KW>(ctx);
    trace("finish done");
    return rc;
}

// multi-GPU: the partitions this rank owns are W segments each -- its own chunk in lrecs, the others' chunks in its receive buffer
template <int KW>
static int finish_owned(dskgpu_ctx* ctx)
{
    if (!ctx->xchg_scattered) FAIL(DSKGPU_ERR_STATE, "world_size > 1: run the dskgpu_xchg_* sequence before dskgpu_finish");
    return stage_count<KW>(ctx);
}

extern "C" {

int dskgpu_finish(dskgpu_ctx* ctx)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (ctx->state != 0) FAIL(DSKGPU_ERR_STATE, "finish called twice");
    if (ctx->cfg.world_size > 1) return KW_DISPATCH(ctx, finish_owned, ctx);
    return KW_DISPATCH(ctx, finish_single, ctx);
}

// ---- multi-GPU exchange ----------------------------------------------------------------------------------------------
int dskgpu_xchg_local_totals(dskgpu_ctx* ctx, uint64_t* kmers, uint64_t* records)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    int rc = stage_totals(ctx); if (rc) return rc;
    if (kmers) *kmers = ctx->local_nkm; if (records) *records = ctx->local_nrec;
    return DSKGPU_OK;
}

int dskgpu_xchg_prepare(dskgpu_ctx* ctx, uint64_t* local4)
{
    if (!ctx || !local4) return DSKGPU_ERR_ARG;
    use_device(ctx);
    int rc = stage_totals(ctx); if (rc) return rc;
    local4[0] = ctx->local_nkm; local4[1] = ctx->local_nrec; local4[2] = ctx->sample_nkm; local4[3] = ctx->sample_distinct;
    return DSKGPU_OK;
}

int dskgpu_xchg_sketch(dskgpu_ctx* ctx, uint32_t* sketch)
{
    if (!ctx || !sketch) return DSKGPU_ERR_ARG;
    use_device(ctx);
    int rc = stage_totals(ctx); if (rc) return rc;
    memcpy(sketch, ctx->h_hll, sizeof(u32) * HLL_M);
    return DSKGPU_OK;
}

int dskgpu_xchg_set_sketch(dskgpu_ctx* ctx, const uint32_t* merged)
{
    if (!ctx || !merged) return DSKGPU_ERR_ARG;
    if (ctx->global_set) FAIL(DSKGPU_ERR_STATE, "xchg_set_sketch after xchg_set_global");
    ctx->sketch_distinct = hll_estimate(merged);
    return DSKGPU_OK;
}

int dskgpu_xchg_set_global(dskgpu_ctx* ctx, const uint64_t* global4, int* log2_bins)
{
    if (!ctx || !global4) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (!ctx->totals_done) FAIL(DSKGPU_ERR_STATE, "xchg_set_global before xchg_prepare");
    set_global(ctx, global4[0], global4[1], global4[2], global4[3]);
    if (log2_bins) *log2_bins = ctx->bin_level;
    return DSKGPU_OK;
}

int dskgpu_xchg_hist(dskgpu_ctx* ctx, void* d_out)
{
    if (!ctx || !d_out) return DSKGPU_ERR_ARG;
    use_device(ctx);
    int rc = stage_totals(ctx); if (rc) return rc;
    if (!ctx->global_set) FAIL(DSKGPU_ERR_STATE, "xchg_hist before xchg_set_global (the ranks must agree on the bin level)");
    const void* src = nullptr;
    if ((rc = fold_local_hist(ctx, &src))) return rc;
    CK(cudaMemcpyAsync(d_out, src, sizeof(unsigned long long) * ((size_t)2 << ctx->bin_level), cudaMemcpyDeviceToDevice, ctx->stream));
    return DSKGPU_OK;
}

int dskgpu_xchg_plan(dskgpu_ctx* ctx, const void* d_global_hist, uint32_t* nparts, uint32_t* parts_per_rank, uint64_t* need_records)
{
    if (!ctx || !d_global_hist) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (!ctx->hist_fetched) FAIL(DSKGPU_ERR_STATE, "xchg_plan before xchg_hist");
    if (ctx->cfg.world_size > PLAN_MAXW) FAIL(DSKGPU_ERR_ARG, "world_size %d beyond the %d ranks of one exchange", ctx->cfg.world_size, PLAN_MAXW);
    if (!ctx->planned) {
        trace(nullptr);
        int rc = plan_device(ctx, d_global_hist); if (rc) return rc;
        trace("xchg_plan: planned on the device, header read");
    }
    if (nparts) *nparts = ctx->h_hdr->P;
    if (parts_per_rank) *parts_per_rank = ctx->h_hdr->PW;
    if (need_records) for (int r = 0; r < ctx->cfg.world_size; r++) need_records[r] = ctx->h_hdr->need_recs[r];
    return DSKGPU_OK;
}

int dskgpu_xchg_counts(dskgpu_ctx* ctx, void* d_out)
{
    if (!ctx || !d_out) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (!ctx->planned) FAIL(DSKGPU_ERR_STATE, "xchg_counts before xchg_plan");
    const u32 W = (u32)ctx->cfg.world_size;
    const u64 nq = (u64)W * ctx->h_hdr->PW, qcap = ((u64)1 << ctx->bin_level) + W;
    const u64* lcnt_q = (const u64*)ctx->gk_q.p + 2 * qcap;
    CK(cudaMemcpyAsync(d_out, lcnt_q, nq * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync((u64*)d_out + nq, ((const PlanHdr*)ctx->pl_hdr.p)->send_recs, (size_t)W * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return DSKGPU_OK;
}

int dskgpu_xchg_ensure_recv(dskgpu_ctx* ctx, uint64_t capacity_records)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (!ctx->planned) FAIL(DSKGPU_ERR_STATE, "xchg_ensure_recv before xchg_plan");
    const u64 want = std::max<u64>(capacity_records, ctx->h_hdr->need_recs[ctx->cfg.rank]);
    int rc = ensure(ctx, ctx->precs, want * (u64)ctx->RW * 8 + 64); if (rc) return rc;
    ctx->xchg_planned = true;
    return DSKGPU_OK;
}

int dskgpu_xchg_recv_buffer(dskgpu_ctx* ctx, void** d_recv, size_t* bytes)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (!ctx->xchg_planned) FAIL(DSKGPU_ERR_STATE, "xchg_recv_buffer before xchg_ensure_recv");
    if (d_recv) *d_recv = ctx->precs.p; if (bytes) *bytes = ctx->precs.cap;
    return DSKGPU_OK;
}

int dskgpu_xchg_ipc_handle(dskgpu_ctx* ctx, void* handle64)
{
    if (!ctx || !handle64) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (!ctx->xchg_planned) FAIL(DSKGPU_ERR_STATE, "xchg_ipc_handle before xchg_ensure_recv");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->precs.p));
    memcpy(handle64, &h, 64);
    return DSKGPU_OK;
}

int dskgpu_xchg_open_peer(dskgpu_ctx* ctx, const void* handle64, void** d_ptr)
{
    if (!ctx || !handle64 || !d_ptr) return DSKGPU_ERR_ARG;
    use_device(ctx);
    cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->ipc_opened.push_back(*d_ptr);
    return DSKGPU_OK;
}

int dskgpu_xchg_set_peers(dskgpu_ctx* ctx, void* const* d_peer_recv)
{
    if (!ctx || !d_peer_recv) return DSKGPU_ERR_ARG;
    use_device(ctx);
    ctx->peer_recv.assign(d_peer_recv, d_peer_recv + ctx->cfg.world_size);
    ctx->peer_recv[ctx->cfg.rank] = ctx->precs.p;
    return DSKGPU_OK;
}

// Local scatter into owner-major order, then this rank's chunk for every other rank crosses NVLink as ONE contiguous copy
// per peer (k_xchg_send: peer stores through the CUDA-IPC / peer-access pointers, no NCCL on the data path).  The rank's own
// chunk is not copied at all: the counting kernel reads it where the scatter put it.
int dskgpu_xchg_scatter(dskgpu_ctx* ctx, const void* d_recv_counts, const void* d_send_matrix)
{
    if (!ctx || !d_recv_counts || !d_send_matrix) return DSKGPU_ERR_ARG;
    use_device(ctx);
    const u32 W = (u32)ctx->cfg.world_size, me = (u32)ctx->cfg.rank;
    if (W < 2) FAIL(DSKGPU_ERR_STATE, "xchg_scatter needs world_size > 1");
    if (!ctx->xchg_planned || ctx->peer_recv.size() != (size_t)W) FAIL(DSKGPU_ERR_STATE, "xchg_scatter before xchg_ensure_recv / xchg_set_peers");
    if (ctx->precs.cap < ctx->h_hdr->need_recs[me] * (u64)ctx->RW * 8) FAIL(DSKGPU_ERR_STATE, "receive buffer smaller than the planned layout");
    int rc = KW_DISPATCH(ctx, stage_scatter, ctx);
    if (rc) return rc;
    if ((rc = stage_bases(ctx, d_recv_counts, d_send_matrix))) return rc;
    u64 remote = 0;
    for (u32 o = 0; o < W; o++) if (o != me) remote += ctx->h_hdr->send_recs[o];
    if (remote) {
        SpanGuard g(ctx, SPAN_PART);
        const unsigned per_peer = (unsigned)std::max<u64>(1, std::min<u64>((remote / (W - 1) * (u64)(ctx->RW / 2) + 1023) / 1024, (u64)ctx->num_sms * 8 / (W - 1)));
        cudaEvent_t xa = get_event(ctx), xb = get_event(ctx);
        cudaEventRecord(xa, ctx->stream);
        k_xchg_send<<<per_peer * (W - 1), 256, 0, ctx->stream>>>((const ulonglong2*)ctx->qrecs, (const XchgTab*)ctx->xtab.p, W, me, (u32)ctx->RW / 2); LAUNCHED();
        cudaEventRecord(xb, ctx->stream);
        ctx->spans.push_back({xa, xb, SPAN_XCHG});
        CK(cudaGetLastError());
    }
    ctx->xchg_bytes_out = remote * (u64)ctx->RW * 8;                 // bytes stored into OTHER ranks' HBM (NVLink)
    ctx->xchg_scattered = true;
    return DSKGPU_OK;
}

int dskgpu_xchg_close_peer(dskgpu_ctx* ctx, void* d_ptr)
{
    if (!ctx || !d_ptr) return DSKGPU_ERR_ARG;
    use_device(ctx);
    for (size_t i = 0; i < ctx->ipc_opened.size(); i++)
        if (ctx->ipc_opened[i] == d_ptr) {
            CK(cudaStreamSynchronize(ctx->stream));                    // nothing of ours may still be storing through the mapping
            CK(cudaIpcCloseMemHandle(d_ptr));
            ctx->ipc_opened.erase(ctx->ipc_opened.begin() + (long)i);
            return DSKGPU_OK;
        }
    FAIL(DSKGPU_ERR_ARG, "xchg_close_peer: pointer was not opened by this context");
}

int dskgpu_set_pass(dskgpu_ctx* ctx, int pass_id, int nb_passes)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    if (nb_passes < 1 || pass_id < 0 || pass_id >= nb_passes) FAIL(DSKGPU_ERR_ARG, "pass_id %d outside [0, nb_passes = %d)", pass_id, nb_passes);
    if (ctx->state != 0 || ctx->bytes_pushed != 0) FAIL(DSKGPU_ERR_STATE, "set_pass on a context that holds data (reset it first)");
    ctx->nb_passes = nb_passes; ctx->pass_id = pass_id;
    return DSKGPU_OK;
}

int dskgpu_push_sync(dskgpu_ctx* ctx)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    CK(cudaStreamSynchronize(ctx->copy_stream));
    return DSKGPU_OK;
}

// Sizing rule behind the pass count (the reference derives nb_passes from the estimated volume against -max-disk,
// K/ConfigurationAlgorithm.cpp:245-467; here the bound is HBM).  Per k-mer of one pass on one GPU:
//   records, input order + partition order (+ the receive buffer when the job is sharded) ... (2 | 3) x record bytes / s
//   record meta (bin | nk) ............................................................... 4 / s
//   solid-set ping-pong buffers at the default estimate (8 % of the k-mers + margin) ...... 2 x 0.10 x (8 KW + 4)
// with s = k-mers per record (~ 9 at m = 12..14, the lengths multi-G k-mer jobs get), plus 2 GiB of fixed buffers.
int dskgpu_suggest_nb_passes(uint64_t expected_kmers, int kmer_size, int world_size, uint64_t hbm_bytes, int device)
{
    if (world_size < 1) world_size = 1;
    if (hbm_bytes == 0) {
        size_t fr = 0, tot = 0;
        int cur = -1; cudaGetDevice(&cur);
        if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
        if (cur >= 0 && cur != device) cudaSetDevice(cur);
        hbm_bytes = fr;
    }
    const int KW = kw_of(kmer_size);
    const double s = kmer_size < 32 ? 9.0 : 18.0;
    const double per_kmer = ((world_size > 1 ? 3.0 : 2.0) * 16.0 * KW + 4.0) / s + 2.0 * 0.10 * (8.0 * KW + 4.0);
    const double fixed = 2.0 * 1024 * 1024 * 1024;
    const double budget = (double)hbm_bytes * 0.90 - fixed;
    if (budget <= 0) return 1 << 10;
    const double need = (double)expected_kmers / world_size * per_kmer;
    int passes = (int)(need / budget) + 1;
    return passes < 1 ? 1 : passes;
}

int dskgpu_xchg_sync(dskgpu_ctx* ctx)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    return DSKGPU_OK;
}

// ---- several contexts (ranks) in one process: the exchange protocol with host-side sums in place of the collectives ------
}  // extern "C"
#include <thread>
extern "C" {

int dskgpu_multi_finish(dskgpu_ctx* const* ctxs, int n)
{
    dskgpu_ctx* ctx = (ctxs && n > 0) ? ctxs[0] : nullptr;
    if (!ctx) return DSKGPU_ERR_ARG;
    if (n > PLAN_MAXW) FAIL(DSKGPU_ERR_ARG, "multi_finish: %d contexts beyond the %d ranks of one exchange", n, PLAN_MAXW);
    for (int r = 0; r < n; r++) {
        if (!ctxs[r]) return DSKGPU_ERR_ARG;
        if (ctxs[r]->cfg.world_size != n || ctxs[r]->cfg.rank != r) FAIL(DSKGPU_ERR_ARG, "multi_finish: ctxs[%d] must have rank %d of world_size %d", r, r, n);
        if (ctxs[r]->KW != ctx->KW || ctxs[r]->k != ctx->k || ctxs[r]->m != ctx->m) FAIL(DSKGPU_ERR_ARG, "multi_finish: contexts differ in k / minimizer size");
    }
    if (n == 1) return dskgpu_finish(ctx);
    int rc;
    // peer access between the distinct devices (both directions; contexts sharing a device need nothing)
    for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) {
        const int da = ctxs[a]->cfg.device, db = ctxs[b]->cfg.device;
        if (da == db) continue;
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, da, db));
        if (!can) FAIL(DSKGPU_ERR_CUDA, "multi_finish: device %d cannot access device %d (no P2P path)", da, db);
        CK(cudaSetDevice(da));
        cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError(); else CK(e);
    }
    auto sync_all = [&]() -> int { for (int r = 0; r < n; r++) { ctx = ctxs[r]; use_device(ctx); CK(cudaStreamSynchronize(ctx->stream)); } ctx = ctxs[0]; return 0; };
    // 1. job totals (k-mers, records, density sample) -> the same bin level and partition size everywhere
    uint64_t g4[4] = {0, 0, 0, 0};
    for (int r = 0; r < n; r++) { uint64_t l4[4]; if ((rc = dskgpu_xchg_prepare(ctxs[r], l4))) return rc; for (int i = 0; i < 4; i++) g4[i] += l4[i]; }
    {
        std::vector<uint32_t> merged(HLL_M, 0), one(HLL_M);
        for (int r = 0; r < n; r++) { if ((rc = dskgpu_xchg_sketch(ctxs[r], one.data()))) return rc; for (u32 i = 0; i < HLL_M; i++) merged[i] = std::max(merged[i], one[i]); }
        for (int r = 0; r < n; r++) if ((rc = dskgpu_xchg_set_sketch(ctxs[r], merged.data()))) return rc;
    }
    int level = 0;
    for (int r = 0; r < n; r++) { int lv = 0; if ((rc = dskgpu_xchg_set_global(ctxs[r], g4, &lv))) return rc; if (r && lv != level) FAIL(DSKGPU_ERR_STATE, "multi_finish: ranks disagree on the bin level"); level = lv; }
    // 2. whole-job bin histogram: every device sums the ranks' histograms through peer pointers (the stand-in for the all-reduce)
    const u64 nh = (u64)2 << level;
    PtrList lst; memset(&lst, 0, sizeof lst);
    for (int r = 0; r < n; r++) { ctx = ctxs[r]; use_device(ctx); const void* p = nullptr; if ((rc = fold_local_hist(ctx, &p))) return rc; lst.p[r] = (const u64*)p; }
    if ((rc = sync_all())) return rc;
    for (int r = 0; r < n; r++) {
        ctx = ctxs[r]; use_device(ctx);
        if ((rc = ensure(ctx, ctx->ghist, nh * 8))) return rc;
        k_sum_hists<<<(unsigned)std::min<u64>((nh + 255) / 256, 148 * 8), 256, 0, ctx->stream>>>(lst, n, nh, (u64*)ctx->ghist.p); LAUNCHED();
        CK(cudaGetLastError());
    }
    // 3. every rank plans the same partitions (on its device)
    for (int r = 0; r < n; r++) { ctx = ctxs[r]; use_device(ctx); if ((rc = plan_device(ctx, ctx->ghist.p))) return rc; }
    ctx = ctxs[0];
    const u32 P = ctx->h_hdr->P, PW = ctx->h_hdr->PW;
    for (int r = 1; r < n; r++) if (ctxs[r]->h_hdr->P != P || ctxs[r]->h_hdr->nlight != ctx->h_hdr->nlight) FAIL(DSKGPU_ERR_STATE, "multi_finish: ranks disagree on the plan");
    // 4. per-partition counts to the owners (the stand-in for the all-to-all), chunk sizes to everybody, receive buffers
    std::vector<u64> S((size_t)n * n);
    for (int s = 0; s < n; s++) for (int o = 0; o < n; o++) S[(size_t)s * n + o] = ctxs[s]->h_hdr->send_recs[o];
    std::vector<void*> recv(n, nullptr);
    const u64 qcap = ((u64)1 << level) + (u64)n;
    for (int o = 0; o < n; o++) {
        ctx = ctxs[o]; use_device(ctx);
        if ((rc = ensure(ctx, ctx->ghist, std::max<u64>(nh, (u64)n * PW) * 8))) return rc;          // the histogram is dead: reuse as [W][PW] count rows
        if ((rc = ensure(ctx, ctx->xS, (size_t)n * n * 8))) return rc;
        for (int s = 0; s < n; s++)
            CK(cudaMemcpyAsync((u64*)ctx->ghist.p + (u64)s * PW, (const u64*)ctxs[s]->gk_q.p + 2 * qcap + (u64)o * PW, (size_t)PW * 8, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(ctx->xS.p, S.data(), S.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));                                                       // S is pageable host memory
        if ((rc = dskgpu_xchg_ensure_recv(ctx, 0))) return rc;
        recv[o] = ctx->precs.p;
    }
    // 5. scatter + one contiguous copy per peer, straight into the owners' HBM (same process: plain device pointers)
    for (int r = 0; r < n; r++) { if ((rc = dskgpu_xchg_set_peers(ctxs[r], recv.data()))) return rc; if ((rc = dskgpu_xchg_scatter(ctxs[r], ctxs[r]->ghist.p, ctxs[r]->xS.p))) return rc; }
    if ((rc = sync_all())) return rc;                                                                 // every record has landed before anyone counts
    // 6. every rank counts what it owns (the stage blocks on its own stream: one host thread per rank)
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < n; r++) th.emplace_back([&, r] { rcs[r] = dskgpu_finish(ctxs[r]); });
    for (auto& t : th) t.join();
    for (int r = 0; r < n; r++) if (rcs[r]) { ctx = ctxs[r]; return rcs[r]; }
    return DSKGPU_OK;
}

// host-only mirror of the receive layout of rank `owner` (CPU test-suite): counts[s * W * PW + q] = records rank s holds for
// the partition at position q of the q-ordered tables.  The receive buffer holds the chunks of the senders s != owner in
// rank order (the owner's own chunk never moves): region_base[s] = first record of sender s's chunk (region_base[W] = records
// received); seg_off[s * PW + j] = first record of (sender s, owned job j) -- inside the receive buffer for s != owner,
// inside the owner's own chunk for s == owner.
int dskgpu_xchg_layout(int world_size, uint32_t parts_per_rank, const uint64_t* counts, int owner, uint64_t* region_base, uint64_t* seg_off)
{
    if (world_size < 1 || world_size > PLAN_MAXW || !counts || !region_base || !seg_off || owner < 0 || owner >= world_size) return DSKGPU_ERR_ARG;
    const u32 W = (u32)world_size, PW = parts_per_rank;
    u64 base = 0;
    for (u32 s = 0; s < W; s++) {
        region_base[s] = base;
        u64 o = (s == (u32)owner) ? 0 : base;
        for (u32 j = 0; j < PW; j++) { seg_off[(u64)s * PW + j] = o; o += counts[(u64)s * W * PW + (u64)owner * PW + j]; }
        if (s != (u32)owner) base = o;
    }
    region_base[W] = base;
    return DSKGPU_OK;
}

// second pass of "-abundance-min auto": the partitioned records are still in HBM, only the counting stage runs again
int dskgpu_recount(dskgpu_ctx* ctx, const int64_t* abundance_min)
{
    if (!ctx || !abundance_min) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (ctx->state != 1) FAIL(DSKGPU_ERR_STATE, "recount before finish");
    for (int b = 0; b < DSKGPU_MAX_BANKS; b++) ctx->cfg.abundance_min[b] = abundance_min[b < ctx->cfg.nb_banks ? b : ctx->cfg.nb_banks - 1];
    // the solid set is bounded by the distinct k-mers the first pass found (the first pass of -abundance-min auto dumps nothing)
    const u64 cap_request = ctx->st.kmers_nb_distinct + 1024;
    int rc = reset_count_state(ctx); if (rc) return rc;
    ctx->state = 0;
    rc = KW_DISPATCH(ctx, stage_count, ctx, cap_request);
    if (rc) ctx->state = 2;                                        // failed: only reset / destroy are valid now
    return rc;
}

int dskgpu_bank_histograms(dskgpu_ctx* ctx, uint64_t* hist)
{
    if (!ctx || !hist) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (ctx->state != 1) FAIL(DSKGPU_ERR_STATE, "histograms requested before finish");
    if (ctx->NB == 1) { memcpy(hist, ctx->h_hist, sizeof(uint64_t) * DSKGPU_HISTO_LEN); return DSKGPU_OK; }
    if (!ctx->bank_hist.p) FAIL(DSKGPU_ERR_STATE, "per-bank histograms were not requested (cfg.bank_histograms)");
    CK(cudaMemcpyAsync(hist, ctx->bank_hist.p, sizeof(uint64_t) * DSKGPU_HISTO_LEN * (size_t)ctx->NB, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSKGPU_OK;
}

int dskgpu_num_partitions(dskgpu_ctx* ctx) { if (!ctx || ctx->state != 1) return DSKGPU_ERR_STATE; return 1; }

int dskgpu_partition(dskgpu_ctx* ctx, int p, const uint64_t** kmers, const uint32_t** counts, uint64_t* n, int* words)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (ctx->state != 1) FAIL(DSKGPU_ERR_STATE, "results requested before finish");
    if (p != 0) FAIL(DSKGPU_ERR_ARG, "partition %d out of range", p);
    if (ctx->n_solid && !ctx->results_on_host) FAIL(DSKGPU_ERR_STATE, "results were kept on the device (keep_results_on_device)");
    if (kmers) *kmers = ctx->h_skeys; if (counts) *counts = ctx->h_svals; if (n) *n = ctx->n_solid; if (words) *words = ctx->KW;
    return DSKGPU_OK;
}

int dskgpu_partition_device(dskgpu_ctx* ctx, int p, const void** d_kmers, const void** d_counts, uint64_t* n, int* words)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (ctx->state != 1) FAIL(DSKGPU_ERR_STATE, "results requested before finish");
    if (p != 0) FAIL(DSKGPU_ERR_ARG, "partition %d out of range", p);
    if (d_kmers) *d_kmers = ctx->skeys[ctx->solid_buf].p; if (d_counts) *d_counts = ctx->svals[ctx->solid_buf].p;
    if (n) *n = ctx->n_solid; if (words) *words = ctx->KW;
    return DSKGPU_OK;
}

int dskgpu_histogram(dskgpu_ctx* ctx, uint64_t* hist1d, uint64_t* hist2d)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (ctx->state != 1) FAIL(DSKGPU_ERR_STATE, "histogram requested before finish");
    if (hist1d) memcpy(hist1d, ctx->h_hist, sizeof(uint64_t) * DSKGPU_HISTO_LEN);
    if (hist2d) memcpy(hist2d, ctx->h_hist + DSKGPU_HISTO_LEN, sizeof(uint64_t) * DSKGPU_HISTO_LEN * DSKGPU_HISTO2D_DIM2);
    return DSKGPU_OK;
}

int dskgpu_get_stats(dskgpu_ctx* ctx, dskgpu_stats* out)
{
    if (!ctx || !out) return DSKGPU_ERR_ARG;
    *out = ctx->st;
    return DSKGPU_OK;
}

int dskgpu_record_bytes(dskgpu_ctx* ctx) { return ctx ? ctx->RW * 8 : DSKGPU_ERR_ARG; }

// ---------------------------------------------------------------------------------------------------------------
// host self checks (same functions the kernels run; no GPU needed)
// ---------------------------------------------------------------------------------------------------------------
}  // extern "C"

template <int FMT>
static int64_t selftest_scan_fmt(const u8* raw, u64 lo, u64 hi, uint8_t* out, size_t out_cap)
{
    // replays the device algorithm chunk by chunk (masks -> table -> emission masks) and cross-checks every
    // chunk against the byte-wise reference state machine scan_step(); tables are also composed per tile.
    int state = (FMT == FMT_FASTA) ? ST_HDR : 0;          // true state (mask path)
    int state_ref = state;                                  // byte-wise path
    size_t w = 0; u32 err = 0; int err_ref = 0;
    const u64 c0 = lo / SCAN_BPT, c1 = (hi + SCAN_BPT - 1) / SCAN_BPT;
    Tab tile_tab = tab_identity(); int tile_state0 = state; u32 tile_cnt = 0; u64 in_tile = 0;
    for (u64 ci = c0; ci < c1; ci++) {
        const u64 a = ci * SCAN_BPT;
        u32 words[8] = {0, 0, 0, 0, 0, 0, 0, 0}, active = 0;
        for (int i = 0; i < SCAN_BPT; i++) { u64 x = a + i; bool in = x >= lo && x < hi; u32 c = in ? raw[x] : 0u; words[i >> 2] |= c << (8 * (i & 3)); active |= (in ? 1u : 0u) << i; }
        const int prev = (a > lo && a <= hi) ? raw[a - 1] : '\n';
        const u64 last_end = (a + SCAN_BPT < hi) ? a + SCAN_BPT : hi;
        const int next = (last_end < hi) ? (int)raw[last_end] : -1;
        const bool prev_nl = prev == '\n', next_flag = (next == '\n') || next < 0;
        CMasks m; chunk_masks(words, active, prev_nl, m);
        const Tab t = chunk_table<FMT>(m, prev_nl, next_flag);
        u32 em, sep, e2;
        chunk_emit_masks<FMT>(m, prev_nl, next_flag, state, em, sep, e2);
        err |= e2;
        // byte-wise reference
        u8 ref[SCAN_BPT]; int nref = 0; int pv = prev;
        for (int i = 0; i < SCAN_BPT; i++) {
            if (!((active >> i) & 1)) continue;
            const int ch = (words[i >> 2] >> (8 * (i & 3))) & 0xFF;
            const int nx = (i + 1 < SCAN_BPT && ((active >> (i + 1)) & 1)) ? (int)((words[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xFF) : next;
            const int e = scan_step(FMT, state_ref, pv, ch, nx, err_ref);
            if (e >= 0) ref[nref++] = (u8)e;
            pv = ch;
        }
        // mask path emission, through the same word-level compaction the kernel runs (word_codes: SWAR encode + byte permute)
        int n = 0;
        for (int wi = 0; wi < 8; wi++) {
            const u32 nib = (em >> (4 * wi)) & 0xFu;
            if (!nib) continue;
            const u32 packed = word_codes(words[wi], nib, (sep >> (4 * wi)) & 0xFu);
            if (popc32(nib) < 4 && (packed >> (8 * popc32(nib)))) return -7000000 - (int64_t)ci;   // kernel ORs the words together
            for (int j = 0; j < popc32(nib); j++) {
                const u8 code = (u8)(packed >> (8 * j));
                if (n >= nref || ref[n] != code) return -1000000 - (int64_t)ci;       // emission differs from scan_step
                if (w < out_cap) out[w] = code;
                w++; n++;
            }
        }
        if (n != nref) return -2000000 - (int64_t)ci;
        if ((u32)n != tab_count(t, state)) return -3000000 - (int64_t)ci;          // table count differs
        state = tab_state(t, state);
        if (state != state_ref) return -4000000 - (int64_t)ci;                      // table state differs
        // tile-level composition, as the block scan does
        tile_tab = tab_compose<FMT>(tile_tab, t); tile_cnt += (u32)n; in_tile++;
        if (in_tile == SCAN_THREADS || ci + 1 == c1) {
            if (tab_count(tile_tab, tile_state0) != tile_cnt || tab_state(tile_tab, tile_state0) != state) return -5000000 - (int64_t)ci;
            tile_tab = tab_identity(); tile_state0 = state; tile_cnt = 0; in_tile = 0;
        }
    }
    if (err || err_ref) return ((err != 0) == (err_ref != 0)) ? -(int64_t)(err | (u32)err_ref) : -6000000;
    return (int64_t)w;
}

extern "C" {

int64_t dskgpu_selftest_scan(const char* bytes, size_t n, int format, uint8_t* out, size_t out_cap)
{
    size_t skip = 0; int fmt = format;
    if (fmt != DSKGPU_FMT_LINES) { int det = detect_format(bytes, n, &skip); if (!det) return 0; if (fmt == DSKGPU_FMT_AUTO) fmt = det; }
    const u8* raw = (const u8*)bytes;
    switch (fmt) {
    case FMT_FASTA: return selftest_scan_fmt<FMT_FASTA>(raw, skip, n, out, out_cap);
    case FMT_FASTQ: return selftest_scan_fmt<FMT_FASTQ>(raw, skip, n, out, out_cap);
    default:        return selftest_scan_fmt<FMT_LINES>(raw, skip, n, out, out_cap);
    }
}

int dskgpu_selftest_minimizers(const uint8_t* codes, size_t n, int k, int m, uint32_t* out_min, uint8_t* out_valid)
{
    if (n < (size_t)k) return 0;
    for (size_t p = 0; p + k <= n; p++) {
        bool ok = true; u32 best = 0xFFFFFFFFu;
        for (int i = 0; i < k; i++) if (codes[p + i] >> 2) ok = false;
        for (int j = 0; j + m <= k; j++) {
            u32 x = 0; for (int i = 0; i < m; i++) x = (x << 2) | (codes[p + j + i] & 3);
            u32 v = mmer_order(x, m); if (v < best) best = v;
        }
        out_min[p] = best; out_valid[p] = ok;
    }
    return (int)(n - k + 1);
}

// host model of K2 + K4: split a code stream into records exactly like the kernel (tile = whole stream),
// pack them, then expand them again to canonical k-mers
int64_t dskgpu_selftest_superkmers(const uint8_t* codes, size_t n, int k, int m, uint64_t* out_kmers, size_t cap, uint64_t* n_records)
{
    const int KW = k < 32 ? 1 : 2, RW = 2 * KW;
    const int maxS = rec_max_kmers(KW, k);
    if (n < (size_t)k) { if (n_records) *n_records = 0; return 0; }
    const size_t npos = n - k + 1;
    std::vector<u32> mn(npos); std::vector<u8> valid(npos);
    dskgpu_selftest_minimizers(codes, n, k, m, mn.data(), valid.data());
    size_t w = 0; u64 nrec = 0;
    size_t p = 0;
    while (p < npos) {
        if (!valid[p]) { p++; continue; }
        size_t q = p; while (q < npos && valid[q] && mn[q] == mn[p] && (q - p) < (size_t)maxS) q++;
        const int nk = (int)(q - p);
        // pack
        u64 r[4] = {0, 0, 0, 0};
        for (int i = 0; i < k - 1 + nk; i++) r[i >> 5] |= (u64)(codes[p + i] & 3) << (62 - 2 * (i & 31));
        r[RW - 1] = (r[RW - 1] & ~0xFFFFULL) | ((u64)nk << 8);
        // expand
        if (KW == 1) {
            Kmer<1> f = rec_first_kmer1(r, k), rc = kmer_revcomp(f, k);
            for (int j = 0; j < nk; j++) { if (j) kmer_roll(f, rc, rec_base<2>(r, k - 1 + j), k); Kmer<1> c = kmer_canonical(f, rc); if (w < cap) out_kmers[w] = c.w[0]; w++; }
        } else {
            Kmer<2> f = rec_first_kmer2(r, k), rc = kmer_revcomp(f, k);
            for (int j = 0; j < nk; j++) { if (j) kmer_roll(f, rc, rec_base<4>(r, k - 1 + j), k); Kmer<2> c = kmer_canonical(f, rc);
                if (w < cap) { out_kmers[2 * w] = c.w[0]; out_kmers[2 * w + 1] = c.w[1]; } w++; }
        }
        nrec++; p = q;
    }
    if (n_records) *n_records = nrec;
    return (int64_t)w;
}

// the same round trip with the N-word logic of kmer_wide.cuh (k <= 127: records of 2*KW words, KW = 1..4): split at
// minimizer changes / invalid windows / record capacity, pack MSB first with the [nk:8][bank:8] field in the last word,
// expand by extraction of the first k-mer + rolling.  out_kmers: 4 words per k-mer.
int64_t dskgpu_selftest_wide_superkmers(const uint8_t* codes, size_t n, int k, int m, uint64_t* out_kmers, size_t cap, uint64_t* n_records)
{
    if (k < 2 || k > 127) return DSKGPU_ERR_ARG;
    if (n < (size_t)k) { if (n_records) *n_records = 0; return 0; }
    const size_t npos = n - k + 1;
    std::vector<u32> mn(npos); std::vector<u8> valid(npos);
    dskgpu_selftest_minimizers(codes, n, k, m, mn.data(), valid.data());
    auto run = [&](auto tag) -> int64_t {
        constexpr int KW = decltype(tag)::value, RW = 2 * KW;
        const int maxS = rec_max_kmers(KW, k);
        size_t w = 0; u64 nrec = 0, p = 0;
        while (p < npos) {
            if (!valid[p]) { p++; continue; }
            size_t q = p; while (q < npos && valid[q] && mn[q] == mn[p] && (q - p) < (size_t)maxS) q++;
            const int nk = (int)(q - p);
            u64 r[RW];
            for (int i = 0; i < RW; i++) r[i] = 0;
            for (int i = 0; i < k - 1 + nk; i++) r[i >> 5] |= (u64)(codes[p + i] & 3) << (62 - 2 * (i & 31));
            if ((k - 1 + nk) > rec_capacity_bases(KW)) return -2;                     // would overwrite the nk/bank field
            r[RW - 1] = (r[RW - 1] & ~0xFFFFULL) | ((u64)nk << 8);
            Kmer<KW> f = recn_kmer_at<KW, RW>(r, 0, k), rc = kmern_revcomp<KW>(f, k);
            for (int j = 0; j < nk; j++) {
                if (j) kmern_roll<KW>(f, rc, rec_base<RW>(r, k - 1 + j), k);
                const Kmer<KW> c = kmern_canonical<KW>(f, rc);
                if (w < cap) for (int x = 0; x < 4; x++) out_kmers[4 * w + x] = x < KW ? c.w[x] : 0;
                w++;
            }
            nrec++; p = q;
        }
        if (n_records) *n_records = nrec;
        return (int64_t)w;
    };
    if (k < 32) return run(std::integral_constant<int, 1>());
    if (k < 64) return run(std::integral_constant<int, 2>());
    if (k < 96) return run(std::integral_constant<int, 3>());
    return run(std::integral_constant<int, 4>());
}

// host-only run of the partition planner (plan_host, the sequential mirror of the device planner of plan.cuh): what every
// rank derives from the all-reduced bin histogram.  global_hist / local_hist: [2 << level] (records per bin, then k-mers per
// bin).  Outputs: bin2part[1 << level] (partition ids: light partitions first, heavy ones after), and per partition (capacity
// max_parts) the whole-job k-mers and this rank's records.  Returns the number of partitions, or -1 when max_parts is too
// small.  No device is touched.
int64_t dskgpu_selftest_plan(int level, const uint64_t* global_hist, const uint64_t* local_hist, int world_size, int nb_counts,
                             uint32_t smem_slots, double density, int count_mode, int forced_nb_partitions,
                             uint32_t* bin2part, uint64_t* part_kmers, uint64_t* part_local_recs, size_t max_parts)
{
    if (level < NBINS_LOG2 || level > NBINS_FINE_LOG2_MAX || !global_hist || !local_hist || world_size < 1 || world_size > PLAN_MAXW) return DSKGPU_ERR_ARG;
    dskgpu_ctx* ctx = new dskgpu_ctx();
    dskgpu_config_default(&ctx->cfg);
    ctx->cfg.world_size = world_size; ctx->cfg.count_mode = count_mode; ctx->cfg.nb_partitions = forced_nb_partitions;
    ctx->NB = nb_counts; ctx->smem_cap = smem_slots; ctx->density = density; ctx->density_known = true; ctx->bin_level = level;
    const u32 nb = 1u << level;
    u64 total = 0;
    for (u32 b = 0; b < nb; b++) total += global_hist[nb + b];
    ctx->g_total_kmers = total;
    const PlanParams pp = make_plan_params(ctx);
    const u64 qcap = (u64)nb + world_size;
    std::vector<u32> b2q(nb);
    std::vector<u64> gk(qcap), gr(qcap), lc(qcap);
    const PlanHdr h = plan_host(pp, (const u64*)global_hist, (const u64*)local_hist, b2q.data(), gk.data(), gr.data(), lc.data());
    delete ctx;
    if (h.P > max_parts) return -1;
    auto q2p = [&](u32 q) { return (q % h.PW) * (u32)world_size + q / h.PW; };
    for (u32 b = 0; b < nb; b++) bin2part[b] = q2p(b2q[b]);
    for (u32 p = 0; p < h.P; p++) { const u32 q = plan_q(p, (u32)world_size, h.PW); part_kmers[p] = gk[q]; part_local_recs[p] = lc[q]; }
    return (int64_t)h.P;
}

// the plan a context derived on the DEVICE (after dskgpu_finish / dskgpu_xchg_plan), copied to the host for the tests that
// compare it with the host mirror: bin2part[1 << level] (partition ids), per partition whole-job k-mers / records and this
// rank's records.  Returns the number of partitions (capacity max_parts), the level in *level.
int64_t dskgpu_debug_plan(dskgpu_ctx* ctx, int* level, uint32_t* bin2part, uint64_t* part_kmers, uint64_t* part_recs, uint64_t* part_local_recs, size_t max_parts)
{
    if (!ctx) return DSKGPU_ERR_ARG;
    use_device(ctx);
    if (!ctx->planned) FAIL(DSKGPU_ERR_STATE, "debug_plan before the plan exists");
    const PlanHdr& h = *ctx->h_hdr;
    const u32 W = (u32)ctx->cfg.world_size, nb = 1u << ctx->bin_level;
    if (level) *level = ctx->bin_level;
    if (h.P > max_parts) return -1;
    const u64 qcap = (u64)nb + W, nq = (u64)W * h.PW;
    std::vector<u32> b2q(nb); std::vector<u64> t(3 * qcap);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(b2q.data(), ctx->bin2part.p, (size_t)nb * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(t.data(), ctx->gk_q.p, 3 * qcap * 8, cudaMemcpyDeviceToHost));
    (void)nq;
    auto q2p = [&](u32 q) { return (q % h.PW) * W + q / h.PW; };
    if (bin2part) for (u32 b = 0; b < nb; b++) bin2part[b] = q2p(b2q[b]);
    for (u32 p = 0; p < h.P; p++) {
        const u32 q = plan_q(p, W, h.PW);
        if (part_kmers) part_kmers[p] = t[q]; if (part_recs) part_recs[p] = t[qcap + q]; if (part_local_recs) part_local_recs[p] = t[2 * qcap + q];
    }
    return (int64_t)h.P;
}

// host model of the wide spans (kmer_wide.cuh, k <= 127): canonical k-mers of a code stream computed two ways -- rolling
// (kmern_roll) and by extraction from a packed record (recn_kmer_at + kmern_revcomp) -- which must agree; returns the number
// of windows, or -1 - p at the first window p where the two disagree.  out_words: [n-k+1][4], out_valid: [n-k+1].
int64_t dskgpu_selftest_wide_kmers(const uint8_t* codes, size_t n, int k, uint64_t* out_words, uint8_t* out_valid)
{
    if (k < 2 || k > 127 || n < (size_t)k) return 0;
    const size_t npos = n - k + 1;
    auto run = [&](auto tag) -> int64_t {
        constexpr int KW = decltype(tag)::value, RW = 2 * KW;
        const int cap_bases = 32 * RW - 8;                                   // bases a record of RW words can hold
        Kmer<KW> f, rc;
        for (int i = 0; i < KW; i++) { f.w[i] = 0; rc.w[i] = 0; }
        for (size_t i = 0; i < n; i++) {
            kmern_roll<KW>(f, rc, codes[i] & 3, k);
            if (i + 1 < (size_t)k) continue;
            const size_t p = i + 1 - k;
            bool ok = true;
            for (int j = 0; j < k; j++) if (codes[p + j] >> 2) ok = false;
            const Kmer<KW> c = kmern_canonical<KW>(f, rc);
            for (int q = 0; q < 4; q++) out_words[4 * p + q] = q < KW ? c.w[q] : 0;
            out_valid[p] = ok;
            // second way: a record that starts a few bases before p (so that the extraction offset varies)
            const size_t back = std::min<size_t>(p, (size_t)(p % 7));
            const size_t r0 = p - back;
            const size_t nb = std::min<size_t>((size_t)cap_bases, n - r0);
            if (back + (size_t)k > nb) continue;
            u64 r[RW];
            for (int q = 0; q < RW; q++) r[q] = 0;
            for (size_t b = 0; b < nb; b++) r[b >> 5] |= (u64)(codes[r0 + b] & 3) << (62 - 2 * (b & 31));
            r[RW - 1] &= ~0xFFFFULL;                                         // the [nk:8][bank:8] field
            const Kmer<KW> g = recn_kmer_at<KW, RW>(r, (int)back, k);
            const Kmer<KW> c2 = kmern_canonical<KW>(g, kmern_revcomp<KW>(g, k));
            if (!kmern_eq<KW>(g, f) || !kmern_eq<KW>(c2, c)) return -1 - (int64_t)p;
        }
        return (int64_t)npos;
    };
    if (k < 32) return run(std::integral_constant<int, 1>());
    if (k < 64) return run(std::integral_constant<int, 2>());
    if (k < 96) return run(std::integral_constant<int, 3>());
    return run(std::integral_constant<int, 4>());
}

}  // extern "C"
