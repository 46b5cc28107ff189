// count_smem.cuh -- K4+K7+K6 for partitions that fit one SM: the whole count of a partition happens in the
// 227 KB of shared memory of one CTA.
//
//   reference                                                         here
//   PartitionsByHashCommand::execute (K/PartitionsCommand.cpp:372-739)  one persistent CTA per SM pulls partitions
//   Hash16::insert / OAHash::increment (Hash16.hpp:198-230,             ("jobs") off a queue; expands the super-k-mer
//     OAHash.hpp:92-99)                                                 records straight into an open-addressing table
//   CountProcessor chain (CountProcessorChain.hpp:128-169)              in shared memory (ATOMS.CAS.64/128 claim +
//                                                                       ATOMS.ADD count), sweeps it (histogram +
//                                                                       solidity + compaction), clears it, next job.
//
// Why: a random probe + RED + CAS mix on an L2-resident table runs at ~68 G k-mers/s on B200, the same mix on
// shared memory at ~580 G k-mers/s (tools/ubench.cu, profiles/r01e_ubench.txt).  HBM traffic of the counting stage
// becomes "read every record once, write the solid k-mers once".
//
// Work mapping of the insert phase: a warp takes 32 records; a work item is up to CS_Q consecutive k-mers of one
// record (first k-mer extracted from the packed bases, the others rolled), items are dealt to lanes through a
// prefix sum + binary search by shuffles, so lanes stay busy whatever the super-k-mer lengths are.  A k-mer that
// does not resolve in two probe rounds goes to a per-warp retry queue that is drained with all lanes active
// (otherwise one straggler lane holds the whole warp for every extra probe round).
//
// A partition whose distinct k-mers overflow the table is not an error: the CTA clears the table and redoes the
// partition as two sub-passes that each take half of the k-mer hash space (recursively, up to 2^CS_MAX_SPLIT).
//
// Record staging (round 2): the records of a job are contiguous in HBM, so ONE elected thread brings them into a
// shared-memory job buffer with a 1-D bulk copy (cp.async.bulk.shared::cluster.global + mbarrier complete_tx: the TMA
// unit, no LDG -> register -> STS traffic on the LSU the probes compete for), and the copy for the NEXT job is issued
// as soon as the last insert of the current one is over, so it lands under the sweep.  The next job id is claimed
// (global atomic) at the top of the current job.  With the whole job in shared memory the work items are dealt
// STATICALLY: a block scan over the records' item counts, then every warp takes an equal share of the items --
// the insert phase no longer ends with most warps waiting at the barrier for the warp that drew the last chunk
// (18.7 % of the stall samples of the round-1 kernel, profiles/r01y).
#pragma once
#include "kmer_bits.cuh"
#include "superk.cuh"
#include "count.cuh"
#include "plan.cuh"

namespace dsk {

#ifndef CS_THREADS_N
#define CS_THREADS_N 1024
#endif
#ifndef CS_CTAS_N
#define CS_CTAS_N 1
#endif
constexpr int CS_THREADS = CS_THREADS_N;       // one CTA per SM: the biggest table (measured: 2 x 512 threads with half tables is 25 % slower)
constexpr int CS_CTAS_PER_SM = CS_CTAS_N;
#ifndef CS_CHUNK
#define CS_CHUNK 32                            // records a warp takes at a time (<= 32)
#endif
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_MAXPROBE = 128;
constexpr int CS_MAX_SPLIT = 10;               // up to 1024 sub-passes before the job is reported as failed
constexpr int CS_Q = 4;                        // k-mers per work item
constexpr int CS_QCAP = 64;                    // retry queue entries per warp

// A job is one owned partition: job j of this rank is the partition at position qbase + j of the q-ordered tables
// (plan.cuh).  Its records are W segments, one per rank that parsed reads (segment s of job j = X[s * PW + j + 1] -
// X[s * PW + j] records at tab->segptr[s] + X[s * PW + j] * record bytes: the sender-major receive layout; W = 1 on one GPU).
// Nothing about the jobs is built on the host: the kernel reads the planner's device tables.
// The job starts as 2^split0 sub-passes, derived from the partition's whole-job k-mers (gk_q) and `fit`, the k-mers one pass
// takes at the sampled density -- a partition that cannot fit the table is never tried in one pass first.
struct CsSegs { const XchgTab* tab; const u64* X; const u64* gk_q; u32 W, PW, qbase; float fit; u32 max_split0; };
constexpr int CS_MAX_SPLIT0 = 4;

// bytes of dynamic shared memory for a table of `cap` slots (host + device agree through this one function):
// table keys + counts, record staging [warps][32][RW], retry queue [warps][QCAP] keys + slots
// nb = counts kept per slot (1, or one per bank when the processors need per-bank counts: -histo2D, solidity kinds)
// records the shared-memory job buffer holds (a job with more records is streamed through it in slices); sized so that
// the table and the buffer fill up together at the densities of 30-100x read sets (see stage_count / plan_target_kmers)
#ifndef CS_BUFREC1
#define CS_BUFREC1 2368
#endif
#ifndef CS_BUFREC2
#define CS_BUFREC2 544
#endif
template <int KW> DSK_HD u32 cs_bufrec() { return KW == 1 ? (u32)CS_BUFREC1 : (u32)CS_BUFREC2; }
template <int KW> DSK_HD size_t cs_smem_bytes(u32 cap, int nb = 1)
{
    return (size_t)cap * (8 * KW + 4 * nb)                          // table: keys + counts
         + (size_t)cs_bufrec<KW>() * (2 * KW * 8 + 4)               // job buffer: records + (item prefix | record index) of the records of the sub-pass
         + (size_t)CS_WARPS * CS_QCAP * (8 * KW + 4)                // retry queues
         + 16;                                                      // mbarrier
}
constexpr int CS_H2_I1 = H2_SMEM_I1;           // -histo2D: bins (i1 < 64, any i2) are accumulated in shared memory
constexpr int CS_MAX_BANKS = 4;                // per-bank counts beyond this go to the global-table path

#ifdef __CUDACC__

// 32-bit hash of a canonical k-mer: the high bits pick the home slot (multiply-shift range reduction), bits 8.. pick the
// sub-pass when a partition had to be split
__device__ __forceinline__ u32 cs_hash(const Kmer<1>& a)
{
    const u32 x = (u32)a.w[0] ^ ((u32)(a.w[0] >> 32) * 0x9E3779B1u);
    return (x ^ (x >> 15)) * 0x85EBCA77u;
}
__device__ __forceinline__ u32 cs_hash(const Kmer<2>& a)
{
    u32 x = (u32)a.w[0] ^ ((u32)(a.w[0] >> 32) * 0x9E3779B1u);
    x ^= ((u32)a.w[1] * 0xC2B2AE3Du) ^ ((u32)(a.w[1] >> 32) * 0x27D4EB2Fu);
    return (x ^ (x >> 15)) * 0x85EBCA77u;
}

template <int KW> __device__ __forceinline__ u32 cs_hash_of(const Kmer<KW>& a) { return cs_hash(a); }

// explicit shared-space accesses on 32-bit addresses (generic pointers make the compiler rebuild the window base at
// every access site)
__device__ __forceinline__ u32 cs_saddr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u64 cs_lds64v(u32 a) { u64 v; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ u64 cs_cas64(u32 a, u64 cmp, u64 val) { u64 o; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(o) : "r"(a), "l"(cmp), "l"(val) : "memory"); return o; }
__device__ __forceinline__ void cs_inc32(u32 a) { asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(a) : "memory"); }

// one probe of slot `slot` (keys_a = shared address of the key array): true when the slot holds (or now holds) `key`
__device__ __forceinline__ bool cs_probe(u32 keys_a, u32 slot, const Kmer<1>& key)
{
    const u64 EMPTY = ~0ULL;
    const u32 a = keys_a + slot * 8u;
    const u64 kk = cs_lds64v(a);
    if (kk == key.w[0]) return true;
    if (kk != EMPTY) return false;
    const u64 old = cs_cas64(a, EMPTY, key.w[0]);
    return old == EMPTY || old == key.w[0];
}
__device__ __forceinline__ bool cs_probe(u32 keys_a, u32 slot, const Kmer<2>& key)
{
    const u64 EMPTY = ~0ULL;
    const u32 a = keys_a + slot * 16u;
    ulonglong2 kk;
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(kk.x), "=l"(kk.y) : "r"(a) : "memory");
    if (kk.x == key.w[0] && kk.y == key.w[1]) return true;
    // a snapshot is trusted only when it shows a complete foreign key (no half is all-ones: final, slots only go
    // EMPTY -> key); everything else is decided by the CAS, which returns the truth
    if (kk.x != EMPTY && kk.y != EMPTY) return false;
    u64 olo, ohi;
    asm volatile("{\n\t.reg .b128 c, s, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\t"
                 "atom.shared.cas.b128 d, [%6], c, s;\n\tmov.b128 {%0, %1}, d;\n\t}"
                 : "=l"(olo), "=l"(ohi) : "l"(EMPTY), "l"(EMPTY), "l"(key.w[0]), "l"(key.w[1]), "r"(a) : "memory");
    return (olo == EMPTY && ohi == EMPTY) || (olo == key.w[0] && ohi == key.w[1]);
}

// the same probe split in two, so that the loads of several k-mers can be in flight together: cs_snap reads the slot,
// cs_resolve decides from the snapshot (a complete foreign key is final; anything else is decided by the CAS)
template <int KW> struct CsSnap { u64 w[KW]; };
__device__ __forceinline__ CsSnap<1> cs_snap1(u32 keys_a, u32 slot) { CsSnap<1> s; s.w[0] = cs_lds64v(keys_a + slot * 8u); return s; }
__device__ __forceinline__ CsSnap<2> cs_snap2(u32 keys_a, u32 slot)
{
    CsSnap<2> s;
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(s.w[0]), "=l"(s.w[1]) : "r"(keys_a + slot * 16u) : "memory");
    return s;
}
__device__ __forceinline__ bool cs_resolve(u32 keys_a, u32 slot, const Kmer<1>& key, const CsSnap<1>& sn)
{
    const u64 EMPTY = ~0ULL;
    if (sn.w[0] == key.w[0]) return true;
    if (sn.w[0] != EMPTY) return false;
    const u64 old = cs_cas64(keys_a + slot * 8u, EMPTY, key.w[0]);
    return old == EMPTY || old == key.w[0];
}
__device__ __forceinline__ bool cs_resolve(u32 keys_a, u32 slot, const Kmer<2>& key, const CsSnap<2>& sn)
{
    const u64 EMPTY = ~0ULL;
    if (sn.w[0] == key.w[0] && sn.w[1] == key.w[1]) return true;
    if (sn.w[0] != EMPTY && sn.w[1] != EMPTY) return false;
    u64 olo, ohi;
    asm volatile("{\n\t.reg .b128 c, s, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\t"
                 "atom.shared.cas.b128 d, [%6], c, s;\n\tmov.b128 {%0, %1}, d;\n\t}"
                 : "=l"(olo), "=l"(ohi) : "l"(EMPTY), "l"(EMPTY), "l"(key.w[0]), "l"(key.w[1]), "r"(keys_a + slot * 16u) : "memory");
    return (olo == EMPTY && ohi == EMPTY) || (olo == key.w[0] && ohi == key.w[1]);
}

// snapshot tests: a full match is final (slots only go EMPTY -> key); a complete foreign key is final too
__device__ __forceinline__ bool cs_match(const Kmer<1>& key, const CsSnap<1>& sn) { return sn.w[0] == key.w[0]; }
__device__ __forceinline__ bool cs_match(const Kmer<2>& key, const CsSnap<2>& sn) { return sn.w[0] == key.w[0] && sn.w[1] == key.w[1]; }
__device__ __forceinline__ bool cs_foreign(const Kmer<1>& key, const CsSnap<1>& sn) { return sn.w[0] != ~0ULL && sn.w[0] != key.w[0]; }
__device__ __forceinline__ bool cs_foreign(const Kmer<2>& key, const CsSnap<2>& sn)
{
    return sn.w[0] != ~0ULL && sn.w[1] != ~0ULL && !(sn.w[0] == key.w[0] && sn.w[1] == key.w[1]);
}

// predicated claim of an (apparently) empty slot: the CAS is issued under a predicate, not behind a branch, so the fast
// path of the insert stays straight-line.  Returns true when the slot now holds `key` (claimed, or claimed by an equal key).
__device__ __forceinline__ bool cs_claim_if(bool p, u32 keys_a, u32 slot, const Kmer<1>& key)
{
    const u64 EMPTY = ~0ULL;
    u64 o = 0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %4, 0;\n\t@q atom.shared.cas.b64 %0, [%1], %2, %3;\n\t}"
                 : "+l"(o) : "r"(keys_a + slot * 8u), "l"(EMPTY), "l"(key.w[0]), "r"((u32)p) : "memory");
    return p && (o == EMPTY || o == key.w[0]);
}
__device__ __forceinline__ bool cs_claim_if(bool p, u32 keys_a, u32 slot, const Kmer<2>& key)
{
    const u64 EMPTY = ~0ULL;
    u64 olo = 0, ohi = 0;
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b128 c, s, d;\n\tsetp.ne.u32 q, %7, 0;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\tmov.b128 d, {%0, %1};\n\t"
                 "@q atom.shared.cas.b128 d, [%6], c, s;\n\tmov.b128 {%0, %1}, d;\n\t}"
                 : "+l"(olo), "+l"(ohi) : "l"(EMPTY), "l"(EMPTY), "l"(key.w[0]), "l"(key.w[1]), "r"(keys_a + slot * 16u), "r"((u32)p) : "memory");
    return p && ((olo == EMPTY && ohi == EMPTY) || (olo == key.w[0] && ohi == key.w[1]));
}

// 32 bases of a record starting at base p (top bits first); RW words in registers, no dynamic indexing
template <int RW>
__device__ __forceinline__ u64 cs_window(const u64* r, int p)
{
    const int q = p >> 5, o = p & 31;
    u64 a, b;
    if constexpr (RW == 2) { a = q ? r[1] : r[0]; b = q ? 0ULL : r[1]; }
    else { a = q == 0 ? r[0] : q == 1 ? r[1] : q == 2 ? r[2] : r[3]; b = q == 0 ? r[1] : q == 1 ? r[2] : q == 2 ? r[3] : 0ULL; }
    return o ? ((a << (2 * o)) | (b >> (64 - 2 * o))) : a;
}

// ---- mbarrier + 1-D bulk copy (TMA unit) -------------------------------------------------------------------------------
__device__ __forceinline__ void cs_mbar_init(u32 mbar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// cs_mbar_expect arms the barrier with the byte count of ALL the copies of one slice (one arrival per phase); cs_bulk_copy
// issues one of them: global -> shared, completion by complete_tx on the mbarrier (16-byte aligned addresses, size a
// multiple of 16)
__device__ __forceinline__ void cs_mbar_expect(u32 mbar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cs_bulk_copy(u32 dst, u64 src_addr, u32 bytes, u32 mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src_addr), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cs_mbar_wait(u32 mbar, u32 parity)
{
    u32 ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}

// MB = false: one count per k-mer (banks summed), solidity = abundance range [amin, amax] -- the dsk default.
// MB = true : `nb` counts per slot (the record's bank byte picks the column); the sweep runs the whole CountProcessor
//             chain of count.cuh (`process_counts`: per-bank solidity kinds, -histo2D, per-bank histograms) on them.
// KEYS = false: a job is a partition of super-k-mer records (segs).
// KEYS = true : a job is a hash bucket of flat canonical k-mers written by k_expand_bucket (count.cuh): job j reads
//               min(bucket_n[j], slab) keys at recs + j * slab * KW; `segs` is unused.
template <int KW, bool MB, bool KEYS = false>
__global__ void __launch_bounds__(CS_THREADS, CS_CTAS_PER_SM) k_count_smem(const u64* __restrict__ recs, const CsSegs segs, u32 njobs,
                                                              int k, u32 cap, long long amin, long long amax,
                                                              u64* __restrict__ out_keys, u32* __restrict__ out_vals, u64 out_cap,
                                                              unsigned long long* __restrict__ g_hist, Counters* ctr, u32* work_counter,
                                                              int nb_arg, const SolidityParams spar, unsigned long long* __restrict__ g_hist2d,
                                                              const u32* __restrict__ bucket_n = nullptr, u32 slab = 0)
{
    constexpr int RW = 2 * KW;
    constexpr u32 BUFREC = KW == 1 ? (u32)CS_BUFREC1 : (u32)CS_BUFREC2;            // == cs_bufrec<KW>()
    constexpr u32 RPT = (BUFREC + CS_THREADS - 1) / CS_THREADS;                   // records per thread in the prefix scan
    const u32 nb = MB ? (u32)nb_arg : 1u;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    u64* s_keys = reinterpret_cast<u64*>(s_dyn);                                   // [cap][KW]
    u32* s_counts = reinterpret_cast<u32*>(s_dyn + (size_t)cap * 8 * KW);          // [cap][nb]
    u64* s_jrec = reinterpret_cast<u64*>(s_dyn + (size_t)cap * (8 * KW + 4 * nb)); // [BUFREC][RW]   (cap % 4 == 0 keeps it 16-byte aligned)
    u32* s_pref = reinterpret_cast<u32*>(s_jrec + (size_t)BUFREC * RW);            // [BUFREC] records of the sub-pass, compacted: exclusive item prefix | record << 16
    u64* s_qkey = reinterpret_cast<u64*>(s_pref + BUFREC);                         // [CS_WARPS][CS_QCAP][KW]   (BUFREC % 4 == 0)
    u32* s_qslot = reinterpret_cast<u32*>(s_qkey + (size_t)CS_WARPS * CS_QCAP * KW);   // [CS_WARPS][CS_QCAP]
    u64* s_mbar = reinterpret_cast<u64*>(s_qslot + (size_t)CS_WARPS * CS_QCAP);
    __shared__ u32 s_hist[HIST_SMEM_BINS];
    __shared__ u32 s_h2[MB ? 11 * CS_H2_I1 : 1];                                   // -histo2D bins with dim-1 index < CS_H2_I1, all 11 dim-2 rows
    __shared__ u32 s_wsum[CS_WARPS];
    __shared__ u32 s_job, s_ovf;                                                   // s_ovf: overflow events so far (only ever incremented)
    __shared__ unsigned long long s_drb;                                           // s_drb, s_dnrec, s_dsplit: descriptor of job s_job
    __shared__ u32 s_dnrec, s_dsplit;
    __shared__ unsigned long long s_segsrc[PLAN_MAXW];                             // byte address of the job's segment s
    __shared__ u32 s_segpre[PLAN_MAXW + 1];                                        // records of the segments before s (thread 0 only)

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    const u64 EMPTY = ~0ULL;
    for (u32 i = t; i < cap * KW; i += CS_THREADS) s_keys[i] = EMPTY;
    for (u32 i = t; i < cap * nb; i += CS_THREADS) s_counts[i] = 0;
    for (int i = t; i < HIST_SMEM_BINS; i += CS_THREADS) s_hist[i] = 0;
    if constexpr (MB) for (int i = t; i < 11 * CS_H2_I1; i += CS_THREADS) s_h2[i] = 0;
    u32 n1 = 0, n2 = 0, ndist = 0, nsplit = 0, ovf_seen = 0, mb_parity = 0;
    // solidity of a summed count c (an int32 >= 1 when the slot is occupied): amin <= c <= amax as one unsigned range test
    const long long sol_a = amin < 1 ? 1 : amin, sol_b = amax > 0x7FFFFFFFLL ? 0x7FFFFFFFLL : amax;
    const bool sol_any = sol_b >= sol_a && sol_a <= 0x7FFFFFFFLL;
    const u32 sol_lo = sol_any ? (u32)sol_a : 0xFFFFFFFFu, sol_span = sol_any ? (u32)(sol_b - sol_a) : 0u;   // (no count is 2^32 - 1)
    u64* my_qkey = s_qkey + (size_t)warp * CS_QCAP * KW;
    u32* my_qslot = s_qslot + (size_t)warp * CS_QCAP;
    const u32 keys_a = cs_saddr(s_keys), counts_a = cs_saddr(s_counts);
    const u32 jrec_a = cs_saddr(s_jrec), mbar_a = cs_saddr(s_mbar);

    // thread 0: records [s0, s0 + n) of the published job (its segments taken as one sequence) -> job buffer, one bulk copy
    // per segment the slice overlaps
    auto load_slice = [&](u32 s0, u32 n) {
        cs_mbar_expect(mbar_a, n * (u32)(RW * 8));
        for (u32 s = 0; s < segs.W; s++) {
            const u32 lo = max(s0, s_segpre[s]), hi = min(s0 + n, s_segpre[s + 1]);
            if (lo < hi) cs_bulk_copy(jrec_a + (lo - s0) * (u32)(RW * 8), s_segsrc[s] + (u64)(lo - s_segpre[s]) * (u64)(RW * 8), (hi - lo) * (u32)(RW * 8), mbar_a);
        }
    };
    // WARP 0 (all its lanes): publishes the descriptor of job `job` (held by lane 0) and starts the bulk copies of its first
    // slice of records.  Lane s looks after segment s -- its bounds, its source address, its copy -- so the W segments of a
    // multi-GPU job cost one round trip to the tables, not W.
    auto publish = [&](u32 job_lane0) {
        const u32 job = __shfl_sync(0xFFFFFFFFu, job_lane0, 0);
        if (lane == 0) s_job = job;
        if (job < njobs) {
            if constexpr (KEYS) { if (lane == 0) { s_drb = (u64)job * slab; s_dnrec = min(bucket_n[job], slab); s_dsplit = 0; } }
            else {
                const bool mine = (u32)lane < segs.W;
                u64 a = 0, b = 0, src = 0;
                if (mine) { a = segs.X[(u64)lane * segs.PW + job]; b = segs.X[(u64)lane * segs.PW + job + 1]; src = segs.tab->segptr[lane] + a * (u64)(RW * 8); }
                const u32 c = (u32)(b - a);
                u32 inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
                const u32 pre = inc - c, tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
                if (mine) { s_segsrc[lane] = src; s_segpre[lane] = pre; }
                const u32 n0 = min(tot, BUFREC);
                if (lane == 0) {
                    s_segpre[segs.W] = tot;
                    const float km = (float)segs.gk_q[(u64)segs.qbase + job];
                    u32 sp0 = 0;
                    while (sp0 < segs.max_split0 && km > segs.fit * (float)(1u << sp0)) sp0++;
                    s_dnrec = tot; s_dsplit = sp0;
                    if (tot) cs_mbar_expect(mbar_a, n0 * (u32)(RW * 8));
                }
                __syncwarp();                                                      // the barrier is armed before any copy can complete on it
                const u32 hi = min(n0, pre + c);
                if (mine && pre < hi) cs_bulk_copy(jrec_a + pre * (u32)(RW * 8), src, (hi - pre) * (u32)(RW * 8), mbar_a);
            }
        }
        __syncwarp();
    };
    if (warp == 0) {
        u32 first = 0;
        if (lane == 0) {
            s_ovf = 0;
            if constexpr (!KEYS) cs_mbar_init(mbar_a, 1);
            first = atomicAdd(work_counter, 1u);
        }
        __syncwarp();
        publish(first);
    }

    // The insert is a uniform pipeline: the FAST path of a k-mer is straight-line code -- one snapshot of its home slot, a
    // predicated increment when the slot already holds the key (most k-mers of a 30-100x read set are repeats) -- and every
    // other k-mer (new key, or home slot taken) goes to a per-warp ring of (key, slot, probes) entries with one ballot.  The
    // ring is served 32 entries at a time with all lanes active, ONE probe per entry and round (claim by CAS, or step to the
    // next slot and go back to the ring): no lane ever waits for another lane's probe sequence, and there is no divergent
    // branch on the path of a k-mer (r02g profile: 45 % of the warp instructions of the previous version were control flow,
    // 21 of 32 lanes active on average).
    u32 qh = 0, qc = 0;                                                            // ring head / entries (warp-uniform)
    auto ring_push = [&](bool put, const Kmer<KW>& key, u32 sw) {
        const u32 pm = __ballot_sync(0xFFFFFFFFu, put);
        if (put) {
            const u32 pos = (qh + qc + (u32)__popc(pm & lt_mask)) & (u32)(CS_QCAP - 1);
#pragma unroll
            for (int q = 0; q < KW; q++) my_qkey[pos * KW + q] = key.w[q];
            my_qslot[pos] = sw;
        }
        qc += (u32)__popc(pm);
    };
    // one round over (up to) 32 ring entries: entry = key + (slot | bank << 16 | probes << 24)
    auto ring_serve = [&]() {
        __syncwarp();
        const u32 take = min(qc, 32u);
        const bool on = (u32)lane < take;
        const u32 idx = (qh + (u32)lane) & (u32)(CS_QCAP - 1);
        Kmer<KW> key; u32 sw = 0;
#pragma unroll
        for (int q = 0; q < KW; q++) key.w[q] = on ? my_qkey[idx * KW + q] : 0ULL;
        if (on) sw = my_qslot[idx];
        __syncwarp();                                                              // entries are in registers before the tail is rewritten
        bool again = false;
        if (on) {
            u32 slot = sw & 0xFFFFu;
            const u32 bank = MB ? ((sw >> 16) & 0xFFu) : 0u;
            if (cs_probe(keys_a, slot, key)) cs_inc32(counts_a + (MB ? slot * nb + bank : slot) * 4u);
            else {
                const u32 np = (sw >> 24) + 1u;
                slot = (slot + 1 == cap) ? 0u : slot + 1;
                if (np >= (u32)CS_MAXPROBE) atomicAdd(&s_ovf, 1u);
                else { again = true; sw = slot | (bank << 16) | (np << 24); }
            }
        }
        qh = (qh + take) & (u32)(CS_QCAP - 1); qc -= take;
        ring_push(again, key, sw);
    };

    for (;;) {
        __syncthreads();                                                           // table clean, descriptor of s_job published
        const u32 job = s_job;
        if (job >= njobs) break;
        const u64 rb = s_drb;
        const u32 nrec = s_dnrec;
        const u32 nslices = (nrec + BUFREC - 1) / BUFREC;
        // the next job is claimed now (thread 0 keeps the ticket in a register: nobody waits for the atomic until the
        // records of this job are no longer needed)
        u32 next_job = 0;
        if (t == 0) next_job = atomicAdd(work_counter, 1u);
        if (nrec == 0) {                                                           // an empty partition (uniform branch: nothing was copied, the table is clean)
            __syncthreads();                                                       // everybody has read the descriptor
            if (warp == 0) publish(next_job);
            continue;
        }
        u32 buf_state = 1;                                                         // 1: slice 0 of this job is on its way (publish); 2: resident; 0: neither

        // depth-first over work items (record sub-pass, hash-split level, hash residue); uniform across the CTA.
        // A job too big for one table starts as 2^split0 RECORD sub-passes: sub-pass rp takes the records whose sub-bin
        // (4 hash bits of the minimizer, stored in the record) is rp modulo 2^split0 -- every record, hence every k-mer, in exactly
        // one sub-pass, so the extra passes cost a prefix scan each, not a re-extraction of every k-mer.  A table that still
        // overflows (one minimizer heavier than a table) splits the k-mer hash space in two, recursively (lvl, res).
        u32 stack[CS_MAX_SPLIT + 2 + (1 << CS_MAX_SPLIT0)];
        int sp = 0;
        const u32 rmask = (1u << s_dsplit) - 1u;
        for (u32 r = rmask + 1u; r-- > 0;) stack[sp++] = r << 24;
        while (sp > 0) {
            const u32 item = stack[--sp];
            const u32 rp = item >> 24, lvl = (item >> 16) & 0xFFu, res = item & 0xFFFFu, smask = (1u << lvl) - 1u;

            // ---- insert: flat keys (one per thread and iteration; the probe loop runs to the end, loads stay <= 52 %) ---------
            if constexpr (KEYS) {
                const u64* kp = recs + rb * KW;
                for (u32 i0 = 0; i0 < nrec; i0 += CS_THREADS) {
                    const u32 i = i0 + (u32)t;
                    if (i >= nrec) continue;
                    Kmer<KW> c;
                    if constexpr (KW == 1) c.w[0] = __ldg(kp + i);
                    else { const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(kp) + i); c.w[0] = v.x; c.w[1] = v.y; }
                    u32 bank = 0;
                    if constexpr (MB) { bank = (u32)(c.w[KW - 1] >> 62); c.w[KW - 1] &= ~(3ULL << 62); }
                    const u32 h = cs_hash(c);
                    if (((h >> 8) & smask) != res) continue;
                    u32 slot = __umulhi(h, cap);
                    bool ok = false;
                    for (int p = 0; p < CS_MAXPROBE && !ok; p++) { ok = cs_probe(keys_a, slot, c); if (!ok) slot = (slot + 1 == cap) ? 0u : slot + 1; }
                    if (ok) cs_inc32(counts_a + (MB ? slot * nb + bank : slot) * 4u); else atomicAdd(&s_ovf, 1u);
                }
                __syncthreads();
            }
            // ---- insert: super-k-mer records, one slice of <= BUFREC records in the job buffer at a time ----------------------
            for (u32 sl = 0; !KEYS && sl < nslices; sl++) {
                const u32 s0 = sl * BUFREC, n = min(BUFREC, nrec - s0);
                if (!(sl == 0 && buf_state == 2)) {
                    // (everybody is past the barrier that ended the previous slice: the buffer is free)
                    if (!(sl == 0 && buf_state == 1) && t == 0) { if constexpr (!KEYS) load_slice(s0, n); }
                    cs_mbar_wait(mbar_a, mb_parity); mb_parity ^= 1u;
                }
                buf_state = (nslices == 1) ? 2u : 0u;                              // a one-slice job stays in the buffer for its sub-passes

                // work items of CS_Q k-mers: exclusive prefix over the records of the slice (block scan), then equal shares per warp
                // (one scan for both: items in the low 16 bits -- <= 2368 x 7 --, records of this sub-pass in the high 16)
                u32 my_ni[RPT], tsum = 0;
#pragma unroll
                for (u32 j = 0; j < RPT; j++) {
                    const u32 r = (u32)t * RPT + j;
                    u32 ni = 0;
                    if (r < n) {
                        const u32 lw = (u32)s_jrec[(size_t)r * RW + (RW - 1)], nk = (lw >> 8) & 0xFFu;       // [nk:8][sub-bin:4][bank:4]
                        ni = (((lw >> 4) & rmask) == rp) ? (((nk + CS_Q - 1) / CS_Q) | (1u << 16)) : 0u;       // (rmask <= 15)
                    }
                    my_ni[j] = ni; tsum += ni;
                }
                u32 tinc = tsum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, tinc, d); if (lane >= d) tinc += o; }
                if (lane == 31) s_wsum[warp] = tinc;
                __syncthreads();
                // prefix over the 32 warp totals: every warp scans them with shuffles (one LDS + 5 steps instead of 32 LDS + adds)
                u32 wpre, I;
                {
                    u32 ws = lane < CS_WARPS ? s_wsum[lane] : 0u, wi = ws;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += o; }
                    I = __shfl_sync(0xFFFFFFFFu, wi, 31);
                    wpre = __shfl_sync(0xFFFFFFFFu, wi - ws, warp);
                }
                const u32 nc = I >> 16;                                            // records of this sub-pass (compacted positions 0 .. nc - 1)
                I &= 0xFFFFu;                                                      // their work items
                {
                    u32 run = wpre + tinc - tsum;
#pragma unroll
                    for (u32 j = 0; j < RPT; j++) {
                        if (my_ni[j]) s_pref[run >> 16] = (run & 0xFFFFu) | (((u32)t * RPT + j) << 16);
                        run += my_ni[j];
                    }
                }
                __syncthreads();

                const u32 lo = (I * (u32)warp) / (u32)CS_WARPS, hi = (I * (u32)(warp + 1)) / (u32)CS_WARPS;        // this warp's items
                if (lo < hi) {
                    // (compacted) record owning item `lo`: the largest r with pref[r] <= lo (32-ary search by ballots; pref[0] = 0)
                    u32 rcur;
                    {
                        u32 base = 0, len = nc;
                        while (len > 32) {
                            const u32 stride = (len + 31) >> 5, off = (u32)lane * stride;
                            const bool le = off < len && (s_pref[base + off] & 0xFFFFu) <= lo;  // true for a prefix of the lanes
                            const u32 j = (u32)__popc(__ballot_sync(0xFFFFFFFFu, le)) - 1u;
                            base += j * stride; len = min(stride, len - j * stride);
                        }
                        const bool le = (u32)lane < len && (s_pref[base + lane] & 0xFFFFu) <= lo;
                        rcur = base + (u32)__popc(__ballot_sync(0xFFFFFFFFu, le)) - 1u;
                    }
                    // 32 consecutive items per iteration, whatever records they belong to: 32 items span at most 32 records, so
                    // the owner of item x is found among the prefixes of records rcur .. rcur + 31 (one value per lane)
                    for (u32 x0 = lo; x0 < hi; x0 += 32) {
                        // the table overflowed (somebody ran out of probes): the pass is lost, every further insert into the full
                        // table would walk CS_MAXPROBE slots for nothing -- drop the ring and leave (the split passes redo it all).
                        // The flag is READ here and TESTED at the bottom of the iteration: its shared-memory latency hides behind
                        // the iteration instead of stalling its first instruction (measured: 3 % of the kernel when tested at once)
                        const u32 ovf_peek = *reinterpret_cast<volatile u32*>(&s_ovf);
                        const u32 x = x0 + (u32)lane;
                        const u32 rl = rcur + (u32)lane;
                        const u32 el = rl < nc ? s_pref[rl] : 0xFFFFFFFFu;             // (item prefix | record << 16) of compacted record rcur + lane
                        const u32 pl = rl < nc ? (el & 0xFFFFu) : 0xFFFFFFFFu;
                        u32 ro = 0;                                                    // largest lane l with pref[rcur + l] <= x
#pragma unroll
                        for (int step = 16; step; step >>= 1) { const u32 v = __shfl_sync(0xFFFFFFFFu, pl, (ro + step) & 31); if (v <= x) ro += step; }
                        const u32 en = __shfl_sync(0xFFFFFFFFu, el, ro);
                        const u32 ex_r = en & 0xFFFFu;
                        u64 rw[RW];
                        {
                            const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(s_jrec + (size_t)min(en >> 16, BUFREC - 1u) * RW);
                            const ulonglong2 a = rp[0]; rw[0] = a.x; rw[1] = a.y;
                            if constexpr (RW == 4) { const ulonglong2 b = rp[1]; rw[2] = b.x; rw[3] = b.y; }
                        }
                        const u32 nk_r = (u32)(rw[RW - 1] >> 8) & 0xFFu;
                        const int j0 = (int)(x - ex_r) * CS_Q;
                        const int cnt = x < hi ? min((int)nk_r - j0, CS_Q) : 0;
                        // record owning the first item of the next iteration (lane 31 knows: its own record, or the one after it)
                        rcur = __shfl_sync(0xFFFFFFFFu, rcur + ro + ((x + 1u >= ex_r + (nk_r + CS_Q - 1) / CS_Q) ? 1u : 0u), 31);
                        u32 bank = 0;
                        Kmer<KW> f, rc;
#pragma unroll
                        for (int q = 0; q < KW; q++) { f.w[q] = 0; rc.w[q] = 0; }
                        u64 nextb = 0;
                        if (cnt > 0) {
                            if constexpr (MB) bank = (u32)rw[RW - 1] & 0xFu;
                            if constexpr (KW == 1) f.w[0] = cs_window<RW>(rw, j0) >> (64 - 2 * k);
                            else { const u64 hw[2] = {cs_window<RW>(rw, j0), cs_window<RW>(rw, j0 + 32)}; f = rec_first_kmer2(hw, k); }
                            rc = kmer_revcomp(f, k);
                            nextb = cs_window<RW>(rw, j0 + k);                         // the CS_Q-1 bases that follow the first k-mer
                        }
                        // the CS_Q k-mers of the item: canonical forms, hashes and home slots first, then all the first probes are
                        // issued together (their shared-memory latencies overlap), then every k-mer is resolved
                        // (128-bit keys: two at a time, the register budget of a 1024-thread CTA is 64)
                        constexpr int PH = KW == 1 ? CS_Q : 2;
#pragma unroll
                        for (int g = 0; g < CS_Q; g += PH) {
                            Kmer<KW> cc[PH]; u32 sl0[PH]; u32 actm = 0;
#pragma unroll
                            for (int v = 0; v < PH; v++) {
                                const int u = g + v;                                   // (straight-line: lanes past their item's end compute on and are masked)
                                if (u) { kmer_roll(f, rc, (int)(nextb >> 62), k); nextb <<= 2; }
                                cc[v] = kmer_canonical(f, rc);
                                const u32 h = cs_hash(cc[v]);
                                sl0[v] = __umulhi(h, cap);
                                actm |= (u < cnt && ((h >> 8) & smask) == res) ? (1u << v) : 0u;
                            }
                            CsSnap<KW> sn[PH];
#pragma unroll
                            for (int v = 0; v < PH; v++) {
                                if constexpr (KW == 1) sn[v] = cs_snap1(keys_a, sl0[v]); else sn[v] = cs_snap2(keys_a, sl0[v]);   // (slot 0 for the idle ones: harmless)
                            }
#pragma unroll
                            for (int v = 0; v < PH; v++) {
                                const bool act = (actm >> v) & 1u;
                                bool hit = act && cs_match(cc[v], sn[v]);
                                // home slot not (completely) taken by another key: claim it -- a predicated CAS, no branch.  The
                                // CAS tells the truth whatever the snapshot showed; a k-mer goes to the ring only when its home
                                // slot holds a foreign key (final: slots only go EMPTY -> key), one slot further, one probe done.
                                hit |= cs_claim_if(act && !hit && !cs_foreign(cc[v], sn[v]), keys_a, sl0[v], cc[v]);
                                if (hit) cs_inc32(counts_a + (MB ? sl0[v] * nb + bank : sl0[v]) * 4u);
                                const u32 s1 = (sl0[v] + 1 == cap) ? 0u : sl0[v] + 1;
                                ring_push(act && !hit, cc[v], s1 | (MB ? (bank << 16) : 0u) | (1u << 24));
                                while (qc >= 32u) ring_serve();                        // (a round may hand every entry back: the ring must be under 32 before the next push)
                            }
                        }
                        if (__any_sync(0xFFFFFFFFu, ovf_peek != ovf_seen)) { qc = 0; break; }
                    }
                }
                while (qc) {
                    const u32 ovf_peek = *reinterpret_cast<volatile u32*>(&s_ovf);
                    ring_serve();
                    if (__any_sync(0xFFFFFFFFu, ovf_peek != ovf_seen)) { qc = 0; break; }
                }
                __syncthreads();                                                   // inserts of the slice done, job buffer free
                // (uniform: nobody inserts between this barrier and the next ones)  a lost pass skips its remaining slices; the
                // buffer state of a multi-slice job already says "reload slice 0"
                if (*reinterpret_cast<volatile u32*>(&s_ovf) != ovf_seen) break;
            }
            const u32 ovf_now = *reinterpret_cast<volatile u32*>(&s_ovf);          // stable: nobody inserts until after the next barriers
            const bool overflowed = ovf_now != ovf_seen;
            ovf_seen = ovf_now;
            if (overflowed) {
                // clear, then split this item in two (or give up: reported, never silent)
                for (u32 i = t; i < cap * KW; i += CS_THREADS) s_keys[i] = EMPTY;
                for (u32 i = t; i < cap * nb; i += CS_THREADS) s_counts[i] = 0;
                if (lvl >= (u32)CS_MAX_SPLIT) { if (t == 0) atomicAdd(&ctr->smem_failed, 1u); }
                else { stack[sp++] = (rp << 24) | ((lvl + 1) << 16) | (res + (1u << lvl)); stack[sp++] = (rp << 24) | ((lvl + 1) << 16) | res; nsplit++; }
                if (KEYS || sp == 0) __syncthreads();                              // (the record path has barriers before its next insert)
                if (sp == 0 && warp == 0) publish(next_job);                       // gave up on the last item: move on
                continue;
            }
            // the records of this job are no longer needed after its last work item: the next job's first slice is copied
            // into the buffer while the table is swept
            if (sp == 0 && warp == 0) publish(next_job);

            // ---- sweep: CountProcessor chain over the table, compaction of the solid pairs, clear ---------------------
            // a slot is occupied iff its count is non-zero (a claim is always followed by its increment), so the scan reads
            // only the counts, four per 16-byte load; keys are read for the solid slots alone.  cap % (4 * CS_THREADS) == 0.
            u32 solidm = 0;                                                        // bit (4 * v + q): slot 4 * (v * CS_THREADS + t) + q
            if constexpr (MB) {
                // per-bank counts: thread t owns slots t, t + CS_THREADS, ... (bit v of solidm: slot v * CS_THREADS + t); the
                // abundance to dump (CountProcessorDump: Count(kmer, sum)) replaces the bank-0 count once the chain has run
                // Histogram updates are aggregated per warp (match.any on the bin): most distinct k-mers of a read set share one bin.
                const u32 nv = (cap + CS_THREADS - 1) / CS_THREADS;                // uniform trip count: the warp votes below
                for (u32 v = 0; v < nv; v++) {
                    const u32 sl = v * CS_THREADS + (u32)t;
                    u32 cv[MAXB]; u32 any = 0;
                    if (sl < cap) for (u32 b = 0; b < nb; b++) { cv[b] = s_counts[sl * nb + b]; any |= cv[b]; }
                    u32 bin1 = 0, bin2 = H2_NONE;
                    if (any) {
                        ndist++;
                        int32_t sum;
                        if (eval_counts(cv, spar, &sum, &bin1, &bin2)) solidm |= 1u << v;
                        s_counts[sl * nb] = (u32)sum;
                    }
                    const u32 p1 = __match_any_sync(0xFFFFFFFFu, bin1);
                    if (bin1 && lane == __ffs((int)p1) - 1) {
                        const u32 c = (u32)__popc(p1);
                        if (bin1 < HIST_SMEM_BINS) atomicAdd(&s_hist[bin1], c); else atomicAdd(&g_hist[bin1], (unsigned long long)c);
                    }
                    if (spar.histo2d) {
                        const u32 p2 = __match_any_sync(0xFFFFFFFFu, bin2);
                        if (bin2 != H2_NONE && lane == __ffs((int)p2) - 1) {
                            const u32 c = (u32)__popc(p2), i2 = bin2 / 10001u, i1 = bin2 - i2 * 10001u;
                            if (i1 < CS_H2_I1) atomicAdd(&s_h2[i2 * CS_H2_I1 + i1], c); else atomicAdd(&g_hist2d[bin2], (unsigned long long)c);
                        }
                    }
                }
            } else
            {
                const uint4* c4 = reinterpret_cast<const uint4*>(s_counts);
                int v = 0;
                for (u32 g = t; g < cap / 4; g += CS_THREADS, v++) {
                    const uint4 cc = c4[g];
                    const u32 cv[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const u32 c = cv[q];
                        // counts 0, 1 and 2 -- nine slots out of ten -- are handled without a branch
                        ndist += (c != 0u) ? 1u : 0u; n1 += (c == 1u) ? 1u : 0u; n2 += (c == 2u) ? 1u : 0u;
                        if (c > 2u) { const u32 bin = histo_bin((int32_t)c); if (bin) { if (bin < HIST_SMEM_BINS) atomicAdd(&s_hist[bin], 1u); else atomicAdd(&g_hist[bin], 1ULL); } }
                        solidm |= (c - sol_lo <= sol_span) ? (1u << (4 * v + q)) : 0u;   // amin <= c <= amax, c != 0 (one unsigned range test)
                    }
                }
            }
            const u32 n = (u32)__popc(solidm);
            u32 inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
            // one bump of the global cursor per WARP and job (no block-wide scan, no barrier: the order of the solid pairs is
            // settled by the final sort); while it is in flight every thread clears the slot groups it scanned that hold
            // nothing solid (nobody else reads or writes them in this phase)
            const u32 tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
            unsigned long long wbase = 0;
            if (tot && lane == 31) wbase = atomicAdd(&ctr->solid_n, (unsigned long long)tot);
            if constexpr (MB) {
                u32 v = 0;
                for (u32 sl = t; sl < cap; sl += CS_THREADS, v++) {
                    if ((solidm >> v) & 1u) continue;
#pragma unroll
                    for (int q = 0; q < KW; q++) s_keys[sl * KW + q] = EMPTY;
                    for (u32 b = 0; b < nb; b++) s_counts[sl * nb + b] = 0;
                }
            } else {
                ulonglong2* k2 = reinterpret_cast<ulonglong2*>(s_keys);
                uint4* c4 = reinterpret_cast<uint4*>(s_counts);
                int v = 0;
                for (u32 g = t; g < cap / 4; g += CS_THREADS, v++) {
                    if ((solidm >> (4 * v)) & 0xFu) continue;
#pragma unroll
                    for (int q = 0; q < 2 * KW; q++) k2[2 * KW * g + q] = make_ulonglong2(EMPTY, EMPTY);
                    c4[g] = make_uint4(0, 0, 0, 0);
                }
            }
            if (tot) {
                u64 pos = __shfl_sync(0xFFFFFFFFu, wbase, 31) + inc - n;
                u32 m = solidm;
                while (m) {
                    const int b = __ffs((int)m) - 1; m &= m - 1;
                    const u32 slot = MB ? (u32)b * CS_THREADS + (u32)t : 4u * ((u32)(b >> 2) * CS_THREADS + (u32)t) + (u32)(b & 3);
                    if (pos < out_cap) {
#pragma unroll
                        for (int q = 0; q < KW; q++) out_keys[pos * KW + q] = s_keys[slot * KW + q];
                        out_vals[pos] = s_counts[MB ? slot * nb : slot];
                    } else atomicExch(&ctr->overflow, 2u);
                    pos++;
                }
                // ... and the groups that held solid slots
                if constexpr (MB) {
                    u32 mm = solidm;
                    while (mm) {
                        const int b = __ffs((int)mm) - 1; mm &= mm - 1;
                        const u32 sl = (u32)b * CS_THREADS + (u32)t;
#pragma unroll
                        for (int q = 0; q < KW; q++) s_keys[sl * KW + q] = EMPTY;
                        for (u32 bb = 0; bb < nb; bb++) s_counts[sl * nb + bb] = 0;
                    }
                } else {
                    ulonglong2* k2 = reinterpret_cast<ulonglong2*>(s_keys);
                    uint4* c4 = reinterpret_cast<uint4*>(s_counts);
                    int v = 0;
                    for (u32 g = t; g < cap / 4; g += CS_THREADS, v++) {
                        if (!((solidm >> (4 * v)) & 0xFu)) continue;
#pragma unroll
                        for (int q = 0; q < 2 * KW; q++) k2[2 * KW * g + q] = make_ulonglong2(EMPTY, EMPTY);
                        c4[g] = make_uint4(0, 0, 0, 0);
                    }
                }
            }
            if (sp > 0) __syncthreads();                                           // table clean before the next sub-pass inserts (KEYS path; cheap otherwise)
        }
    }

    // ---- per-CTA totals ---------------------------------------------------------------------------------------------
    n1 = __reduce_add_sync(0xFFFFFFFFu, n1); n2 = __reduce_add_sync(0xFFFFFFFFu, n2);
    ndist = __reduce_add_sync(0xFFFFFFFFu, ndist);
    if (lane == 0) {
        if (n1) atomicAdd(&s_hist[1], n1);
        if (n2) atomicAdd(&s_hist[2], n2);
        if (ndist) atomicAdd(&ctr->distinct_n, (unsigned long long)ndist);
    }
    __syncthreads();
    flush_hist(s_hist, g_hist);
    if constexpr (MB) {
        if (spar.histo2d)
            for (int i = t; i < 11 * CS_H2_I1; i += CS_THREADS) {
                const u32 c = s_h2[i];
                if (c) atomicAdd(&g_hist2d[(u32)(i % CS_H2_I1) + 10001u * (u32)(i / CS_H2_I1)], (unsigned long long)c);
            }
    }
    if (t == 0 && nsplit) atomicAdd(&ctr->smem_splits, nsplit);
}

#endif  // __CUDACC__
}  // namespace dsk
