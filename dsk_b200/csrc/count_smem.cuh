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
#pragma once
#include "kmer_bits.cuh"
#include "superk.cuh"
#include "count.cuh"

namespace dsk {

constexpr int CS_THREADS = 1024;               // one CTA per SM: the biggest table (measured: 2 x 512 threads with half tables is 25 % slower)
constexpr int CS_CTAS_PER_SM = 1;
#ifndef CS_CHUNK
#define CS_CHUNK 32                            // records a warp takes at a time (<= 32)
#endif
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_MAXPROBE = 128;
constexpr int CS_MAX_SPLIT = 10;               // up to 1024 sub-passes before the job is reported as failed
constexpr int CS_Q = 4;                        // k-mers per work item
constexpr int CS_QCAP = 64;                    // retry queue entries per warp

// split0: the job starts as 2^split0 sub-passes (the host knows the partition's k-mers and the sampled density, so a
// partition that cannot fit the table is never tried in one pass first)
struct SmemJob { unsigned long long rec_begin; unsigned int nrec; unsigned int split0; };
constexpr int CS_MAX_SPLIT0 = 4;

// bytes of dynamic shared memory for a table of `cap` slots (host + device agree through this one function):
// table keys + counts, record staging [warps][32][RW], retry queue [warps][QCAP] keys + slots
// nb = counts kept per slot (1, or one per bank when the processors need per-bank counts: -histo2D, solidity kinds)
template <int KW> DSK_HD size_t cs_smem_bytes(u32 cap, int nb = 1)
{
    return (size_t)cap * (8 * KW + 4 * nb) + (size_t)CS_WARPS * 32 * 2 * KW * 8 + (size_t)CS_WARPS * CS_QCAP * (8 * KW + 4);
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

// MB = false: one count per k-mer (banks summed), solidity = abundance range [amin, amax] -- the dsk default.
// MB = true : `nb` counts per slot (the record's bank byte picks the column); the sweep runs the whole CountProcessor
//             chain of count.cuh (`process_counts`: per-bank solidity kinds, -histo2D, per-bank histograms) on them.
// KEYS = false: a job is a partition of super-k-mer records (jobs[]).
// KEYS = true : a job is a hash bucket of flat canonical k-mers written by k_expand_bucket (count.cuh): job j reads
//               min(bucket_n[j], slab) keys at recs + j * slab * KW; `jobs` is unused.
template <int KW, bool MB, bool KEYS = false>
__global__ void __launch_bounds__(CS_THREADS, CS_CTAS_PER_SM) k_count_smem(const u64* __restrict__ recs, const SmemJob* __restrict__ jobs, u32 njobs,
                                                              int k, u32 cap, long long amin, long long amax,
                                                              u64* __restrict__ out_keys, u32* __restrict__ out_vals, u64 out_cap,
                                                              unsigned long long* __restrict__ g_hist, Counters* ctr, u32* work_counter,
                                                              int nb_arg, const SolidityParams spar, unsigned long long* __restrict__ g_hist2d,
                                                              const u32* __restrict__ bucket_n = nullptr, u32 slab = 0)
{
    constexpr int RW = 2 * KW;
    const u32 nb = MB ? (u32)nb_arg : 1u;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    u64* s_keys = reinterpret_cast<u64*>(s_dyn);                                   // [cap][KW]
    u32* s_counts = reinterpret_cast<u32*>(s_dyn + (size_t)cap * 8 * KW);          // [cap][nb]
    u64* s_rec = reinterpret_cast<u64*>(s_dyn + (size_t)cap * (8 * KW + 4 * nb));  // [CS_WARPS][32][RW]   (cap % 4 == 0 keeps it 16-byte aligned)
    u64* s_qkey = s_rec + (size_t)CS_WARPS * 32 * RW;                              // [CS_WARPS][CS_QCAP][KW]
    u32* s_qslot = reinterpret_cast<u32*>(s_qkey + (size_t)CS_WARPS * CS_QCAP * KW);   // [CS_WARPS][CS_QCAP]
    __shared__ u32 s_hist[HIST_SMEM_BINS];
    __shared__ u32 s_h2[MB ? 11 * CS_H2_I1 : 1];                                   // -histo2D bins with dim-1 index < CS_H2_I1, all 11 dim-2 rows
    __shared__ u32 s_wsum[CS_WARPS];
    __shared__ u32 s_job, s_flag, s_chunk;
    __shared__ unsigned long long s_base;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    const u64 EMPTY = ~0ULL;
    for (u32 i = t; i < cap * KW; i += CS_THREADS) s_keys[i] = EMPTY;
    for (u32 i = t; i < cap * nb; i += CS_THREADS) s_counts[i] = 0;
    for (int i = t; i < HIST_SMEM_BINS; i += CS_THREADS) s_hist[i] = 0;
    if constexpr (MB) for (int i = t; i < 11 * CS_H2_I1; i += CS_THREADS) s_h2[i] = 0;
    if (t == 0) s_flag = 0;
    u32 n1 = 0, n2 = 0, ndist = 0, nsplit = 0;
    u64* my_rec = s_rec + (size_t)warp * 32 * RW;
    u64* my_qkey = s_qkey + (size_t)warp * CS_QCAP * KW;
    u32* my_qslot = s_qslot + (size_t)warp * CS_QCAP;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(recs);
    const u32 keys_a = cs_saddr(s_keys), counts_a = cs_saddr(s_counts);

    // resolve everything in this warp's retry queue (all lanes active, full probe sequences)
    auto drain = [&](u32 qn) {
        __syncwarp();
        for (u32 b0 = 0; b0 < qn; b0 += 32) {
            const u32 idx = b0 + lane;
            if (idx < qn) {
                Kmer<KW> key;
#pragma unroll
                for (int q = 0; q < KW; q++) key.w[q] = my_qkey[idx * KW + q];
                u32 slot = my_qslot[idx], bank = 0;
                if constexpr (MB) { bank = slot >> 16; slot &= 0xFFFFu; }             // cap <= 16384
                bool ok = false;
                for (int p = 0; p < CS_MAXPROBE && !ok; p++) { ok = cs_probe(keys_a, slot, key); if (!ok) slot = (slot + 1 == cap) ? 0u : slot + 1; }
                if (ok) cs_inc32(counts_a + (MB ? slot * nb + bank : slot) * 4u); else s_flag = 1u;
            }
        }
        __syncwarp();
    };

    for (;;) {
        __syncthreads();                                                           // table clean, s_job free
        if (t == 0) s_job = atomicAdd(work_counter, 1u);
        __syncthreads();
        const u32 job = s_job;
        if (job >= njobs) break;
        const u64 rb = KEYS ? (u64)job * slab : jobs[job].rec_begin;
        const u32 nrec = KEYS ? min(bucket_n[job], slab) : jobs[job].nrec, nchunks = (nrec + CS_CHUNK - 1) / CS_CHUNK;

        // depth-first over (split level, residue) work items; uniform across the CTA
        u32 stack[CS_MAX_SPLIT + 2 + (1 << CS_MAX_SPLIT0)];
        int sp = 0;
        {
            const u32 l0 = KEYS ? 0u : min(jobs[job].split0, (u32)CS_MAX_SPLIT0);
            for (u32 r = (1u << l0); r-- > 0;) stack[sp++] = (l0 << 16) | r;
        }
        while (sp > 0) {
            const u32 item = stack[--sp];
            const u32 lvl = item >> 16, res = item & 0xFFFFu, smask = (1u << lvl) - 1u;
            if (t == 0) s_chunk = CS_WARPS;                                        // chunks beyond the first one per warp are dealt dynamically
            __syncthreads();

            // ---- insert: flat keys (one per thread and iteration; the probe loop runs to the end, loads stay <= 52 %) ---------
            if constexpr (KEYS) {
                const u64* kp = recs + rb * KW;
                for (u32 i0 = 0; i0 < nrec; i0 += CS_THREADS) {
                    if (*reinterpret_cast<volatile u32*>(&s_flag)) break;
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
                    if (ok) cs_inc32(counts_a + (MB ? slot * nb + bank : slot) * 4u); else s_flag = 1u;
                }
            }
            // ---- insert: warps expand chunks of 32 records -----------------------------------------------------------
            u32 qn = 0;
            for (u32 chunk = warp; !KEYS && chunk < nchunks;) {
                if (*reinterpret_cast<volatile u32*>(&s_flag)) break;              // somebody overflowed: the pass is void
                const u32 ri = chunk * CS_CHUNK + lane;
                u32 nk = 0;
                if (lane < CS_CHUNK && ri < nrec) {
                    const u64 i = rb + ri;
                    ulonglong2* dst = reinterpret_cast<ulonglong2*>(my_rec + lane * RW);
                    if constexpr (RW == 2) { const ulonglong2 v = __ldg(src + i); dst[0] = v; nk = (u32)(v.y >> 8) & 0xFFu; }
                    else { const ulonglong2 v = __ldg(src + 2 * i), u = __ldg(src + 2 * i + 1); dst[0] = v; dst[1] = u; nk = (u32)(u.y >> 8) & 0xFFu; }
                }
                // next chunk of this warp (claimed early so the atomic's latency hides under the work)
                u32 next = 0;
                if (lane == 0) next = atomicAdd(&s_chunk, 1u);
                const u32 ni = (nk + CS_Q - 1) / CS_Q;                             // work items of my record
                u32 inc = ni;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
                const u32 total = __shfl_sync(0xFFFFFFFFu, inc, 31);
                const u32 exc = inc - ni;
                next = __shfl_sync(0xFFFFFFFFu, next, 0);
                __syncwarp();
                for (u32 t0 = 0; t0 < total; t0 += 32) {
                    const u32 x = t0 + lane;
                    // owner record of item x: number of lanes whose inclusive prefix is <= x (binary search by shuffles)
                    u32 r = 0;
#pragma unroll
                    for (int step = 16; step; step >>= 1) { const u32 v = __shfl_sync(0xFFFFFFFFu, inc, (r + step - 1) & 31); if (v <= x) r += step; }
                    const u32 ex_r = __shfl_sync(0xFFFFFFFFu, exc, r & 31);
                    const u32 nk_r = __shfl_sync(0xFFFFFFFFu, nk, r & 31);
                    int cnt = 0, j0 = 0;
                    u32 bank = 0;
                    u64 rw[RW];
                    Kmer<KW> f, rc;
                    u64 nextb = 0;
                    if (x < total) {
                        j0 = (int)(x - ex_r) * CS_Q;
                        cnt = min((int)nk_r - j0, CS_Q);
                        const ulonglong2* rp = reinterpret_cast<const ulonglong2*>(my_rec + r * RW);
                        const ulonglong2 a = rp[0]; rw[0] = a.x; rw[1] = a.y;
                        if constexpr (RW == 4) { const ulonglong2 b = rp[1]; rw[2] = b.x; rw[3] = b.y; }
                        if constexpr (MB) bank = (u32)rw[RW - 1] & 0xFFu;
                        if constexpr (KW == 1) f.w[0] = cs_window<RW>(rw, j0) >> (64 - 2 * k);
                        else { const u64 hw[2] = {cs_window<RW>(rw, j0), cs_window<RW>(rw, j0 + 32)}; f = rec_first_kmer2(hw, k); }
                        rc = kmer_revcomp(f, k);
                        nextb = cs_window<RW>(rw, j0 + k);                         // the CS_Q-1 bases that follow the first k-mer
                    }
#pragma unroll
                    for (int u = 0; u < CS_Q; u++) {
                        bool pending = false, found = false; u32 slot = 0; Kmer<KW> c;
                        if (u < cnt) {
                            if (u) { kmer_roll(f, rc, (int)(nextb >> 62), k); nextb <<= 2; }
                            c = kmer_canonical(f, rc);
                            const u32 h = cs_hash(c);
                            if (((h >> 8) & smask) == res) {
                                slot = __umulhi(h, cap);
                                found = cs_probe(keys_a, slot, c);
                                if (!found) { slot = (slot + 1 == cap) ? 0u : slot + 1; found = cs_probe(keys_a, slot, c); }
                                if (!found) { pending = true; slot = (slot + 1 == cap) ? 0u : slot + 1; }
                            }
                        }
                        const u32 pm = __ballot_sync(0xFFFFFFFFu, pending);         // (also the reconvergence point of the probes)
                        if (found) cs_inc32(counts_a + (MB ? slot * nb + bank : slot) * 4u);
                        if (pm) {
                            if (pending) {
                                const u32 pos = qn + (u32)__popc(pm & lt_mask);
#pragma unroll
                                for (int q = 0; q < KW; q++) my_qkey[pos * KW + q] = c.w[q];
                                my_qslot[pos] = MB ? (slot | (bank << 16)) : slot;
                            }
                            qn += (u32)__popc(pm);
                            if (qn > CS_QCAP - 32) { drain(qn); qn = 0; }
                        }
                    }
                }
                __syncwarp();
                chunk = next;
            }
            if (qn) drain(qn);
            __syncthreads();
            const u32 overflowed = s_flag;
            __syncthreads();
            if (overflowed) {
                // clear, then split this item in two (or give up: reported, never silent)
                for (u32 i = t; i < cap * KW; i += CS_THREADS) s_keys[i] = EMPTY;
                for (u32 i = t; i < cap * nb; i += CS_THREADS) s_counts[i] = 0;
                if (t == 0) s_flag = 0;
                if (lvl >= (u32)CS_MAX_SPLIT) { if (t == 0) atomicAdd(&ctr->smem_failed, 1u); }
                else { stack[sp++] = ((lvl + 1) << 16) | (res + (1u << lvl)); stack[sp++] = ((lvl + 1) << 16) | res; nsplit++; }
                continue;
            }

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
                        if (c == 0u) continue;
                        ndist++;
                        if (c == 1u) n1++;
                        else if (c == 2u) n2++;
                        else { const u32 bin = histo_bin((int32_t)c); if (bin) { if (bin < HIST_SMEM_BINS) atomicAdd(&s_hist[bin], 1u); else atomicAdd(&g_hist[bin], 1ULL); } }
                        const long long sum = (long long)(int32_t)c;
                        if (amin <= sum && sum <= amax) solidm |= 1u << (4 * v + q);
                    }
                }
            }
            const u32 n = (u32)__popc(solidm);
            u32 inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            u32 wpre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < CS_WARPS; w++) { const u32 x = s_wsum[w]; if (w < warp) wpre += x; tot += x; }
            if (tot) {
                if (t == 0) s_base = atomicAdd(&ctr->solid_n, (unsigned long long)tot);
                __syncthreads();
                u64 pos = s_base + wpre + inc - n;
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
            }
            if constexpr (MB) {
                for (u32 sl = t; sl < cap; sl += CS_THREADS) {
#pragma unroll
                    for (int q = 0; q < KW; q++) s_keys[sl * KW + q] = EMPTY;
                    for (u32 b = 0; b < nb; b++) s_counts[sl * nb + b] = 0;
                }
            } else
            {
                // every thread clears exactly the slot groups it scanned (nobody else reads or writes them in this phase)
                ulonglong2* k2 = reinterpret_cast<ulonglong2*>(s_keys);
                uint4* c4 = reinterpret_cast<uint4*>(s_counts);
                for (u32 g = t; g < cap / 4; g += CS_THREADS) {
#pragma unroll
                    for (int q = 0; q < 2 * KW; q++) k2[2 * KW * g + q] = make_ulonglong2(EMPTY, EMPTY);
                    c4[g] = make_uint4(0, 0, 0, 0);
                }
            }
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
