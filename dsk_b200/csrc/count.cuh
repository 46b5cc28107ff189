// count.cuh -- K4/K7/K6: per-partition counting.
//
//   hash path : k_hash_insert (expand super-k-mers -> canonical k-mers -> open-addressing table with
//               atomicCAS claim + atomic increment)  then  k_hash_scan (filter + histogram + compaction).
//               Replaces PartitionsByHashCommand::execute / Hash16::insert / OAHash::increment
//               (K/PartitionsCommand.cpp:372-739, G/src/gatb/tools/collections/impl/Hash16.hpp:198-230,
//               OAHash.hpp:92-99).
//   sort path : k_expand_keys (ReadSuperKCommand::execute, K/PartitionsCommand.cpp:944-1128) -> LSD radix
//               sort (radix.cuh) -> k_rle_emit (executeDump's run-length count, K/PartitionsCommand.cpp:1751-1801).
//   both end in the fused CountProcessor chain: histogram (K/CountProcessorHistogram.hpp:173-184 with the
//   Histogram.hpp:92,221 quirks) -> solidity (K/CountProcessorSolidity.hpp:186-300) -> dump of Count(kmer,sum)
//   (K/CountProcessorDump.hpp:85-152).
#pragma once
#include <cmath>
#include "kmer_bits.cuh"
#include "kmer_wide.cuh"
#include "superk.cuh"

namespace dsk {

constexpr int MAXB = 16;                                         // DSKGPU_MAX_BANKS
constexpr u32 HASH_MAX_PROBE = 4096;

struct SolidityParams {
    int kind;                      // 0 sum 1 min 2 max 3 one 4 all 5 custom
    int nbanks;                    // counts kept per k-mer (1 when banks are summed)
    int histo2d;
    long long amin[MAXB];
    long long amax;
    unsigned char solid_vec[MAXB];
    unsigned long long* bank_hist; // optional [nbanks][10001]: histogram of every bank's own count (CountProcessorCutoff with N banks)
};

#ifdef __CUDACC__

constexpr int HIST_SMEM_BINS = 1024;

// ---- the CountProcessor chain for one distinct k-mer --------------------------------------------------------
// returns true if solid; *sum_out = abundance to dump.  Histograms: bins < HIST_SMEM_BINS go to the block's smem
// histogram, the rest straight to global.
constexpr int H2_SMEM_I1 = 64;        // -histo2D bins with dim-1 index < 64 (all 11 dim-2 rows) are accumulated per block in
                                      // shared memory: three quarters of the distinct k-mers of a read set land in ONE bin, and
                                      // same-address global atomics serialise in L2 (measured: 0.6 s per 600 M k-mers)
__device__ __forceinline__ void flush_hist2d(const u32* s_h2, unsigned long long* g_hist2d)
{
    for (int i = threadIdx.x; i < 11 * H2_SMEM_I1; i += blockDim.x) {
        const u32 c = s_h2[i];
        if (c) atomicAdd(&g_hist2d[(u32)(i % H2_SMEM_I1) + 10001u * (u32)(i / H2_SMEM_I1)], (unsigned long long)c);
    }
}

__device__ __forceinline__ bool process_counts(const u32* cv, const SolidityParams& sp, u32* s_hist,
                                               unsigned long long* g_hist, unsigned long long* g_hist2d, int32_t* sum_out,
                                               u32* s_h2 = nullptr)
{
    const int nb = sp.nbanks;
    int32_t sum = 0;
    if (nb == 1) sum = (int32_t)cv[0];
    else for (int b = 0; b < nb; b++) if (sp.kind != 5 || sp.solid_vec[b]) sum += (int32_t)cv[b];   // CountProcessorChain.hpp:158-169
    *sum_out = sum;
    u32 bin = histo_bin(sum);
    if (bin) { if (bin < HIST_SMEM_BINS) atomicAdd(&s_hist[bin], 1u); else atomicAdd(&g_hist[bin], 1ULL); }
    if (sp.bank_hist) {                                              // CountProcessorCutoff.hpp:113-116: bank i sees count[i]
        for (int b = 0; b < nb; b++) { const u32 bb = histo_bin((int32_t)cv[b]); if (bb) atomicAdd(&sp.bank_hist[(size_t)b * 10001u + bb], 1ULL); }
    }
    if (sp.histo2d) {
        u32 i1 = (u32)(sum - (int32_t)cv[0]) & 0xFFFFu, i2 = cv[0] & 0xFFFFu;
        if (i1 >= 10000u) i1 = 10000u;
        if (i2 >= 10u) i2 = 10u;
        if (s_h2 && i1 < (u32)H2_SMEM_I1) atomicAdd(&s_h2[i2 * H2_SMEM_I1 + i1], 1u);
        else atomicAdd(&g_hist2d[i1 + 10001u * i2], 1ULL);
    }
    auto inr = [&](long long x, long long lo) { return lo <= x && x <= sp.amax; };
    switch (sp.kind) {
    case 0: return inr(sum, sp.amin[0]);
    case 1: { u32 v = cv[0]; for (int b = 1; b < nb; b++) v = min(v, cv[b]); return inr(v, sp.amin[0]); }
    case 2: { u32 v = cv[0]; for (int b = 1; b < nb; b++) v = max(v, cv[b]); return inr(v, sp.amin[0]); }
    case 3: { for (int b = 0; b < nb; b++) if (inr(cv[b], sp.amin[b])) return true; return false; }
    case 4: { for (int b = 0; b < nb; b++) if (!inr(cv[b], sp.amin[b])) return false; return true; }
    default: { for (int b = 0; b < nb; b++) { bool in = inr(cv[b], sp.amin[b]); if (sp.solid_vec[b] != in) return false; } return true; }
    }
}

// The same chain without its side effects: returns solidity, the abundance to dump, the 1-D histogram bin (0 = not
// recorded) and the flat 2-D bin (i1 + 10001 * i2, or ~0u without -histo2D).  The caller applies the histogram updates
// (count_smem.cuh aggregates them per warp: three quarters of the distinct k-mers of a read set land in ONE bin).
constexpr u32 H2_NONE = 0xFFFFFFFFu;
__device__ __forceinline__ bool eval_counts(const u32* cv, const SolidityParams& sp, int32_t* sum_out, u32* bin1, u32* bin2)
{
    const int nb = sp.nbanks;
    int32_t sum = 0;
    if (nb == 1) sum = (int32_t)cv[0];
    else for (int b = 0; b < nb; b++) if (sp.kind != 5 || sp.solid_vec[b]) sum += (int32_t)cv[b];   // CountProcessorChain.hpp:158-169
    *sum_out = sum;
    *bin1 = histo_bin(sum);
    *bin2 = H2_NONE;
    if (sp.bank_hist) {                                              // CountProcessorCutoff.hpp:113-116: bank i sees count[i]
        for (int b = 0; b < nb; b++) { const u32 bb = histo_bin((int32_t)cv[b]); if (bb) atomicAdd(&sp.bank_hist[(size_t)b * 10001u + bb], 1ULL); }
    }
    if (sp.histo2d) {                                                // Histogram.hpp:92-98 inc2D, clamps of SURVEY 8(a)-13(v)
        u32 i1 = (u32)(sum - (int32_t)cv[0]) & 0xFFFFu, i2 = cv[0] & 0xFFFFu;
        if (i1 >= 10000u) i1 = 10000u;
        if (i2 >= 10u) i2 = 10u;
        *bin2 = i1 + 10001u * i2;
    }
    auto inr = [&](long long x, long long lo) { return lo <= x && x <= sp.amax; };
    switch (sp.kind) {
    case 0: return inr(sum, sp.amin[0]);
    case 1: { u32 v = cv[0]; for (int b = 1; b < nb; b++) v = min(v, cv[b]); return inr(v, sp.amin[0]); }
    case 2: { u32 v = cv[0]; for (int b = 1; b < nb; b++) v = max(v, cv[b]); return inr(v, sp.amin[0]); }
    case 3: { for (int b = 0; b < nb; b++) if (inr(cv[b], sp.amin[b])) return true; return false; }
    case 4: { for (int b = 0; b < nb; b++) if (!inr(cv[b], sp.amin[b])) return false; return true; }
    default: { for (int b = 0; b < nb; b++) { bool in = inr(cv[b], sp.amin[b]); if (sp.solid_vec[b] != in) return false; } return true; }
    }
}

__device__ __forceinline__ void flush_hist(const u32* s_hist, unsigned long long* g_hist)
{
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) { u32 v = s_hist[i]; if (v) atomicAdd(&g_hist[i], (unsigned long long)v); }
}

// append the solid (k-mer, abundance) pairs of a whole block with ONE bump of the global cursor: every thread
// contributes n <= MAXN entries; a same-address atomic per warp would serialise in L2 (measured: 10 ms per step).
// Must be called by all threads of the block; `par` alternates between consecutive calls (double-buffered smem).
struct AppendSmem { u32 wsum[2][8]; unsigned long long base[2]; };

template <int KW, int MAXN>
__device__ __forceinline__ void block_append(u32 mask /*bit i: entry i is valid*/, const Kmer<KW>* k, const int32_t* v, u64* out_keys,
                                             u32* out_vals, u64 out_cap, Counters* ctr, AppendSmem* sm, int par)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n = (u32)__popc(mask);
    u32 inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) sm->wsum[par][warp] = inc;
    __syncthreads();
    u32 wpre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { u32 x = sm->wsum[par][w]; if (w < warp) wpre += x; tot += x; }
    if (tot == 0) return;                                          // uniform across the block
    if (threadIdx.x == 0) sm->base[par] = atomicAdd(&ctr->solid_n, (unsigned long long)tot);
    __syncthreads();
    u64 pos = sm->base[par] + wpre + inc - n;
#pragma unroll
    for (int i = 0; i < MAXN; i++) {
        if ((mask >> i) & 1u) {
            if (pos < out_cap) {
#pragma unroll
                for (int q = 0; q < KW; q++) out_keys[pos * KW + q] = k[i].w[q];
                out_vals[pos] = (u32)v[i];
            } else atomicExch(&ctr->overflow, 2u);
            pos++;
        }
    }
}

// ---- 128-bit compare-and-swap (sm_90+: ATOMG.E.CAS.128) ------------------------------------------------------
__device__ __forceinline__ void cas128(u64* addr, u64 clo, u64 chi, u64 slo, u64 shi, u64& olo, u64& ohi)
{
    asm volatile("{\n\t.reg .b128 c, s, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\t"
                 "atom.global.relaxed.gpu.cas.b128 d, [%6], c, s;\n\tmov.b128 {%0, %1}, d;\n\t}"
                 : "=l"(olo), "=l"(ohi) : "l"(clo), "l"(chi), "l"(slo), "l"(shi), "l"(addr) : "memory");
}

// 32-byte (one L2 sector) load, L2-coherent: LDG.E.ENL2.256 on sm_100
__device__ __forceinline__ void ld256cg(const u64* p, u64& a, u64& b, u64& c, u64& d)
{
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}

// find-or-claim the slot of `key`; returns slot or 0xFFFFFFFF on overflow.
// Bucketed linear probing: a bucket is one 32-byte sector (4 x 64-bit keys or 2 x 128-bit keys), fetched with a
// single 256-bit load, so a probe costs one L2 round trip and ~all k-mers resolve in their first bucket.
// Slots only ever go EMPTY -> key during the insert phase, which makes a stale snapshot safe: a non-empty slot is
// final, an empty-looking slot is validated by the CAS that claims it (the CAS returns the truth).
__device__ __forceinline__ u32 table_slot(u64* keys, u32 smask, const Kmer<1>& key)
{
    const u64 EMPTY = ~0ULL;
    const u32 bmask = smask >> 2;
    u32 b = (u32)kmer_hash(key) & bmask;
    for (u32 probe = 0; probe < HASH_MAX_PROBE; probe++) {
        u64 kk[4];
        ld256cg(keys + 4 * (u64)b, kk[0], kk[1], kk[2], kk[3]);
#pragma unroll
        for (int i = 0; i < 4; i++) if (kk[i] == key.w[0]) return 4 * b + i;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (kk[i] == EMPTY) {
                u64 old = atomicCAS((unsigned long long*)&keys[4 * (u64)b + i], EMPTY, key.w[0]);
                if (old == EMPTY || old == key.w[0]) return 4 * b + i;
            }
        }
        b = (b + 1) & bmask;
    }
    return 0xFFFFFFFFu;
}
__device__ __forceinline__ u32 table_slot(u64* keys, u32 smask, const Kmer<2>& key)
{
    const u64 EMPTY = ~0ULL;
    const u32 bmask = smask >> 1;
    u32 b = (u32)kmer_hash(key) & bmask;
    for (u32 probe = 0; probe < HASH_MAX_PROBE; probe++) {
        u64 kk[4];
        ld256cg(keys + 4 * (u64)b, kk[0], kk[1], kk[2], kk[3]);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            // a 16-byte key may be read torn against a concurrent claim; the snapshot is only trusted when it shows
            // a complete foreign key (neither half all-ones), which can never change again
            const u64 lo = kk[2 * i], hi = kk[2 * i + 1];
            const bool foreign = (lo != EMPTY) && (hi != EMPTY) && !(lo == key.w[0] && hi == key.w[1]);
            if (!foreign) {
                u64 olo, ohi;
                cas128(keys + 4 * (u64)b + 2 * i, EMPTY, EMPTY, key.w[0], key.w[1], olo, ohi);
                if ((olo == EMPTY && ohi == EMPTY) || (olo == key.w[0] && ohi == key.w[1])) return 2 * b + i;
            }
        }
        b = (b + 1) & bmask;
    }
    return 0xFFFFFFFFu;
}

// ---- K4+K7: expand records and count into the table ------------------------------------------------------------
// Flat k-mer parallelism: a warp stages 32 records in shared memory, builds the (record, offset) owner map of
// their k-mers, then every lane extracts ONE k-mer per iteration straight from the packed bases (no rolling,
// no divergence on the super-k-mer length).  Canonical form via bit-reversal (K/Model.hpp:294, :877-884).
template <int KW> struct InsCfg { static constexpr int MAXNK = 32 * 2 * KW - 8; };   // k-mers per record upper bound

template <int KW>
__device__ __forceinline__ Kmer<KW> rec_kmer_at(const u64* r, int j, int k)
{
    // k-mer starting at base j of the record
    if constexpr (KW == 1) {
        const int q = j >> 5, o = j & 31;                      // j >= 32 only happens for very small k
        u64 a = r[q], b = q ? 0 : r[1];
        u64 v = o ? ((a << (2 * o)) | (b >> (64 - 2 * o))) : a;
        Kmer<1> x; x.w[0] = v >> (64 - 2 * k); return x;
    } else {
        const int q = j >> 5, o = j & 31;
        u64 a = r[q], b = (q + 1 < 4) ? r[q + 1] : 0, c = (q + 2 < 4) ? r[q + 2] : 0;
        u64 sh[2];
        sh[0] = o ? ((a << (2 * o)) | (b >> (64 - 2 * o))) : a;
        sh[1] = o ? ((b << (2 * o)) | (c >> (64 - 2 * o))) : b;
        return rec_first_kmer2(sh, k);
    }
}

template <int KW>
__global__ void __launch_bounds__(256) k_hash_insert(const u64* __restrict__ recs, u64 rec_begin, u64 rec_end, int k,
                                                     u64* keys, u32* counts, u32 smask, int nbanks, Counters* ctr,
                                                     const unsigned long long* nrec_dev = nullptr /*optional: record count that lives on the device*/)
{
    constexpr int RW = 2 * KW;
    constexpr int MAXNK = InsCfg<KW>::MAXNK;
    if (nrec_dev) { const u64 e = rec_begin + *nrec_dev; if (e < rec_end) rec_end = e; }
    __shared__ __align__(16) u64 s_rec[8][32 * RW];
    __shared__ u8 s_owner[8][32 * MAXNK];
    __shared__ u16 s_off[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 nwarps = (u64)gridDim.x * 8, gwarp = (u64)blockIdx.x * 8 + warp;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(recs);
    for (u64 base = rec_begin + gwarp * 32; base < rec_end; base += nwarps * 32) {
        const u64 i = base + lane;
        u32 nk = 0;
        if (i < rec_end) {
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(&s_rec[warp][lane * RW]);
            if constexpr (RW == 2) { ulonglong2 v = __ldg(src + i); dst[0] = v; nk = (u32)(v.y >> 8) & 0xFFu; }
            else { ulonglong2 v = __ldg(src + 2 * i), u = __ldg(src + 2 * i + 1); dst[0] = v; dst[1] = u; nk = (u32)(u.y >> 8) & 0xFFu; }
        }
        u32 inc = nk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
        const u32 total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const u32 off = inc - nk;
        s_off[warp][lane] = (u16)off;
        for (u32 j = 0; j < nk; j++) s_owner[warp][off + j] = (u8)lane;
        __syncwarp();
        for (u32 t = lane; t < total; t += 32) {
            const u32 r = s_owner[warp][t];
            const int j = (int)(t - s_off[warp][r]);
            const u64* rw = &s_rec[warp][r * RW];
            Kmer<KW> f = rec_kmer_at<KW>(rw, j, k);
            Kmer<KW> c = kmer_canonical(f, kmer_revcomp(f, k));
            const int bank = (nbanks > 1) ? (int)(rw[RW - 1] & 0xFu) : 0;
            u32 slot = table_slot(keys, smask, c);
            if (slot == 0xFFFFFFFFu) { atomicExch(&ctr->hash_overflow, 1u); break; }
            atomicAdd(&counts[(u64)slot * nbanks + bank], 1u);
        }
        __syncwarp();
    }
}

// ---- K6 (hash flavour): sweep the table, run the processor chain, reset the slots ---------------------------------
// each thread owns SV consecutive slots per iteration; keys and (single-bank) counts come in as independent
// 16-byte vector loads so one iteration costs one memory round trip
constexpr int HLL_BITS = 12;
constexpr u32 HLL_M = 1u << HLL_BITS;
// cardinality from HyperLogLog registers (host side; 64-bit hashes: no large-range correction; linear counting when sparse)
inline double hll_estimate(const u32* reg)
{
    double z = 0.0; u32 zeros = 0;
    for (u32 i = 0; i < HLL_M; i++) { z += std::ldexp(1.0, -(int)reg[i]); zeros += reg[i] == 0; }
    const double m = (double)HLL_M, alpha = 0.7213 / (1.0 + 1.079 / m);
    double e = alpha * m * m / z;
    if (e <= 2.5 * m && zeros) e = m * std::log(m / (double)zeros);
    return e;
}

template <int KW, bool NB1>
__global__ void __launch_bounds__(256) k_hash_scan(u64* keys, u32* counts, u32 nslots, SolidityParams sp, int discard,
                                                   u64* out_keys, u32* out_vals, u64 out_cap,
                                                   unsigned long long* g_hist, unsigned long long* g_hist2d, Counters* ctr,
                                                   u32* __restrict__ hll = nullptr /*density sample: HyperLogLog registers [HLL_M] of its distinct k-mers*/)
{
    constexpr int SV = (KW == 1) ? 8 : 4;                          // slots per thread-iteration
    constexpr int KV = SV * KW / 2;                                // 16-byte key vectors per iteration (4)
    constexpr int CV = SV / 4;                                     // 16-byte count vectors (NB1 only)
    __shared__ u32 s_hist[HIST_SMEM_BINS];
    __shared__ u32 s_h2[NB1 ? 1 : 11 * H2_SMEM_I1];
    __shared__ u32 s_distinct;
    __shared__ AppendSmem s_app;
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) s_hist[i] = 0;
    if (!NB1) for (int i = threadIdx.x; i < 11 * H2_SMEM_I1; i += blockDim.x) s_h2[i] = 0;
    if (threadIdx.x == 0) s_distinct = 0;
    __syncthreads();
    const u64 EMPTY = ~0ULL;
    u32 ndist = 0, nsamp_solid = 0;
    unsigned long long sumsq = 0;
    const u32 ngroups = nslots / SV;                               // nslots is a power of two >= 1024
    const u32 nthreads = blockDim.x * gridDim.x;
    const u32 nloop = (ngroups + nthreads - 1) / nthreads;
    ulonglong2* k2 = reinterpret_cast<ulonglong2*>(keys);
    uint4* c4 = reinterpret_cast<uint4*>(counts);
    for (u32 it = 0; it < nloop; it++) {
        const u32 g = it * nthreads + blockIdx.x * blockDim.x + threadIdx.x;
        const bool in = g < ngroups;
        ulonglong2 v[KV]; uint4 cvec[CV > 0 ? CV : 1];
#pragma unroll
        for (int q = 0; q < KV; q++) v[q] = in ? k2[(u64)g * KV + q] : make_ulonglong2(EMPTY, EMPTY);
        if (NB1) {
#pragma unroll
            for (int q = 0; q < CV; q++) cvec[q] = in ? c4[(u64)g * CV + q] : make_uint4(0, 0, 0, 0);
        }
        u32 occm = 0;
        Kmer<KW> sk[SV]; int32_t sv[SV]; u32 solidm = 0;
#pragma unroll
        for (int q = 0; q < SV; q++) {
            if constexpr (KW == 1) { sk[q].w[0] = (q & 1) ? v[q >> 1].y : v[q >> 1].x; if (sk[q].w[0] != EMPTY) occm |= 1u << q; }
            else { sk[q].w[0] = v[q].x; sk[q].w[1] = v[q].y; if (!(v[q].x == EMPTY && v[q].y == EMPTY)) occm |= 1u << q; }
            sv[q] = 0;
        }
        if (occm) {
#pragma unroll
            for (int q = 0; q < KV; q++) k2[(u64)g * KV + q] = make_ulonglong2(EMPTY, EMPTY);
            if (NB1) {
#pragma unroll
                for (int q = 0; q < CV; q++) c4[(u64)g * CV + q] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int q = 0; q < SV; q++) {
            if (!((occm >> q) & 1u)) continue;
            u32 cv[MAXB];
            if (NB1) { const uint4 c = cvec[q >> 2]; cv[0] = (q & 3) == 0 ? c.x : (q & 3) == 1 ? c.y : (q & 3) == 2 ? c.z : c.w; }
            else {
                const u64 slot = (u64)g * SV + q;
                for (int b = 0; b < sp.nbanks; b++) { cv[b] = counts[slot * sp.nbanks + b]; counts[slot * sp.nbanks + b] = 0; }
            }
            if (discard == 2) {                                    // density sample: occupied slots, multiplicities, share reaching the threshold
                ndist++; sumsq += (unsigned long long)cv[0] * cv[0];
                if ((long long)cv[0] >= sp.amin[0] && (long long)cv[0] <= sp.amax) nsamp_solid++;
                if (hll) {                                         // the ranks of a multi-GPU job merge these (max) to get the distinct k-mers of the UNION of their samples
                    const u64 h = kmer_hash(sk[q]), w = h << HLL_BITS;
                    atomicMax(&hll[h >> (64 - HLL_BITS)], w ? (u32)__clzll((long long)w) + 1u : (u32)(64 - HLL_BITS + 1));
                }
            }
            if (!discard) {
                ndist++;
                int32_t sum = 0;
                if (process_counts(cv, sp, s_hist, g_hist, g_hist2d, &sum, NB1 ? nullptr : s_h2)) { sv[q] = sum; solidm |= 1u << q; }
            }
        }
        block_append<KW, SV>(solidm, sk, sv, out_keys, out_vals, out_cap, ctr, &s_app, (int)(it & 1));
    }
    ndist = __reduce_add_sync(0xFFFFFFFFu, ndist);
    if ((threadIdx.x & 31) == 0 && ndist) atomicAdd(&s_distinct, ndist);
    __syncthreads();
    flush_hist(s_hist, g_hist);
    if (!NB1 && sp.histo2d) flush_hist2d(s_h2, g_hist2d);
    if (threadIdx.x == 0 && s_distinct) atomicAdd(discard == 2 ? &ctr->sample_distinct : &ctr->distinct_n, (unsigned long long)s_distinct);
    if (discard == 2) {
        for (int d = 16; d; d >>= 1) sumsq += __shfl_xor_sync(0xFFFFFFFFu, sumsq, d);
        if ((threadIdx.x & 31) == 0 && sumsq) atomicAdd(&ctr->sample_sumsq, sumsq);
        nsamp_solid = __reduce_add_sync(0xFFFFFFFFu, nsamp_solid);
        if ((threadIdx.x & 31) == 0 && nsamp_solid) atomicAdd(&ctr->sample_solid, (unsigned long long)nsamp_solid);
    }
}

__global__ void k_fill_u64(u64* p, u64 n, u64 v)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) p[i] = v;
}

// ---- K4 (sort flavour): expand records into a flat key array -------------------------------------------------------
template <int KW>
__global__ void __launch_bounds__(256) k_expand_keys(const u64* __restrict__ recs, u64 rec_begin, u64 rec_end, int k,
                                                     u64* __restrict__ out_keys, u32* __restrict__ out_bank, int nbanks, Counters* ctr)
{
    constexpr int RW = 2 * KW;
    const int lane = threadIdx.x & 31;
    const u64 nrec = rec_end - rec_begin;
    const u64 nloop = (nrec + (u64)blockDim.x * gridDim.x - 1) / ((u64)blockDim.x * gridDim.x);
    for (u64 it = 0; it < nloop; it++) {
        u64 i = rec_begin + (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        u64 r[RW]; int nk = 0;
        if (i < rec_end) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(recs);
            if constexpr (RW == 2) { ulonglong2 v = __ldg(src + i); r[0] = v.x; r[1] = v.y; }
            else if constexpr (RW == 4) { ulonglong2 v = __ldg(src + 2 * i), u = __ldg(src + 2 * i + 1); r[0] = v.x; r[1] = v.y; r[2] = u.x; r[3] = u.y; }
            else {
#pragma unroll
                for (int x = 0; x < RW / 2; x++) { const ulonglong2 v = __ldg(src + (RW / 2) * i + x); r[2 * x] = v.x; r[2 * x + 1] = v.y; }
            }
            nk = (int)((r[RW - 1] >> 8) & 0xFFu);
        }
        u32 inc = (u32)nk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
        u32 wtot = __shfl_sync(0xFFFFFFFFu, inc, 31);
        unsigned long long base = 0;
        if (lane == 0 && wtot) base = atomicAdd(&ctr->expand_cursor, (unsigned long long)wtot);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (nk == 0) continue;
        u64 pos = base + inc - nk;
        const u32 bank = (u32)(r[RW - 1] & 0xFu);
        Kmer<KW> f, rc;
        if constexpr (KW == 1) f = rec_first_kmer1(r, k); else if constexpr (KW == 2) f = rec_first_kmer2(r, k); else f = rec_first_kmer<KW>(r, k);
        rc = kmer_revcomp(f, k);
        for (int j = 0; j < nk; j++) {
            if (j) kmer_roll(f, rc, rec_base<RW>(r, k - 1 + j), k);
            Kmer<KW> c = kmer_canonical(f, rc);
#pragma unroll
            for (int q = 0; q < KW; q++) out_keys[(pos + j) * KW + q] = c.w[q];
            if (nbanks > 1) out_bank[pos + j] = bank;
        }
    }
}

// ---- heavy minimizer bins: expand records into hash buckets of flat keys ------------------------------------------------
// A partition is bounded below by its heaviest minimizer bin (9e-5 of the job at k=31, m=10): in multi-G k-mer jobs most
// bins outgrow a shared-memory table.  Those partitions are expanded into canonical k-mers (what ReadSuperKCommand does,
// K/PartitionsCommand.cpp:944-1128) and every k-mer is appended to the slab of bucket = hash(k-mer) * S >> 32: equal
// k-mers meet in one bucket, a bucket holds what one shared-memory table takes, and the buckets are then counted by
// k_count_smem<KW, MB, true> straight from the flat keys.  Extra HBM traffic: one key written + read per k-mer (16 B at
// k <= 31) instead of the global table's random L2 atomics.  With per-bank counts the bank id rides in the two spare top
// bits of the key (2k <= 62 resp. 126 bits are used, nb <= 4).
__device__ __forceinline__ u32 bucket_hash(u32 h)      // independent of the bits that pick the slot (top) and the sub-pass (8..)
{
    h = (h ^ (h >> 16)) * 0xC2B2AE3Du; h = (h ^ (h >> 13)) * 0x27D4EB2Fu; return h ^ (h >> 16);
}

template <int KW> __device__ __forceinline__ u32 cs_hash_of(const Kmer<KW>& a);   // count_smem.cuh

template <int KW>
__global__ void __launch_bounds__(256) k_expand_bucket(const u64* __restrict__ recs, u64 rec_begin, u64 rec_end, int k, int nbanks,
                                                       u32 nbuckets, u32 slab, u64* __restrict__ out_keys, u32* __restrict__ cursors, Counters* ctr)
{
    constexpr int RW = 2 * KW;
    constexpr int MAXNK = InsCfg<KW>::MAXNK;
    __shared__ __align__(16) u64 s_rec[8][32 * RW];
    __shared__ u8 s_owner[8][32 * MAXNK];
    __shared__ u16 s_off[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 nwarps = (u64)gridDim.x * 8, gwarp = (u64)blockIdx.x * 8 + warp;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(recs);
    for (u64 base = rec_begin + gwarp * 32; base < rec_end; base += nwarps * 32) {
        const u64 i = base + lane;
        u32 nk = 0;
        if (i < rec_end) {
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(&s_rec[warp][lane * RW]);
            if constexpr (RW == 2) { ulonglong2 v = __ldg(src + i); dst[0] = v; nk = (u32)(v.y >> 8) & 0xFFu; }
            else { ulonglong2 v = __ldg(src + 2 * i), u = __ldg(src + 2 * i + 1); dst[0] = v; dst[1] = u; nk = (u32)(u.y >> 8) & 0xFFu; }
        }
        u32 inc = nk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
        const u32 total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const u32 off = inc - nk;
        s_off[warp][lane] = (u16)off;
        for (u32 j = 0; j < nk; j++) s_owner[warp][off + j] = (u8)lane;
        __syncwarp();
        for (u32 t = lane; t < total; t += 32) {
            const u32 r = s_owner[warp][t];
            const int j = (int)(t - s_off[warp][r]);
            const u64* rw = &s_rec[warp][r * RW];
            Kmer<KW> f = rec_kmer_at<KW>(rw, j, k);
            Kmer<KW> c = kmer_canonical(f, kmer_revcomp(f, k));
            const u32 b = __umulhi(bucket_hash(cs_hash_of<KW>(c)), nbuckets);
            const u32 pos = atomicAdd(&cursors[b], 1u);
            if (pos < slab) {
                if (nbanks > 1) c.w[KW - 1] |= (u64)(rw[RW - 1] & 3u) << 62;
                u64* o = out_keys + ((u64)b * slab + pos) * KW;
                if constexpr (KW == 1) o[0] = c.w[0];
                else *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(c.w[0], c.w[1]);
            } else atomicExch(&ctr->bucket_overflow, 1u);
        }
        __syncwarp();
    }
}

// ---- K6 (sort flavour): run-length count over sorted keys ------------------------------------------------------------
// the thread at the head of a run gallops to its end (sorted input => O(log count) probes)
template <int KW>
__global__ void __launch_bounds__(256) k_rle_emit(const u64* __restrict__ keys, const u32* __restrict__ banks, u64 n,
                                                  SolidityParams sp, u64* out_keys, u32* out_vals, u64 out_cap,
                                                  unsigned long long* g_hist, unsigned long long* g_hist2d, Counters* ctr)
{
    __shared__ u32 s_hist[HIST_SMEM_BINS];
    __shared__ u32 s_h2[11 * H2_SMEM_I1];
    __shared__ u32 s_distinct;
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) s_hist[i] = 0;
    for (int i = threadIdx.x; i < 11 * H2_SMEM_I1; i += blockDim.x) s_h2[i] = 0;
    if (threadIdx.x == 0) s_distinct = 0;
    __syncthreads();
    auto load = [&](u64 i) { Kmer<KW> x;
#pragma unroll
        for (int q = 0; q < KW; q++) x.w[q] = keys[i * KW + q];
        return x; };
    constexpr int RI = 4;                                          // keys per thread per iteration
    __shared__ AppendSmem s_app;
    u32 ndist = 0;
    const u64 per_it = (u64)blockDim.x * gridDim.x * RI;
    const u64 nloop = (n + per_it - 1) / per_it;
    for (u64 it = 0; it < nloop; it++) {
        Kmer<KW> sk[RI]; int32_t sv[RI]; u32 solidm = 0;
#pragma unroll
        for (int r = 0; r < RI; r++) {
            const u64 i = it * per_it + ((u64)r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
            if (i >= n) continue;
            const Kmer<KW> key = load(i);
            sk[r] = key; sv[r] = 0;
            const bool head = (i == 0) || !kmer_eq(load(i - 1), key);
            if (!head) continue;
            u64 step = 1, lo = i;                                  // keys[lo] == key ; find last equal
            while (lo + step < n && kmer_eq(load(lo + step), key)) { lo += step; step <<= 1; }
            u64 hi = (lo + step < n) ? lo + step : n;              // keys[hi] != key or hi == n
            while (hi - lo > 1) { u64 mid = lo + (hi - lo) / 2; if (kmer_eq(load(mid), key)) lo = mid; else hi = mid; }
            const u64 cnt = lo - i + 1;
            u32 cv[MAXB]; int32_t sum = 0;
            if (sp.nbanks == 1) cv[0] = (u32)cnt;
            else { for (int b = 0; b < sp.nbanks; b++) cv[b] = 0; for (u64 j = i; j <= lo; j++) cv[banks[j]]++; }
            ndist++;
            if (process_counts(cv, sp, s_hist, g_hist, g_hist2d, &sum, s_h2)) { sv[r] = sum; solidm |= 1u << r; }
        }
        block_append<KW, RI>(solidm, sk, sv, out_keys, out_vals, out_cap, ctr, &s_app, (int)(it & 1));
    }
    ndist = __reduce_add_sync(0xFFFFFFFFu, ndist);
    if ((threadIdx.x & 31) == 0 && ndist) atomicAdd(&s_distinct, ndist);
    __syncthreads();
    flush_hist(s_hist, g_hist);
    if (sp.histo2d) flush_hist2d(s_h2, g_hist2d);
    if (threadIdx.x == 0 && s_distinct) atomicAdd(&ctr->distinct_n, (unsigned long long)s_distinct);
}

// ---- sort path of the wide spans (192 / 256-bit keys): order by a 64-bit hash, verify, count ---------------------------------
// A full-width LSD sort of 24 / 32-byte keys costs 24 - 32 passes over the whole array (k = 95: 104 of the 116 ms of a 187 M
// k-mer step, profiles/r04j).  Equal k-mers only have to MEET, not to be in k-mer order (the solid set is ordered afterwards):
// the keys stay where k_expand_keys wrote them, a (64-bit hash, index) pair per key is sorted in 8 passes, and
//   k_verify_hashed : every element compares its FULL key with its predecessor's when their hashes are equal -- if any pair
//                     differs, two different k-mers share a hash (probability ~2^-8 per 2^28-key group) and the group is
//                     redone by the full-width sort: the result is exact either way;
//   k_rle_emit_hashed: heads of hash runs gallop to the run end on the hash array, gather their key through the index, run the
//                     processor chain (per-bank counts through the index as well).
template <int KW>
__global__ void __launch_bounds__(256) k_hash_keys(const u64* __restrict__ keys, u64 n, u64* __restrict__ hashes, u32* __restrict__ idx, u64 hmask)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        Kmer<KW> x;
#pragma unroll
        for (int q = 0; q < KW; q++) x.w[q] = keys[i * KW + q];
        hashes[i] = kmer_hash(x) & hmask;                          // (hmask: all ones; the tests narrow it to provoke collisions)
        idx[i] = (u32)i;
    }
}

template <int KW>
__global__ void __launch_bounds__(256) k_verify_hashed(const u64* __restrict__ hashes, const u32* __restrict__ idx, const u64* __restrict__ keys, u64 n,
                                                       unsigned int* flag)
{
    bool bad = false;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (hashes[i] != hashes[i - 1]) continue;
        const u64 a = idx[i], b = idx[i - 1];
#pragma unroll
        for (int q = 0; q < KW; q++) bad |= keys[a * KW + q] != keys[b * KW + q];
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1u);
}

template <int KW>
__global__ void __launch_bounds__(256) k_rle_emit_hashed(const u64* __restrict__ hashes, const u32* __restrict__ idx, const u64* __restrict__ keys,
                                                         const u32* __restrict__ banks, u64 n, SolidityParams sp, u64* out_keys, u32* out_vals, u64 out_cap,
                                                         unsigned long long* g_hist, unsigned long long* g_hist2d, Counters* ctr)
{
    __shared__ u32 s_hist[HIST_SMEM_BINS];
    __shared__ u32 s_h2[11 * H2_SMEM_I1];
    __shared__ u32 s_distinct;
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) s_hist[i] = 0;
    for (int i = threadIdx.x; i < 11 * H2_SMEM_I1; i += blockDim.x) s_h2[i] = 0;
    if (threadIdx.x == 0) s_distinct = 0;
    __syncthreads();
    constexpr int RI = 4;
    __shared__ AppendSmem s_app;
    u32 ndist = 0;
    const u64 per_it = (u64)blockDim.x * gridDim.x * RI;
    const u64 nloop = (n + per_it - 1) / per_it;
    for (u64 it = 0; it < nloop; it++) {
        Kmer<KW> sk[RI]; int32_t sv[RI]; u32 solidm = 0;
#pragma unroll
        for (int r = 0; r < RI; r++) {
            const u64 i = it * per_it + ((u64)r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
            sv[r] = 0;
            if (i >= n) continue;
            const u64 h = hashes[i];
            if (i != 0 && hashes[i - 1] == h) continue;            // not the head of its run
            u64 step = 1, lo = i;                                  // hashes[lo] == h ; find the last equal
            while (lo + step < n && hashes[lo + step] == h) { lo += step; step <<= 1; }
            u64 hi = (lo + step < n) ? lo + step : n;
            while (hi - lo > 1) { const u64 mid = lo + (hi - lo) / 2; if (hashes[mid] == h) lo = mid; else hi = mid; }
            const u64 cnt = lo - i + 1;
            u32 cv[MAXB]; int32_t sum = 0;
            if (sp.nbanks == 1) cv[0] = (u32)cnt;
            else { for (int b = 0; b < sp.nbanks; b++) cv[b] = 0; for (u64 j = i; j <= lo; j++) cv[banks[idx[j]]]++; }
            ndist++;
            if (process_counts(cv, sp, s_hist, g_hist, g_hist2d, &sum, s_h2)) {
                const u64 a = idx[i];
#pragma unroll
                for (int q = 0; q < KW; q++) sk[r].w[q] = keys[a * KW + q];
                sv[r] = sum; solidm |= 1u << r;
            }
        }
        block_append<KW, RI>(solidm, sk, sv, out_keys, out_vals, out_cap, ctr, &s_app, (int)(it & 1));
    }
    ndist = __reduce_add_sync(0xFFFFFFFFu, ndist);
    if ((threadIdx.x & 31) == 0 && ndist) atomicAdd(&s_distinct, ndist);
    __syncthreads();
    flush_hist(s_hist, g_hist);
    if (sp.histo2d) flush_hist2d(s_h2, g_hist2d);
    if (threadIdx.x == 0 && s_distinct) atomicAdd(&ctr->distinct_n, (unsigned long long)s_distinct);
}

#endif  // __CUDACC__
}  // namespace dsk
