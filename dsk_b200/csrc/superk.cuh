// superk.cuh -- K2: code stream -> minimizers -> super-k-mer records;  K3: partition histogram + scatter.
//
// Replaces (K/ = thirdparty/gatb-core/gatb-core/src/gatb/kmer/impl/):
//   ModelCanonical::next / ModelMinimizer::next       K/Model.hpp:877-884, :1106-1139
//   Sequence2SuperKmer::KmerFunctor                   K/Sequence2SuperKmer.hpp:90-133
//   FillPartitions::processSuperkmer, SuperKmer::save K/SortingCountAlgorithm.cpp:1084-1154, K/Model.hpp:1386-1471
//
// The reference walks each read sequentially (rolling k-mer + "did the minimizer fall out" state).  Here the
// minimizer of a window is recomputed as a pure function (it is one: K/Model.hpp:1254-1287), so every
// position of the code stream is independent: a tile of 2048 positions is packed to 2 bits in shared memory,
// every thread evaluates 8 consecutive windows (m-mer values shared through smem, sliding minimum with a
// prefix/suffix split), and super-k-mer boundaries are found with bit scans over per-tile bitmaps.
// Super-k-mers never span tiles or invalid windows; their boundaries are not observable in the output.
#pragma once
#include "kmer_bits.cuh"
#include "scan.cuh"

namespace dsk {

constexpr int SK_THREADS = 256;
constexpr int SK_PPT = 8;                         // positions per thread
constexpr int SK_TP = SK_THREADS * SK_PPT;        // 2048 positions per tile
constexpr int SK_HALO = 64;                       // >= k-1 (k <= 63); the wide spans (k <= 127) take 128: sk_halo<KW>()
template <int KW> DSK_HD constexpr int sk_halo() { return KW <= 2 ? SK_HALO : 128; }
constexpr u32 SK_NOMIN = 0xFFFFFFFFu;

struct Counters {
    unsigned long long nrec;          // super-k-mer records written
    unsigned long long kmers_valid;   // valid k-mers seen
    unsigned long long kmers_in_recs; // sum of nk over records (must equal kmers_valid)
    unsigned long long solid_n;       // solid (k-mer,count) pairs emitted
    unsigned long long distinct_n;    // distinct k-mers seen by the counters
    unsigned int overflow;            // record buffer overflow
    unsigned int hash_overflow;
    unsigned long long expand_cursor; // sort path: keys written
    unsigned int smem_splits;         // shared-memory path: table overflows answered by splitting a pass in two
    unsigned int smem_failed;         // shared-memory path: passes that still overflowed at the deepest split
    unsigned long long sample_nrec;   // density sample: records / k-mers selected, distinct k-mers found among them
    unsigned long long sample_nkm;
    unsigned long long sample_distinct;
    unsigned int bucket_overflow;     // key-bucket path: a hash bucket outgrew its slab (one k-mer with a huge multiplicity)
    unsigned int pad0;
    unsigned long long sample_sumsq;  // density sample: sum of count^2 over its distinct k-mers (occurrence-weighted multiplicity)
    unsigned long long kmers_pass;    // valid k-mers whose minimizer belongs to the current pass (== kmers_in_recs)
    unsigned long long sample_solid;  // density sample: distinct k-mers whose summed count reaches the smallest abundance-min
    unsigned int hist_suspect;        // a fine bin holds so many k-mers that its packed record count may have wrapped (k_check_bins)
    unsigned int sort_fallback;       // ordering of the solid set: a prefix group too long for the neighbourhood fix-up (k_rs_fix)
};

// fine-bin histogram entry: records in the top 28 bits, k-mers in the low 36 -- ONE 64-bit RED per record instead of two, and
// half the L2 footprint (32 MB for 2^22 bins).  A record holds at least one k-mer, so a bin whose k-mer field is below 2^28
// cannot have wrapped its record field; anything else (one minimizer with > 268 M k-mers on one GPU: degenerate input) is
// flagged by k_check_bins and the histogram is rebuilt exactly from the record meta (k_rebuild_hist).
constexpr int BINH_KBITS = 36;
constexpr unsigned long long BINH_KMASK = (1ULL << BINH_KBITS) - 1ULL;

#ifdef __CUDACC__

__device__ __forceinline__ u32 sk_pidx(u32 p) { return p + (p >> 3); }     // padded smem index (conflict-free strips)

template <int KW>
__global__ void __launch_bounds__(SK_THREADS) k_superkmers(const u8* __restrict__ codes, const StreamState* __restrict__ ss,
                                                           int k, int m, int bank, u64* __restrict__ recs,
                                                           u32* __restrict__ rec_meta, u64 rec_cap, Counters* ctr,
                                                           unsigned long long* __restrict__ bin_hist /*[1 << fine_log2] packed: records << 36 | k-mers*/,
                                                           u32 nb_passes, u32 pass_id, int fine_shift /*24 - fine_log2*/)
{
    constexpr int RW = 2 * KW;
    constexpr int HALO = sk_halo<KW>();
    __shared__ u64 s_pk[(SK_TP + HALO) / 32 + 6];                // 2-bit bases, MSB first, 32 per word
    __shared__ u64 s_bad[(SK_TP + HALO) / 64 + 2 + (KW > 2 ? 1 : 0)];   // invalid flags, MSB first, 64 per word
    __shared__ u32 s_mv[SK_TP + HALO + (SK_TP + HALO) / 8 + 8];
    __shared__ u32 s_start[SK_TP / 32];                          // super-k-mer run starts, MSB first
    __shared__ u32 s_brk[SK_TP / 32 + 9];                        // run start or invalid window
    __shared__ u32 s_last[SK_THREADS];                           // minimizer of each thread's last window
    __shared__ u32 s_wsum[SK_THREADS / 32];
    __shared__ unsigned long long s_goff;
    __shared__ u32 s_nvalid;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const u64 total = ss->total;
    const u64 limit = (total >= (u64)k) ? total - (u64)(k - 1) : 0;   // windows starting at >= limit are not complete
    const u64 tile0 = (u64)blockIdx.x * SK_TP;
    if (tile0 >= limit) return;
    const int w = k - m + 1;                                     // m-mers per window
    const int maxS = rec_max_kmers(KW, k);
    const u32 mmask = (1u << (2 * m)) - 1u;

    // ---- 1. load 8 codes per thread, pack to 2 bits + invalid bitmap ---------------------------------
    auto load8 = [&](int ti) {
        const uint2 v = *reinterpret_cast<const uint2*>(codes + tile0 + 8 * (u64)ti);
        // SWAR: gather the 2-bit codes / the "not a valid base" flags of 4 bytes with one multiply each
        auto pack4 = [](u32 w) -> u32 { return ((w & 0x03030303u) * 0x40100401u) >> 24; };               // byte j -> bits 7-2j..6-2j
        auto bad4 = [](u32 w) -> u32 { u32 nz = ((((w >> 2) & 0x03030303u) + 0x03030303u) >> 2) & 0x01010101u; return (nz * 0x80402010u) >> 28; };
        const u32 p16 = (pack4(v.x) << 8) | pack4(v.y);
        const u32 b8 = (bad4(v.x) << 4) | bad4(v.y);
        reinterpret_cast<u16*>(s_pk)[(ti >> 2) * 4 + (3 - (ti & 3))] = (u16)p16;
        reinterpret_cast<u8*>(s_bad)[(ti >> 3) * 8 + (7 - (ti & 7))] = (u8)b8;
    };
    load8(t);
    if (t < HALO / 8) load8(SK_THREADS + t);
    if (t < 6) s_pk[(SK_TP + HALO) / 32 + t] = 0;
    if (t < 2 + (KW > 2 ? 1 : 0)) s_bad[(SK_TP + HALO) / 64 + t] = ~0ULL;
    if (t < 9) s_brk[SK_TP / 32 + t] = 0xFFFFFFFFu;
    if (t == 0) s_nvalid = 0;
    __syncthreads();

    auto get64 = [&](u32 p) -> u64 {                              // 32 bases starting at position p
        u32 wi = p >> 5, o = p & 31;
        u64 a = s_pk[wi], b = s_pk[wi + 1];
        return o ? ((a << (2 * o)) | (b >> (64 - 2 * o))) : a;
    };
    auto getbad64 = [&](u32 p) -> u64 {                           // 64 invalid flags starting at position p
        u32 wi = p >> 6, o = p & 63;
        u64 a = s_bad[wi], b = s_bad[wi + 1];
        return o ? ((a << o) | (b >> (64 - o))) : a;
    };

    // ---- 2. m-mer selection values (the reference's mmer_lut, computed instead of looked up) --------------
    auto mvals8 = [&](int ti) {
        u32 p0 = 8 * (u32)ti;
        u64 W = get64(p0);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            u32 x = (u32)(W >> (64 - 2 * (j + m))) & mmask;
            s_mv[sk_pidx(p0 + j)] = mmer_order(x, m);
        }
    };
    mvals8(t);
    if (t < HALO / 8) mvals8(SK_THREADS + t);
    __syncthreads();

    // ---- 3. sliding minimum over w m-mers for my 8 windows, validity, run-start flags ----------------------
    const u32 p0 = 8 * (u32)t;
    u32 mn[8];
    if (w >= 8) {
        u32 a[8];
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = s_mv[sk_pidx(p0 + j)];
        u32 common = a[7];
        for (int i = 8; i < w; i++) common = min(common, s_mv[sk_pidx(p0 + i)]);
        u32 sfx[8]; sfx[7] = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 6; j >= 0; j--) sfx[j] = min(a[j], sfx[j + 1]);
        u32 pfx = 0xFFFFFFFFu;
        mn[0] = min(common, sfx[0]);
#pragma unroll
        for (int j = 1; j < 8; j++) {
            pfx = min(pfx, s_mv[sk_pidx(p0 + w + j - 1)]);
            mn[j] = min(min(common, sfx[j]), pfx);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            u32 v = 0xFFFFFFFFu;
            for (int i = 0; i < w; i++) v = min(v, s_mv[sk_pidx(p0 + j + i)]);
            mn[j] = v;
        }
    }
    u32 validmask = 0;                                            // bit (7-j): window j is a valid k-mer
    {
        u64 b0 = getbad64(p0), b1 = getbad64(p0 + 64);
        if constexpr (KW <= 2) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                u64 x = j ? ((b0 << j) | (b1 >> (64 - j))) : b0;
                bool ok = ((x >> (64 - k)) == 0) && (tile0 + p0 + j < limit);
                validmask |= (ok ? 1u : 0u) << (7 - j);
            }
        } else {
            // 64 < k <= 127: the window's first 64 flags, then its last k - 64
            const u64 b2 = getbad64(p0 + 128);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const u64 x = j ? ((b0 << j) | (b1 >> (64 - j))) : b0;
                const u64 y = j ? ((b1 << j) | (b2 >> (64 - j))) : b1;
                const bool ok = x == 0 && (k == 64 || (y >> (128 - k)) == 0) && (tile0 + p0 + j < limit);
                validmask |= (ok ? 1u : 0u) << (7 - j);
            }
        }
    }
    s_last[t] = ((validmask & 1u) ? mn[7] : SK_NOMIN);
    __syncthreads();
    u32 prevmn = t ? s_last[t - 1] : SK_NOMIN;
    u32 startmask = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        bool v = (validmask >> (7 - j)) & 1u;
        u32 cur = v ? mn[j] : SK_NOMIN;
        if (v && cur != prevmn) startmask |= 1u << (7 - j);
        prevmn = cur;
    }
    reinterpret_cast<u8*>(s_start)[(t >> 2) * 4 + (3 - (t & 3))] = (u8)startmask;
    reinterpret_cast<u8*>(s_brk)[(t >> 2) * 4 + (3 - (t & 3))] = (u8)(startmask | (~validmask & 0xFFu));
    __syncthreads();

    // ---- 4. split runs longer than maxS, count records, reserve output space -----------------------------------
    // pass filter (the reference's `minimizer % nbPass == pass`, K/SortingCountAlgorithm.cpp:1086, on the fine bin of the
    // minimizer): windows of other passes are valid k-mers of the bank but produce no record in this pass
    u32 passmask = 0xFFu;
    if (nb_passes > 1) {
        passmask = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) if (((validmask >> (7 - j)) & 1u) && bin_of(mn[j]) % nb_passes == pass_id) passmask |= 1u << (7 - j);
    }
    u32 chunkmask = 0;
    {
        u32 rel = 0;                                              // index of the window inside its run, modulo maxS
        if ((validmask & 0x80u) && !(startmask & 0x80u)) {       // my first window continues a run: find its start
            u32 q = p0 - 1;                                       // p0 > 0 here (thread 0 always starts a run)
            int wi = (int)(q >> 5);
            u32 bits = s_start[wi] & (0xFFFFFFFFu << (31 - (q & 31)));
            while (bits == 0) { wi--; bits = s_start[wi]; }
            u32 rs = (u32)wi * 32 + 31 - (u32)(__ffs((int)bits) - 1);
            rel = (p0 - rs) % (u32)maxS;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const u32 bit = 1u << (7 - j);
            if (startmask & bit) rel = 0;
            if (validmask & bit) { if (rel == 0) chunkmask |= bit; rel++; if (rel == (u32)maxS) rel = 0; }
        }
    }
    chunkmask &= passmask;
    u32 nch = __popc(chunkmask);
    u32 nval = __popc(validmask) | ((u32)__popc(validmask & passmask) << 16);   // all valid windows | those of this pass
    // block exclusive scan of nch
    u32 inc = nch;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_wsum[warp] = inc;
    u32 wval = __reduce_add_sync(0xFFFFFFFFu, nval);                          // two 16-bit lanes: <= 256 per warp each
    if (lane == 0 && wval) atomicAdd(&s_nvalid, wval);                        // <= 2048 per block each
    __syncthreads();
    u32 wpre = 0, btotal = 0;
#pragma unroll
    for (int i = 0; i < SK_THREADS / 32; i++) { u32 s = s_wsum[i]; if (i < warp) wpre += s; btotal += s; }
    u32 myidx = wpre + inc - nch;
    if (t == 0) {
        s_goff = atomicAdd(&ctr->nrec, (unsigned long long)btotal);
        atomicAdd(&ctr->kmers_valid, (unsigned long long)(s_nvalid & 0xFFFFu));
        atomicAdd(&ctr->kmers_pass, (unsigned long long)(s_nvalid >> 16));
    }
    __syncthreads();
    const u64 goff = s_goff;
    if (goff + btotal > rec_cap) { if (t == 0) atomicExch(&ctr->overflow, 1u); return; }   // uniform: whole block leaves

    // ---- 5. emit records: chunk starts are listed in smem, then thread i builds record i (coalesced stores) ------
    __shared__ u16 s_list[SK_TP];                                 // tile-relative position of every record start
    u32* s_lmn = s_mv;                                            // its minimizer (s_mv is dead after step 3)
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if ((chunkmask >> (7 - j)) & 1u) { s_list[myidx] = p0 + j; s_lmn[myidx] = mn[j]; myidx++; }
    }
    __syncthreads();
    u32 nk_sum = 0;
    for (u32 i = t; i < btotal; i += SK_THREADS) {
        const u32 p = s_list[i];
        // next break strictly after p (bounded: s_brk is all ones past the tile)
        u32 q = p + 1; u32 wi = q >> 5;
        u32 bits = s_brk[wi] & (0xFFFFFFFFu >> (q & 31));
        u32 steps = 0;
        while (bits == 0 && steps < 9) { wi++; bits = s_brk[wi]; steps++; }
        const u32 nb = bits ? (wi * 32 + (u32)__clz((int)bits)) : (p + (u32)maxS);
        const u32 nk = min((u32)maxS, nb - p);
        u64 rw[RW];
#pragma unroll
        for (int x = 0; x < RW; x++) rw[x] = get64(p + 32 * x);
        const u32 hmin = bin_hash(s_lmn[i]);
        rw[RW - 1] = (rw[RW - 1] & ~0xFFFFULL) | ((u64)nk << 8) | (u64)(((hmin >> 4) & 15u) << 4) | (u64)bank;     // [nk:8][sub-bin:4][bank:4]
        const u64 ri = goff + i;
        if constexpr (RW == 2) {
            reinterpret_cast<ulonglong2*>(recs)[ri] = make_ulonglong2(rw[0], rw[1]);
        } else if constexpr (RW == 4) {
            reinterpret_cast<ulonglong2*>(recs)[2 * ri] = make_ulonglong2(rw[0], rw[1]);
            reinterpret_cast<ulonglong2*>(recs)[2 * ri + 1] = make_ulonglong2(rw[2], rw[3]);
        } else {
#pragma unroll
            for (int x = 0; x < RW / 2; x++) reinterpret_cast<ulonglong2*>(recs)[(RW / 2) * ri + x] = make_ulonglong2(rw[2 * x], rw[2 * x + 1]);
        }
        // fine histogram of the bins while the records are produced (fire-and-forget REDs under an ALU-bound kernel)
        const u32 bin = hmin >> (32 - META_BIN_BITS);
        rec_meta[ri] = bin | (nk << 24);
        atomicAdd(&bin_hist[bin >> fine_shift], (1ULL << BINH_KBITS) | (unsigned long long)nk);
        nk_sum += nk;
    }
    nk_sum = __reduce_add_sync(0xFFFFFFFFu, nk_sum);
    if (lane == 0 && nk_sum) atomicAdd(&ctr->kmers_in_recs, (unsigned long long)nk_sum);
}

// ---- K3: scatter records into partition order --------------------------------------------------------------------
// bin2q[bin >> bin_shift] = position of the bin's partition in q order (plan.cuh: owner-major, so that every owner's
// partitions -- and after this kernel every owner's RECORDS -- are one contiguous chunk); cursor[q] starts at the first
// record slot of partition q (a copy of the exclusive prefix of this rank's per-partition record counts), so a record takes
// one L2 atomic that returns its final position and one 16/32-byte vector store.  There are 20 K (C2) to millions of
// partitions against ~8 K records per CTA, so there is nothing to aggregate per (CTA, partition).
constexpr int SC_THREADS = 256;
template <int KW>
__global__ void __launch_bounds__(SC_THREADS) k_part_scatter(const u64* __restrict__ recs, const u32* __restrict__ rec_meta, u64 nrec,
                                                             const u32* __restrict__ bin2q, int bin_shift, u64* __restrict__ out,
                                                             unsigned long long* __restrict__ cursor)
{
    constexpr int RW = 2 * KW;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(recs);
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(out);
    for (u64 i = (u64)blockIdx.x * SC_THREADS + threadIdx.x; i < nrec; i += (u64)gridDim.x * SC_THREADS) {
        const u32 q = __ldg(bin2q + ((rec_meta[i] & META_BIN_MASK) >> bin_shift));
        if constexpr (RW <= 4) {
            ulonglong2 a = src[(RW / 2) * i], b;
            if constexpr (RW == 4) b = src[2 * i + 1];
            const u64 d = atomicAdd(&cursor[q], 1ULL);
            if constexpr (RW == 2) dst[d] = a;
            else { dst[2 * d] = a; dst[2 * d + 1] = b; }
        } else {
            ulonglong2 v[RW / 2];
#pragma unroll
            for (int x = 0; x < RW / 2; x++) v[x] = src[(RW / 2) * i + x];
            const u64 d = atomicAdd(&cursor[q], 1ULL);
#pragma unroll
            for (int x = 0; x < RW / 2; x++) dst[(RW / 2) * d + x] = v[x];
        }
    }
}

// ---- K3 for jobs with millions of partitions: MSD multi-split with block-level binning in shared memory ------------------
// The single-pass scatter above pays one L2 atomic and one isolated 16-byte store per record; beyond ~1 M partitions (multi-G
// k-mer jobs, multi-GPU jobs) the open destinations no longer fit the L2 and it falls to 10 G records/s.  Here the q-ordered
// layout is reached in ceil(bits / 8) passes, most significant digit first.  In a pass a CTA takes a tile of 2048 records,
// bins them by `q >> shift` in shared memory (one shared atomic per record gives its rank inside the tile's bin), reserves
// ONE run per (tile, bin) on the bin's global cursor, and stores the records of a bin next to each other (runs of ~8 records
// = one 128-byte line when a pass splits 256 ways).  The destination of every bin is known beforehand -- the planner's
// exclusive prefix `loff` -- so there is no counting pass.  Input of a later pass is grouped by the previous digit, so the
// bins a tile can meet are the 256 children of at most two parents (512 counters); anything beyond that (tiny jobs) takes the
// per-record cursor path.
constexpr int MS_THREADS = 256;
constexpr u32 MS_CAP = 512;

// cursor[g] = first record slot of group g = q >> shift (groups 0 .. (nq - 1) >> shift)
__global__ void k_msd_cursor(const u64* __restrict__ loff, u64 nq, int shift, unsigned long long* __restrict__ cursor)
{
    const u64 ng = ((nq - 1) >> shift) + 1;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (u64)gridDim.x * blockDim.x) cursor[g] = loff[min(g << shift, nq)];
}

template <int KW> DSK_HD u32 ms_tile() { return KW == 1 ? 2048u : 1024u; }
template <int KW> DSK_HD size_t ms_smem_bytes() { return (size_t)ms_tile<KW>() * (KW * 16 + 4) + MS_CAP * (4 + 4 + 8) + 64; }

template <int KW, bool FIRST, bool LAST>
__global__ void __launch_bounds__(MS_THREADS) k_msd_pass(const u64* __restrict__ src, const u32* __restrict__ src_key /*FIRST: record meta*/, u64 nrec,
                                                         const u32* __restrict__ bin2q, int bin_shift, int shift, int parent_shift,
                                                         unsigned long long* __restrict__ cursor, u64* __restrict__ dst, u32* __restrict__ dst_key)
{
    constexpr int RPT = KW == 1 ? 8 : 4;
    constexpr int VPR = KW;                                        // 16-byte vectors per record
    constexpr u32 TILE = MS_THREADS * RPT;
    extern __shared__ __align__(16) unsigned char ms_dyn[];
    ulonglong2* s_rec = reinterpret_cast<ulonglong2*>(ms_dyn);                               // [TILE][VPR] records grouped by bin
    unsigned long long* s_base = reinterpret_cast<unsigned long long*>(s_rec + (size_t)TILE * VPR);   // [MS_CAP] global slot of the bin's run
    u32* s_key = reinterpret_cast<u32*>(s_base + MS_CAP);                                    // [TILE]
    u32* s_cnt = s_key + TILE;                                                               // [MS_CAP]
    u32* s_off = s_cnt + MS_CAP;                                                             // [MS_CAP] exclusive prefix of s_cnt
    __shared__ u32 s_g0, s_wsum[MS_THREADS / 32];
    const ulonglong2* in = reinterpret_cast<const ulonglong2*>(src);
    ulonglong2* out = reinterpret_cast<ulonglong2*>(dst);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (u64 base = (u64)blockIdx.x * TILE; base < nrec; base += (u64)gridDim.x * TILE) {
        for (u32 i = t; i < MS_CAP; i += MS_THREADS) s_cnt[i] = 0;
        u32 q[RPT]; ulonglong2 ra[RPT], rb[RPT], rx[KW > 2 ? RPT : 1][KW > 2 ? VPR - 2 : 1];   // rx: vectors 2 .. of a wide record
#pragma unroll
        for (int j = 0; j < RPT; j++) {
            const u64 i = base + (u64)j * MS_THREADS + t;
            q[j] = 0xFFFFFFFFu;
            if (i < nrec) {
                q[j] = FIRST ? __ldg(bin2q + ((src_key[i] & META_BIN_MASK) >> bin_shift)) : src_key[i];
                ra[j] = in[VPR * i];
                if constexpr (KW >= 2) rb[j] = in[VPR * i + 1];
                if constexpr (KW > 2) {
#pragma unroll
                    for (int x = 2; x < VPR; x++) rx[j][x - 2] = in[VPR * i + x];
                }
            }
        }
        // bins of this tile start at the first child of the first record's parent (the input is grouped by parent, parents in
        // increasing order; the first pass has one parent: everything)
        if (t == 0) s_g0 = FIRST ? 0u : ((q[0] >> parent_shift) << (parent_shift - shift));
        __syncthreads();
        const u32 g0 = s_g0;
        u32 r[RPT];
#pragma unroll
        for (int j = 0; j < RPT; j++) {
            r[j] = 0xFFFFFFFFu;
            if (q[j] != 0xFFFFFFFFu) { const u32 idx = (q[j] >> shift) - g0; if (idx < MS_CAP) r[j] = atomicAdd(&s_cnt[idx], 1u); }
        }
        __syncthreads();
        // exclusive prefix of the bin counts (two bins per thread) + one global reservation per non-empty bin
        {
            const u32 c0 = s_cnt[2 * t], c1 = s_cnt[2 * t + 1];
            u32 inc = c0 + c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
            if (lane == 31) s_wsum[warp] = inc;
            if (c0) s_base[2 * t] = atomicAdd(&cursor[g0 + 2 * t], (unsigned long long)c0);
            if (c1) s_base[2 * t + 1] = atomicAdd(&cursor[g0 + 2 * t + 1], (unsigned long long)c1);
            __syncthreads();
            u32 wpre = 0;
#pragma unroll
            for (int w = 0; w < MS_THREADS / 32; w++) wpre += w < warp ? s_wsum[w] : 0u;
            const u32 ex = wpre + inc - (c0 + c1);
            s_off[2 * t] = ex; s_off[2 * t + 1] = ex + c0;
        }
        __syncthreads();
        u32 staged = 0;
#pragma unroll
        for (int w = 0; w < MS_THREADS / 32; w++) staged += s_wsum[w];
#pragma unroll
        for (int j = 0; j < RPT; j++) {
            if (q[j] == 0xFFFFFFFFu) continue;
            const u32 g = q[j] >> shift;
            if (r[j] != 0xFFFFFFFFu) {
                const u32 p = s_off[g - g0] + r[j];
                s_rec[(size_t)p * VPR] = ra[j];
                if constexpr (KW >= 2) s_rec[(size_t)p * VPR + 1] = rb[j];
                if constexpr (KW > 2) {
#pragma unroll
                    for (int x = 2; x < VPR; x++) s_rec[(size_t)p * VPR + x] = rx[j][x - 2];
                }
                s_key[p] = q[j];
            } else {                                               // a bin outside the window of this tile (tiny jobs): per-record cursor
                const u64 pos = atomicAdd(&cursor[g], 1ULL);
                out[VPR * pos] = ra[j];
                if constexpr (KW >= 2) out[VPR * pos + 1] = rb[j];
                if constexpr (KW > 2) {
#pragma unroll
                    for (int x = 2; x < VPR; x++) out[VPR * pos + x] = rx[j][x - 2];
                }
                if constexpr (!LAST) dst_key[pos] = q[j];
            }
        }
        __syncthreads();
        // the staged records leave bin by bin: consecutive threads store consecutive 16/32-byte records of a run
        for (u32 i = t; i < staged; i += MS_THREADS) {
            const u32 qq = s_key[i], idx = (qq >> shift) - g0;
            const u64 pos = s_base[idx] + (u64)(i - s_off[idx]);
            out[VPR * pos] = s_rec[(size_t)i * VPR];
            if constexpr (KW >= 2) out[VPR * pos + 1] = s_rec[(size_t)i * VPR + 1];
            if constexpr (KW > 2) {
#pragma unroll
                for (int x = 2; x < VPR; x++) out[VPR * pos + x] = s_rec[(size_t)i * VPR + x];
            }
            if constexpr (!LAST) dst_key[pos] = qq;
        }
        __syncthreads();                                           // staging buffer and counters are reused by the next tile
    }
}

// packed fine bin histogram [nfine] -> level histogram [2][nfine >> shift] (records per bin, then k-mers per bin)
__global__ void k_fold_bins(const unsigned long long* __restrict__ fine, u32 nfine, int shift, unsigned long long* __restrict__ out)
{
    const u32 nb = nfine >> shift, per = 1u << shift;
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
        const unsigned long long* src = fine + ((size_t)b << shift);
        unsigned long long r = 0, k = 0;
        for (u32 j = 0; j < per; j++) { const unsigned long long v = src[j]; r += v >> BINH_KBITS; k += v & BINH_KMASK; }
        out[b] = r; out[nb + b] = k;
    }
}

// flags a fine bin whose k-mer field reached `limit` (2^28 in production): its record field may have wrapped
__global__ void k_check_bins(const unsigned long long* __restrict__ fine, u32 nfine, unsigned long long limit, Counters* ctr)
{
    bool bad = false;
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < nfine; b += gridDim.x * blockDim.x) bad |= (fine[b] & BINH_KMASK) >= limit;
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicExch(&ctr->hist_suspect, 1u);
}

// the exact level histogram from the record meta (bin24 | nk << 24), for the flagged case: two 64-bit atomics per record
// (shift = 24 - level, nb = bins of the level)
__global__ void k_rebuild_hist(const u32* __restrict__ rec_meta, u64 nrec, int shift, u32 nb, unsigned long long* __restrict__ out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += (u64)gridDim.x * blockDim.x) {
        const u32 mt = rec_meta[i], b = (mt & META_BIN_MASK) >> shift;
        atomicAdd(&out[b], 1ULL);
        atomicAdd(&out[nb + b], (unsigned long long)(mt >> 24));
    }
}

// density sample: the records of the fine bins below `thresh` are copied out (all occurrences of a k-mer share their
// bin, so distinct / total of the sample estimates distinct / total of the job)
template <int KW>
__global__ void __launch_bounds__(256) k_sample_select(const u64* __restrict__ recs, const u32* __restrict__ rec_meta, const unsigned long long* nrec_dev,
                                                       u32 thresh, u64* __restrict__ out, u64 km_cap, Counters* ctr)
{
    constexpr int RW = 2 * KW;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(recs);
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(out);
    const u64 nrec = *nrec_dev;
    auto take = [&](u64 i, u32 mt) {
        const unsigned long long nk = mt >> 24;
        if (atomicAdd(&ctr->sample_nkm, nk) + nk > km_cap) { atomicAdd(&ctr->sample_nkm, ~nk + 1ULL); return; }   // sample full (a record holds >= 1 k-mer,
        const u64 d = atomicAdd(&ctr->sample_nrec, 1ULL);                                                            //  so records <= km_cap as well)
#pragma unroll
        for (int q = 0; q < RW / 2; q++) dst[(RW / 2) * d + q] = src[(RW / 2) * i + q];
    };
    // four metas per 16-byte load (the buffer is 256-byte aligned); one record in ~750 is selected
    const uint4* m4 = reinterpret_cast<const uint4*>(rec_meta);
    const u64 n4 = nrec / 4;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (u64)gridDim.x * blockDim.x) {
        const uint4 v = m4[g];
        if ((v.x & META_BIN_MASK) < thresh) take(4 * g, v.x);
        if ((v.y & META_BIN_MASK) < thresh) take(4 * g + 1, v.y);
        if ((v.z & META_BIN_MASK) < thresh) take(4 * g + 2, v.z);
        if ((v.w & META_BIN_MASK) < thresh) take(4 * g + 3, v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (nrec & 3)) { const u64 i = 4 * n4 + threadIdx.x; const u32 mt = rec_meta[i]; if ((mt & META_BIN_MASK) < thresh) take(i, mt); }
}

#endif  // __CUDACC__
}  // namespace dsk
