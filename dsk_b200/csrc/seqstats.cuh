// seqstats.cuh -- per-sequence statistics of the banks, from the code stream the scanner emits.
//
// Replaces BankStats::update / operator+= (K/BankKmers.hpp:166-215) and the kmersNbInvalid counter of
// Sequence2SuperKmer (K/Sequence2SuperKmer.hpp:95-108): the keys seq_size_min / seq_size_max / seq_size_deviation and
// kmers_nb_invalid of SortingCountAlgorithm::getInfo() (K/SortingCountAlgorithm.cpp:733-742).  Optional
// (dskgpu_config.sequence_stats): the reference gathers them for free while it walks every sequence; here they cost one more
// read of the code stream per chunk, so the C++ host adapter asks for them and the benchmark does not.
//
// A sequence is a maximal run of codes between record separators (code 8).  FASTA puts the separator IN FRONT of a record
// (at its '>' line), FASTQ and the one-sequence-per-line format BEHIND it (the newline that ends the sequence line):
//   FASTA        : a run is a sequence when a separator precedes it; the last one is closed by the end of the stream
//   FASTQ / lines: a run is a sequence when a separator follows it; a non-empty run at the end of the stream counts too
// Lengths include the non-ACGT letters (Sequence::getDataSize).  A sequence of length n holds max(0, n - k + 1) k-mer
// windows; the ones that are not valid k-mers (kmers_nb_valid comes from k_superkmers) are the invalid ones.
//
// Per chunk: k_seqstat_tiles -- every CTA takes 4096 codes of the new region [carry, total), finds the separators (16 codes
// per thread), closes the sequences whose two ends lie inside the tile and leaves (first separator, last separator) of the
// tile in a table; k_seqstat_stitch -- one CTA closes the sequences that span tiles (a max-scan over the table) and carries
// the open run (`open_len`, `seen_sep`) to the next chunk in the stream state.
#pragma once
#include "scan.cuh"

namespace dsk {

struct SeqStats {                    // device accumulators (zeroed by dskgpu_reset)
    unsigned long long n;            // sequences closed
    unsigned long long sum;          // sum of their lengths
    unsigned long long sumsq;        // sum of squares
    unsigned long long windows;      // sum of max(0, len - k + 1)
    unsigned long long max_len;
    unsigned long long min_inv;      // max over sequences of ~len  (min length = ~min_inv once n > 0)
    unsigned long long open_len;     // codes since the last boundary (separator or stream start) -- stream state
    unsigned int seen_sep, pad;      // a separator was seen in this stream
};

#ifdef __CUDACC__

constexpr int SQ_THREADS = 256;
constexpr int SQ_TILE = SQ_THREADS * 16;
constexpr unsigned long long SQ_NONE = ~0ULL;

struct SqAcc { unsigned long long n, sum, sumsq, windows, mx, mninv; };
__device__ __forceinline__ void sq_add(SqAcc& a, unsigned long long len, int k)
{
    a.n++; a.sum += len; a.sumsq += len * len;
    if (len >= (unsigned long long)k) a.windows += len - (unsigned long long)k + 1;
    if (len > a.mx) a.mx = len;
    if (~len > a.mninv) a.mninv = ~len;
}
__device__ __forceinline__ void sq_flush(const SqAcc& a, SeqStats* st)
{
    if (!a.n) return;
    atomicAdd(&st->n, a.n); atomicAdd(&st->sum, a.sum); atomicAdd(&st->sumsq, a.sumsq); atomicAdd(&st->windows, a.windows);
    atomicMax(&st->max_len, a.mx); atomicMax(&st->min_inv, a.mninv);
}

// tab[2 * tile] = absolute position of the first separator of the tile (SQ_NONE: none), tab[2 * tile + 1] = of the last
__global__ void __launch_bounds__(SQ_THREADS) k_seqstat_tiles(const u8* __restrict__ codes, const StreamState* __restrict__ ss, int k,
                                                               unsigned long long* __restrict__ tab, SeqStats* st)
{
    const u64 carry = ss->carry, total = ss->total;
    const u64 base = (carry & ~(u64)15) + (u64)blockIdx.x * SQ_TILE;
    if (base >= total) { if (threadIdx.x == 0) { tab[2 * (u64)blockIdx.x] = SQ_NONE; tab[2 * (u64)blockIdx.x + 1] = SQ_NONE; } return; }
    __shared__ unsigned long long s_wlast[SQ_THREADS / 32], s_first;
    __shared__ unsigned long long s_acc[6];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_first = SQ_NONE;
    if (t < 6) s_acc[t] = 0;
    const u64 p0 = base + 16 * (u64)t;
    u32 sep = 0;                                                           // bit j: code p0 + j is a separator of the new region
    if (p0 < total) {
        const uint4 v = *reinterpret_cast<const uint4*>(codes + p0);       // (the code buffer is 16-byte aligned and padded)
        const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; j++) sep |= (((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) == 8u ? 1u : 0u) << j;
        for (int j = 0; j < 16; j++) { const u64 p = p0 + j; if (p < carry || p >= total) sep &= ~(1u << j); }
    }
    // last separator at or before the end of my 16 codes, over the threads before me (exclusive max-scan; SQ_NONE = none yet)
    unsigned long long mylast = sep ? p0 + (31 - __clz((int)sep)) : 0ULL;  // 0 = none (positions are compared + 1)
    unsigned long long inc = sep ? mylast + 1 : 0ULL;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d && o > inc) inc = o; }
    if (lane == 31) s_wlast[warp] = inc;
    __syncthreads();
    unsigned long long before = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) before = 0;
    for (int w2 = 0; w2 < warp; w2++) { const unsigned long long o = s_wlast[w2]; if (o > before) before = o; }
    // `before` = 1 + position of the last separator of the tile before my codes (0 = none)
    SqAcc a = {0, 0, 0, 0, 0, 0};
    u32 m = sep;
    unsigned long long prev = before;
    while (m) {
        const int j = __ffs((int)m) - 1; m &= m - 1;
        const u64 p = p0 + j;
        if (prev) sq_add(a, p - prev, k);                                  // codes strictly between the two separators: p - (prev - 1) - 1
        else atomicMin(&s_first, (unsigned long long)p);                   // first separator of the tile: closed by the stitch pass
        prev = p + 1;
    }
    // block reduction of the accumulators
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        a.n += __shfl_xor_sync(0xFFFFFFFFu, a.n, d); a.sum += __shfl_xor_sync(0xFFFFFFFFu, a.sum, d);
        a.sumsq += __shfl_xor_sync(0xFFFFFFFFu, a.sumsq, d); a.windows += __shfl_xor_sync(0xFFFFFFFFu, a.windows, d);
        const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, a.mx, d), o2 = __shfl_xor_sync(0xFFFFFFFFu, a.mninv, d);
        if (o1 > a.mx) a.mx = o1; if (o2 > a.mninv) a.mninv = o2;
    }
    if (lane == 0 && a.n) {
        atomicAdd(&s_acc[0], a.n); atomicAdd(&s_acc[1], a.sum); atomicAdd(&s_acc[2], a.sumsq); atomicAdd(&s_acc[3], a.windows);
        atomicMax(&s_acc[4], a.mx); atomicMax(&s_acc[5], a.mninv);
    }
    __syncthreads();
    if (t == 0) {
        SqAcc b = {s_acc[0], s_acc[1], s_acc[2], s_acc[3], s_acc[4], s_acc[5]};
        sq_flush(b, st);
        unsigned long long last = 0;
        for (int w2 = 0; w2 < SQ_THREADS / 32; w2++) if (s_wlast[w2] > last) last = s_wlast[w2];
        tab[2 * (u64)blockIdx.x] = s_first;
        tab[2 * (u64)blockIdx.x + 1] = last ? last - 1 : SQ_NONE;
    }
}

// closes the sequences that end at the FIRST separator of a tile (their start lies in an earlier tile or chunk) and carries
// the open run to the next chunk.  fasta: separators are in front of the records (see the header).
__global__ void __launch_bounds__(1024) k_seqstat_stitch(const StreamState* __restrict__ ss, int k, int fasta, u64 ntiles_max,
                                                         const unsigned long long* __restrict__ tab, SeqStats* st)
{
    const u64 carry = ss->carry, total = ss->total;
    if (total <= carry) return;                                            // nothing new in this chunk
    const u64 abase = carry & ~(u64)15;
    u64 ntiles = (total - abase + SQ_TILE - 1) / SQ_TILE;
    if (ntiles > ntiles_max) ntiles = ntiles_max;
    __shared__ unsigned long long s_w[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const u64 per = (ntiles + 1023) / 1024;
    const u64 b = min((u64)t * per, ntiles), e = min(b + per, ntiles);
    // 1 + last separator position over my tiles (0 = none), then exclusive max-scan over the threads
    unsigned long long mine = 0;
    for (u64 i = b; i < e; i++) { const unsigned long long l = tab[2 * i + 1]; if (l != SQ_NONE && l + 1 > mine) mine = l + 1; }
    unsigned long long inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d && o > inc) inc = o; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned long long prev = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) prev = 0;
    for (int w2 = 0; w2 < warp; w2++) { const unsigned long long o = s_w[w2]; if (o > prev) prev = o; }
    const unsigned long long open_in = st->open_len;                       // (read by everybody before thread 1023 rewrites it: barrier below)
    const unsigned int seen_in = st->seen_sep;
    SqAcc a = {0, 0, 0, 0, 0, 0};
    for (u64 i = b; i < e; i++) {
        const unsigned long long f = tab[2 * i], l = tab[2 * i + 1];
        if (f != SQ_NONE) {
            if (prev) sq_add(a, f - prev, k);                              // previous separator inside this chunk
            else if (!fasta || seen_in) sq_add(a, open_in + (f - carry), k);   // the run came in open from the previous chunk(s)
        }
        if (l != SQ_NONE) prev = l + 1;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        a.n += __shfl_xor_sync(0xFFFFFFFFu, a.n, d); a.sum += __shfl_xor_sync(0xFFFFFFFFu, a.sum, d);
        a.sumsq += __shfl_xor_sync(0xFFFFFFFFu, a.sumsq, d); a.windows += __shfl_xor_sync(0xFFFFFFFFu, a.windows, d);
        const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, a.mx, d), o2 = __shfl_xor_sync(0xFFFFFFFFu, a.mninv, d);
        if (o1 > a.mx) a.mx = o1; if (o2 > a.mninv) a.mninv = o2;
    }
    if (lane == 0) sq_flush(a, st);
    __syncthreads();                                                       // open_len / seen_sep have been read by everybody
    if (t == 1023) {
        // `prev` after my (the last) run = 1 + last separator of the whole chunk, or 0
        if (prev) { st->open_len = total - prev; st->seen_sep = 1u; }
        else st->open_len = open_in + (total - carry);
    }
}

// end of a stream (bank finished, or finish): the run still open is a sequence under the rules of the header
__global__ void k_seqstat_close(int k, int fasta, SeqStats* st)
{
    const unsigned long long len = st->open_len;
    if (fasta ? st->seen_sep != 0u : len > 0) { SqAcc a = {0, 0, 0, 0, 0, 0}; sq_add(a, len, k); sq_flush(a, st); }
    st->open_len = 0; st->seen_sep = 0;
}

#endif  // __CUDACC__
}  // namespace dsk
