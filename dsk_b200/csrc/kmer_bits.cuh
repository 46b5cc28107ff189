// kmer_bits.cuh -- bit-level k-mer logic shared by device kernels and the host self-checks.
//
// Everything here is a pure function, compiled for both host and device, so the CPU test-suite can pin
// the exact code the kernels run against the oracle (tests/test_host_logic.py) without a GPU.
//
// Reference semantics restated (G/ = thirdparty/gatb-core/gatb-core/, K/ = G/src/gatb/kmer/impl/):
//   encoding        G/src/gatb/tools/misc/api/Data.hpp:185      A=0 C=1 T=2 G=3, (c>>1)&3, only ACGTacgt valid
//   canonical       K/Model.hpp:294, :857-884                   min(forward, revcomp), complement = code ^ 2
//   minimizer       K/Model.hpp:1010-1070, :1220-1251, :1254-1287  min over m-mers of lut[] (canonical m-mer,
//                                                               4^m-1 if it has "AA" beyond its first 2 letters)
//   record scanning G/src/gatb/bank/impl/BankFasta.cpp:485-572
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DSK_HD __host__ __device__ __forceinline__
#else
#define DSK_HD inline
#endif

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t  u8;

namespace dsk {

// ---- code stream alphabet (one byte per emitted symbol) --------------------------------------------
enum : int { CODE_INVALID = 4, CODE_SEP = 8 };   // 0..3 base; 4|code = non-ACGT base; 8 = record separator

DSK_HD int encode_base(int c)
{
    int u = c & 0xDF;                                     // upper-case
    bool ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
    return ((c >> 1) & 3) | (ok ? 0 : CODE_INVALID);
}

// ---- record scanner state machines -----------------------------------------------------------------
enum : int { FMT_FASTA = 1, FMT_FASTQ = 2, FMT_LINES = 3 };
enum : int { ST_SEQ = 0, ST_HDR = 1 };                    // FASTA: type of the current line
enum : int { SCAN_ERR_PLUS_IN_FASTA = 1, SCAN_ERR_FASTQ_AT = 2, SCAN_ERR_FASTQ_PLUS = 4 };

// One byte of the scanner.  `prev` = previous byte of the stream ('\n' at stream start), `next` = following
// byte (-1 at end of stream).  Returns the emitted code (0..8) or -1; updates state / err.
DSK_HD int scan_step(int fmt, int& state, int prev, int c, int next, int& err)
{
    const bool ls = (prev == '\n');
    if (fmt == FMT_FASTA) {
        if (ls) {
            if (c == '>' || c == '@') { state = ST_HDR; return CODE_SEP; }   // BankFasta.cpp:528 next record
            if (c == '+') { err |= SCAN_ERR_PLUS_IN_FASTA; state = ST_HDR; return -1; }
            state = ST_SEQ;
        }
        if (state == ST_HDR) return -1;
        if (c == '\n') return -1;                                           // :530 empty line / end of line
        if (c == '\r' && (next == '\n' || next < 0)) return -1;             // :471 trailing CR
        return encode_base(c);
    }
    if (fmt == FMT_FASTQ) {                                                 // state = line index mod 4
        switch (state) {
        case 0: if (ls && c != '@' && c != '\n') err |= SCAN_ERR_FASTQ_AT;
                if (c == '\n') state = 1;
                return -1;
        case 1: if (c == '\n') { state = 2; return CODE_SEP; }
                if (c == '\r' && (next == '\n' || next < 0)) return -1;
                return encode_base(c);
        case 2: if (ls && c != '+') err |= SCAN_ERR_FASTQ_PLUS;
                if (c == '\n') state = 3;
                return -1;
        default: if (c == '\n') state = 0;
                return -1;
        }
    }
    // FMT_LINES: one sequence per line
    if (c == '\n') return CODE_SEP;
    return encode_base(c);
}

// ---- 2-bit arithmetic ---------------------------------------------------------------------------------
// reverse the 32 2-bit groups of a 64-bit word
DSK_HD u64 rev2_64(u64 x)
{
#ifdef __CUDA_ARCH__
    u64 y = __brevll(x);                                    // bit reversal, then swap the two bits of every pair back
    return ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
#else
    x = ((x >> 2)  & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4)  & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8)  & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
#endif
}
DSK_HD u32 rev2_32(u32 x)
{
#ifdef __CUDA_ARCH__
    u32 y = __brev(x);
    return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
#else
    x = ((x >> 2)  & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4)  & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8)  & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
// reverse complement of a k-mer (k <= 32) held in the low 2k bits, first base most significant
DSK_HD u64 revcomp64(u64 x, int k)
{
    u64 r = rev2_64(x ^ 0xAAAAAAAAAAAAAAAAULL);            // complement = ^2 per base, then reverse groups
    return r >> (64 - 2 * k);
}
// 128-bit variant: value = hi:lo, k <= 64
DSK_HD void revcomp128(u64 lo, u64 hi, int k, u64& rlo, u64& rhi)
{
    u64 a = rev2_64(lo ^ 0xAAAAAAAAAAAAAAAAULL);           // becomes the high word
    u64 b = rev2_64(hi ^ 0xAAAAAAAAAAAAAAAAULL);           // becomes the low word
    int sh = 128 - 2 * k;                                   // 0..126
    if (sh == 0)       { rhi = a; rlo = b; }
    else if (sh < 64)  { rlo = (b >> sh) | (a << (64 - sh)); rhi = a >> sh; }
    else if (sh == 64) { rlo = a; rhi = 0; }
    else               { rlo = a >> (sh - 64); rhi = 0; }
}

// m-mer value used for minimizer selection (the reference's _mmer_lut entry), 2 <= m <= 15, 32-bit arithmetic
DSK_HD u32 mmer_value(u32 x, int m)
{
    const u32 mmask = (1u << (2 * m)) - 1u;
    u32 rc = rev2_32(x ^ 0xAAAAAAAAu) >> (32 - 2 * m);      // complement every base, reverse, drop the padding pairs
    u32 v = rc < x ? rc : x;
    // K/Model.hpp:1220-1251 is_allowed: ban "AA" anywhere except as the first two letters
    const u32 mask_ma1 = 0x55555555u & ((1u << ((m - 2) * 2)) - 1u);
    u32 a1 = ~(v | (v >> 2));
    a1 = ((a1 >> 1) & a1) & mask_ma1;
    return a1 ? mmask : v;
}

// What the kernels minimise over.  Identical to mmer_value for every allowed m-mer, so a k-mer that has an allowed m-mer gets
// the reference's minimizer; the banned m-mers, which the reference all maps to ONE default value (mmask), keep their own
// value above every allowed one.  The k-mers without any allowed m-mer (1e-3 of a random genome at m = 14: 144 M k-mers of
// a 72 G k-mer job) are then spread over many minimizer bins instead of forming one giant partition on one rank.  Which
// partition a k-mer lands in is unobservable in the results.
DSK_HD u32 mmer_order(u32 x, int m)
{
    u32 rc = rev2_32(x ^ 0xAAAAAAAAu) >> (32 - 2 * m);
    u32 v = rc < x ? rc : x;
    const u32 mask_ma1 = 0x55555555u & ((1u << ((m - 2) * 2)) - 1u);
    u32 a1 = ~(v | (v >> 2));
    a1 = ((a1 >> 1) & a1) & mask_ma1;
    return a1 ? (v | (1u << (2 * m))) : v;                    // m <= 15: 31 bits
}

// minimizer -> bin (the role of Repartitor::operator(), K/PartiInfo.hpp:323; any deterministic map is legal --
// SURVEY.md appendix C).  Records are histogrammed into 2^NBINS_FINE_LOG2 fine bins while they are produced.  At finish
// the histogram is folded to the level the job needs (2^16 bins for the 400 M k-mer configuration, up to 2^22 for
// multi-G k-mer jobs: level = what keeps the average bin well under one shared-memory table) and the host packs
// consecutive bins of that level into partitions of the size the counting kernel wants (balanced on exact counts,
// the job the reference gives to its sampled LPT table, K/PartiInfo.cpp:48-106).
// A record carries the top META_BIN_BITS bits of the minimizer hash ("bin24"); the fine histogram has 2^fine_log2 bins
// (fine bin = bin24 >> (24 - fine_log2)): 2^22 (32 MB, L2-resident under the record stream; DSKGPU_FINE_LOG2 picks another
// level up to 2^24, which costs k_superkmers 3x on multi-G k-mer jobs: the REDs then miss L2).  At 2^22 bins the AVERAGE bin of
// a 72 G k-mer job already is a whole shared-memory table: such partitions are counted as record sub-passes (sub-bins below).
constexpr int META_BIN_BITS = 24;
constexpr u32 META_BIN_MASK = (1u << META_BIN_BITS) - 1u;
constexpr int NBINS_FINE_LOG2_MAX = 24;
constexpr int NBINS_LOG2 = 16;                               // coarsest level (DSKGPU_NBINS)
constexpr u32 NBINS = 1u << NBINS_LOG2;
// 32-bit hash of a minimizer: the top 24 bits are the record's bin ("bin24"), the next 4 its SUB-BIN -- hash bits no
// partition is built from, stored in the record itself ([nk:8][sub:4][bank:4] in the low 16 bits of its last word): a
// partition too big for one shared-memory table is counted as 2^s sub-passes that each take the records of their sub-bins,
// every record in exactly one of them (all occurrences of a k-mer share minimizer, hence sub-bin).
DSK_HD u32 bin_hash(u32 minimizer)
{
    u32 h = minimizer * 0x9E3779B1u;
    h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
    return h;
}
DSK_HD u32 bin_of(u32 minimizer)
{
    const u32 h = bin_hash(minimizer);
    return h >> (32 - META_BIN_BITS);
}

// 64-bit finalizer (murmur3 fmix64) for hash-table slots
DSK_HD u64 mix64(u64 x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// ---- super-k-mer records ---------------------------------------------------------------------------------
// A record is RW = 2*KW 64-bit words.  Bases are packed MSB-first (base i of the super-k-mer sits in word
// i/32 at bits 62-2*(i%32)); the low 16 bits of the LAST word hold [nk:8][bank:8], so a record carries at
// most 32*RW-8 bases = maxS + k - 1 with maxS k-mers.  (Role of SuperKmer::save, K/Model.hpp:1386-1471.)
DSK_HD int rec_capacity_bases(int kw) { return 64 * kw - 8; }
DSK_HD int rec_max_kmers(int kw, int k)
{
    int s = rec_capacity_bases(kw) - k + 1;
    return s > 255 ? 255 : s;
}

// base i (0-based) of a record
template <int RW> DSK_HD int rec_base(const u64* w, int i) { return (int)((w[i >> 5] >> (62 - 2 * (i & 31))) & 3); }

// first k-mer (forward strand) of a record, k <= 32*KW (KW=1: k<=31 in practice, KW=2: k<=63)
template <int KW> struct Kmer { u64 w[KW]; };   // w[0] = least significant word

DSK_HD Kmer<1> rec_first_kmer1(const u64* r, int k) { Kmer<1> x; x.w[0] = r[0] >> (64 - 2 * k); return x; }
DSK_HD Kmer<2> rec_first_kmer2(const u64* r, int k)
{
    Kmer<2> x;                                             // top 2k bits of r[0]:r[1]:(r[2]) , 32 < k <= 63 or k == 32
    if (k <= 32) { x.w[0] = r[0] >> (64 - 2 * k); x.w[1] = 0; }
    else { int sh = 128 - 2 * k; x.w[1] = r[0] >> sh; x.w[0] = (r[0] << (64 - sh)) | (r[1] >> sh); }
    return x;
}

DSK_HD bool kmer_less(const Kmer<1>& a, const Kmer<1>& b) { return a.w[0] < b.w[0]; }
DSK_HD bool kmer_less(const Kmer<2>& a, const Kmer<2>& b) { return a.w[1] < b.w[1] || (a.w[1] == b.w[1] && a.w[0] < b.w[0]); }
DSK_HD bool kmer_eq(const Kmer<1>& a, const Kmer<1>& b) { return a.w[0] == b.w[0]; }
DSK_HD bool kmer_eq(const Kmer<2>& a, const Kmer<2>& b) { return a.w[0] == b.w[0] && a.w[1] == b.w[1]; }

// rolling update (K/Model.hpp:877-884): fwd = ((fwd<<2)+c)&mask ; rc = (rc>>2) + (comp(c) << 2(k-1))
DSK_HD void kmer_roll(Kmer<1>& f, Kmer<1>& r, int c, int k)
{
    const u64 mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
    f.w[0] = ((f.w[0] << 2) | (u64)c) & mask;
    r.w[0] = (r.w[0] >> 2) | ((u64)(c ^ 2) << (2 * (k - 1)));
}
DSK_HD void kmer_roll(Kmer<2>& f, Kmer<2>& r, int c, int k)
{
    const int hb = 2 * k - 64;                              // bits used in the high word (k > 32) or <= 0
    f.w[1] = (f.w[1] << 2) | (f.w[0] >> 62);
    f.w[0] = (f.w[0] << 2) | (u64)c;
    r.w[0] = (r.w[0] >> 2) | (r.w[1] << 62);
    r.w[1] = r.w[1] >> 2;
    if (hb > 0) {
        f.w[1] &= (hb >= 64) ? ~0ULL : ((1ULL << hb) - 1ULL);
        r.w[1] |= (u64)(c ^ 2) << (hb - 2);
    } else {
        f.w[1] = 0;
        f.w[0] &= (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
        r.w[0] |= (u64)(c ^ 2) << (2 * (k - 1));
    }
}
DSK_HD Kmer<1> kmer_revcomp(const Kmer<1>& f, int k) { Kmer<1> r; r.w[0] = revcomp64(f.w[0], k); return r; }
DSK_HD Kmer<2> kmer_revcomp(const Kmer<2>& f, int k) { Kmer<2> r; revcomp128(f.w[0], f.w[1], k, r.w[0], r.w[1]); return r; }
template <int KW> DSK_HD Kmer<KW> kmer_canonical(const Kmer<KW>& f, const Kmer<KW>& r) { return kmer_less(r, f) ? r : f; }

DSK_HD u64 kmer_hash(const Kmer<1>& a) { return mix64(a.w[0]); }
DSK_HD u64 kmer_hash(const Kmer<2>& a) { return mix64(a.w[0] ^ mix64(a.w[1] + 0x9E3779B97F4A7C15ULL)); }

// abundance -> 1-D histogram bin with the reference's quirks (Histogram.hpp:92 u16 truncation + clamp;
// Histogram.hpp:221 cache merge drops bin `length`).  Returns 0 when the count is not recorded.
DSK_HD u32 histo_bin(int32_t sum)
{
    u32 idx = (u32)sum & 0xFFFFu;
    return (idx >= 10000u) ? 0u : idx;                      // bin 10000 is never merged; bin 0 never printed
}

}  // namespace dsk
