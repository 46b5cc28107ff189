// kmer_wide.cuh -- N-word k-mer logic for the spans beyond 64 (KSIZE_LIST 96 / 128: k <= 95 in three 64-bit words,
// k <= 127 in four).  The wide spans run on the device through the sort path (k_superkmers<3|4>, partition
// scatter, k_expand_keys, one-sweep radix sort, k_rle_emit); the overloads at the end of this header plug the N-word logic
// into the names the kernels use (kmer_roll, kmer_revcomp, kmer_less, ...).  Like kmer_bits.cuh everything is a pure
// host+device function, pinned on the CPU (tests/test_host_logic.py::test_wide_kmer_logic_matches_the_wide_oracle) against
// oracle/liboracle_wide.so, which is itself pinned against the reference built with KSIZE_LIST "32 64 96 128".
//
// Reference semantics restated: LargeInt<precision> value = sum code_i * 4^(k-1-i) (K/Model.hpp:636-657), compared most
// significant word first (LargeInt.hpp:502-509); rolling update K/Model.hpp:877-884; revcomp LargeInt.hpp:722-735.
#pragma once
#include "kmer_bits.cuh"

namespace dsk {

// w[0] = least significant word, as Kmer<KW> in kmer_bits.cuh
template <int KW> DSK_HD bool kmern_less(const Kmer<KW>& a, const Kmer<KW>& b)
{
#pragma unroll
    for (int i = KW - 1; i >= 0; i--) { if (a.w[i] != b.w[i]) return a.w[i] < b.w[i]; }
    return false;
}
template <int KW> DSK_HD bool kmern_eq(const Kmer<KW>& a, const Kmer<KW>& b)
{
    bool e = true;
#pragma unroll
    for (int i = 0; i < KW; i++) e = e && (a.w[i] == b.w[i]);
    return e;
}

// keep the low 2k bits
template <int KW> DSK_HD void kmern_mask(Kmer<KW>& x, int k)
{
    const int top = (2 * k - 1) >> 6, bits = 2 * k - 64 * top;            // bits used in word `top`: 2..64
#pragma unroll
    for (int i = 0; i < KW; i++) {
        if (i > top) x.w[i] = 0;
        else if (i == top && bits < 64) x.w[i] &= (1ULL << bits) - 1ULL;
    }
}

// rolling update: fwd = ((fwd << 2) | c) & mask ; rc = (rc >> 2) | (comp(c) << 2(k-1)), comp(c) = c ^ 2
template <int KW> DSK_HD void kmern_roll(Kmer<KW>& f, Kmer<KW>& r, int c, int k)
{
#pragma unroll
    for (int i = KW - 1; i > 0; i--) f.w[i] = (f.w[i] << 2) | (f.w[i - 1] >> 62);
    f.w[0] = (f.w[0] << 2) | (u64)c;
    kmern_mask<KW>(f, k);
#pragma unroll
    for (int i = 0; i < KW - 1; i++) r.w[i] = (r.w[i] >> 2) | (r.w[i + 1] << 62);
    r.w[KW - 1] >>= 2;
    const int pos = 2 * (k - 1);
#pragma unroll
    for (int i = 0; i < KW; i++) if (i == (pos >> 6)) r.w[i] |= (u64)(c ^ 2) << (pos & 63);
}

// reverse complement: complement every base, reverse the 32*KW two-bit groups, drop the padding groups
template <int KW> DSK_HD Kmer<KW> kmern_revcomp(const Kmer<KW>& f, int k)
{
    Kmer<KW> t;
#pragma unroll
    for (int i = 0; i < KW; i++) t.w[KW - 1 - i] = rev2_64(f.w[i] ^ 0xAAAAAAAAAAAAAAAAULL);
    const int sh = 64 * KW - 2 * k;                                       // >= 2 (k < 32*KW), < 64*KW
    const int ws = sh >> 6, bs = sh & 63;
    Kmer<KW> r;
#pragma unroll
    for (int i = 0; i < KW; i++) {
        const u64 lo = (i + ws < KW) ? t.w[i + ws] : 0ULL;
        const u64 hi = (i + ws + 1 < KW) ? t.w[i + ws + 1] : 0ULL;
        r.w[i] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
    return r;
}
template <int KW> DSK_HD Kmer<KW> kmern_canonical(const Kmer<KW>& f, const Kmer<KW>& r) { return kmern_less<KW>(r, f) ? r : f; }

// 64 bits of a record starting at stream bit b (records pack bases MSB first: base i sits in word i/32 at bits 62-2(i%32));
// bits past the RW words read as zero
template <int RW> DSK_HD u64 recn_bits64(const u64* r, int b)
{
    const int q = b >> 6, o = b & 63;
    const u64 a = (q < RW) ? r[q] : 0ULL, c = (q + 1 < RW) ? r[q + 1] : 0ULL;
    return o ? ((a << o) | (c >> (64 - o))) : a;
}

// the k-mer that starts at base j of a record (forward strand): stream bits [2j, 2j + 2k)
template <int KW, int RW> DSK_HD Kmer<KW> recn_kmer_at(const u64* r, int j, int k)
{
    Kmer<KW> x;
    const int s = 2 * j, e = 2 * j + 2 * k;
#pragma unroll
    for (int i = 0; i < KW; i++) {
        const int b = e - 64 * (i + 1);                                   // stream bit where value word i starts
        if (b >= s) x.w[i] = recn_bits64<RW>(r, b);
        else {
            const int nv = e - 64 * i - s;                                // bits of the k-mer left for this word
            x.w[i] = nv > 0 ? (recn_bits64<RW>(r, s) >> (64 - nv)) : 0ULL;
        }
    }
    return x;
}

template <int KW> DSK_HD u64 kmern_hash(const Kmer<KW>& a)
{
    u64 h = 0x9E3779B97F4A7C15ULL;
#pragma unroll
    for (int i = 0; i < KW; i++) h = mix64(h ^ a.w[i]);
    return h;
}

// ---- the names the kernels use, for the wide key types (Kmer<1> / Kmer<2> keep their hand-written versions in kmer_bits.cuh) ----
DSK_HD bool kmer_less(const Kmer<3>& a, const Kmer<3>& b) { return kmern_less<3>(a, b); }
DSK_HD bool kmer_less(const Kmer<4>& a, const Kmer<4>& b) { return kmern_less<4>(a, b); }
DSK_HD bool kmer_eq(const Kmer<3>& a, const Kmer<3>& b) { return kmern_eq<3>(a, b); }
DSK_HD bool kmer_eq(const Kmer<4>& a, const Kmer<4>& b) { return kmern_eq<4>(a, b); }
DSK_HD void kmer_roll(Kmer<3>& f, Kmer<3>& r, int c, int k) { kmern_roll<3>(f, r, c, k); }
DSK_HD void kmer_roll(Kmer<4>& f, Kmer<4>& r, int c, int k) { kmern_roll<4>(f, r, c, k); }
DSK_HD Kmer<3> kmer_revcomp(const Kmer<3>& f, int k) { return kmern_revcomp<3>(f, k); }
DSK_HD Kmer<4> kmer_revcomp(const Kmer<4>& f, int k) { return kmern_revcomp<4>(f, k); }
DSK_HD u64 kmer_hash(const Kmer<3>& a) { return kmern_hash<3>(a); }
DSK_HD u64 kmer_hash(const Kmer<4>& a) { return kmern_hash<4>(a); }

// first k-mer (forward strand) of a record, any key width
template <int KW> DSK_HD Kmer<KW> rec_first_kmer(const u64* r, int k)
{
    if constexpr (KW == 1) return rec_first_kmer1(r, k);
    else if constexpr (KW == 2) return rec_first_kmer2(r, k);
    else return recn_kmer_at<KW, 2 * KW>(r, 0, k);
}

}  // namespace dsk
