// scan.cuh -- K1: device record scanner.  Raw FASTA/FASTQ bytes -> compacted code stream
// (one byte per base: 0..3, 4|code for non-ACGT, 8 = record separator).
//
// Replaces BankFasta::Iterator::get_next_seq_from_file + ConvertASCII
// (G/src/gatb/bank/impl/BankFasta.cpp:485-572, G/src/gatb/tools/misc/api/Data.hpp:185), which the reference
// runs single-threaded under the iterate lock (ICommand.hpp:304-331).
//
// The scanner is a tiny state machine (scan_step in kmer_bits.cuh: line type for FASTA, line index mod 4 for
// FASTQ).  Every 32-byte thread chunk is turned into bit masks (newline, header start, ...) held in registers;
// from the masks a transition table  state_in -> (state_out, #codes emitted)  follows in a few bit operations.
// Tables compose associatively, so a block scan (k_scan_tables), a scan over tiles (k_scan_tiles) and a second
// block scan (k_scan_emit) give every thread its true input state and output offset.  The mask evaluators are
// host+device functions; dskgpu_selftest_scan replays them on the host against the byte-wise scan_step.
#pragma once
#include "kmer_bits.cuh"

namespace dsk {

constexpr int SCAN_BPT = 32;                 // bytes per thread
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = SCAN_BPT * SCAN_THREADS;   // 8192 bytes per tile

DSK_HD int popc32(u32 x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
DSK_HD int ctz32(u32 x)      // x != 0
{
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
DSK_HD int msb32(u32 x)      // x != 0
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}

// ---- per-chunk bit masks (bit i <-> byte i of the chunk) ---------------------------------------------------
struct CMasks { u32 nl, gt, plus, at, cr, active; };

DSK_HD u32 bits_below(int b) { return b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u); }     // bits [0, b)

// 4-bit mask of the bytes of w equal to the byte replicated in pat4 (exact SWAR zero-byte test + multiply gather)
DSK_HD u32 eqmask4(u32 w, u32 pat4)
{
    const u32 x = w ^ pat4;
    const u32 z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;     // 0x80 in every byte of x that is zero
    return ((z >> 7) * 0x10204080u) >> 28;
}
DSK_HD u32 byte_at(const u32* w, int p)                      // byte p of the chunk without dynamic register indexing
{
    u32 x = w[0];
#pragma unroll
    for (int i = 1; i < 8; i++) x = ((p >> 2) == i) ? w[i] : x;
    return (x >> (8 * (p & 3))) & 0xFFu;
}

// newline / CR masks for all bytes (SWAR); '>', '@', '+' only matter at line starts, so they are looked up there only
DSK_HD void chunk_masks(const u32* w /*[8]*/, u32 active, bool prev_nl, CMasks& m)
{
    u32 nl = 0, cr = 0, anycr = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        nl |= eqmask4(w[i], 0x0A0A0A0Au) << (4 * i);
        const u32 x = w[i] ^ 0x0D0D0D0Du;
        anycr |= ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
    }
    if (anycr) {
#pragma unroll
        for (int i = 0; i < 8; i++) cr |= eqmask4(w[i], 0x0D0D0D0Du) << (4 * i);
    }
    nl &= active; cr &= active;
    u32 gt = 0, at = 0, plus = 0;
    u32 ls = ((nl << 1) | ((prev_nl && active) ? (1u << ctz32(active)) : 0u)) & active;
    while (ls) {
        const int p = ctz32(ls);
        const u32 c = byte_at(w, p);
        gt   |= (((c == '>') | (c == '@')) ? 1u : 0u) << p;
        at   |= (c == '@' ? 1u : 0u) << p;
        plus |= (c == '+' ? 1u : 0u) << p;
        ls &= ls - 1;
    }
    m.nl = nl; m.gt = gt; m.plus = plus; m.at = at; m.cr = cr; m.active = active;     // gt/at/plus: line starts only
}

// CR bytes that are dropped: followed by '\n' or by the end of the stream (BankFasta.cpp:471)
DSK_HD u32 cr_dropped(const CMasks& m, bool next_is_nl_or_eof)
{
    if (m.cr == 0) return 0;
    u32 follow = (m.nl >> 1);
    if (m.active && next_is_nl_or_eof) follow |= 1u << msb32(m.active);
    return m.cr & follow;
}

// FASTA evaluation for input state `st_in` (ST_SEQ / ST_HDR): emitted-byte mask, separator mask, output state.
// Line starts: byte after '\n' (prev_nl tells it for byte 0).  A line starting with '>' or '@' is a header
// (BankFasta.cpp:528); one starting with '+' would switch the reference to FASTQ parsing -> flagged.
DSK_HD void fasta_eval(const CMasks& m, bool prev_nl, bool next_flag, int st_in, u32& em, u32& sep, int& st_out, u32& err_ls)
{
    const u32 ls = ((m.nl << 1) | ((prev_nl && m.active) ? (1u << ctz32(m.active)) : 0u)) & m.active;   // prev_nl applies to the first active byte
    const u32 hdr_like = ls & (m.gt | m.plus);
    u32 hm = 0; int cur = st_in; int start = 0; u32 it = ls;
    while (it) {
        const int p = ctz32(it);
        if (cur == ST_HDR) hm |= bits_below(p) & ~bits_below(start);
        cur = ((hdr_like >> p) & 1) ? ST_HDR : ST_SEQ;
        start = p; it &= it - 1;
    }
    if (cur == ST_HDR) hm |= ~bits_below(start);
    sep = ls & m.gt;
    em = (m.active & ~hm & ~m.nl & ~cr_dropped(m, next_flag)) | sep;
    st_out = cur;
    err_ls = ls & m.plus;
}

// FASTQ evaluation for input phase `ph_in` (line index mod 4): sequence lines are the segments of phase 1;
// their '\n' becomes the record separator.
DSK_HD void fastq_eval(const CMasks& m, bool prev_nl, bool next_flag, int ph_in, u32& em, u32& sep, int& ph_out, u32& err)
{
    const u32 ls = ((m.nl << 1) | ((prev_nl && m.active) ? (1u << ctz32(m.active)) : 0u)) & m.active;   // prev_nl applies to the first active byte
    u32 seqm = 0; int ph = ph_in; int start = 0; u32 it = m.nl; err = 0;
    for (;;) {
        const int p = it ? ctz32(it) : 32;                              // next newline (inclusive end of the segment)
        const u32 seg = (p >= 31 ? 0xFFFFFFFFu : bits_below(p + 1)) & ~bits_below(start);
        if (ph == 1) seqm |= seg;
        const u32 seg_ls = seg & ls;                                    // the segment's first byte if it starts a line
        if (seg_ls) {
            if (ph == 0 && !(seg_ls & (m.at | m.nl))) err |= SCAN_ERR_FASTQ_AT;
            if (ph == 2 && !(seg_ls & m.plus)) err |= SCAN_ERR_FASTQ_PLUS;
        }
        if (p >= 32) break;
        ph = (ph + 1) & 3; start = p + 1; it &= it - 1;
        if (start >= 32) break;
    }
    seqm &= m.active;
    em = seqm & ~cr_dropped(m, next_flag);
    sep = em & m.nl;
    ph_out = ph;
}

// ---- transition tables ------------------------------------------------------------------------------------------
// generic form over <= 4 states: st = 4 x 2 bits (state_out per state_in), cnt = 4 x u16 (codes emitted)
struct Tab { u32 st; u64 cnt; };
DSK_HD Tab tab_identity() { Tab t; t.st = 0xE4u; t.cnt = 0; return t; }
DSK_HD int tab_state(const Tab& t, int s) { return (t.st >> (2 * s)) & 3; }
DSK_HD u32 tab_count(const Tab& t, int s) { return (u32)((t.cnt >> (16 * s)) & 0xFFFFu); }
// apply f first, then g.  FASTQ tables are rotations (state_out = state_in + q), FASTA tables have 2 states.
template <int FMT>
DSK_HD Tab tab_compose(const Tab& f, const Tab& g)
{
    Tab r;
    if (FMT == FMT_LINES) { r.st = 0xE4u; r.cnt = f.cnt + g.cnt; return r; }
    if (FMT == FMT_FASTQ) {
        const u32 qf = f.st & 3u, qg = g.st & 3u, q = (qf + qg) & 3u;       // st of state 0 is the rotation
        const u64 gc = qf ? ((g.cnt >> (16 * qf)) | (g.cnt << (64 - 16 * qf))) : g.cnt;
        r.cnt = f.cnt + gc;                                                   // 4 x u16 lanes, no lane overflows (<= 8192)
        r.st = q | (((q + 1) & 3u) << 2) | (((q + 2) & 3u) << 4) | (((q + 3) & 3u) << 6);
        return r;
    }
    // FASTA: states 0 (SEQ) and 1 (HDR)
    const u32 f0 = f.st & 1u, f1 = (f.st >> 2) & 1u;
    const u32 g0 = g.st & 1u, g1 = (g.st >> 2) & 1u;
    const u32 gc0 = (u32)(g.cnt & 0xFFFFu), gc1 = (u32)((g.cnt >> 16) & 0xFFFFu);
    const u32 r0 = f0 ? g1 : g0, r1 = f1 ? g1 : g0;
    const u32 c0 = (u32)(f.cnt & 0xFFFFu) + (f0 ? gc1 : gc0), c1 = (u32)((f.cnt >> 16) & 0xFFFFu) + (f1 ? gc1 : gc0);
    r.st = r0 | (r1 << 2) | (2u << 4) | (3u << 6);
    r.cnt = (u64)c0 | ((u64)c1 << 16);
    return r;
}

template <int FMT>
DSK_HD Tab chunk_table(const CMasks& m, bool prev_nl, bool next_flag)
{
    Tab t;
    if (FMT == FMT_FASTA) {
        u32 em0, em1, sep, e; int so0, so1;
        fasta_eval(m, prev_nl, next_flag, ST_SEQ, em0, sep, so0, e);
        fasta_eval(m, prev_nl, next_flag, ST_HDR, em1, sep, so1, e);
        t.cnt = (u64)popc32(em0) | ((u64)popc32(em1) << 16);
        t.st = (u32)so0 | ((u32)so1 << 2) | (2u << 4) | (3u << 6);
        return t;
    }
    if (FMT == FMT_FASTQ) {
        t.st = 0; t.cnt = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            u32 em, sep, e; int so;
            fastq_eval(m, prev_nl, next_flag, s, em, sep, so, e);
            t.cnt |= (u64)popc32(em) << (16 * s);
            t.st |= (u32)so << (2 * s);
        }
        return t;
    }
    t.cnt = (u64)popc32(m.active);
    t.st = 0xE4u;
    return t;
}

// masks of the bytes to emit for the true input state
template <int FMT>
DSK_HD void chunk_emit_masks(const CMasks& m, bool prev_nl, bool next_flag, int state, u32& em, u32& sep, u32& err)
{
    int so;
    if (FMT == FMT_FASTA) { u32 e; fasta_eval(m, prev_nl, next_flag, state, em, sep, so, e); err = e ? (u32)SCAN_ERR_PLUS_IN_FASTA : 0u; }
    else if (FMT == FMT_FASTQ) { fastq_eval(m, prev_nl, next_flag, state, em, sep, so, err); }
    else { em = m.active; sep = m.nl; err = 0; }
}

// code of byte c (0..3 | 4 when not ACGTacgt)
DSK_HD u32 encode_fast(u32 c)
{
    const u32 u = (c & 0xDFu) - 65u;                               // 'A' -> 0, 'C' -> 2, 'G' -> 6, 'T' -> 19
    const bool ok = (u < 20u) && ((0x80045u >> u) & 1u);
    return ((c >> 1) & 3u) | (ok ? 0u : (u32)CODE_INVALID);
}

// codes of 4 bytes at once: (c >> 1) & 3 per byte, | 4 in the bytes that are not one of ACGTacgt
DSK_HD u32 encode4(u32 w)
{
    const u32 codes = (w >> 1) & 0x03030303u;
    // expected upper-case letter of each code (A C T G for 0..3): bytes of 0x47544341 selected by the codes
    u32 exp;
#ifdef __CUDA_ARCH__
    const u32 t = (codes | (codes >> 4)) & 0x00330033u;             // codes of bytes 0,1 / 2,3 as nibbles
    exp = __byte_perm(0x47544341u, 0u, (t | (t >> 8)) & 0x3333u);
#else
    exp = 0;
    for (int j = 0; j < 4; j++) exp |= ((0x47544341u >> (8 * ((codes >> (8 * j)) & 3u))) & 0xFFu) << (8 * j);
#endif
    const u32 x = (w & 0xDFDFDFDFu) ^ exp;                          // zero byte <=> valid base
    const u32 nz = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
    return codes | (nz >> 5);
}
// byte-permute selector that moves the bytes whose bit is set in `nib` to the low end, in order
DSK_HD u32 compact_sel(u32 nib)
{
    u32 sel = 0, pos = 0;
#pragma unroll
    for (u32 j = 0; j < 4; j++) { if ((nib >> j) & 1u) { sel |= j << (4 * pos); pos++; } }
    return sel;
}
// host + device byte permute (low 4 selector nibbles, selecting among the 4 bytes of a)
DSK_HD u32 perm4(u32 a, u32 sel)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(a, 0u, sel);
#else
    u32 r = 0;
    for (int j = 0; j < 4; j++) r |= ((a >> (8 * ((sel >> (4 * j)) & 3u))) & 0xFFu) << (8 * j);
    return r;
#endif
}

// the emitted codes of one 4-byte word, packed to the low end: `nib` = bytes to emit, `sepnib` = those that are separators
DSK_HD u32 word_codes(u32 w, u32 nib, u32 sepnib)
{
    u32 codes = encode4(w);
    const u32 sm4 = ((sepnib * 0x00204081u) & 0x01010101u) * 0xFFu;        // 0xFF in the separator bytes
    codes = (codes & ~sm4) | (0x08080808u & sm4);
    if (nib == 0xFu) return codes;
    return perm4(codes, compact_sel(nib)) & ((1u << (8 * popc32(nib))) - 1u);      // bytes beyond the emitted ones must be zero: callers OR them
}

// ---- device side ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// per-stream scanner state, lives in device memory so chunks chain without host round trips
struct StreamState {
    u32 fsm_state;       // scanner state at the start of the next chunk
    int prev_byte;       // last byte of the previous chunk ('\n' at stream start)
    int prev_byte_next;  // last byte of the chunk being scanned (committed by k_scan_carry)
    u32 pad0;
    u64 carry;           // codes already in the code buffer (tail of the previous chunk, < k)
    u64 total;           // codes in the buffer after the current chunk was scanned
    u32 err;             // SCAN_ERR_* flags
    u32 pad;
    u64 nsep, nbase;     // stats: records / nucleotides seen
};

struct TileTab { u32 st; u32 cnt[4]; };          // tile-level table (counts fit u32: <= 8192)
struct TileIn  { u64 base; u32 state; u32 pad; };

// loads the 32 bytes of a thread chunk (absolute byte index a of the 16-byte aligned buffer `raw`); bytes outside
// the stream window [lo, hi) are inactive.  Also returns the neighbour information the scanner needs.
__device__ __forceinline__ void load_chunk(const u8* __restrict__ raw, u64 a, u64 lo, u64 hi, const StreamState* ss,
                                           int next_after, u32* w, u32& active, bool& prev_nl, bool& next_flag)
{
    if (a >= lo && a + SCAN_BPT <= hi) {
        const uint4* p = reinterpret_cast<const uint4*>(raw + a);
        uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
        w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        active = 0xFFFFFFFFu;
    } else {
        active = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = 0;
        if (a + SCAN_BPT > lo && a < hi) {
#pragma unroll
            for (int i = 0; i < SCAN_BPT; i++) {
                u64 x = a + i; bool in = (x >= lo && x < hi);
                u32 c = in ? raw[x] : 0u;
                w[i >> 2] |= c << (8 * (i & 3));
                active |= (in ? 1u : 0u) << i;
            }
        }
    }
    int prev = (a > lo && a <= hi) ? (int)raw[a - 1] : ss->prev_byte;
    prev_nl = (prev == '\n');
    // byte following the last active byte of this chunk
    u64 last_end = (a + SCAN_BPT < hi) ? a + SCAN_BPT : hi;
    int next = (last_end < hi) ? (int)raw[last_end] : (next_after == -2 ? (int)raw[hi] : next_after);   // -2: stream continues in place
    next_flag = (next == '\n') || (next < 0);
}

__device__ __forceinline__ Tab tab_shfl_up(const Tab& t, int d)
{
    Tab o; o.st = __shfl_up_sync(0xFFFFFFFFu, t.st, d); o.cnt = __shfl_up_sync(0xFFFFFFFFu, t.cnt, d); return o;
}

// block-wide exclusive scan of tables (compose order = thread order). Returns the exclusive prefix of this
// thread; *block_total gets the composition of the whole block (valid in all threads).
template <int FMT>
__device__ __forceinline__ Tab block_scan_tabs(Tab mine, Tab* s_warp /*[8]*/, Tab* block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Tab inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Tab o = tab_shfl_up(inc, d);
        if (lane >= d) inc = tab_compose<FMT>(o, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    Tab pre = tab_identity(), tot = tab_identity();
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        Tab x = s_warp[w];
        if (w < warp) pre = tab_compose<FMT>(pre, x);
        tot = tab_compose<FMT>(tot, x);
    }
    Tab exl = tab_shfl_up(inc, 1);
    if (lane == 0) exl = tab_identity();
    *block_total = tot;
    return tab_compose<FMT>(pre, exl);
}

// pass A: one table per tile
template <int FMT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tables(const u8* __restrict__ raw, u64 lo, u64 hi, u64 tile_first,
                                                               const StreamState* ss, int next_after, TileTab* tabs)
{
    __shared__ Tab s_warp[SCAN_THREADS / 32];
    const u64 tile = tile_first + blockIdx.x;
    const u64 a = tile * SCAN_TILE + (u64)threadIdx.x * SCAN_BPT;
    u32 w[8], active; bool prev_nl, next_flag;
    load_chunk(raw, a, lo, hi, ss, next_after, w, active, prev_nl, next_flag);
    CMasks m; chunk_masks(w, active, prev_nl, m);
    Tab mine = chunk_table<FMT>(m, prev_nl, next_flag);
    Tab tot; block_scan_tabs<FMT>(mine, s_warp, &tot);
    if (threadIdx.x == 0) {
        TileTab tt; tt.st = tot.st;
        for (int s = 0; s < 4; s++) tt.cnt[s] = tab_count(tot, s);
        tabs[blockIdx.x] = tt;
    }
}

// pass B: chain the tile tables (single block of 1024 threads, each owning a contiguous run of tiles; the 1024 run
// tables are chained by a block-wide scan -- tables compose associatively)
struct RunTab { u32 st; u32 cnt[4]; };          // st: 4 x 2 bits state_out per state_in; cnt: codes emitted per state_in (no overflow: u32 per chunk)
__device__ __forceinline__ RunTab run_compose(const RunTab& f, const RunTab& g)   // f first, then g
{
    RunTab r; r.st = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const u32 m = (f.st >> (2 * s)) & 3u;
        // no dynamic indexing of g.cnt (would go to local memory)
        const u32 gc = m == 0 ? g.cnt[0] : m == 1 ? g.cnt[1] : m == 2 ? g.cnt[2] : g.cnt[3];
        r.cnt[s] = f.cnt[s] + gc;
        r.st |= ((g.st >> (2 * m)) & 3u) << (2 * s);
    }
    return r;
}
__device__ __forceinline__ RunTab run_shfl_up(const RunTab& t, int d)
{
    RunTab o; o.st = __shfl_up_sync(0xFFFFFFFFu, t.st, d);
#pragma unroll
    for (int s = 0; s < 4; s++) o.cnt[s] = __shfl_up_sync(0xFFFFFFFFu, t.cnt[s], d);
    return o;
}

__global__ void __launch_bounds__(1024) k_scan_tiles(const TileTab* __restrict__ tabs, u64 ntiles, TileIn* tin, StreamState* ss,
                                                     const u8* raw, u64 lo, u64 hi)
{
    __shared__ RunTab s_warp[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    u64 per = (ntiles + 1023) / 1024;
    u64 b = (u64)t * per, e = b + per; if (e > ntiles) e = ntiles; if (b > ntiles) b = ntiles;
    // phase 1: table of my run, for all 4 start states
    u32 st[4] = {0, 1, 2, 3}; u32 cn[4] = {0, 0, 0, 0};
#pragma unroll 8                                                  // the loads of a batch do not depend on the chain: issued together
    for (u64 i = b; i < e; i++) {
        TileTab tt = tabs[i];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const u32 m = st[s];
            cn[s] += m == 0 ? tt.cnt[0] : m == 1 ? tt.cnt[1] : m == 2 ? tt.cnt[2] : tt.cnt[3];
            st[s] = (tt.st >> (2 * m)) & 3;
        }
    }
    RunTab mine; mine.st = st[0] | (st[1] << 2) | (st[2] << 4) | (st[3] << 6);
#pragma unroll
    for (int s = 0; s < 4; s++) mine.cnt[s] = cn[s];
    // phase 2: inclusive scan of the run tables (warp shuffles, then the 32 warp totals)
    RunTab inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { RunTab o = run_shfl_up(inc, d); if (lane >= d) inc = run_compose(o, inc); }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        RunTab w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { RunTab o = run_shfl_up(w, d); if (lane >= d) w = run_compose(o, w); }
        s_warp[lane] = w;                                        // inclusive over warps
    }
    __syncthreads();
    RunTab exl = run_shfl_up(inc, 1);                            // exclusive within the warp
    if (lane == 0) { exl.st = 0xE4u; exl.cnt[0] = exl.cnt[1] = exl.cnt[2] = exl.cnt[3] = 0; }
    if (warp > 0) exl = run_compose(s_warp[warp - 1], exl);
    // my true start state and output base
    const u32 s0 = ss->fsm_state;
    u32 s = (exl.st >> (2 * s0)) & 3;
    u64 base = s0 == 0 ? exl.cnt[0] : s0 == 1 ? exl.cnt[1] : s0 == 2 ? exl.cnt[2] : exl.cnt[3];
    // phase 3: replay my run with the true start state
#pragma unroll 8
    for (u64 i = b; i < e; i++) {
        TileTab tt = tabs[i];
        TileIn ti; ti.base = base; ti.state = s; ti.pad = 0; tin[i] = ti;
        base += s == 0 ? tt.cnt[0] : s == 1 ? tt.cnt[1] : s == 2 ? tt.cnt[2] : tt.cnt[3];
        s = (tt.st >> (2 * s)) & 3;
    }
    __syncthreads();                                             // everybody has read ss->fsm_state
    if (t == 1023) {
        ss->fsm_state = s;
        ss->total = ss->carry + base;
        ss->prev_byte_next = (hi > lo) ? (int)raw[hi - 1] : ss->prev_byte;
    }
}

// pass C: emit codes into the code buffer behind the carry of the previous chunk
template <int FMT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_emit(const u8* __restrict__ raw, u64 lo, u64 hi, u64 tile_first,
                                                             const StreamState* ss_ro, StreamState* ss, int next_after,
                                                             const TileIn* __restrict__ tin, u8* __restrict__ codes)
{
    __shared__ Tab s_warp[SCAN_THREADS / 32];
    __shared__ __align__(16) u8 s_out[SCAN_TILE + 32];
    __shared__ u32 s_red[3];
    const u64 tile = tile_first + blockIdx.x;
    const u64 a = tile * SCAN_TILE + (u64)threadIdx.x * SCAN_BPT;
    if (threadIdx.x < 3) s_red[threadIdx.x] = 0;
    u32 w[8], active; bool prev_nl, next_flag;
    load_chunk(raw, a, lo, hi, ss_ro, next_after, w, active, prev_nl, next_flag);
    CMasks m; chunk_masks(w, active, prev_nl, m);
    Tab mine = chunk_table<FMT>(m, prev_nl, next_flag);
    Tab tot; Tab exl = block_scan_tabs<FMT>(mine, s_warp, &tot);
    const TileIn ti = tin[blockIdx.x];
    const int state = tab_state(exl, ti.state);
    const u32 tile_cnt = tab_count(tot, ti.state);
    u8* dst = codes + ss_ro->carry + ti.base;
    const u32 pad = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);           // smem mirrors the 16-byte phase of dst
    u32 off = pad + tab_count(exl, ti.state);
    u32 em, sep, err;
    chunk_emit_masks<FMT>(m, prev_nl, next_flag, state, em, sep, err);
    // compaction, a 4-byte word at a time: SWAR encode, byte-permute the emitted bytes to the front, append to a 64-bit
    // accumulator; whole aligned words go out with one 32-bit store, only the ragged head / tail bytes are stored singly
    {
        u32* s_out32 = reinterpret_cast<u32*>(s_out);
        const u32 head = off & 3u;                                 // bytes of the first aligned word that belong to my predecessor
        u32 wpos = off >> 2;
        u64 acc = 0; u32 fill = head; bool first = head != 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 nib = (em >> (4 * i)) & 0xFu;
            if (nib == 0) continue;
            const u32 packed = word_codes(w[i], nib, (sep >> (4 * i)) & 0xFu);
            acc |= (u64)packed << (8 * fill);
            fill += (u32)__popc(nib);
            if (fill >= 4) {
                const u32 word = (u32)acc;
                if (first) {                                       // bytes head..3 are mine
                    for (u32 j = head; j < 4; j++) s_out[4 * wpos + j] = (u8)(word >> (8 * j));
                    first = false;
                } else s_out32[wpos] = word;
                wpos++; acc >>= 32; fill -= 4;
            }
        }
        // tail: bytes [first ? head : 0, fill) of the last word
        for (u32 j = first ? head : 0u; j < fill; j++) s_out[4 * wpos + j] = (u8)((u32)acc >> (8 * j));
    }
    // stats / errors
    const u32 nsep = __popc(sep & em), nbase = __popc(em) - nsep;
    const u32 wsep = __reduce_add_sync(0xFFFFFFFFu, nsep), wbase = __reduce_add_sync(0xFFFFFFFFu, nbase);
    const u32 werr = __reduce_or_sync(0xFFFFFFFFu, err);
    if ((threadIdx.x & 31) == 0) { if (wsep) atomicAdd(&s_red[0], wsep); if (wbase) atomicAdd(&s_red[1], wbase); if (werr) atomicOr(&s_red[2], werr); }
    __syncthreads();
    // copy out: 16-byte vectors where a whole vector is ours, bytes at the two ragged ends
    u8* dst0 = dst - pad;
    const u32 end = pad + tile_cnt;
    const u32 nvec = (end + 15) / 16;
    for (u32 q = threadIdx.x; q < nvec; q += SCAN_THREADS) {
        const u32 b0 = q * 16, b1 = b0 + 16;
        if (b0 >= pad && b1 <= end) reinterpret_cast<uint4*>(dst0)[q] = reinterpret_cast<const uint4*>(s_out)[q];
        else for (u32 j = (b0 > pad ? b0 : pad); j < (b1 < end ? b1 : end); j++) dst0[j] = s_out[j];
    }
    if (threadIdx.x == 0) {
        if (s_red[0]) atomicAdd((unsigned long long*)&ss->nsep, (unsigned long long)s_red[0]);
        if (s_red[1]) atomicAdd((unsigned long long*)&ss->nbase, (unsigned long long)s_red[1]);
        if (s_red[2]) atomicOr(&ss->err, s_red[2]);
    }
}

// after the super-k-mer pass: keep the last k-1 codes as the carry of the next chunk
__global__ void k_scan_carry(u8* codes, StreamState* ss, int k)
{
    __shared__ u8 s[128];                                          // k - 1 <= 126 codes, one per thread (launched with 128 threads)
    u64 total = ss->total;
    u64 c = (total < (u64)(k - 1)) ? total : (u64)(k - 1);
    if (threadIdx.x < c) s[threadIdx.x] = codes[total - c + threadIdx.x];
    __syncthreads();
    if (threadIdx.x < c) codes[threadIdx.x] = s[threadIdx.x];
    if (threadIdx.x == 0) { ss->carry = c; ss->total = c; ss->prev_byte = ss->prev_byte_next; }
}

// end of a bank: nothing left can form a k-mer; restart the scanner
__global__ void k_scan_reset_stream(StreamState* ss, int fmt)
{
    ss->fsm_state = (fmt == FMT_FASTA) ? ST_HDR : 0;
    ss->prev_byte = '\n'; ss->prev_byte_next = '\n';
    ss->carry = 0; ss->total = 0;
}

#endif  // __CUDACC__
}  // namespace dsk
