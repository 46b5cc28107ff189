// scan.cuh -- K1: device record scanner.  Raw FASTA/FASTQ bytes -> compacted code stream
// (one byte per base: 0..3, 4|code for non-ACGT, 8 = record separator).
//
// Replaces BankFasta::Iterator::get_next_seq_from_file + ConvertASCII
// (G/src/gatb/bank/impl/BankFasta.cpp:485-572, G/src/gatb/tools/misc/api/Data.hpp:185), which the reference
// runs single-threaded under the iterate lock (ICommand.hpp:304-331).
//
// Parallelisation: the scanner is a tiny state machine (line type for FASTA, line index mod 4 for FASTQ).
// Every 32-byte thread chunk is summarised as a transition table  state_in -> (state_out, #codes emitted);
// tables compose associatively, so a block scan (k_scan_tables), a scan over tiles (k_scan_tiles) and a
// second block scan (k_scan_emit) give every thread its true input state and output offset.  The same
// table builder / composer runs on the host in dskgpu_selftest_scan (tests pin it against the oracle).
#pragma once
#include "kmer_bits.cuh"

namespace dsk {

constexpr int SCAN_BPT = 32;                 // bytes per thread
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = SCAN_BPT * SCAN_THREADS;   // 8192 bytes per tile

// transition table over <=4 states: st = 4 x 2 bits (state_out per state_in), cnt = 4 x u16
struct Tab { u32 st; u64 cnt; };
DSK_HD Tab tab_identity() { Tab t; t.st = 0xE4u; t.cnt = 0; return t; }
DSK_HD int tab_state(const Tab& t, int s) { return (t.st >> (2 * s)) & 3; }
DSK_HD u32 tab_count(const Tab& t, int s) { return (u32)((t.cnt >> (16 * s)) & 0xFFFFu); }
// apply f first, then g
DSK_HD Tab tab_compose(const Tab& f, const Tab& g)
{
    Tab r; r.st = 0; r.cnt = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        int fs = tab_state(f, s);
        r.st |= (u32)tab_state(g, fs) << (2 * s);
        r.cnt |= (u64)((tab_count(f, s) + tab_count(g, fs)) & 0xFFFFu) << (16 * s);
    }
    return r;
}

// bytes of one thread chunk plus its neighbours; inactive bytes (outside the stream window) are skipped
struct Chunk {
    u8 b[SCAN_BPT];
    u32 active;          // bit i: byte i belongs to the stream window
    int prev, next;      // byte before b[0] / after b[31] as seen by the scanner
};

// table of one chunk.  One pass, exploiting the structure of each format (see file header).
DSK_HD Tab chunk_table(int fmt, const Chunk& c)
{
    Tab t;
    if (fmt == FMT_FASTA) {
        // run from ST_SEQ; both states converge at the first line start inside the chunk
        int state = ST_SEQ, err = 0; u32 cnt = 0, cnt_prefix = 0; bool had_ls = false;
        int prev = c.prev;
        for (int i = 0; i < SCAN_BPT; i++) {
            if (!((c.active >> i) & 1)) continue;
            int ch = c.b[i];
            int nx = (i + 1 < SCAN_BPT && ((c.active >> (i + 1)) & 1)) ? c.b[i + 1] : c.next;
            if (prev == '\n' && !had_ls) { had_ls = true; cnt_prefix = cnt; }
            if (scan_step(FMT_FASTA, state, prev, ch, nx, err) >= 0) cnt++;
            prev = ch;
        }
        if (!had_ls) cnt_prefix = cnt;
        u32 st_hdr = had_ls ? (u32)state : (u32)ST_HDR;
        t.st = (u32)state | (st_hdr << 2) | (2u << 4) | (3u << 6);
        t.cnt = (u64)cnt | ((u64)(cnt - cnt_prefix) << 16);
        return t;
    }
    if (fmt == FMT_FASTQ) {
        // segment q (after q newlines) is the sequence line for start phase s = (1-q)&3
        u32 cn[4] = {0, 0, 0, 0}; int q = 0;
        for (int i = 0; i < SCAN_BPT; i++) {
            if (!((c.active >> i) & 1)) continue;
            int ch = c.b[i];
            int nx = (i + 1 < SCAN_BPT && ((c.active >> (i + 1)) & 1)) ? c.b[i + 1] : c.next;
            int s = (1 - q) & 3;
            if (ch == '\n') { cn[s]++; q = (q + 1) & 3; }
            else if (!(ch == '\r' && (nx == '\n' || nx < 0))) cn[s]++;
        }
        t.st = 0; t.cnt = 0;
        for (int s = 0; s < 4; s++) { t.st |= (u32)((s + q) & 3) << (2 * s); t.cnt |= (u64)cn[s] << (16 * s); }
        return t;
    }
    // FMT_LINES: every byte emits exactly one code, single state
    u32 n = 0;
    for (int i = 0; i < SCAN_BPT; i++) n += (c.active >> i) & 1;
    t.st = 0xE4u; t.cnt = (u64)n * 0x0001000100010001ULL;
    return t;
}

// emission of one chunk given its true input state; out[] receives the codes (<= 32). returns count.
DSK_HD int chunk_emit(int fmt, const Chunk& c, int state, u8* out, int& err, u32& nsep, u32& nbase)
{
    int n = 0, prev = c.prev;
    for (int i = 0; i < SCAN_BPT; i++) {
        if (!((c.active >> i) & 1)) continue;
        int ch = c.b[i];
        int nx = (i + 1 < SCAN_BPT && ((c.active >> (i + 1)) & 1)) ? c.b[i + 1] : c.next;
        int e = scan_step(fmt, state, prev, ch, nx, err);
        if (e >= 0) { out[n++] = (u8)e; if (e == CODE_SEP) nsep++; else nbase++; }
        prev = ch;
    }
    return n;
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

__device__ __forceinline__ void load_chunk(const u8* __restrict__ raw, u64 a, u64 lo, u64 hi, const StreamState* ss,
                                           int next_after, Chunk& c)
{
    c.active = 0;
    if (a >= lo && a + SCAN_BPT <= hi) {
        const uint4* p = reinterpret_cast<const uint4*>(raw + a);
        uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
        u32 w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int i = 0; i < SCAN_BPT; i++) c.b[i] = (u8)(w[i >> 2] >> (8 * (i & 3)));
        c.active = 0xFFFFFFFFu;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_BPT; i++) {
            u64 x = a + i; bool in = (x >= lo && x < hi);
            c.b[i] = in ? raw[x] : 0;
            c.active |= (in ? 1u : 0u) << i;
        }
    }
    // neighbours
    if (a > lo && a <= hi) c.prev = raw[a - 1]; else c.prev = ss->prev_byte;
    // `next` is consulted after the last ACTIVE byte of the chunk
    u64 last_end = (a + SCAN_BPT < hi) ? a + SCAN_BPT : hi;
    c.next = (last_end < hi) ? (int)raw[last_end] : (next_after == -2 ? (int)raw[hi] : next_after);   // -2: stream continues in place
}

__device__ __forceinline__ Tab tab_shfl_up(const Tab& t, int d)
{
    Tab o; o.st = __shfl_up_sync(0xFFFFFFFFu, t.st, d); o.cnt = __shfl_up_sync(0xFFFFFFFFu, t.cnt, d); return o;
}

// block-wide exclusive scan of tables (compose order = thread order). Returns the exclusive prefix of this
// thread; *block_total gets the composition of the whole block (valid in all threads).
__device__ __forceinline__ Tab block_scan_tabs(Tab mine, Tab* s_warp /*[8]*/, Tab* block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Tab inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Tab o = tab_shfl_up(inc, d);
        if (lane >= d) inc = tab_compose(o, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    Tab pre = tab_identity();
    for (int w = 0; w < warp; w++) pre = tab_compose(pre, s_warp[w]);
    Tab exl = tab_shfl_up(inc, 1);
    if (lane == 0) exl = tab_identity();
    Tab tot = tab_identity();
    for (int w = 0; w < SCAN_THREADS / 32; w++) tot = tab_compose(tot, s_warp[w]);
    *block_total = tot;
    return tab_compose(pre, exl);
}

// pass A: one table per tile
template <int FMT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tables(const u8* __restrict__ raw, u64 lo, u64 hi, u64 tile_first,
                                                               const StreamState* ss, int next_after, TileTab* tabs)
{
    __shared__ Tab s_warp[SCAN_THREADS / 32];
    u64 tile = tile_first + blockIdx.x;
    u64 a = tile * SCAN_TILE + (u64)threadIdx.x * SCAN_BPT;
    Chunk c; load_chunk(raw, a, lo, hi, ss, next_after, c);
    Tab mine = chunk_table(FMT, c);
    Tab tot; block_scan_tabs(mine, s_warp, &tot);
    if (threadIdx.x == 0) {
        TileTab tt; tt.st = tot.st;
        for (int s = 0; s < 4; s++) tt.cnt[s] = tab_count(tot, s);
        tabs[blockIdx.x] = tt;
    }
}

// pass B: chain the tile tables (single block of 1024 threads, each owning a contiguous run of tiles)
__global__ void __launch_bounds__(1024) k_scan_tiles(const TileTab* __restrict__ tabs, u64 ntiles, TileIn* tin, StreamState* ss,
                                                     const u8* raw, u64 lo, u64 hi)
{
    __shared__ u32 s_st[1024];
    __shared__ u32 s_cnt[1024][4];
    __shared__ u64 s_base[1024];
    __shared__ u32 s_in[1024];
    const int t = threadIdx.x;
    u64 per = (ntiles + 1023) / 1024;
    u64 b = (u64)t * per, e = b + per; if (e > ntiles) e = ntiles; if (b > ntiles) b = ntiles;
    // phase 1: table of my run, for all 4 start states
    u32 st[4] = {0, 1, 2, 3}; u32 cn[4] = {0, 0, 0, 0};
    for (u64 i = b; i < e; i++) {
        TileTab tt = tabs[i];
        for (int s = 0; s < 4; s++) { cn[s] += tt.cnt[st[s]]; st[s] = (tt.st >> (2 * st[s])) & 3; }
    }
    s_st[t] = st[0] | (st[1] << 2) | (st[2] << 4) | (st[3] << 6);
    for (int s = 0; s < 4; s++) s_cnt[t][s] = cn[s];
    __syncthreads();
    // phase 2: thread 0 chains the 1024 run tables
    if (t == 0) {
        u32 s = ss->fsm_state; u64 base = 0;
        for (int i = 0; i < 1024; i++) {
            s_in[i] = s; s_base[i] = base;
            base += s_cnt[i][s]; s = (s_st[i] >> (2 * s)) & 3;
        }
        ss->fsm_state = s;
        ss->total = ss->carry + base;
        ss->prev_byte_next = (hi > lo) ? (int)raw[hi - 1] : ss->prev_byte;
    }
    __syncthreads();
    // phase 3: replay my run with the true start state
    u32 s = s_in[t]; u64 base = s_base[t];
    for (u64 i = b; i < e; i++) {
        TileTab tt = tabs[i];
        TileIn ti; ti.base = base; ti.state = s; ti.pad = 0; tin[i] = ti;
        base += tt.cnt[s]; s = (tt.st >> (2 * s)) & 3;
    }
}

// pass C: emit codes.  `carry0` = number of codes already at the front of the code buffer.
template <int FMT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_emit(const u8* __restrict__ raw, u64 lo, u64 hi, u64 tile_first,
                                                             const StreamState* ss_ro, StreamState* ss, int next_after,
                                                             const TileIn* __restrict__ tin, u8* __restrict__ codes)
{
    __shared__ Tab s_warp[SCAN_THREADS / 32];
    __shared__ u8 s_out[SCAN_TILE];
    __shared__ u32 s_red[3];
    u64 tile = tile_first + blockIdx.x;
    u64 a = tile * SCAN_TILE + (u64)threadIdx.x * SCAN_BPT;
    if (threadIdx.x < 3) s_red[threadIdx.x] = 0;
    Chunk c; load_chunk(raw, a, lo, hi, ss_ro, next_after, c);
    Tab mine = chunk_table(FMT, c);
    Tab tot; Tab exl = block_scan_tabs(mine, s_warp, &tot);
    TileIn ti = tin[blockIdx.x];
    int state = tab_state(exl, ti.state);
    u32 off = tab_count(exl, ti.state);
    u32 tile_cnt = tab_count(tot, ti.state);
    int err = 0; u32 nsep = 0, nbase = 0;
    u8 tmp[SCAN_BPT];
    int n = chunk_emit(FMT, c, state, tmp, err, nsep, nbase);
    for (int i = 0; i < n; i++) s_out[off + i] = tmp[i];
    // stats / errors
    u32 wsep = __reduce_add_sync(0xFFFFFFFFu, nsep), wbase = __reduce_add_sync(0xFFFFFFFFu, nbase);
    u32 werr = __reduce_or_sync(0xFFFFFFFFu, (u32)err);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_red[0], wsep); atomicAdd(&s_red[1], wbase); atomicOr(&s_red[2], werr); }
    __syncthreads();
    u8* dst = codes + ss_ro->carry + ti.base;
    for (u32 i = threadIdx.x; i < tile_cnt; i += SCAN_THREADS) dst[i] = s_out[i];
    if (threadIdx.x == 0) {
        if (s_red[0]) atomicAdd((unsigned long long*)&ss->nsep, (unsigned long long)s_red[0]);
        if (s_red[1]) atomicAdd((unsigned long long*)&ss->nbase, (unsigned long long)s_red[1]);
        if (s_red[2]) atomicOr(&ss->err, s_red[2]);
    }
}

// after the super-k-mer pass: keep the last k-1 codes as the carry of the next chunk
__global__ void k_scan_carry(u8* codes, StreamState* ss, int k)
{
    __shared__ u8 s[64];
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
    ss->fsm_state = (fmt == FMT_FASTA) ? ST_HDR : 0;   // FASTA: bytes before the first header are skipped
    ss->prev_byte = '\n'; ss->prev_byte_next = '\n';
    ss->carry = 0; ss->total = 0;
}

#endif  // __CUDACC__
}  // namespace dsk
