// radix.cuh -- K5: LSD radix sort, 8-bit digits, one-sweep (single pass per digit: chained-scan look-back
// fused into the scatter), keys of KW 64-bit words with an optional 32-bit payload.
//
// Replaces SortCommand::execute / executeSort (std::sort per kx-mer bucket, K/PartitionsCommand.cpp:1400-1504).
// HBM traffic: one histogram read of the keys for all digits, then per digit pass one read + one write.
//
// Layout per pass: tile = 256 threads x ITEMS keys, warp-striped.  Ranking inside the tile is stable
// (match_any per warp, per-warp digit counters, then cross-warp prefix), tile digit counts are published to
// `status[tile][digit]` (aggregate, then inclusive prefix) and predecessors are looked back, keys are staged
// through shared memory in tile-sorted order so that global writes are runs of consecutive addresses per digit.
#pragma once
#include "kmer_bits.cuh"

namespace dsk {
#ifdef __CUDACC__

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr u32 RS_FLAG_AGG = 1u << 30, RS_FLAG_INC = 2u << 30, RS_MASK = (1u << 30) - 1u;
#ifndef RS_ITEMS1
#define RS_ITEMS1 12
#endif
#ifndef RS_ITEMS2
#define RS_ITEMS2 6
#endif
#ifndef RS_MINB
#define RS_MINB 4
#endif
template <int KW> struct RsCfg { static constexpr int ITEMS = (KW == 1) ? RS_ITEMS1 : (KW == 2) ? RS_ITEMS2 : 4; static constexpr int TILE = RS_THREADS * ITEMS; };

template <int KW> __device__ __forceinline__ u32 rs_digit(const u64* key, int pass)
{
    // no dynamic indexing: it would push the caller's key registers into local memory
    u64 w = key[0];
    if constexpr (KW == 2) w = (pass & 8) ? key[1] : key[0];
    if constexpr (KW == 3) { const int q = pass >> 3; w = q == 0 ? key[0] : q == 1 ? key[1] : key[2]; }
    if constexpr (KW == 4) { const int q = pass >> 3; w = q == 0 ? key[0] : q == 1 ? key[1] : q == 2 ? key[2] : key[3]; }
    return (u32)(w >> ((pass & 7) * 8)) & 0xFFu;
}

// histogram of the digits [first, npass) of every key in one read: hist[pass][256]
template <int KW>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const u64* __restrict__ keys, u64 n, int npass, unsigned long long* __restrict__ hist, int first = 0)
{
    extern __shared__ u32 s_h[];                                  // [npass][256]
    for (int i = threadIdx.x; i < npass * 256; i += RS_THREADS) s_h[i] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (u64)gridDim.x * RS_THREADS) {
        u64 key[KW];
#pragma unroll
        for (int q = 0; q < KW; q++) key[q] = keys[i * KW + q];
        for (int p = first; p < npass; p++) atomicAdd(&s_h[p * 256 + rs_digit<KW>(key, p)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += RS_THREADS) { u32 v = s_h[i]; if (v) atomicAdd(&hist[i], (unsigned long long)v); }
}

// exclusive scan of each pass's 256 bins (one block per pass)
__global__ void __launch_bounds__(256) k_rs_scan(unsigned long long* hist)
{
    __shared__ unsigned long long s[256];
    unsigned long long* h = hist + (u64)blockIdx.x * 256;
    unsigned long long v = h[threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    unsigned long long acc = 0;
    for (int i = 0; i < (int)threadIdx.x; i++) acc += s[i];
    h[threadIdx.x] = acc;
}

template <int KW, bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS, RS_MINB) k_rs_onesweep(const u64* __restrict__ in_keys, u64* __restrict__ out_keys,
                                                            const u32* __restrict__ in_vals, u32* __restrict__ out_vals,
                                                            u64 n, int pass, const unsigned long long* __restrict__ gbase /*[256]*/,
                                                            u32* status /*[ntiles][256]*/, u32* tile_counter)
{
    constexpr int ITEMS = RsCfg<KW>::ITEMS;
    constexpr int TILE = RsCfg<KW>::TILE;
    __shared__ u32 s_wc[RS_WARPS][256];                           // per-warp digit counters -> exclusive over warps
    __shared__ u32 s_dstart[256];                                 // first tile-sorted index of each digit
    __shared__ unsigned long long s_goff[256];                    // global index of tile-sorted index 0 of each digit
    extern __shared__ __align__(16) unsigned char s_dyn[];          // tile-sorted staging: keys then payloads
    u64* s_keys = reinterpret_cast<u64*>(s_dyn);
    u32* s_vals = reinterpret_cast<u32*>(s_dyn + (size_t)TILE * KW * 8);
    __shared__ u32 s_tile;
    __shared__ u32 s_wsum[RS_WARPS];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = t; i < RS_WARPS * 256; i += RS_THREADS) (&s_wc[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u64 tile0 = (u64)tile * TILE;
    const u32 nvalid = (u32)((n - tile0 < (u64)TILE) ? (n - tile0) : (u64)TILE);

    // 1. load (warp-striped) and rank within the warp, stable
    u64 key[ITEMS][KW]; u32 rd[ITEMS];                             // rd = rank | digit << 16
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        u32 li = (u32)warp * 32 * ITEMS + r * 32 + lane;           // index inside the tile
        bool ok = li < nvalid;
#pragma unroll
        for (int q = 0; q < KW; q++) key[r][q] = ok ? in_keys[(tile0 + li) * KW + q] : ~0ULL;
        u32 d = ok ? rs_digit<KW>(key[r], pass) : 256u;            // padding sorts after everything, never written
        u32 peers = __match_any_sync(0xFFFFFFFFu, d);
        u32 pre = (d < 256u) ? s_wc[warp][d] : 0u;
        rd[r] = (pre + __popc(peers & ((1u << lane) - 1u))) | (d << 16);
        __syncwarp();
        if (d < 256u && lane == (__ffs((int)peers) - 1)) s_wc[warp][d] = pre + __popc(peers);
        __syncwarp();
    }
    __syncthreads();

    // 2. thread d: exclusive prefix of digit d over warps, tile count, publish + look back
    {
        const int d = t;
        u32 s = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { u32 c = s_wc[w][d]; s_wc[w][d] = s; s += c; }
        volatile u32* st = status + (u64)tile * 256 + d;
        u32 excl = 0;
        if (tile == 0) { *st = s | RS_FLAG_INC; }
        else {
            *st = s | RS_FLAG_AGG;
            int lt = (int)tile - 1;
            while (lt >= 0) {
                u32 v = *(volatile u32*)(status + (u64)lt * 256 + d);
                if (v & RS_FLAG_INC) { excl += v & RS_MASK; break; }
                if (v & RS_FLAG_AGG) { excl += v & RS_MASK; lt--; }
            }
            *st = (excl + s) | RS_FLAG_INC;
        }
        // exclusive scan of s over digits -> s_dstart
        u32 inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += x; }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        u32 wpre = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) if (w < warp) wpre += s_wsum[w];
        u32 dstart = wpre + inc - s;
        s_dstart[d] = dstart;
        s_goff[d] = gbase[d] + (unsigned long long)excl - (unsigned long long)dstart;
    }
    __syncthreads();

    // 3. stage in tile-sorted order
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        u32 d = rd[r] >> 16;
        if (d < 256u) {
            u32 idx = s_dstart[d] + s_wc[warp][d] + (rd[r] & 0xFFFFu);
#pragma unroll
            for (int q = 0; q < KW; q++) s_keys[idx * KW + q] = key[r][q];
            if (HAS_VAL) s_vals[idx] = in_vals[tile0 + (u32)warp * 32 * ITEMS + r * 32 + lane];   // payload fetched late: keeps 16 registers free
        }
    }
    __syncthreads();

    // 4. write out: consecutive tile-sorted indices of one digit are consecutive in global memory
    for (u32 i = t; i < nvalid; i += RS_THREADS) {
        u64 kk[KW];
#pragma unroll
        for (int q = 0; q < KW; q++) kk[q] = s_keys[i * KW + q];
        u32 d = rs_digit<KW>(kk, pass);
        u64 g = s_goff[d] + i;
#pragma unroll
        for (int q = 0; q < KW; q++) out_keys[g * KW + q] = kk[q];
        if (HAS_VAL) out_vals[g] = s_vals[i];
    }
}

// ---- ordering of the solid set: top digits by LSD passes, the rest by a neighbourhood fix-up ------------------------------
// The solid k-mers of a read set are n distinct, nearly uniform keys: once they are ordered by their top ceil(log2 n) bits
// (ceil(log2 n / 8) one-sweep passes instead of 2k/8), the keys that share a prefix are neighbours, groups of one or two
// (Poisson, mean < 1).  Every thread ranks its key inside its group by scanning the neighbours with the same prefix and
// stores it at its final position: one more pass instead of four or five (k = 31) or twelve (k = 63).  Keys that do not
// behave (a group longer than RS_FIX_LIMIT: low-complexity sets, tiny k) raise `*flag`; the caller then runs the plain
// full-width LSD sort from the same input, which is still intact.
constexpr u32 RS_FIX_LIMIT = 48;
template <int KW> __device__ __forceinline__ bool rs_same_prefix(const u64* a, const u64* b, int shift)
{
    if constexpr (KW == 1) return ((a[0] ^ b[0]) >> shift) == 0;
    else return shift >= 64 ? (((a[1] ^ b[1]) >> (shift - 64)) == 0) : (a[1] == b[1] && ((a[0] ^ b[0]) >> shift) == 0);
}
template <int KW> __device__ __forceinline__ bool rs_key_less(const u64* a, const u64* b)
{
    if constexpr (KW == 1) return a[0] < b[0];
    else return a[1] < b[1] || (a[1] == b[1] && a[0] < b[0]);
}
template <int KW>
__global__ void __launch_bounds__(256) k_rs_fix(const u64* __restrict__ in_keys, const u32* __restrict__ in_vals, u64* __restrict__ out_keys,
                                                u32* __restrict__ out_vals, u64 n, int shift, unsigned int* flag)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    u64 my[KW], o[KW];
#pragma unroll
    for (int q = 0; q < KW; q++) my[q] = in_keys[i * KW + q];
    u32 smaller = 0, left = 0, right = 0;
    for (u64 j = i; j > 0;) {
        j--;
#pragma unroll
        for (int q = 0; q < KW; q++) o[q] = in_keys[j * KW + q];
        if (!rs_same_prefix<KW>(my, o, shift)) break;
        left++;
        smaller += rs_key_less<KW>(my, o) ? 0u : 1u;              // (an equal key on the left stays on the left)
        if (left > RS_FIX_LIMIT) { atomicExch(flag, 1u); return; }
    }
    for (u64 j = i + 1; j < n; j++) {
#pragma unroll
        for (int q = 0; q < KW; q++) o[q] = in_keys[j * KW + q];
        if (!rs_same_prefix<KW>(my, o, shift)) break;
        right++;
        smaller += rs_key_less<KW>(o, my) ? 1u : 0u;
        if (right > RS_FIX_LIMIT) { atomicExch(flag, 1u); return; }
    }
    const u64 pos = i - left + smaller;
#pragma unroll
    for (int q = 0; q < KW; q++) out_keys[pos * KW + q] = my[q];
    out_vals[pos] = in_vals[i];
}

#endif  // __CUDACC__
}  // namespace dsk
