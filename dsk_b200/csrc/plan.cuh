// plan.cuh -- the partition planner and the exchange layout, computed ON THE DEVICE.
//
// Replaces (K/ = thirdparty/gatb-core/gatb-core/src/gatb/kmer/impl/):
//   Repartitor::computeDistrib / justGroup          K/PartiInfo.cpp:48-106   (minimizer bins -> partitions)
//   RepartitorAlgorithm (sampling pass)             K/RepartitionAlgorithm.cpp:395-492
//   the nb_partitions sizing of ConfigurationAlgorithm  K/ConfigurationAlgorithm.cpp:245-467
//
// Round 1 planned on the host (bin histogram D2H, a sequential greedy packing, tables H2D): 0.8 ms of an 11.9 ms step on one
// GPU and 350 ms of an 823 ms step on the 4-GPU configs[2] shape (2^22 bins, 2.8 M partitions).  Here nothing but one small
// header crosses PCIe:
//
//   rule      bin b belongs to the partition that holds the k-mer offset at which the bin starts:
//               cut(b) = ex[b] / T          (ex = exclusive prefix of the whole-job k-mers per bin, T = target k-mers)
//             and partitions are numbered densely in bin order (a bin heavier than T does not leave empty ids behind).
//             A partition therefore holds T k-mers on average and at most T + its last bin.  Every rank derives the same
//             plan from the same all-reduced histogram; which partition a k-mer lands in is unobservable in the results.
//   heavy     partitions beyond the reach of the shared-memory table are renumbered after all the others (stable), so
//             that the host-driven paths (global table / sort) see them as one tail.
//   owners    partition p belongs to rank p % W.  Everything per partition is stored in "q order",
//               q(p) = (p % W) * PW + p / W,   PW = ceil(P / W)
//             so the partitions of one owner are a contiguous chunk [o * PW, (o + 1) * PW) of every table AND of the local
//             records after the scatter: the exchange is W - 1 contiguous copies (sender-major receive layout).
//
// The same rule runs sequentially on the host in plan_host() (CPU tests, and the GPU test that compares the two).
#pragma once
#include "kmer_bits.cuh"

namespace dsk {

constexpr int PLAN_MAXW = 16;                 // ranks of one job (one NVSwitch domain)

struct PlanParams {                           // scalars the host derives from the job totals (the same on every rank)
    u64 T;                                    // k-mers per partition
    u64 lim;                                  // a partition with more k-mers is "heavy" (host-driven global paths)
    u32 nbins;                                // bins of the planning level
    u32 W, me;
    u32 smem_ok;                              // 0: there is no shared-memory path in this mode, every partition is "heavy"
};

struct PlanHdr {                              // what the host reads back (one small D2H copy)
    u32 P, PW, nlight, nheavy;                // partitions, partitions per owner, ids [0, nlight) light, [nlight, P) heavy
    u64 need_recs[PLAN_MAXW];                 // whole-job records / k-mers of the partitions rank r owns
    u64 need_kmers[PLAN_MAXW];
    u64 send_recs[PLAN_MAXW];                 // records THIS rank holds for the partitions of rank r
};

// device table of the exchange + the segments the counting kernels read (written by k_xchg_bases)
struct XchgTab {
    u64 segptr[PLAN_MAXW];                    // segment s of owned job j starts at byte address segptr[s] + X[s * PW + j] * record bytes
    u64 dst[PLAN_MAXW];                       // byte address inside rank o's receive buffer where this rank's chunk goes
    u64 src_off[PLAN_MAXW];                   // first record of rank o's chunk in the local q-ordered records
    u64 cnt[PLAN_MAXW];                       // records of that chunk
    u32 bad, pad;                             // layout inconsistency detected on the device (reported by the host)
};

DSK_HD u32 plan_q(u32 p, u32 W, u32 PW) { return (p % W) * PW + p / W; }
DSK_HD bool plan_cut(u64 ex_prev, u64 ex_cur, u64 T) { return ex_prev / T != ex_cur / T; }
DSK_HD bool plan_heavy(u64 kmers, u64 recs, u64 lim, u32 smem_ok) { return !smem_ok || kmers > lim || recs >= 0xFFFFFFFFull; }

// ---- host mirror (sequential; CPU tests + GPU cross-check) -----------------------------------------------------------
// gh / lh: [2 * nbins] = records per bin, then k-mers per bin (whole job / this rank).  Outputs in q order, capacity
// nbins + W each; bin2q[nbins].  Returns the header.
inline PlanHdr plan_host(const PlanParams& pp, const u64* gh, const u64* lh, u32* bin2q, u64* gk_q, u64* gr_q, u64* lcnt_q)
{
    const u32 nb = pp.nbins, W = pp.W;
    const u64* gk = gh + nb;
    std::vector<u32> raw(nb);
    u32 P = 1;
    {
        u64 ex_prev = 0, ex = 0;
        for (u32 b = 0; b < nb; b++) {
            if (b && plan_cut(ex_prev, ex, pp.T)) P++;
            raw[b] = P - 1;
            ex_prev = ex; ex += gk[b];
        }
    }
    std::vector<u64> pk(P, 0), pr(P, 0), pl(P, 0);
    for (u32 b = 0; b < nb; b++) { pk[raw[b]] += gk[b]; pr[raw[b]] += gh[b]; pl[raw[b]] += lh[b]; }
    std::vector<u32> newid(P);
    u32 nheavy = 0;
    for (u32 p = 0; p < P; p++) if (plan_heavy(pk[p], pr[p], pp.lim, pp.smem_ok)) nheavy++;
    const u32 nlight = P - nheavy;
    { u32 l = 0, h = 0; for (u32 p = 0; p < P; p++) newid[p] = plan_heavy(pk[p], pr[p], pp.lim, pp.smem_ok) ? nlight + h++ : l++; }
    PlanHdr hdr; memset(&hdr, 0, sizeof hdr);
    hdr.P = P; hdr.PW = (P + W - 1) / W; hdr.nlight = nlight; hdr.nheavy = nheavy;
    const u32 PW = hdr.PW;
    for (u64 i = 0; i < (u64)W * PW; i++) { gk_q[i] = 0; gr_q[i] = 0; lcnt_q[i] = 0; }
    for (u32 p = 0; p < P; p++) {
        const u32 q = plan_q(newid[p], W, PW);
        gk_q[q] = pk[p]; gr_q[q] = pr[p]; lcnt_q[q] = pl[p];
        hdr.need_recs[newid[p] % W] += pr[p]; hdr.need_kmers[newid[p] % W] += pk[p]; hdr.send_recs[newid[p] % W] += pl[p];
    }
    for (u32 b = 0; b < nb; b++) bin2q[b] = plan_q(newid[raw[b]], W, PW);
    return hdr;
}

#ifdef __CUDACC__

// ---- device-wide exclusive prefix sum of u64 values produced by a functor (3 launches) ----------------------------------
constexpr int PS_T = 1024, PS_IPT = 4, PS_ITEMS = PS_T * PS_IPT;

__device__ __forceinline__ u64 ps_block_inclusive(u64 v, u64* s_w /*[32]*/, u64* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    u64 ws = s_w[lane], wi = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 o = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += o; }
    const u64 wpre = __shfl_sync(0xFFFFFFFFu, wi - ws, warp);
    if (total) *total = __shfl_sync(0xFFFFFFFFu, wi, 31);
    __syncthreads();                                               // s_w may be reused by the caller
    return wpre + inc;
}

// n = min(*n_dev, n_cap) when n_dev is given (sizes that only exist on the device), else n_cap
template <class F>
__global__ void __launch_bounds__(PS_T) k_ps_sums(F f, const u64* n_dev, u64 n_cap, u64* __restrict__ bsum)
{
    __shared__ u64 s_w[32];
    const u64 n = n_dev ? min(*n_dev, n_cap) : n_cap;
    const u64 base = (u64)blockIdx.x * PS_ITEMS + (u64)threadIdx.x * PS_IPT;
    u64 s = 0;
#pragma unroll
    for (int j = 0; j < PS_IPT; j++) if (base + j < n) s += f(base + j);
    u64 tot;
    ps_block_inclusive(s, s_w, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_ps_bsums(u64* __restrict__ bsum, u32 nblk)       // in place -> exclusive; bsum[nblk] = total
{
    __shared__ u64 s_w[32];
    u64 carry = 0;
    for (u32 c0 = 0; c0 < nblk; c0 += 1024) {
        const u32 i = c0 + threadIdx.x;
        const u64 v = i < nblk ? bsum[i] : 0;
        u64 tot;
        const u64 inc = ps_block_inclusive(v, s_w, &tot);
        if (i < nblk) bsum[i] = carry + inc - v;
        carry += tot;
    }
    if (threadIdx.x == 0) bsum[nblk] = carry;
}

template <class F>
__global__ void __launch_bounds__(PS_T) k_ps_apply(F f, const u64* n_dev, u64 n_cap, const u64* __restrict__ bsum, u32 nblk, u64* __restrict__ out)
{
    __shared__ u64 s_w[32];
    const u64 n = n_dev ? min(*n_dev, n_cap) : n_cap;
    const u64 base = (u64)blockIdx.x * PS_ITEMS + (u64)threadIdx.x * PS_IPT;
    u64 v[PS_IPT], s = 0;
#pragma unroll
    for (int j = 0; j < PS_IPT; j++) { v[j] = base + j < n ? f(base + j) : 0; s += v[j]; }
    const u64 inc = ps_block_inclusive(s, s_w, nullptr);
    u64 run = bsum[blockIdx.x] + inc - s;
#pragma unroll
    for (int j = 0; j < PS_IPT; j++) { if (base + j < n) out[base + j] = run; run += v[j]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nblk];      // the total, one past the end
}

struct PsLoad { const u64* p; __device__ u64 operator()(u64 i) const { return p[i]; } };
struct PsCutFlag { const u64* ex; u64 T; __device__ u64 operator()(u64 i) const { return (i && plan_cut(ex[i - 1], ex[i], T)) ? 1 : 0; } };
struct PsHeavyFlag { const u64* pk; const u64* pr; u64 lim; u32 smem_ok; __device__ u64 operator()(u64 i) const { return plan_heavy(pk[i], pr[i], lim, smem_ok) ? 1 : 0; } };

// returns the number of kernels launched
template <class F>
static int ps_scan(cudaStream_t st, F f, const u64* n_dev, u64 n_cap, u64* bsum /*[blocks + 1]*/, u64* out /*[n_cap + 1]*/)
{
    const u32 nblk = (u32)((n_cap + PS_ITEMS - 1) / PS_ITEMS);
    k_ps_sums<F><<<nblk ? nblk : 1, PS_T, 0, st>>>(f, n_dev, n_cap, bsum);
    k_ps_bsums<<<1, 1024, 0, st>>>(bsum, nblk ? nblk : 1);
    k_ps_apply<F><<<nblk ? nblk : 1, PS_T, 0, st>>>(f, n_dev, n_cap, bsum, nblk ? nblk : 1, out);
    return 3;
}
static inline size_t ps_bsum_bytes(u64 n_cap) { return ((n_cap + PS_ITEMS - 1) / PS_ITEMS + 2) * 8; }

// ---- planner kernels -------------------------------------------------------------------------------------------------
// raw partition of a bin: number of cuts at or before it
__device__ __forceinline__ u32 plan_raw(const u64* ex, const u64* E, u64 T, u32 b) { return (u32)E[b] + ((b && plan_cut(ex[b - 1], ex[b], T)) ? 1u : 0u); }

// per-partition sums of the raw partitions (k-mers / records of the whole job, records of this rank); nvals[0] = raw partitions
__global__ void __launch_bounds__(256) k_plan_sums(const u64* __restrict__ gh, const u64* __restrict__ lh, const u64* __restrict__ ex, const u64* __restrict__ E,
                                                   PlanParams pp, unsigned long long* pk, unsigned long long* pr, unsigned long long* pl, u64* nvals)
{
    const u32 nb = pp.nbins;
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
        const u32 p = plan_raw(ex, E, pp.T, b);
        const u64 k = gh[nb + b], r = gh[b], l = lh[b];
        if (k) atomicAdd(&pk[p], (unsigned long long)k);
        if (r) atomicAdd(&pr[p], (unsigned long long)r);
        if (l) atomicAdd(&pl[p], (unsigned long long)l);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) nvals[0] = E[nb] + 1;
}

// heavy partitions after the light ones (stable); everything per partition goes to q order.  H = exclusive scan of the heavy
// flags over the raw partitions.  nvals[1] = W * PW (length of the q-ordered tables).
__global__ void __launch_bounds__(256) k_plan_renumber(const u64* __restrict__ pk, const u64* __restrict__ pr, const u64* __restrict__ pl, const u64* __restrict__ H,
                                                       PlanParams pp, u64* nvals, u32* __restrict__ newid, u64* __restrict__ gk_q, u64* __restrict__ gr_q,
                                                       u64* __restrict__ lcnt_q, PlanHdr* hdr)
{
    const u32 P = (u32)nvals[0], W = pp.W, PW = (P + W - 1) / W;
    const u32 nheavy = (u32)H[P], nlight = P - nheavy;
    for (u32 p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        const bool hv = plan_heavy(pk[p], pr[p], pp.lim, pp.smem_ok);
        const u32 n = hv ? nlight + (u32)H[p] : p - (u32)H[p];
        newid[p] = n;
        const u32 q = plan_q(n, W, PW);
        gk_q[q] = pk[p]; gr_q[q] = pr[p]; lcnt_q[q] = pl[p];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { hdr->P = P; hdr->PW = PW; hdr->nlight = nlight; hdr->nheavy = nheavy; nvals[1] = (u64)W * PW; }
}

__global__ void __launch_bounds__(256) k_plan_bin2q(const u64* __restrict__ ex, const u64* __restrict__ E, const u32* __restrict__ newid, PlanParams pp,
                                                    const u64* nvals, u32* __restrict__ bin2q)
{
    const u32 P = (u32)nvals[0], W = pp.W, PW = (P + W - 1) / W;
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < pp.nbins; b += gridDim.x * blockDim.x)
        bin2q[b] = plan_q(newid[plan_raw(ex, E, pp.T, b)], W, PW);
}

// block r: what rank r receives (whole-job records / k-mers of its chunk) and what this rank holds for it
__global__ void __launch_bounds__(256) k_plan_hdr(const u64* __restrict__ gk_q, const u64* __restrict__ gr_q, const u64* __restrict__ loff, PlanHdr* hdr)
{
    __shared__ u64 s_a[256], s_b[256];
    const u32 r = blockIdx.x, PW = hdr->PW;
    u64 a = 0, b = 0;
    for (u32 j = threadIdx.x; j < PW; j += 256) { a += gr_q[(u64)r * PW + j]; b += gk_q[(u64)r * PW + j]; }
    s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
    __syncthreads();
    for (int d = 128; d; d >>= 1) { if ((int)threadIdx.x < d) { s_a[threadIdx.x] += s_a[threadIdx.x + d]; s_b[threadIdx.x] += s_b[threadIdx.x + d]; } __syncthreads(); }
    if (threadIdx.x == 0) { hdr->need_recs[r] = s_a[0]; hdr->need_kmers[r] = s_b[0]; hdr->send_recs[r] = loff[(u64)(r + 1) * PW] - loff[(u64)r * PW]; }
}

// several contexts in one process (dskgpu_multi_finish): out = sum of the ranks' bin histograms, read through peer pointers
struct PtrList { const u64* p[PLAN_MAXW]; };
__global__ void __launch_bounds__(256) k_sum_hists(PtrList src, int n, u64 len, u64* __restrict__ out)
{
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < len; i += (u64)gridDim.x * 256) {
        u64 a = 0;
        for (int r = 0; r < n; r++) a += src.p[r][i];
        out[i] = a;
    }
}

// ---- exchange: bases, send, gather -------------------------------------------------------------------------------------
// S[s * W + o] = records rank s holds for rank o (all-gathered send_recs); X = exclusive scan over the received count rows
// [W][PW] (row s = rank s's counts of MY partitions; row `me` = my own); peers[o] = byte address of rank o's receive buffer.
// The receive buffer of rank o holds the chunks of the senders s != o in rank order; a rank's own chunk stays where the
// scatter put it (lrecs) and is counted from there.
__global__ void k_xchg_bases(const PlanHdr* hdr, const u64* __restrict__ loff, const u64* __restrict__ X, const u64* __restrict__ S,
                             const u64* __restrict__ peers, u64 lrecs_addr, u32 W, u32 me, u32 RB, XchgTab* out)
{
    if (threadIdx.x || blockIdx.x) return;
    const u32 PW = hdr->PW;
    u32 bad = 0;
    u64 got = 0;
    for (u32 s = 0; s < W; s++) got += S[s * W + me];
    if (got != hdr->need_recs[me]) bad |= 1u;                                   // the ranks disagree on the plan
    if (X[(u64)W * PW] != got) bad |= 2u;                                       // count rows do not add up to the totals
    u64 rbase = 0;
    for (u32 s = 0; s < W; s++) {
        const u64 first = X[(u64)s * PW];
        if (s == me) out->segptr[s] = lrecs_addr + (loff[(u64)me * PW] - first) * RB;
        else { out->segptr[s] = peers[me] + (rbase - first) * RB; rbase += S[s * W + me]; }
    }
    for (u32 o = 0; o < W; o++) {
        u64 before = 0;
        for (u32 s = 0; s < me; s++) if (s != o) before += S[s * W + o];
        out->dst[o] = (o == me) ? 0 : peers[o] + before * RB;
        out->src_off[o] = loff[(u64)o * PW];
        out->cnt[o] = S[me * W + o];
        if (out->cnt[o] != loff[(u64)(o + 1) * PW] - loff[(u64)o * PW]) bad |= 4u;
    }
    out->bad = bad;
}

// this rank's chunk for every other rank: W - 1 contiguous copies into peer HBM (NVLink), all peers served at once
// (block b works for peer b % (W - 1)); 16-byte vector loads / stores, four in flight per thread
__global__ void __launch_bounds__(256) k_xchg_send(const ulonglong2* __restrict__ lrecs, const XchgTab* __restrict__ tab, u32 W, u32 me, u32 v_per_rec)
{
    const u32 npeer = W - 1;
    const u32 pi = blockIdx.x % npeer, sub = blockIdx.x / npeer, nsub = gridDim.x / npeer;
    if (sub >= nsub) return;
    const u32 o = pi >= me ? pi + 1 : pi;
    const ulonglong2* src = lrecs + tab->src_off[o] * v_per_rec;
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(tab->dst[o]);
    const u64 nv = tab->cnt[o] * v_per_rec;
    for (u64 i = (u64)sub * 1024 + threadIdx.x; i < nv; i += (u64)nsub * 1024) {
        ulonglong2 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) if (i + 256u * j < nv) v[j] = src[i + 256u * j];
#pragma unroll
        for (int j = 0; j < 4; j++) if (i + 256u * j < nv) dst[i + 256u * j] = v[j];
    }
}

// owned heavy partitions (jobs [j0, j0 + nj) of my chunk): their W segments are gathered into one partition-major buffer
// for the host-driven paths.  hoff[i] = first record of heavy job i in `out` (host-computed prefix of the whole-job records).
// The grid is split into groups of G blocks; group g takes the segments w = g, g + ngroups, ... and its G blocks share the
// vectors of a segment (a single hot minimizer = one partition of millions of records: every block must help copying it).
__global__ void __launch_bounds__(256) k_gather_heavy(const XchgTab* __restrict__ tab, const u64* __restrict__ X, u32 PW, u32 W, u32 j0, u32 nj,
                                                      const u64* __restrict__ hoff, ulonglong2* __restrict__ out, u32 v_per_rec, u32 G)
{
    const u32 ngroups = gridDim.x / G, grp = blockIdx.x / G, sub = blockIdx.x % G;
    if (grp >= ngroups) return;
    for (u32 w = grp; w < nj * W; w += ngroups) {
        const u32 i = w / W, s = w % W, j = j0 + i;
        u64 before = 0;
        for (u32 s2 = 0; s2 < s; s2++) before += X[(u64)s2 * PW + j + 1] - X[(u64)s2 * PW + j];
        const u64 a = X[(u64)s * PW + j], n = X[(u64)s * PW + j + 1] - a;
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(tab->segptr[s]) + a * v_per_rec;
        ulonglong2* dst = out + (hoff[i] + before) * v_per_rec;
        const u64 nv = n * v_per_rec;
        for (u64 x = (u64)sub * 256 + threadIdx.x; x < nv; x += (u64)G * 256) dst[x] = src[x];
    }
}

#endif  // __CUDACC__
}  // namespace dsk
