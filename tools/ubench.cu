// ubench.cu -- micro-benchmarks that size the design choices of the counting kernels on B200 (run under gpurun):
// random 32-byte probes / RED / CAS on an L2-resident table, the same in shared memory, MLP variants.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef uint64_t u64; typedef uint32_t u32;

__device__ __forceinline__ u64 mix64(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__device__ __forceinline__ void ld256cg(const u64* p, u64& a, u64& b, u64& c, u64& d)
{ asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory"); }

// MODE bits: 1 = probe load (32 B), 2 = RED on counts, 4 = CAS on keys (25 % of items), 8 = atomicAdd with return instead of RED
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k_global(u64* keys, u32* counts, u32 bmask, u64 items_per_thread, u64* sink)
{
    u64 acc = 0;
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (u64 it = 0; it < items_per_thread; it += ILP) {
        u64 h[ILP]; u64 v[ILP][4];
#pragma unroll
        for (int i = 0; i < ILP; i++) h[i] = mix64(tid * 0x9E3779B97F4A7C15ULL + it + i);
#pragma unroll
        for (int i = 0; i < ILP; i++) { if (MODE & 1) ld256cg(keys + 4 * (u64)((u32)h[i] & bmask), v[i][0], v[i][1], v[i][2], v[i][3]); else { v[i][0] = h[i]; v[i][1] = v[i][2] = v[i][3] = 0; } }
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const u32 b = (u32)h[i] & bmask;
            u32 sl = 4 * b + ((MODE & 1) ? (u32)((v[i][0] ^ v[i][1] ^ v[i][2] ^ v[i][3]) & 3) : (u32)(h[i] >> 40) & 3);
            if ((MODE & 4) && ((h[i] >> 50) & 3) == 0) { u64 old = atomicCAS((unsigned long long*)&keys[sl], ~0ULL, ~0ULL); acc += old; }
            if (MODE & 2) atomicAdd(&counts[sl], 1u);
            if (MODE & 8) acc += atomicAdd(&counts[sl], 1u);
            acc += v[i][0];
        }
    }
    if (acc == 0x1234567) *sink = acc;
}

// packed slot: 64-bit word, RED.64 add on it after a probe
template <int ILP>
__global__ void __launch_bounds__(256) k_packed(u64* keys, u32 bmask, u64 items_per_thread, u64* sink)
{
    u64 acc = 0;
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (u64 it = 0; it < items_per_thread; it += ILP) {
        u64 h[ILP]; u64 v[ILP][4];
#pragma unroll
        for (int i = 0; i < ILP; i++) h[i] = mix64(tid * 0x9E3779B97F4A7C15ULL + it + i);
#pragma unroll
        for (int i = 0; i < ILP; i++) ld256cg(keys + 4 * (u64)((u32)h[i] & bmask), v[i][0], v[i][1], v[i][2], v[i][3]);
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const u32 b = (u32)h[i] & bmask;
            u32 sl = 4 * b + (u32)((v[i][0] ^ v[i][1] ^ v[i][2] ^ v[i][3]) & 3);
            atomicAdd((unsigned long long*)&keys[sl], 1ULL);
        }
    }
    if (acc == 0x1234567) *sink = acc;
}

// shared-memory table: MODE 1 = LDS.64 probe, 2 = ATOMS.ADD u32, 4 = ATOMS.CAS.64 (25 %)
template <int MODE>
__global__ void __launch_bounds__(1024) k_shared(u32 nslots_mask, u64 items_per_thread, u64* sink)
{
    extern __shared__ u64 s_tab[];                       // keys [nslots] then counts [nslots] (u32)
    u32* s_cnt = (u32*)(s_tab + nslots_mask + 1);
    for (u32 i = threadIdx.x; i <= nslots_mask; i += blockDim.x) { s_tab[i] = ~0ULL; s_cnt[i] = 0; }
    __syncthreads();
    u64 acc = 0;
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (u64 it = 0; it < items_per_thread; it++) {
        const u64 h = mix64(tid * 0x9E3779B97F4A7C15ULL + it);
        const u32 sl = (u32)h & nslots_mask;
        if (MODE & 1) acc += s_tab[sl];
        if ((MODE & 4) && ((h >> 50) & 3) == 0) acc += atomicCAS((unsigned long long*)&s_tab[sl], ~0ULL, ~0ULL);
        if (MODE & 2) atomicAdd(&s_cnt[sl], 1u);
    }
    if (acc == 0x1234567) *sink = acc;
}

template <typename F> static float timeit(F f, int reps = 5)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return best;
}

int main(int argc, char** argv)
{
    const int log2slots = argc > 1 ? atoi(argv[1]) : 23;
    const u64 nslots = 1ULL << log2slots;
    u64* keys; u32* counts; u64* sink;
    cudaMalloc(&keys, nslots * 8); cudaMalloc(&counts, nslots * 4); cudaMalloc(&sink, 8);
    cudaMemset(keys, 0xFF, nslots * 8); cudaMemset(counts, 0, nslots * 4);
    const u32 bmask = (u32)(nslots / 4 - 1);
    const int blocks = 148 * 8;
    const u64 ipt = 64;                                  // items per thread
    const double items = (double)blocks * 256 * ipt;
    printf("table 2^%d slots (%.0f MB keys + %.0f MB counts), %d blocks x 256 thr x %llu items = %.1f M items\n", log2slots, nslots * 8 / 1e6, nslots * 4 / 1e6,
           blocks, (unsigned long long)ipt, items / 1e6);
#define RUN(name, ...) { float ms = timeit([&] { __VA_ARGS__; }); printf("%-44s %8.3f ms  %7.1f G items/s\n", name, ms, items / ms / 1e6); }
    RUN("global probe ld256              ILP1", (k_global<1, 1><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global probe ld256              ILP2", (k_global<1, 2><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global probe ld256              ILP4", (k_global<1, 4><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global RED.add u32              ILP1", (k_global<2, 1><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global RED.add u32              ILP4", (k_global<2, 4><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global ATOM.add u32 (return)    ILP1", (k_global<8, 1><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global ATOM.add u32 (return)    ILP4", (k_global<8, 4><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global probe + RED              ILP1", (k_global<3, 1><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global probe + RED              ILP2", (k_global<3, 2><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global probe + RED              ILP4", (k_global<3, 4><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global probe + RED + 25% CAS    ILP1", (k_global<7, 1><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global probe + RED + 25% CAS    ILP2", (k_global<7, 2><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    RUN("global probe + RED + 25% CAS    ILP4", (k_global<7, 4><<<blocks, 256>>>(keys, counts, bmask, ipt, sink)));
    cudaMemset(keys, 0, nslots * 8);
    RUN("global probe + RED.64 same word ILP1", (k_packed<1><<<blocks, 256>>>(keys, bmask, ipt, sink)));
    RUN("global probe + RED.64 same word ILP4", (k_packed<4><<<blocks, 256>>>(keys, bmask, ipt, sink)));
    // shared memory: 16K slots (128 KB keys + 64 KB counts), 1 CTA of 1024 threads per SM
    {
        const u32 sm_slots = 16384; const size_t smem = sm_slots * 12;
        cudaFuncSetAttribute(k_shared<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_shared<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_shared<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_shared<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_shared<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int sblocks = 148; const u64 sipt = 512; const double sitems = (double)sblocks * 1024 * sipt;
#define RUNS(name, M) { float ms = timeit([&] { k_shared<M><<<sblocks, 1024, smem>>>(sm_slots - 1, sipt, sink); }); printf("%-44s %8.3f ms  %7.1f G items/s\n", name, ms, sitems / ms / 1e6); }
        RUNS("shared LDS.64 probe", 1);
        RUNS("shared ATOMS.ADD u32", 2);
        RUNS("shared LDS.64 + ATOMS.ADD", 3);
        RUNS("shared ATOMS.CAS.64 (25%)", 4);
        RUNS("shared LDS + ATOMS.ADD + 25% CAS.64", 7);
    }
    return 0;
}
