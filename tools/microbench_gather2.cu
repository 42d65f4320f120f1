// microbench_gather2.cu — random 32-byte lookups into an L2-resident table WHILE a read / result stream flows through L2
// (the k_probe4 situation): every lane streams 16 bytes in (coalesced LDG.128) and 16 bytes out per 4 lookups... i.e. the
// cfg 5 ratio of 12 B in + 4 B out per lookup.  Reports lookups/s by table size, with and without L2 policy hints, and the
// DRAM traffic is read off ncu if wanted.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_gather2 tools/microbench_gather2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_normal() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }

// HINTS: 0 none, 1 table evict_last + stream evict_first.  INFLIGHT: lookups in flight per lane (1, 2, 4)
template <int HINTS, int INFLIGHT>
__global__ void __launch_bounds__(1024) k_mix(const uint32_t* __restrict__ table, uint32_t n_buckets,
                                              const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t n_groups,
                                              uint32_t* sink) {
    const uint64_t pk = HINTS ? pol_last() : pol_normal(), ps = HINTS ? pol_first() : pol_normal();
    uint32_t acc = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        // 4 reads of 12 bytes = 3 x 16 bytes in, 16 bytes out
        uint4 w[3];
#pragma unroll
        for (int v = 0; v < 3; v++)
            asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=r"(w[v].x), "=r"(w[v].y), "=r"(w[v].z), "=r"(w[v].w) : "l"(in + (size_t)g * 3 + v), "l"(ps));
        uint32_t h[4] = {w[0].x * 0x9E3779B1u + w[0].y, w[0].w * 0x85EBCA77u + w[1].x, w[1].z * 0xC2B2AE3Du + w[1].w,
                         w[2].y * 0x27D4EB2Fu + w[2].z};
#pragma unroll
        for (int q = 0; q < 4; q++) h[q] += (g * 4u + q) * 0x9E3779B9u;  // the filler input repeats: make every lookup's bucket its own
        uint32_t r[4];
#pragma unroll
        for (int b = 0; b < 4; b += INFLIGHT) {
            uint32_t e[INFLIGHT][8];
#pragma unroll
            for (int q = 0; q < INFLIGHT; q++) {
                uint32_t x = h[b + q];
                x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
                const uint32_t bucket = (uint32_t)(((uint64_t)x * n_buckets) >> 32);
                const uint32_t* p = table + (size_t)bucket * 8;
                asm volatile("ld.global.nc.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                             : "=r"(e[q][0]), "=r"(e[q][1]), "=r"(e[q][2]), "=r"(e[q][3]), "=r"(e[q][4]), "=r"(e[q][5]),
                               "=r"(e[q][6]), "=r"(e[q][7]) : "l"(p), "l"(pk));
            }
#pragma unroll
            for (int q = 0; q < INFLIGHT; q++)
                r[b + q] = min(min(min(e[q][0], e[q][1]), min(e[q][2], e[q][3])), min(min(e[q][4], e[q][5]), min(e[q][6], e[q][7])));
        }
        acc += r[0] ^ r[1] ^ r[2] ^ r[3];
        asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(out + g), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "l"(ps) : "memory");
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int HINTS, int INFLIGHT>
void run(uint32_t mb, const uint4* in, uint4* out, uint32_t n_groups, uint32_t* sink) {
    const uint32_t n_buckets = (mb << 20) / 32;
    uint32_t* table;
    cudaMalloc(&table, (size_t)n_buckets * 32);
    cudaMemset(table, 0x5a, (size_t)n_buckets * 32);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_mix<HINTS, INFLIGHT><<<sms, 1024>>>(table, n_buckets, in, out, n_groups, sink);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 3; rep++) k_mix<HINTS, INFLIGHT><<<sms, 1024>>>(table, n_buckets, in, out, n_groups, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 3;
    printf("table %3u MB  hints %d  in flight %d: %7.3f ms per %u M lookups = %6.1f G lookups/s, stream %.0f GB/s  (%s)\n", mb,
           HINTS, INFLIGHT, ms, n_groups * 4 / 1000000, n_groups * 4.0 / (ms * 1e-3) / 1e9, n_groups * 64.0 / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(table);
}

int main() {
    const uint32_t n_groups = 64u << 20;  // 256 Mi lookups, 3 GB in, 1 GB out
    uint4 *in, *out;
    uint32_t* sink;
    cudaMalloc(&in, (size_t)n_groups * 48);
    cudaMalloc(&out, (size_t)n_groups * 16);
    cudaMalloc(&sink, 4);
    cudaMemset(in, 0x33, (size_t)n_groups * 48);
    {   // make the words vary so the buckets are random
        uint32_t* h = new uint32_t[1 << 20];
        for (uint32_t i = 0; i < (1u << 20); i++) h[i] = i * 2654435761u ^ (i << 7);
        for (size_t off = 0; off < (size_t)n_groups * 48; off += 4u << 20) cudaMemcpy((char*)in + off, h, 4u << 20, cudaMemcpyHostToDevice);
        delete[] h;
    }
    // the chunk repeats every 4 MB, so perturb: add the group index into the hash via a pass of the kernel? keep simple:
    for (uint32_t mb : {4u, 8u, 16u, 24u, 32u, 48u, 64u, 96u}) {
        run<0, 2>(mb, in, out, n_groups, sink);
        run<1, 2>(mb, in, out, n_groups, sink);
    }
    run<1, 1>(16, in, out, n_groups, sink);
    run<1, 4>(16, in, out, n_groups, sink);
    run<0, 4>(16, in, out, n_groups, sink);
    run<1, 4>(32, in, out, n_groups, sink);
    return 0;
}
