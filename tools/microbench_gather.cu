// microbench_gather.cu — cost of RANDOM global-memory lookups (one slot per lane) into an L2-resident table, by access
// width and by the number of active lanes: the k_probe4 / k_probe5 table-probe primitive.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_gather tools/microbench_gather.cu
// Prints clocks per warp-level load per SM (1024 threads per SM, one resident CTA) and random loads per second chip-wide.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 512

__device__ __forceinline__ uint32_t lcg(uint32_t x) { return x * 1664525u + 1013904223u; }

// WIDTH: bytes per lane per load (4, 8, 16, 32; 64 = two 16-byte loads of one 32-byte bucket, the round-1 k_probe4 form)
// ACTIVE: lanes of a warp that issue the load (predicated off for the rest)
// UNROLL: independent loads in flight per lane
template <int WIDTH, int ACTIVE, int UNROLL>
__global__ void __launch_bounds__(1024) k_gather(const uint32_t* __restrict__ table, uint32_t mask32, uint32_t* out,
                                                 unsigned long long* cyc) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t a = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u, acc = 0;
    const bool on = lane < ACTIVE;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it += UNROLL) {
        uint32_t idx[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            a = lcg(a);
            idx[u] = ((a >> 4) ^ (a << 11)) & mask32;  // 32-byte bucket index
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint32_t* p = table + (size_t)idx[u] * 8;
            if (on) {
                if (WIDTH == 4) {
                    uint32_t v;
                    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
                    acc += v;
                } else if (WIDTH == 8) {
                    uint32_t v0, v1;
                    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "l"(p));
                    acc += v0 ^ v1;
                } else if (WIDTH == 16) {
                    uint32_t v0, v1, v2, v3;
                    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(p));
                    acc += v0 ^ v1 ^ v2 ^ v3;
                } else if (WIDTH == 32) {
                    uint32_t e[8];
                    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]), "=r"(e[4]), "=r"(e[5]), "=r"(e[6]), "=r"(e[7])
                                 : "l"(p));
                    acc += e[0] ^ e[1] ^ e[2] ^ e[3] ^ e[4] ^ e[5] ^ e[6] ^ e[7];
                } else {
                    uint32_t v0, v1, v2, v3, w0, w1, w2, w3;
                    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(p));
                    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "l"(p + 4));
                    acc += v0 ^ v1 ^ v2 ^ v3 ^ w0 ^ w1 ^ w2 ^ w3;
                }
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) out[0] = acc;
}

template <int WIDTH, int ACTIVE, int UNROLL>
void run(const char* name, const uint32_t* table, uint32_t n_buckets) {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* out;
    unsigned long long* cyc;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, 8 * sms);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) k_gather<WIDTH, ACTIVE, UNROLL><<<sms, 1024>>>(table, n_buckets - 1, out, cyc);
    cudaEventRecord(e0);
    k_gather<WIDTH, ACTIVE, UNROLL><<<sms, 1024>>>(table, n_buckets - 1, out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[256];
    cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; i++) mean += (double)h[i];
    mean /= sms;
    const double loads = (double)sms * 32 * ACTIVE * ITERS;  // lane-level lookups
    printf("%-40s table %4u MB  active %2d  unroll %d: %7.1f clk per warp-load per SM, %6.2f G lookups/s  (%s)\n", name,
           (unsigned)((size_t)n_buckets * 32 >> 20), ACTIVE, UNROLL, mean / (32.0 * ITERS), loads / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (uint32_t mb : {8u, 16u, 32u, 64u}) {
        const uint32_t n_buckets = (mb << 20) / 32;
        uint32_t* table;
        cudaMalloc(&table, (size_t)n_buckets * 32);
        cudaMemset(table, 1, (size_t)n_buckets * 32);
        run<4, 32, 4>("LDG.32", table, n_buckets);
        run<8, 32, 4>("LDG.64", table, n_buckets);
        run<16, 32, 4>("LDG.128", table, n_buckets);
        run<32, 32, 4>("LDG.256 (v8)", table, n_buckets);
        run<64, 32, 4>("2 x LDG.128 (one 32-byte bucket)", table, n_buckets);
        run<16, 32, 1>("LDG.128", table, n_buckets);
        run<16, 32, 2>("LDG.128", table, n_buckets);
        run<16, 32, 8>("LDG.128", table, n_buckets);
        run<16, 16, 4>("LDG.128", table, n_buckets);
        run<16, 8, 4>("LDG.128", table, n_buckets);
        run<16, 4, 4>("LDG.128", table, n_buckets);
        run<32, 8, 4>("LDG.256 (v8)", table, n_buckets);
        cudaFree(table);
    }
    return 0;
}
