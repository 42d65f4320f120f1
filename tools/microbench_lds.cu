// microbench_lds.cu — cost of random shared-memory lookups by access width, with bank-tiled replicas.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ uint32_t lcg(uint32_t x) { return x * 1664525u + 1013904223u; }

// MODE 0: LDS.128 replicas x8 | 1: LDS.64 replicas x16 | 2: LDS.32 replicas x32 | 3: LDS.32 no replicas (random banks)
// MODE 4: 2x LDS.64 (x16) + 1x LDS.32 (x4 replicas)   | 5: 2x LDS.128 (x8)      | 6: LDS.64 no replicas
template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t slots, uint32_t* out, unsigned long long* cyc) {
    extern __shared__ uint32_t s_w[];
    for (uint32_t t = threadIdx.x; t < 40960; t += blockDim.x) s_w[t] = t * 2654435761u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t a = threadIdx.x * 2654435761u + blockIdx.x * 97u, acc = 0;
    const uint4* t128 = reinterpret_cast<const uint4*>(s_w);
    const uint2* t64 = reinterpret_cast<const uint2*>(s_w);
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        a = lcg(a);
        const uint32_t s = (a >> 10) & (slots - 1), s2 = (a >> 20) & (slots - 1);
        if (MODE == 0) { const uint4 v = t128[s * 8 + (lane & 7)]; acc += v.x ^ v.z; }
        if (MODE == 1) { const uint2 v = t64[s * 16 + (lane & 15)]; acc += v.x ^ v.y; }
        if (MODE == 2) { acc += s_w[s * 32 + lane]; }
        if (MODE == 3) { acc += s_w[s]; }
        if (MODE == 4) {
            const uint2 v = t64[s * 16 + (lane & 15)], u = t64[s2 * 16 + (lane & 15)];
            const uint32_t pick = (v.x < u.x) ? s : s2;
            acc += v.y ^ u.y ^ s_w[32768 + (pick & 1023) * 4 + (lane & 3)];
        }
        if (MODE == 5) {
            const uint4 v = t128[s * 8 + (lane & 7)], u = t128[s2 * 8 + (lane & 7)];
            acc += v.x ^ v.z ^ u.y ^ u.w;
        }
        if (MODE == 6) { const uint2 v = t64[s]; acc += v.x ^ v.y; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
void run(const char* name, uint32_t slots) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    unsigned long long* cyc;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, 8 * sms);
    const size_t smem = 40960 * 4;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) k<MODE><<<sms, 1024, smem>>>(slots, out, cyc);
    cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; i++) mean += (double)h[i];
    mean /= sms;
    printf("%-52s %6.1f clk per warp-iteration per SM  err=%s\n", name, 32.0 * mean / (1024.0 * ITERS),
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("LDS.128, 8 bank-tiled replicas", 1024);
    run<1>("LDS.64, 16 bank-tiled replicas", 1024);
    run<2>("LDS.32, 32 bank-tiled replicas", 1024);
    run<3>("LDS.32, random banks (no replicas)", 1024);
    run<6>("LDS.64, random banks (no replicas)", 1024);
    run<4>("2x LDS.64 (x16) + LDS.32 (x4 replicas)", 1024);
    run<5>("2x LDS.128 (x8)", 1024);
    return 0;
}
