// microbench_tex.cu — can random 16-byte table lookups go through the texture pipe instead of the LSU pipe?
// Measures lookups/clk/SM for: LDS.128 (bank-tiled replicas), tex1Dfetch<uint4> from a 16 KB linear texture,
// LDG.128 (L1-resident), and LDS + TEX interleaved (do the two pipes overlap?).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
__device__ __forceinline__ uint32_t lcg(uint32_t x) { return x * 1664525u + 1013904223u; }

template <int MODE>
__global__ void __launch_bounds__(1024) k(cudaTextureObject_t tex, const uint4* __restrict__ gtab, uint32_t slots,
                                          uint32_t* out, unsigned long long* cyc) {
    extern __shared__ uint4 s_tab[];  // slots * 8 replicas
    for (uint32_t t = threadIdx.x; t < slots * 8; t += blockDim.x) s_tab[t] = gtab[t / 8];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t a = threadIdx.x * 2654435761u + blockIdx.x * 97u, acc = 0;
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        a = lcg(a);
        const uint32_t s = (a >> 10) & (slots - 1);
        if (MODE == 0 || MODE == 3) {
            const uint4 v = s_tab[s * 8 + (lane & 7)];
            acc += v.x ^ v.z;
        }
        if (MODE == 1 || MODE == 3) {
            const uint32_t s2 = (a >> 3) & (slots - 1);
            const uint4 v = tex1Dfetch<uint4>(tex, (int)s2);
            acc += v.y ^ v.w;
        }
        if (MODE == 2) {
            const uint4 v = __ldg(gtab + s);
            acc += v.x ^ v.w;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
void run(const char* name, cudaTextureObject_t tex, const uint4* gtab, uint32_t slots, double lookups_per_iter) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    unsigned long long* cyc;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, 8 * sms);
    const size_t smem = (size_t)slots * 8 * 16;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) k<MODE><<<sms, 1024, smem>>>(tex, gtab, slots, out, cyc);
    cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; i++) mean += (double)h[i];
    mean /= sms;
    printf("%-46s %7.2f lookups/clk/SM  (%.1f clk per warp-lookup)  err=%s\n", name, 1024.0 * ITERS * lookups_per_iter / mean,
           32.0 * mean / (1024.0 * ITERS * lookups_per_iter), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    const uint32_t slots = 1024;
    uint4* gtab;
    cudaMalloc(&gtab, slots * 16);
    cudaMemset(gtab, 1, slots * 16);
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = gtab;
    rd.res.linear.desc = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindUnsigned);
    rd.res.linear.sizeInBytes = slots * 16;
    cudaTextureDesc td{};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    run<0>("LDS.128 random slot, bank-tiled replicas", tex, gtab, slots, 1);
    run<1>("tex1Dfetch<uint4> random texel (16 KB table)", tex, gtab, slots, 1);
    run<2>("LDG.128 random slot (16 KB table, L1)", tex, gtab, slots, 1);
    run<3>("LDS.128 + tex1Dfetch interleaved (2 lookups/iter)", tex, gtab, slots, 2);
    return 0;
}
