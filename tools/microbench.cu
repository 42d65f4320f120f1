// microbench.cu — per-SM issue rates of the primitives the matcher kernels lean on (B200, sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Prints thread-ops per clock per SM, measured with clock64() inside one resident wave.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__device__ __forceinline__ uint32_t lcg(uint32_t x) { return x * 1664525u + 1013904223u; }

template <int MODE>
__global__ void __launch_bounds__(1024) k_rate(uint32_t* out, unsigned long long* cyc, uint32_t* gbins, uint32_t nbins) {
    extern __shared__ uint32_t s_bins[];
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) s_bins[b] = 0;
    __syncthreads();
    uint32_t a = threadIdx.x * 2654435761u + blockIdx.x, b = a ^ 0x9E3779B9u, c = a + 77u, d = a * 3u;
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {  // LOP3: 4 independent chains
            a = (a & b) ^ c; b = (b | c) ^ d; c = (c & d) ^ a; d = (d | a) ^ b;
        } else if (MODE == 1) {  // POPC: 4 independent
            acc0 += __popc(a); acc1 += __popc(b); acc2 += __popc(c); acc3 += __popc(d);
            a += acc0; b += acc1; c += acc2; d += acc3;  // 4 IADD ride along (counted separately below)
        } else if (MODE == 2) {  // min/max pairs
            acc0 = min(acc0 ^ a, b); acc1 = max(acc1 ^ b, c); acc2 = min(acc2 ^ c, d); acc3 = max(acc3 ^ d, a);
        } else if (MODE == 3) {  // shared-memory atomicAdd, pseudo-random bins (no return value)
            a = lcg(a);
            atomicAdd(&s_bins[(a >> 8) % nbins], 1u);
        } else if (MODE == 4) {  // shared-memory atomicAdd, conflict-free (bin = lane)
            atomicAdd(&s_bins[threadIdx.x & 31u], 1u);
        } else if (MODE == 5) {  // global red, pseudo-random bins
            a = lcg(a);
            atomicAdd(&gbins[(a >> 8) % nbins], 1u);
        } else if (MODE == 6) {  // ballot + popc (warp-aggregated counting)
            a = lcg(a);
            acc0 += __popc(__ballot_sync(0xFFFFFFFFu, a & 0x100u));
        } else if (MODE == 7) {  // match_any
            a = lcg(a);
            acc0 += __match_any_sync(0xFFFFFFFFu, (a >> 8) % nbins);
        } else if (MODE == 8) {  // shared-memory plain RMW on random bins (what a tag-then-add scheme would cost)
            a = lcg(a);
            uint32_t i = (a >> 8) % nbins;
            s_bins[i] = s_bins[i] + 1u;
        } else if (MODE == 9) {  // LDS.128 broadcast (panel plane fetch in k_brute)
            const uint4 v = reinterpret_cast<const uint4*>(s_bins)[it & 63];
            acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    uint32_t sum = a ^ b ^ c ^ d ^ acc0 ^ acc1 ^ acc2 ^ acc3;
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sum ^= s_bins[i];
    if (sum == 0x12345678u) out[0] = sum;
}

template <int MODE>
void run(const char* name, double ops_per_iter, int threads, int blocks_per_sm, uint32_t nbins) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * blocks_per_sm;
    uint32_t *out, *gb;
    unsigned long long* cyc;
    cudaMalloc(&out, 4);
    cudaMalloc(&gb, 4 * 65536);
    cudaMemset(gb, 0, 4 * 65536);
    cudaMalloc(&cyc, 8 * grid);
    for (int rep = 0; rep < 2; rep++) k_rate<MODE><<<grid, threads, 4 * 8192>>>(out, cyc, gb, nbins);
    cudaDeviceSynchronize();
    unsigned long long* h = new unsigned long long[grid];
    cudaMemcpy(h, cyc, 8 * grid, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; i++) mean += (double)h[i];
    mean /= grid;
    const double per_sm = (double)threads * blocks_per_sm * ITERS * ops_per_iter / mean;
    printf("%-44s threads/SM=%4d  %8.2f thread-ops/clk/SM  (%.1f clk per warp-instr per SM)  err=%s\n", name,
           threads * blocks_per_sm, per_sm, 32.0 / per_sm, cudaGetErrorString(cudaGetLastError()));
    delete[] h;
    cudaFree(out);
    cudaFree(gb);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device: %s, %d SMs, clockRate %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<0>("LOP3 (4 chains)", 4, 1024, 2, 385);
    run<1>("POPC (+1 IADD each)", 4, 1024, 2, 385);
    run<2>("IMNMX (+1 LOP each)", 4, 1024, 2, 385);
    run<3>("ATOMS.ADD random bins (385)", 1, 1024, 2, 385);
    run<3>("ATOMS.ADD random bins (6145)", 1, 1024, 2, 6145);
    run<4>("ATOMS.ADD conflict-free (bin = lane)", 1, 1024, 2, 385);
    run<5>("RED.global random bins (385)", 1, 1024, 2, 385);
    run<5>("RED.global random bins (6145)", 1, 1024, 2, 6145);
    run<6>("VOTE.BALLOT + POPC", 1, 1024, 2, 385);
    run<7>("MATCH.ANY", 1, 1024, 2, 385);
    run<8>("LDS+STS random bins (non-atomic RMW)", 1, 1024, 2, 385);
    run<9>("LDS.128 broadcast", 1, 1024, 2, 385);
    return 0;
}
