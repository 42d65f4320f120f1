// ingest_kernels.cu — the FASTQ record scanner on the device (SURVEY.md 8f "next" #1 / #2; ReadSetIterator::next,
// src/bin/commands/demux.rs:288-342, seq_io record parsing :289-294).
//
// fqtk_b200_fastq_scan (group.cu) walks a chunk line by line on one host thread (~19 M records/s); the chunk crosses PCIe
// anyway (the B segments are gathered out of it on the GPU), so the same table — where every record's sequence line
// starts and how long it is — can be built there at memory speed: a 4-line FASTQ record ends at every fourth '\n', so
// record r's lines are delimited by newlines 4r-1 .. 4r+3.
//   k_nl_count    newlines per 16 KB tile
//   k_nl_scan     exclusive scan of the tile counts (one CTA)
//   k_nl_fill     position of newline k -> nl[k]  (tile prefix + block scan + the thread's own running count)
//   k_fq_records  per record: sequence offset / length (a '\r' before the '\n' is not part of the line), header offset,
//                 and the same three checks as the host scanner ('@', '+', quality length), first offender by atomicMin
//   k_fq_vet      per read: the reference's "too few bases" rule (demux.rs:298-315) and BarcodeMatcher::assign's
//                 too-long rule (barcode_matching.rs:95-106,170-172) for the given B segments, first offender by atomicMin
#include <algorithm>
#include <cstdint>

#include "common.cuh"
#include "kernels.h"

namespace fq {

constexpr int NL_THREADS = 256;
constexpr uint32_t NL_BYTES_PER_THREAD = 64;
constexpr uint32_t NL_TILE = NL_THREADS * NL_BYTES_PER_THREAD;  // 16 KB

FQ_D uint32_t nl_mask16(uint4 v) {  // bit i set iff byte i of the 16 is '\n'
    uint32_t m = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t x = w[k] ^ 0x0A0A0A0Au;                            // zero byte <=> newline
        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);  // 0x80 in every zero byte (exact)
        m |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
    }
    return m;
}

// newline mask of the thread's 64 bytes (4 x 16), bytes past the end do not count
FQ_D void nl_masks(const uint8_t* __restrict__ chunk, uint64_t bytes, uint64_t base, uint32_t (&m)[4]) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint64_t p = base + 16u * q;
        m[q] = 0;
        if (p + 16u <= bytes && ((reinterpret_cast<uintptr_t>(chunk + p) & 15u) == 0)) {
            m[q] = nl_mask16(__ldg(reinterpret_cast<const uint4*>(chunk + p)));
        } else {
            for (uint32_t b = 0; b < 16u && p + b < bytes; b++) m[q] |= (__ldg(chunk + p + b) == '\n') ? (1u << b) : 0u;
        }
    }
}

__global__ void __launch_bounds__(NL_THREADS) k_nl_count(const uint8_t* __restrict__ chunk, uint64_t bytes,
                                                         uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t s_w[NL_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * NL_TILE + (uint64_t)threadIdx.x * NL_BYTES_PER_THREAD;
    uint32_t m[4];
    nl_masks(chunk, bytes, base, m);
    uint32_t c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31u) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < NL_THREADS / 32; k++) t += s_w[k];
        tile_counts[blockIdx.x] = t;
    }
}

// exclusive scan in place (u64 prefixes), total -> prefix[n_tiles]
__global__ void __launch_bounds__(1024) k_nl_scan(const uint32_t* __restrict__ counts, uint32_t n_tiles,
                                                   unsigned long long* __restrict__ prefix) {
    __shared__ unsigned long long part[1024];
    const uint32_t per = (n_tiles + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, n_tiles), hi = min(lo + per, n_tiles);
    unsigned long long sum = 0;
    for (uint32_t k = lo; k < hi; k++) sum += counts[k];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) {
        const unsigned long long v = threadIdx.x >= off ? part[threadIdx.x - off] : 0ull;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = part[threadIdx.x] - sum;
    for (uint32_t k = lo; k < hi; k++) {
        prefix[k] = run;
        run += counts[k];
    }
    if (threadIdx.x == 1023u) prefix[n_tiles] = part[1023];
}

__global__ void __launch_bounds__(NL_THREADS) k_nl_fill(const uint8_t* __restrict__ chunk, uint64_t bytes,
                                                        const unsigned long long* __restrict__ prefix, uint64_t max_nl,
                                                        unsigned long long* __restrict__ nl) {
    __shared__ uint32_t s_w[NL_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * NL_TILE + (uint64_t)threadIdx.x * NL_BYTES_PER_THREAD;
    uint32_t m[4];
    nl_masks(chunk, bytes, base, m);
    const uint32_t c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31u) s_w[w] = incl;
    __syncthreads();
    unsigned long long k = prefix[blockIdx.x] + (incl - c);
    for (uint32_t q = 0; q < w; q++) k += s_w[q];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t mm = m[q];
        while (mm) {
            const uint32_t b = (uint32_t)__ffs(mm) - 1u;
            mm &= mm - 1u;
            if (k < max_nl) nl[k] = base + 16u * q + b;
            k++;
        }
    }
}

// error word: record << 2 | code (0 header, 1 separator, 2 lengths); atomicMin keeps the first record, and within it the
// first check of the host scanner's order
__global__ void __launch_bounds__(256) k_fq_records(const uint8_t* __restrict__ chunk, const unsigned long long* __restrict__ nl,
                                                    uint64_t n_records, unsigned long long* __restrict__ head_offsets,
                                                    unsigned long long* __restrict__ seq_offsets, uint32_t* __restrict__ seq_lengths,
                                                    unsigned long long* __restrict__ err) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_records; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s0 = r ? nl[4 * r - 1] + 1u : 0u;
        const uint64_t n0 = nl[4 * r], n1 = nl[4 * r + 1], n2 = nl[4 * r + 2], n3 = nl[4 * r + 3];
        auto line_len = [&](uint64_t start, uint64_t end) -> uint64_t {
            uint64_t len = end - start;
            if (len && chunk[end - 1] == '\r') len--;
            return len;
        };
        const uint64_t l0 = line_len(s0, n0), l1 = line_len(n0 + 1u, n1), l2 = line_len(n1 + 1u, n2), l3 = line_len(n2 + 1u, n3);
        unsigned long long e = ~0ull;
        if (l0 == 0 || chunk[s0] != '@') e = r << 2 | 0u;
        else if (l2 == 0 || chunk[n1 + 1u] != '+') e = r << 2 | 1u;
        else if (l1 != l3 || l1 > 0xFFFFFFFFull) e = r << 2 | 2u;
        if (e != ~0ull) atomicMin(err, e);
        if (head_offsets) head_offsets[r] = s0;
        seq_offsets[r] = n0 + 1u;
        seq_lengths[r] = (uint32_t)l1;
    }
}

// err[0]: first read with too few bases for a segment; err[1]: first read whose barcode comes out longer than L and is not
// made None by the no-call pre-filter first
__global__ void __launch_bounds__(256) k_fq_vet(const OffsetSource os, uint64_t n, uint32_t L, uint32_t max_nocalls,
                                                unsigned long long* __restrict__ err) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t total = 0;
        bool few = false;
        for (uint32_t k = 0; k < os.n_segments; k++) {
            const uint32_t s = os.source_of[k];
            const uint64_t len = os.seq_lengths[s][i];
            const bool rest = os.length[k] == OffsetSource::REST;
            const uint64_t need = (uint64_t)os.offset[k] + (rest ? 1u : os.length[k]);
            if (len < need) few = true;
            else total += rest ? len - os.offset[k] : os.length[k];
        }
        if (few) {
            atomicMin(&err[0], (unsigned long long)i);
            continue;
        }
        if (total > L) {
            uint32_t nocalls = 0;
            for (uint32_t k = 0; k < os.n_segments; k++) {
                const uint32_t s = os.source_of[k];
                const uint64_t len = os.length[k] == OffsetSource::REST ? os.seq_lengths[s][i] - os.offset[k] : os.length[k];
                const uint8_t* p = os.base[s] + os.seq_offsets[s][i] + os.offset[k];
                for (uint64_t b = 0; b < len; b++) nocalls += byte_is_nocall(p[b]) ? 1u : 0u;
            }
            if (nocalls <= max_nocalls) atomicMin(&err[1], (unsigned long long)i);
        }
    }
}

// first read whose sequence line is shorter than its read structure needs (ReadSetIterator::next, demux.rs:298-315)
__global__ void __launch_bounds__(256) k_fq_min_len(const uint32_t* __restrict__ seq_lengths, uint64_t n, uint32_t min_len,
                                                    unsigned long long* __restrict__ err) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (seq_lengths[i] < min_len) atomicMin(err, (unsigned long long)i);
}

// ---- launchers -------------------------------------------------------------------------------------------------------
uint32_t fastq_scan_tiles(uint64_t bytes) { return (uint32_t)((bytes + NL_TILE - 1) / NL_TILE); }

// tile_counts: u32[tiles]; prefix: u64[tiles + 1] (prefix[tiles] = number of newlines in the chunk)
cudaError_t launch_nl_count(const uint8_t* d_chunk, uint64_t bytes, uint32_t* d_tile_counts, unsigned long long* d_prefix,
                            cudaStream_t stream) {
    const uint32_t tiles = fastq_scan_tiles(bytes);
    if (tiles) {
        k_nl_count<<<tiles, NL_THREADS, 0, stream>>>(d_chunk, bytes, d_tile_counts);
        count_launch();
    }
    k_nl_scan<<<1, 1024, 0, stream>>>(d_tile_counts, tiles, d_prefix);
    count_launch();
    return cudaGetLastError();
}

// nl: u64[max_nl]; then the per-record tables for n_records <= max_nl / 4 records; d_err: u64, preset to ~0
cudaError_t launch_fq_records(const uint8_t* d_chunk, uint64_t bytes, const unsigned long long* d_prefix, uint64_t max_nl,
                              unsigned long long* d_nl, uint64_t n_records, unsigned long long* d_head_offsets,
                              unsigned long long* d_seq_offsets, uint32_t* d_seq_lengths, unsigned long long* d_err,
                              const LaunchGeometry& g, cudaStream_t stream) {
    const uint32_t tiles = fastq_scan_tiles(bytes);
    if (tiles && max_nl) {
        k_nl_fill<<<tiles, NL_THREADS, 0, stream>>>(d_chunk, bytes, d_prefix, max_nl, d_nl);
        count_launch();
    }
    if (n_records) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>((n_records + 255) / 256, (uint64_t)g.sm_count * 16);
        k_fq_records<<<grid, 256, 0, stream>>>(d_chunk, d_nl, n_records, d_head_offsets, d_seq_offsets, d_seq_lengths, d_err);
        count_launch();
    }
    return cudaGetLastError();
}

cudaError_t launch_fq_vet(const OffsetSource& os, uint64_t n, uint32_t L, uint32_t max_nocalls, unsigned long long* d_err2,
                          const LaunchGeometry& g, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)g.sm_count * 16);
    k_fq_vet<<<grid, 256, 0, stream>>>(os, n, L, max_nocalls, d_err2);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_min_len(const uint32_t* d_seq_lengths, uint64_t n, uint32_t min_len, unsigned long long* d_err,
                           const LaunchGeometry& g, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)g.sm_count * 16);
    k_fq_min_len<<<grid, 256, 0, stream>>>(d_seq_lengths, n, min_len, d_err);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fq
