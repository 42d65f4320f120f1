// route_kernels.cu — per-sample routing of a matched batch (SURVEY.md 8f "next" #3).
//
// The reference routes one read at a time: `sample_writers[best_match].write(&read_set)` or the unmatched writer
// (src/bin/commands/demux.rs:970-975), so every sample's output keeps INPUT ORDER (the tests index fq_reads[0],
// demux.rs:1505-1523).  For a batch that is a stable partition of the read indices by assignment: S + 1 groups
// (samples 0..S-1 in sheet order, then unmatched), each in input order, so the host can hand every writer one
// contiguous, ordered run per batch instead of S + 1 scattered per-read calls.
//
// Stable counting sort, three kernels, no sorting library:
//   k_route_hist     every warp owns one contiguous range of reads and histograms it into a warp-private
//                    shared-memory table; the table goes to global memory bucket-major: hist[bucket][warp]
//   k_route_scan_*   exclusive scan of hist in (bucket, warp) order = the first output slot of every (bucket, warp);
//                    the per-bucket totals' scan is the `offsets` table handed back to the caller
//   k_route_scatter  every warp walks its range again, 32 reads per step in input order; a read's rank among the
//                    step's reads of the same bucket comes from ballots over the bucket's bits (no MATCH, no atomics,
//                    deterministic), the warp's running per-bucket cursor lives in shared memory
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fq {

constexpr int ROUTE_THREADS = 256;        // 8 warps per CTA
constexpr int ROUTE_SCAN_THREADS = 1024;
constexpr int ROUTE_UNROLL = 4;           // 32-read steps whose loads are issued together

// result words are read once per pass: no L1 allocation
FQ_D uint32_t ld_stream_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
FQ_D uint4 ld_stream_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

FQ_D uint32_t bucket_of_result(uint32_t r, uint32_t S) {
    const uint32_t b = r >> 16;
    return (r == NONE || b >= S) ? S : b;
}

// reads [lo, hi) of warp `gw` of `n_warps`: contiguous, multiples of 32 except the very end
FQ_D void warp_range(uint64_t n, uint32_t gw, uint32_t n_warps, uint64_t& lo, uint64_t& hi) {
    const uint64_t steps = (n + 31) / 32;
    const uint64_t s_lo = steps * gw / n_warps, s_hi = steps * (gw + 1) / n_warps;
    lo = s_lo * 32;
    hi = s_hi * 32 < n ? s_hi * 32 : n;
    if (lo > n) lo = n;
}

__global__ void __launch_bounds__(ROUTE_THREADS) k_route_hist(const uint32_t* __restrict__ results, uint64_t n, uint32_t S,
                                                              uint32_t n_warps, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_tab[];  // [warps_per_cta][S + 1]
    const uint32_t B = S + 1u, lane = threadIdx.x & 31u, wic = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint32_t* tab = s_tab + (size_t)wic * B;
    for (uint32_t b = lane; b < B; b += 32u) tab[b] = 0u;
    __syncwarp();
    const uint32_t gw = blockIdx.x * wpc + wic;
    if (gw < n_warps) {
        uint64_t lo, hi;
        warp_range(n, gw, n_warps, lo, hi);
        for (uint64_t base = lo; base < hi; base += 32u * ROUTE_UNROLL) {  // ROUTE_UNROLL independent loads in flight
            uint32_t r[ROUTE_UNROLL];
#pragma unroll
            for (int u = 0; u < ROUTE_UNROLL; u++) {
                const uint64_t i = base + 32u * u + lane;
                r[u] = i < hi ? __ldg(results + i) : 0u;
            }
#pragma unroll
            for (int u = 0; u < ROUTE_UNROLL; u++)
                if (base + 32u * u + lane < hi) atomicAdd(&tab[bucket_of_result(r[u], S)], 1u);
        }
        __syncwarp();
        for (uint32_t b = lane; b < B; b += 32u) hist[(size_t)b * n_warps + gw] = tab[b];
    }
}

// one CTA per bucket: exclusive scan of its row hist[b][0..n_warps) in place, row total to totals[b]
__global__ void __launch_bounds__(ROUTE_SCAN_THREADS) k_route_scan_rows(uint32_t* __restrict__ hist, uint32_t n_warps,
                                                                        unsigned long long* __restrict__ totals) {
    __shared__ uint32_t s_part[ROUTE_SCAN_THREADS];
    uint32_t* row = hist + (size_t)blockIdx.x * n_warps;
    const uint32_t per = (n_warps + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, n_warps);
    uint32_t sum = 0;
    for (uint32_t k = lo; k < hi; k++) sum += row[k];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < blockDim.x; off <<= 1) {  // Hillis-Steele inclusive scan of the per-thread sums
        const uint32_t v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;  // exclusive prefix of this thread's chunk
    for (uint32_t k = lo; k < hi; k++) {
        const uint32_t c = row[k];
        row[k] = run;
        run += c;
    }
    if (threadIdx.x == blockDim.x - 1) totals[blockIdx.x] = s_part[threadIdx.x];
}

// single CTA: offsets[0..B] = exclusive scan of totals[0..B), offsets[B] = n
__global__ void __launch_bounds__(ROUTE_SCAN_THREADS) k_route_scan_totals(const unsigned long long* __restrict__ totals,
                                                                          uint32_t B,
                                                                          unsigned long long* __restrict__ offsets) {
    __shared__ unsigned long long s_part[ROUTE_SCAN_THREADS];
    const uint32_t per = (B + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, B);
    unsigned long long sum = 0;
    for (uint32_t k = lo; k < hi; k++) sum += totals[k];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < blockDim.x; off <<= 1) {
        const unsigned long long v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0ull;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = s_part[threadIdx.x] - sum;
    for (uint32_t k = lo; k < hi; k++) {
        offsets[k] = run;
        run += totals[k];
    }
    if (threadIdx.x == blockDim.x - 1) offsets[B] = s_part[threadIdx.x];
}

__global__ void __launch_bounds__(ROUTE_THREADS) k_route_scatter(const uint32_t* __restrict__ results, uint64_t n, uint32_t S,
                                                                 uint32_t n_warps, const uint32_t* __restrict__ hist,
                                                                 const unsigned long long* __restrict__ offsets,
                                                                 uint32_t bucket_bits, uint32_t* __restrict__ order) {
    extern __shared__ uint32_t s_tab[];  // [warps_per_cta][S + 1] running output cursors
    const uint32_t B = S + 1u, lane = threadIdx.x & 31u, wic = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint32_t* cur = s_tab + (size_t)wic * B;
    const uint32_t gw = blockIdx.x * wpc + wic;
    if (gw >= n_warps) return;
    for (uint32_t b = lane; b < B; b += 32u) cur[b] = (uint32_t)offsets[b] + hist[(size_t)b * n_warps + gw];
    __syncwarp();
    uint64_t lo, hi;
    warp_range(n, gw, n_warps, lo, hi);
    const uint32_t lane_lt = (1u << lane) - 1u;
    for (uint64_t base0 = lo; base0 < hi; base0 += 32u * ROUTE_UNROLL) {
        uint32_t r[ROUTE_UNROLL];  // ROUTE_UNROLL steps' result words in flight before the first is consumed
#pragma unroll
        for (int u = 0; u < ROUTE_UNROLL; u++) {
            const uint64_t i = base0 + 32u * u + lane;
            r[u] = i < hi ? __ldg(results + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < ROUTE_UNROLL; u++) {
            const uint64_t i = base0 + 32u * u + lane;
            const bool valid = i < hi;
            const uint32_t b = valid ? bucket_of_result(r[u], S) : 0xFFFFFFFFu;
            // lanes of this step in the same bucket: AND over the bucket's bits of (ballot(bit) XNOR my bit)
            uint32_t same = __ballot_sync(0xFFFFFFFFu, valid);
            for (uint32_t k = 0; k < bucket_bits; k++) {
                const uint32_t bit = (b >> k) & 1u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
                same &= bit ? bal : ~bal;
            }
            const uint32_t start = valid ? cur[b] : 0u;  // every lane reads its bucket's cursor ...
            __syncwarp();
            if (valid) {
                const uint32_t rank = __popc(same & lane_lt);
                order[start + rank] = (uint32_t)i;
                if (rank == 0u) cur[b] = start + __popc(same);  // ... before the group's first lane advances it
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Tile version (S + 1 <= RT_MAX_BUCKETS): a CTA owns a contiguous range and walks it in tiles of RT_TILE reads.
// Each tile is counting-sorted inside shared memory (stable: warp w owns reads [512 w, 512 (w+1)) of the tile, ranks
// inside a 32-read step come from ballots), then every bucket's run of the tile leaves as one contiguous burst at the
// CTA's running cursor for that bucket — DRAM sees full sectors instead of scattered 4-byte stores.
// ------------------------------------------------------------------------------------------------------
constexpr int RT_THREADS = 512;
constexpr int RT_WARPS = RT_THREADS / 32;
constexpr int RT_TILE = 8192;
constexpr int RT_STEPS = RT_TILE / RT_THREADS;  // 32-read steps per warp and tile
constexpr uint32_t RT_MAX_BUCKETS = 2048;

FQ_D void cta_range(uint64_t n, uint32_t c, uint32_t n_ctas, uint64_t& lo, uint64_t& hi) {
    const uint64_t tiles = (n + RT_TILE - 1) / RT_TILE;
    lo = tiles * c / n_ctas * RT_TILE;
    hi = tiles * (c + 1) / n_ctas * RT_TILE;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
}

__global__ void __launch_bounds__(RT_THREADS) k_route_hist_cta(const uint32_t* __restrict__ results, uint64_t n, uint32_t S,
                                                               uint32_t n_ctas, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_tab[];  // [S + 1]
    const uint32_t B = S + 1u;
    for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) s_tab[b] = 0u;
    __syncthreads();
    uint64_t lo, hi;
    cta_range(n, blockIdx.x, n_ctas, lo, hi);
    for (uint64_t base = lo; base < hi; base += (uint64_t)RT_THREADS * ROUTE_UNROLL) {
        uint32_t r[ROUTE_UNROLL];
#pragma unroll
        for (int u = 0; u < ROUTE_UNROLL; u++) {
            const uint64_t i = base + (uint64_t)RT_THREADS * u + threadIdx.x;
            r[u] = i < hi ? __ldg(results + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < ROUTE_UNROLL; u++)
            if (base + (uint64_t)RT_THREADS * u + threadIdx.x < hi) atomicAdd(&s_tab[bucket_of_result(r[u], S)], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) hist[(size_t)b * n_ctas + blockIdx.x] = s_tab[b];
}

__global__ void __launch_bounds__(RT_THREADS) k_route_scatter_tile(const uint32_t* __restrict__ results, uint64_t n, uint32_t S,
                                                                   uint32_t n_ctas, const uint32_t* __restrict__ hist,
                                                                   const unsigned long long* __restrict__ offsets,
                                                                   uint32_t bucket_bits, uint32_t* __restrict__ order) {
    extern __shared__ uint32_t s_mem[];
    const uint32_t B = S + 1u;
    uint32_t* cursor = s_mem;                       // [B]  next output slot of every bucket for this CTA
    uint32_t* ttotal = cursor + B;                  // [B]  reads of the tile per bucket
    uint32_t* tbase = ttotal + B;                   // [B]  first sorted-tile position of every bucket
    uint32_t* whist = tbase + B;                    // [RT_WARPS][B] per-warp counts, then per-warp running cursors
    uint32_t* part = whist + (size_t)RT_WARPS * B;  // [RT_THREADS] scan partials
    uint16_t* tbucket = reinterpret_cast<uint16_t*>(part + RT_THREADS);  // [RT_TILE] bucket of every read of the tile
    uint16_t* sorted = tbucket + RT_TILE;                                // [RT_TILE] tile-local read index by position
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5, lane_lt = (1u << lane) - 1u;
    uint32_t* my_hist = whist + (size_t)w * B;

    for (uint32_t b = threadIdx.x; b < B; b += RT_THREADS)
        cursor[b] = (uint32_t)offsets[b] + hist[(size_t)b * n_ctas + blockIdx.x];
    uint64_t lo, hi;
    cta_range(n, blockIdx.x, n_ctas, lo, hi);
    for (uint64_t tile_lo = lo; tile_lo < hi; tile_lo += RT_TILE) {
        const uint32_t tile_n = (uint32_t)min((uint64_t)RT_TILE, hi - tile_lo);
        for (uint32_t t = threadIdx.x; t < RT_WARPS * B; t += RT_THREADS) whist[t] = 0u;
        __syncthreads();
        // A: buckets of the tile + per-warp counts
        {
            uint32_t r[RT_STEPS];
#pragma unroll
            for (int st = 0; st < RT_STEPS; st++) {
                const uint32_t idx = w * (RT_STEPS * 32u) + st * 32u + lane;
                r[st] = idx < tile_n ? __ldg(results + tile_lo + idx) : 0u;
            }
#pragma unroll
            for (int st = 0; st < RT_STEPS; st++) {
                const uint32_t idx = w * (RT_STEPS * 32u) + st * 32u + lane;
                if (idx < tile_n) {
                    const uint32_t b = bucket_of_result(r[st], S);
                    tbucket[idx] = (uint16_t)b;
                    atomicAdd(&my_hist[b], 1u);
                }
            }
        }
        __syncthreads();
        // B: per bucket, exclusive prefix over the warps (in place) and the tile total
        for (uint32_t b = threadIdx.x; b < B; b += RT_THREADS) {
            uint32_t acc = 0;
#pragma unroll
            for (int ww = 0; ww < RT_WARPS; ww++) {
                const uint32_t c = whist[(size_t)ww * B + b];
                whist[(size_t)ww * B + b] = acc;
                acc += c;
            }
            ttotal[b] = acc;
        }
        __syncthreads();
        // C: exclusive scan of the tile totals over the buckets -> first sorted position of every bucket
        {
            const uint32_t per = (B + RT_THREADS - 1) / RT_THREADS;
            const uint32_t b_lo = threadIdx.x * per, b_hi = min(b_lo + per, B);
            uint32_t sum = 0;
            for (uint32_t b = b_lo; b < b_hi; b++) sum += ttotal[b];
            part[threadIdx.x] = sum;
            __syncthreads();
            for (uint32_t off = 1; off < RT_THREADS; off <<= 1) {
                const uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
                __syncthreads();
                part[threadIdx.x] += v;
                __syncthreads();
            }
            uint32_t run = part[threadIdx.x] - sum;
            for (uint32_t b = b_lo; b < b_hi; b++) {
                tbase[b] = run;
                run += ttotal[b];
            }
        }
        __syncthreads();
        // D: stable rank of every read inside the tile -> sorted[]
        for (int st = 0; st < RT_STEPS; st++) {
            const uint32_t idx = w * (RT_STEPS * 32u) + st * 32u + lane;
            const bool valid = idx < tile_n;
            const uint32_t b = valid ? tbucket[idx] : 0xFFFFu;
            uint32_t same = __ballot_sync(0xFFFFFFFFu, valid);
            for (uint32_t k = 0; k < bucket_bits; k++) {
                const uint32_t bit = (b >> k) & 1u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
                same &= bit ? bal : ~bal;
            }
            const uint32_t start = valid ? my_hist[b] : 0u;
            __syncwarp();
            if (valid) {
                const uint32_t rank = __popc(same & lane_lt);
                sorted[tbase[b] + start + rank] = (uint16_t)idx;
                if (rank == 0u) my_hist[b] = start + __popc(same);
            }
            __syncwarp();
        }
        __syncthreads();
        // E: every bucket's run leaves as one contiguous burst at the CTA's cursor
        for (uint32_t pos = threadIdx.x; pos < tile_n; pos += RT_THREADS) {
            const uint32_t idx = sorted[pos];
            const uint32_t b = tbucket[idx];
            order[cursor[b] + (pos - tbase[b])] = (uint32_t)(tile_lo + idx);
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < B; b += RT_THREADS) cursor[b] += ttotal[b];
        // (the loop-top zeroing + barrier orders this against the next tile)
    }
}

// ------------------------------------------------------------------------------------------------------
// Tile version 2 (S + 2 buckets next to a 32 KB tile in shared memory, two CTAs per SM): the rank of a read among the
// reads of its bucket in the same 32-read step comes from ONE shared-memory atomic instead of a ballot per bucket bit:
// every lane ORs its lane bit into the warp-private cell {mask, count} of its bucket, the cell read back is the set
// of lanes of the step in that bucket (rank = popc of the lower lanes) and the warp's running count of the bucket;
// the group's first lane then clears the mask and advances the count.  Counting and ranking are one pass (the rank is
// relative to the warp's 512-read chunk; the warp / bucket prefixes are added when the read is placed), the tile
// leaves through shared memory as before so that every bucket's run is one contiguous burst.  A dummy bucket behind
// the unmatched one takes the lanes past the end of the last tile, so the ranking loop has no predicates.
// ------------------------------------------------------------------------------------------------------
constexpr int R2_THREADS = 512;
constexpr int R2_WARPS = R2_THREADS / 32;
constexpr int R2_STEPS = 16;  // 32-read steps per warp and tile: a tile is 16 reads per thread (4 096 or 8 192 reads)

// results -> per-CTA bucket counts, 128-bit loads (the CTA ranges are tile-aligned, so 16-byte alignment only
// depends on the base pointer, checked by the host)
__global__ void __launch_bounds__(R2_THREADS) k_route_hist_cta4(const uint32_t* __restrict__ results, uint64_t n, uint32_t S,
                                                                uint32_t n_ctas, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_tab[];  // [S + 1]
    const uint32_t B = S + 1u;
    for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) s_tab[b] = 0u;
    __syncthreads();
    uint64_t lo, hi;
    cta_range(n, blockIdx.x, n_ctas, lo, hi);
    const uint64_t hi4 = lo + ((hi - lo) & ~(uint64_t)3);
    const uint4* v = reinterpret_cast<const uint4*>(results + lo);
    const uint64_t n4 = (hi4 - lo) >> 2;
    for (uint64_t base = 0; base < n4; base += (uint64_t)R2_THREADS * ROUTE_UNROLL) {
        uint4 r[ROUTE_UNROLL];
#pragma unroll
        for (int u = 0; u < ROUTE_UNROLL; u++) {
            const uint64_t i = base + (uint64_t)R2_THREADS * u + threadIdx.x;
            r[u] = i < n4 ? ld_stream_u4(v + i) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < ROUTE_UNROLL; u++)
            if (base + (uint64_t)R2_THREADS * u + threadIdx.x < n4) {
                atomicAdd(&s_tab[bucket_of_result(r[u].x, S)], 1u);
                atomicAdd(&s_tab[bucket_of_result(r[u].y, S)], 1u);
                atomicAdd(&s_tab[bucket_of_result(r[u].z, S)], 1u);
                atomicAdd(&s_tab[bucket_of_result(r[u].w, S)], 1u);
            }
    }
    for (uint64_t i = hi4 + threadIdx.x; i < hi; i += R2_THREADS) atomicAdd(&s_tab[bucket_of_result(__ldg(results + i), S)], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) hist[(size_t)b * n_ctas + blockIdx.x] = s_tab[b];
}

struct Tile2Smem {
    uint32_t *cursor, *delta, *part, *mask, *cnt, *sorted;
};

// one tile of k_route_scatter_tile2; FULL = all TILE reads exist here and in the tile that is prefetched
template <int THREADS, bool FULL>
FQ_D void route_tile2(const Tile2Smem& sm, const uint32_t* __restrict__ results, uint32_t* __restrict__ order, uint32_t S,
                      uint64_t tile_lo, uint32_t tile_n, uint32_t next_n, uint32_t (&r)[R2_STEPS]) {
    constexpr int WARPS = THREADS / 32;
    constexpr uint32_t TILE = THREADS * R2_STEPS, IDX_BITS = THREADS == 1024 ? 14 : THREADS == 512 ? 13 : 12;
    const uint32_t B = S + 1u, B1 = B + 1u;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5, lane_bit = 1u << lane, lane_lt = lane_bit - 1u;
    uint32_t* my_mask = sm.mask + w * B1;
    uint32_t* my_cnt = sm.cnt + w * B1;
    const uint32_t per = (B1 + THREADS - 1) / THREADS;  // buckets per thread in the prefix phase
    const uint32_t b_lo = min(threadIdx.x * per, B1), b_hi = min(b_lo + per, B1);
    // 1: count and rank inside the warp's chunk: the lanes of the step in my bucket come back from the mask cell, the
    //    group's first lane takes the warp's running count of the bucket with one returning atomic and hands it round
#pragma unroll
    for (int st = 0; st < R2_STEPS; st++) {
        const uint32_t idx = w * (R2_STEPS * 32u) + st * 32u + lane;
        const uint32_t b = (FULL || idx < tile_n) ? bucket_of_result(r[st], S) : B;
        atomicOr(&my_mask[b], lane_bit);
        __syncwarp();
        const uint32_t m = my_mask[b];
        __syncwarp();
        uint32_t c = 0;
        if ((m & lane_lt) == 0u) {
            c = atomicAdd(&my_cnt[b], (uint32_t)__popc(m));
            my_mask[b] = 0u;
        }
        __syncwarp();
        c = __shfl_sync(0xFFFFFFFFu, c, __ffs(m) - 1);
        r[st] = b << IDX_BITS | (c + __popc(m & lane_lt));
    }
    __syncthreads();
    // 2: per bucket the exclusive prefix over the warps, per tile the exclusive prefix over the buckets
    uint32_t mine = 0;
    for (uint32_t b = b_lo; b < b_hi; b++)
#pragma unroll
        for (int ww = 0; ww < WARPS; ww++) mine += sm.cnt[ww * B1 + b];
    uint32_t incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= (uint32_t)off) incl += v;
    }
    if (lane == 31u) sm.part[w] = incl;
    __syncthreads();
    uint32_t run = incl - mine;
    for (uint32_t k = 0; k < w; k++) run += sm.part[k];
    for (uint32_t b = b_lo; b < b_hi; b++) {
        const uint32_t base = run;
#pragma unroll
        for (int ww = 0; ww < WARPS; ww++) {
            const uint32_t c = sm.cnt[ww * B1 + b];
            sm.cnt[ww * B1 + b] = run;
            run += c;
        }
        const uint32_t cur = sm.cursor[b];
        sm.delta[b] = cur - base;
        sm.cursor[b] = cur + (run - base);
    }
    __syncthreads();
    // 3: place (bucket, tile index) at its sorted position
#pragma unroll
    for (int st = 0; st < R2_STEPS; st++) {
        const uint32_t idx = w * (R2_STEPS * 32u) + st * 32u + lane;
        const uint32_t b = r[st] >> IDX_BITS, rank = r[st] & (TILE - 1u);
        sm.sorted[my_cnt[b] + rank] = b << IDX_BITS | idx;
    }
    // the next tile's words are on their way while this one leaves
#pragma unroll
    for (int st = 0; st < R2_STEPS; st++) {
        const uint32_t idx = w * (R2_STEPS * 32u) + st * 32u + lane;
        r[st] = (FULL || idx < next_n) ? ld_stream_u32(results + tile_lo + TILE + idx) : 0u;
    }
    __syncthreads();
    // 4: every bucket's run leaves as one contiguous burst at the CTA's cursor; counts back to zero
    if (FULL) {
#pragma unroll
        for (int k = 0; k < R2_STEPS; k++) {
            const uint32_t pos = k * THREADS + threadIdx.x;
            const uint32_t v = sm.sorted[pos];
            order[sm.delta[v >> IDX_BITS] + pos] = (uint32_t)tile_lo + (v & (TILE - 1u));
        }
    } else {
        for (uint32_t pos = threadIdx.x; pos < tile_n; pos += THREADS) {
            const uint32_t v = sm.sorted[pos];
            order[sm.delta[v >> IDX_BITS] + pos] = (uint32_t)tile_lo + (v & (TILE - 1u));
        }
    }
    for (uint32_t t = threadIdx.x; t < WARPS * B1; t += THREADS) sm.cnt[t] = 0u;
    __syncthreads();
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 1024 ? 1 : THREADS == 512 ? 2 : 4)
    k_route_scatter_tile2(const uint32_t* __restrict__ results, uint64_t n, uint32_t S, uint32_t n_ctas,
                          const uint32_t* __restrict__ hist, const unsigned long long* __restrict__ offsets,
                          uint32_t* __restrict__ order) {
    extern __shared__ uint32_t s_mem[];
    constexpr int WARPS = THREADS / 32;
    constexpr uint32_t TILE = THREADS * R2_STEPS;
    const uint32_t B = S + 1u, B1 = B + 1u;  // bucket B = the dummy for lanes past the end of the batch
    Tile2Smem sm;
    sm.cursor = s_mem;                       // [B1] next output slot of every bucket for this CTA
    sm.delta = sm.cursor + B1;               // [B1] cursor - first sorted-tile position (this tile)
    sm.part = sm.delta + B1;                 // [WARPS] scan partials
    sm.mask = sm.part + WARPS;            // [WARPS][B1] lanes of the current step per bucket
    sm.cnt = sm.mask + WARPS * B1;        // [WARPS][B1] reads of the warp's chunk per bucket -> first sorted position
    sm.sorted = sm.cnt + WARPS * B1;      // [TILE] bucket << 13 | tile index
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;

    for (uint32_t b = threadIdx.x; b < B1; b += THREADS)
        sm.cursor[b] = b < B ? (uint32_t)offsets[b] + hist[(size_t)b * n_ctas + blockIdx.x] : 0u;
    for (uint32_t t = threadIdx.x; t < 2u * WARPS * B1; t += THREADS) sm.mask[t] = 0u;  // masks and counts
    uint64_t lo, hi;
    cta_range(n, blockIdx.x, n_ctas, lo, hi);
    uint32_t r[R2_STEPS];  // this tile's result words, then bucket << 13 | rank inside the warp's chunk
    if (lo < hi) {
        const uint32_t tile_n = (uint32_t)min((uint64_t)TILE, hi - lo);
#pragma unroll
        for (int st = 0; st < R2_STEPS; st++) {
            const uint32_t idx = w * (R2_STEPS * 32u) + st * 32u + lane;
            r[st] = idx < tile_n ? ld_stream_u32(results + lo + idx) : 0u;
        }
    }
    __syncthreads();
    for (uint64_t tile_lo = lo; tile_lo < hi; tile_lo += TILE) {
        const uint64_t left = hi - tile_lo;
        if (left >= 2u * TILE) {
            route_tile2<THREADS, true>(sm, results, order, S, tile_lo, TILE, TILE, r);
        } else {
            const uint32_t tile_n = (uint32_t)min((uint64_t)TILE, left);
            route_tile2<THREADS, false>(sm, results, order, S, tile_lo, tile_n, (uint32_t)(left - tile_n), r);
        }
    }
}

static int route_tile2_threads() {  // CTA shape of the mask-rank kernel: 256 / 512 / 1 024 threads x 16 reads per tile (A/B: FQTK_B200_ROUTE_T)
    static const int t = [] {
        const char* e = getenv("FQTK_B200_ROUTE_T");
        const int v = e ? atoi(e) : 512;
        return (v == 256 || v == 1024) ? v : 512;
    }();
    return t;
}
static size_t route_tile2_smem(uint32_t S, int threads) {
    const size_t B1 = S + 2u, warps = threads / 32;
    return (2 * B1 + warps) * 4 + warps * B1 * 8 + (size_t)threads * R2_STEPS * 4;  // mask table + counts + sorted tile
}

static size_t route_tile_smem(uint32_t S) {
    const size_t B = S + 1u;
    return (3 * B + (size_t)RT_WARPS * B + RT_THREADS) * 4 + (size_t)RT_TILE * 2 * 2;
}

struct RoutePlan {
    uint32_t n_warps, warps_per_cta, grid, bucket_bits;  // n_warps = histogram columns (warps, or CTAs in tile mode)
    size_t smem;
    bool tile, tile2;
    int threads;
};

static RoutePlan plan_route(uint64_t n, uint32_t S, const LaunchGeometry& g) {
    RoutePlan p{};
    p.bucket_bits = 1;
    while ((1u << p.bucket_bits) < S + 1u) p.bucket_bits++;
    {
        int threads = route_tile2_threads();
        if (threads != 512 && route_tile2_smem(S, threads) + 1024 > (size_t)g.max_smem_optin) threads = 512;
        if (!getenv("FQTK_B200_ROUTE_V1") && !getenv("FQTK_B200_ROUTE_V2") && S + 2u <= (1u << 19) &&
            route_tile2_smem(S, threads) + 1024 <= (size_t)g.max_smem_optin) {
            p.tile = p.tile2 = true;
            p.threads = threads;
            p.smem = route_tile2_smem(S, threads);
            const uint64_t tiles = (n + RT_TILE - 1) / RT_TILE;  // (CTA ranges are cut at multiples of RT_TILE for every shape)
            const uint32_t fit = (uint32_t)(((size_t)g.max_smem_optin + 1024) / (p.smem + 1024));
            const uint32_t per_sm = std::max(1u, std::min(fit, threads == 1024 ? 1u : threads == 512 ? 2u : 4u));
            uint64_t ctas = (uint64_t)g.sm_count * per_sm;
            if (ctas > tiles) ctas = tiles ? tiles : 1;
            p.n_warps = p.grid = (uint32_t)ctas;
            p.warps_per_cta = (uint32_t)threads / 32u;
            return p;
        }
    }
    if (S + 1u <= RT_MAX_BUCKETS && route_tile_smem(S) + 1024 <= (size_t)g.max_smem_optin && !getenv("FQTK_B200_ROUTE_V1")) {
        p.tile = true;
        p.smem = route_tile_smem(S);
        const uint64_t tiles = (n + RT_TILE - 1) / RT_TILE;
        const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(3, ((size_t)g.max_smem_optin) / (p.smem + 1024)));
        uint64_t ctas = (uint64_t)g.sm_count * per_sm;
        if (ctas > tiles) ctas = tiles ? tiles : 1;
        p.n_warps = p.grid = (uint32_t)ctas;
        p.warps_per_cta = RT_WARPS;
        return p;
    }
    const size_t per_warp = (size_t)(S + 1u) * 4u;
    uint32_t wpc = ROUTE_THREADS / 32;
    while (wpc > 1 && per_warp * wpc > (size_t)g.max_smem_optin - 1024) wpc >>= 1;
    p.warps_per_cta = wpc;
    p.smem = per_warp * wpc;
    // enough warps to fill the machine, few enough that (warps x buckets) open output runs stay L2-resident
    // 8 warps per SM: measured best on B200 (r01q sweep) — more warps only multiply the open (warp, bucket) output
    // runs, whose partially written sectors then bounce between L2 and DRAM (5x write amplification at 32 warps/SM)
    uint32_t warps_per_sm = 8;
    if (const char* e = getenv("FQTK_B200_ROUTE_WARPS_PER_SM")) warps_per_sm = (uint32_t)atoi(e);
    if (warps_per_sm < wpc) wpc = warps_per_sm ? warps_per_sm : 1;
    p.warps_per_cta = wpc;
    p.smem = per_warp * wpc;
    uint32_t ctas = (uint32_t)g.sm_count * ((warps_per_sm + wpc - 1) / wpc);
    const uint64_t steps = (n + 31) / 32;
    uint64_t warps = (uint64_t)ctas * wpc;
    if (warps > steps) warps = steps ? steps : 1;
    p.n_warps = (uint32_t)warps;
    p.grid = (uint32_t)((warps + wpc - 1) / wpc);
    p.bucket_bits = 1;
    while ((1u << p.bucket_bits) < S + 1u) p.bucket_bits++;
    return p;
}

size_t route_workspace_bytes(uint64_t n, uint32_t S, const LaunchGeometry& g) {
    const RoutePlan p = plan_route(n, S, g);
    return (size_t)(S + 1u) * p.n_warps * 4u + (size_t)(S + 1u) * 8u + 256;
}

bool route_supported(uint32_t S, const LaunchGeometry& g) { return (size_t)(S + 1u) * 4u + 1024 <= (size_t)g.max_smem_optin; }

// d_offsets: u64[S + 2] on the device.  d_workspace: route_workspace_bytes() bytes.
cudaError_t launch_route(const uint32_t* d_results, uint64_t n, uint32_t S, uint32_t* d_order,
                         unsigned long long* d_offsets, void* d_workspace, const LaunchGeometry& g, cudaStream_t stream) {
    const RoutePlan p = plan_route(n, S, g);
    uint32_t* hist = reinterpret_cast<uint32_t*>(d_workspace);
    unsigned long long* totals =
        reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(d_workspace) + (((size_t)(S + 1u) * p.n_warps * 4u + 7) & ~(size_t)7));
    if (p.tile2) {
        const size_t hsmem = (size_t)(S + 1u) * 4u;
        if ((reinterpret_cast<uintptr_t>(d_results) & 15u) == 0)
            k_route_hist_cta4<<<p.grid, R2_THREADS, hsmem, stream>>>(d_results, n, S, p.n_warps, hist);
        else
            k_route_hist_cta<<<p.grid, RT_THREADS, hsmem, stream>>>(d_results, n, S, p.n_warps, hist);
        count_launch();
        k_route_scan_rows<<<S + 1u, ROUTE_SCAN_THREADS, 0, stream>>>(hist, p.n_warps, totals);
        count_launch();
        k_route_scan_totals<<<1, ROUTE_SCAN_THREADS, 0, stream>>>(totals, S + 1u, d_offsets);
        count_launch();
        if (p.threads == 1024) {
            cudaFuncSetAttribute(k_route_scatter_tile2<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
            k_route_scatter_tile2<1024><<<p.grid, 1024, p.smem, stream>>>(d_results, n, S, p.n_warps, hist, d_offsets, d_order);
        } else if (p.threads == 512) {
            cudaFuncSetAttribute(k_route_scatter_tile2<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
            k_route_scatter_tile2<512><<<p.grid, 512, p.smem, stream>>>(d_results, n, S, p.n_warps, hist, d_offsets, d_order);
        } else {
            cudaFuncSetAttribute(k_route_scatter_tile2<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
            k_route_scatter_tile2<256><<<p.grid, 256, p.smem, stream>>>(d_results, n, S, p.n_warps, hist, d_offsets, d_order);
        }
        count_launch();
        return cudaGetLastError();
    }
    if (p.tile) {
        const size_t hsmem = (size_t)(S + 1u) * 4u;
        cudaFuncSetAttribute(k_route_scatter_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        k_route_hist_cta<<<p.grid, RT_THREADS, hsmem, stream>>>(d_results, n, S, p.n_warps, hist);
        count_launch();
        k_route_scan_rows<<<S + 1u, ROUTE_SCAN_THREADS, 0, stream>>>(hist, p.n_warps, totals);
        count_launch();
        k_route_scan_totals<<<1, ROUTE_SCAN_THREADS, 0, stream>>>(totals, S + 1u, d_offsets);
        count_launch();
        k_route_scatter_tile<<<p.grid, RT_THREADS, p.smem, stream>>>(d_results, n, S, p.n_warps, hist, d_offsets,
                                                                      p.bucket_bits, d_order);
        count_launch();
        return cudaGetLastError();
    }
    const int threads = (int)p.warps_per_cta * 32;
    cudaFuncSetAttribute(k_route_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    cudaFuncSetAttribute(k_route_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    k_route_hist<<<p.grid, threads, p.smem, stream>>>(d_results, n, S, p.n_warps, hist);
    count_launch();
    k_route_scan_rows<<<S + 1u, ROUTE_SCAN_THREADS, 0, stream>>>(hist, p.n_warps, totals);
    count_launch();
    k_route_scan_totals<<<1, ROUTE_SCAN_THREADS, 0, stream>>>(totals, S + 1u, d_offsets);
    count_launch();
    k_route_scatter<<<p.grid, threads, p.smem, stream>>>(d_results, n, S, p.n_warps, hist, d_offsets, p.bucket_bits, d_order);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fq
