// match_kernels.cu — the sm_100a kernels of the demux barcode matcher.
//
// What they compute is the closed form of BarcodeMatcher::assign (src/lib/barcode_matching.rs:119-186 of the
// reference; derivation in SURVEY.md Appendix A.2): exact IUPAC-aware distance d_j = #{i : obs_i & !exp_ji != 0}
// (src/lib/bitenc.rs:441-452) to every barcode j, best = min d_j with the FIRST index on ties, next = second
// smallest with multiplicity, None iff best > max_mismatches or next - best < min_mismatch_delta; then the
// caller's count rule (src/bin/commands/demux.rs:970-974).
//
// Two kernel families, bit-identical results:
//   k_brute  — thread per read, whole panel as "forbidden base" bit-planes in shared memory (broadcast LDS.128),
//              4 logic ops + 1 popcount per (read, barcode) pair, running (distance<<16 | index) min / second-min.
//   k_probe  — memo-table kernel: the read's packed words are the key of a pre-computed open-addressing table
//              holding the result of every A/C/G/T/N string within max_mismatches of some barcode (the device
//              analogue of the reference's AHashMap cache, barcode_matching.rs:173-182, pre-filled).  A miss on
//              an in-alphabet read is None; a miss on any other read (IUPAC / junk / lowercase-only symbols) is
//              resolved by the whole warp brute-forcing that one read with shuffle min / second-min reduction.
// Per-sample counts: CTA-private shared-memory histogram (atomics), flushed once per CTA with 64-bit REDs.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"
#include "sliced_adders.cuh"

namespace fq {

static std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launches() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }

constexpr uint32_t HIST_SMEM_BINS = 8192;  // S + 1 <= this: CTA-private histogram lives in shared memory
constexpr int BRUTE_THREADS = 512;
constexpr int PROBE_THREADS = 256;
constexpr int PROBE2_THREADS = 1024;

// ------------------------------------------------------------------------------------------------------
// read loaders
// ------------------------------------------------------------------------------------------------------
template <int W>
FQ_D void load_packed(const uint32_t* __restrict__ packed, uint64_t i, uint32_t (&w)[W]) {
    if constexpr (W == 2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(packed) + i);
        w[0] = v.x;
        w[1] = v.y;
    } else if constexpr (W == 4) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(packed) + i);
        w[0] = v.x;
        w[1] = v.y;
        w[2] = v.z;
        w[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < W; k++) w[k] = __ldg(packed + i * W + k);
    }
}

// encode() of one ASCII row (mod.rs:49-61) through the shared-memory LUT.  Returns false when the row's
// length differs from L (short rows are None, barcode_matching.rs:167-169; long rows were vetted by the host).
template <int W>
FQ_D bool load_ascii(const ReadSource& src, uint64_t i, uint32_t L, const uint8_t* __restrict__ lut,
                     uint32_t (&w)[W]) {
#pragma unroll
    for (int k = 0; k < W; k++) w[k] = 0u;
    if (src.lengths != nullptr && __ldg(src.lengths + i) != L) return false;
    const uint8_t* row = src.ascii + i * src.stride;
    const bool aligned4 = ((reinterpret_cast<uintptr_t>(row) & 3u) == 0u);
#pragma unroll
    for (int wi = 0; wi < W; wi++) {
        uint32_t acc = 0u;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t k0 = wi * 8u + half * 4u;
            if (k0 + 4u <= L && aligned4) {
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(row + k0));
                acc |= (uint32_t)lut[v & 0xFFu] << (16 * half);
                acc |= (uint32_t)lut[(v >> 8) & 0xFFu] << (16 * half + 4);
                acc |= (uint32_t)lut[(v >> 16) & 0xFFu] << (16 * half + 8);
                acc |= (uint32_t)lut[v >> 24] << (16 * half + 12);
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t k = k0 + b;
                    if (k < L) acc |= (uint32_t)lut[__ldg(row + k)] << (16 * half + 4 * b);
                }
            }
        }
        w[wi] = acc;
    }
    return true;
}

FQ_D void init_lut(uint8_t* lut) {
    for (uint32_t t = threadIdx.x; t < 256u; t += blockDim.x) lut[t] = (uint8_t)encode_byte(t);
}

// shared-memory access by 32-bit window address (keeps address arithmetic to one IMAD / IADD per access)
FQ_D uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FQ_D uint4 lds128_ro(uint32_t a) {  // read-only data (tier): may be scheduled freely
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
FQ_D uint2 lds64_ro(uint32_t a) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
FQ_D uint32_t lds32_ro(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
FQ_D uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
FQ_D void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
template <int W>
FQ_D void sts_key(uint32_t a, const uint32_t (&w)[W]) {
    if constexpr (W == 1) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(w[0]) : "memory");
    } else if constexpr (W == 2) {
        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(w[0]), "r"(w[1]) : "memory");
    } else if constexpr (W == 3) {
        asm volatile("st.shared.u32 [%0], %1;\n\tst.shared.u32 [%0+4], %2;\n\tst.shared.u32 [%0+8], %3;" ::"r"(a), "r"(w[0]),
                     "r"(w[W > 1 ? 1 : 0]), "r"(w[W > 2 ? 2 : 0])
                     : "memory");
    } else {
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(w[0]), "r"(w[W > 1 ? 1 : 0]),
                     "r"(w[W > 2 ? 2 : 0]), "r"(w[W > 3 ? 3 : 0])
                     : "memory");
    }
}
template <int W>
FQ_D void lds_key(uint32_t a, uint32_t (&w)[W]) {
    if constexpr (W == 1) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(a) : "memory");
    } else if constexpr (W == 2) {
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(a) : "memory");
    } else if constexpr (W == 3) {
        asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                     : "=r"(w[0]), "=r"(w[W > 1 ? 1 : 0]), "=r"(w[W > 2 ? 2 : 0])
                     : "r"(a)
                     : "memory");
    } else {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(w[0]), "=r"(w[W > 1 ? 1 : 0]), "=r"(w[W > 2 ? 2 : 0]), "=r"(w[W > 3 ? 3 : 0])
                     : "r"(a)
                     : "memory");
    }
}
// ------------------------------------------------------------------------------------------------------
// counting
// ------------------------------------------------------------------------------------------------------
struct Counter {
    uint32_t* s_hist;           // nullptr -> straight to global
    unsigned long long* g_counts;
    uint32_t S;
    uint32_t none_local;

    FQ_D void init(uint32_t* smem_hist, const MatchParams& p) {
        S = p.S;
        g_counts = p.counts;
        none_local = 0u;
        s_hist = (p.S + 1u <= HIST_SMEM_BINS) ? smem_hist : nullptr;
        if (s_hist)
            for (uint32_t b = threadIdx.x; b <= S; b += blockDim.x) s_hist[b] = 0u;
    }
    FQ_D void add(uint32_t result) {
        if (result == NONE) {
            none_local++;  // unmatched is the hot bin: keep it in a register, reduce per warp at the end
        } else if (s_hist) {
            atomicAdd(&s_hist[result >> 16], 1u);
        } else {
            atomicAdd(&g_counts[result >> 16], 1ull);
        }
    }
    // all threads of the CTA must call this
    FQ_D void flush() {
        uint32_t v = none_local;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
        if ((threadIdx.x & 31u) == 0u && v) {
            if (s_hist)
                atomicAdd(&s_hist[S], v);
            else
                atomicAdd(&g_counts[S], (unsigned long long)v);
        }
        if (s_hist) {
            __syncthreads();
            for (uint32_t b = threadIdx.x; b <= S; b += blockDim.x) {
                const uint32_t c = s_hist[b];
                if (c) atomicAdd(&g_counts[b], (unsigned long long)c);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// k_brute: thread per read, panel planes in shared memory (or global when they do not fit)
//
// Barcodes are taken two at a time so the running best / second-best lives in u16x2 SIMD registers
// (VIMNMX.U16x2: 3 instructions per TWO barcodes).  PK = 2 (L <= 16) additionally packs the two barcodes' forbidden
// planes into the two 16-bit halves of one plane word, so the 4 logic ops are shared as well.
// The SIMD lanes track distances only; the sample index comes from a coarse tag: after every chunk of
// BRUTE_CHUNK_PAIRS pairs a strict improvement of the best distance records the chunk, and that one chunk is
// re-scanned at the end for the FIRST index with the best distance (strict '<' at barcode_matching.rs:132).
// ------------------------------------------------------------------------------------------------------
constexpr uint32_t BRUTE_CHUNK_PAIRS = 8;  // 16 barcodes per chunk

FQ_D uint32_t vmin2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
FQ_D uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }

// distances of the pair's two barcodes, packed {d_odd : d_even} as u16x2
template <int PK, bool PANEL_SMEM>
FQ_D uint32_t pair_distances(const uint4* __restrict__ planes, uint32_t pair, const uint32_t (&pl)[4]) {
    if constexpr (PK == 2) {
        const uint4 nb = PANEL_SMEM ? planes[pair] : __ldg(planes + pair);
        const uint32_t m = (pl[0] & nb.x) | (pl[1] & nb.y) | (pl[2] & nb.z) | (pl[3] & nb.w);
        return (uint32_t)__popc(m >> 16) * 65536u + (uint32_t)__popc(m & 0xFFFFu);
    } else {
        const uint4 n0 = PANEL_SMEM ? planes[2u * pair] : __ldg(planes + 2u * pair);
        const uint4 n1 = PANEL_SMEM ? planes[2u * pair + 1u] : __ldg(planes + 2u * pair + 1u);
        const uint32_t m0 = (pl[0] & n0.x) | (pl[1] & n0.y) | (pl[2] & n0.z) | (pl[3] & n0.w);
        const uint32_t m1 = (pl[0] & n1.x) | (pl[1] & n1.y) | (pl[2] & n1.z) | (pl[3] & n1.w);
        return (uint32_t)__popc(m1) * 65536u + (uint32_t)__popc(m0);
    }
}

template <int W, bool ASCII, bool PANEL_SMEM, int PK>
__global__ void __launch_bounds__(BRUTE_THREADS) k_brute(const MatchParams p, const ReadSource src,
                                                         uint32_t* __restrict__ results) {
    extern __shared__ uint4 s_dyn[];
    __shared__ uint8_t s_lut[256];
    uint4* s_planes = s_dyn;
    const uint32_t n_pairs = (p.S + 1u) / 2u;
    const uint32_t n_plane_words = (PK == 2) ? n_pairs : 2u * n_pairs;  // uint4 entries of the pair-ordered panel
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_dyn + (PANEL_SMEM ? n_plane_words : 0u));
    const uint4* __restrict__ g_pairs = (PK == 2) ? p.planes2 : p.planes;  // planes[] is padded to an even count

    if constexpr (PANEL_SMEM) {
        for (uint32_t j = threadIdx.x; j < n_plane_words; j += blockDim.x) s_planes[j] = __ldg(g_pairs + j);
    }
    if constexpr (ASCII) init_lut(s_lut);
    Counter cnt;
    cnt.init(s_hist, p);
    __syncthreads();

    const uint4* __restrict__ planes = PANEL_SMEM ? s_planes : g_pairs;
    const uint32_t odd_tail = (p.S & 1u) ? 0xFFFF0000u : 0u;  // the last pair's second barcode does not exist
    const uint32_t n_full_chunks = (n_pairs - (odd_tail ? 1u : 0u)) / BRUTE_CHUNK_PAIRS;
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src.n; i += total) {
        uint32_t w[W];
        bool row_ok = true;
        if constexpr (ASCII)
            row_ok = load_ascii<W>(src, i, p.L, s_lut, w);
        else
            load_packed<W>(src.packed, i, w);
        uint32_t pl[4];
        planes_from_words<W>(w, pl);
        if constexpr (PK == 2) {
#pragma unroll
            for (int k = 0; k < 4; k++) pl[k] |= pl[k] << 16;  // the read's planes, once per packed barcode
        }
        uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;  // u16x2 running best / second best (even | odd barcodes)
        uint32_t best = 0xFFFFu, best_chunk = 0u;
        uint32_t pair = 0;
        for (uint32_t c = 0; c < n_full_chunks; c++) {
#pragma unroll
            for (uint32_t u = 0; u < BRUTE_CHUNK_PAIRS; u++, pair++) {
                const uint32_t d2 = pair_distances<PK, PANEL_SMEM>(planes, pair, pl);
                const uint32_t hi = vmax2(k1, d2);
                k1 = vmin2(k1, d2);
                k2 = vmin2(k2, hi);
            }
            const uint32_t m = min(k1 & 0xFFFFu, k1 >> 16);
            if (m < best) {
                best = m;
                best_chunk = c;
            }
        }
        for (uint32_t c = n_full_chunks; pair < n_pairs; c++) {  // last, partial chunk (and the odd tail barcode)
            const uint32_t stop = min(pair + BRUTE_CHUNK_PAIRS, n_pairs);
            for (; pair < stop; pair++) {
                uint32_t d2 = pair_distances<PK, PANEL_SMEM>(planes, pair, pl);
                if (pair == n_pairs - 1u) d2 |= odd_tail;
                const uint32_t hi = vmax2(k1, d2);
                k1 = vmin2(k1, d2);
                k2 = vmin2(k2, hi);
            }
            const uint32_t m = min(k1 & 0xFFFFu, k1 >> 16);
            if (m < best) {
                best = m;
                best_chunk = c;
            }
        }
        // second smallest with multiplicity across the two SIMD lanes (barcode_matching.rs:140)
        const uint32_t a1 = k1 & 0xFFFFu, b1 = k1 >> 16, a2 = k2 & 0xFFFFu, b2 = k2 >> 16;
        uint32_t next = min(max(a1, b1), min(a2, b2));
        next = (next == 0xFFFFu) ? 255u : next;  // single-sample panel keeps the 255 sentinel (:122)
        uint32_t res = NONE;
        if (row_ok && best <= p.max_mm && (next - best) >= p.min_delta) {
            // first index with the best distance inside the recorded chunk
            uint32_t idx = 0xFFFFu;
            if constexpr (PK == 2) {
#pragma unroll 1
                for (uint32_t q = 0; q < BRUTE_CHUNK_PAIRS; q++) {
                    const uint32_t pr = best_chunk * BRUTE_CHUNK_PAIRS + q;
                    if (pr < n_pairs) {
                        uint32_t d2 = pair_distances<PK, PANEL_SMEM>(planes, pr, pl);
                        if (pr == n_pairs - 1u) d2 |= odd_tail;
                        if ((d2 >> 16) == best) idx = min(idx, 2u * pr + 1u);
                        if ((d2 & 0xFFFFu) == best) idx = min(idx, 2u * pr);
                    }
                }
            } else {
#pragma unroll 1
                for (uint32_t q = 0; q < BRUTE_CHUNK_PAIRS; q++) {
                    const uint32_t pr = best_chunk * BRUTE_CHUNK_PAIRS + q;
                    if (pr < n_pairs) {
                        uint32_t d2 = pair_distances<PK, PANEL_SMEM>(planes, pr, pl);
                        if (pr == n_pairs - 1u) d2 |= odd_tail;
                        if ((d2 >> 16) == best) idx = min(idx, 2u * pr + 1u);
                        if ((d2 & 0xFFFFu) == best) idx = min(idx, 2u * pr);
                    }
                }
            }
            res = (idx << 16) | (best << 8) | next;
        }
        results[i] = res;
        cnt.add(res);
    }
    cnt.flush();
}

// ------------------------------------------------------------------------------------------------------
// k_brute_sliced: every read x every barcode, BIT-SLICED ACROSS BARCODES (SURVEY 7 option iii).  Thread per read; the panel
// is transposed into groups of 32 barcodes: word (g, i, v) has bit j set iff barcode 32g + j mismatches a read symbol with
// 4-bit mask v at position i (bitenc.rs:441-452, for ALL 16 masks: IUPAC codes, no-calls and junk bytes in reads are exact
// by construction).  For one group a thread fetches the 8W words its own symbols select (conflict-free: the lanes of a
// warp differ only in v, 16 consecutive words) and adds them with a carry-save adder tree into vertical counters (bit p of
// 32 distances at once: ~2 logic ops per position instead of ~8 instructions per barcode).
// The running minima stay bit-sliced too, and the group loop has no branch: bit j of the planes mn / ls holds the smallest
// and the second smallest distance seen so far among the barcodes 32g + j of all groups g ("slot" j), gi the group that
// gave the smallest one.  One group costs a compare-exchange of the counters with mn (the borrow chain of c - mn is one
// LOP3 per plane, the two selects one each), a min of the loser with ls, and one LOP3 per plane of gi.  Only after the
// last group are the 32 slots reduced: smallest mn, among equals the smallest group, then the smallest slot — the FIRST
// index with the best distance (barcode_matching.rs:132: strict <) — and the second smallest of the multiset (:140) is
// the smallest of every other slot's mn and the winning slot's ls.  (Round 2's first version compared every group with
// the running second best and extracted the barcodes below it one by one: that branch was taken by some lane of nearly
// every warp, 48 % of the kernel's instructions at S = 384.)
// Shared memory holds a chunk of groups; a panel larger than a chunk is walked chunk by chunk for every batch of reads.
// ------------------------------------------------------------------------------------------------------
// two CTAs per SM: 512 threads x <= 64 registers (W <= 2), 384 threads x <= 85 registers (W = 3, 4: off[] alone is 8W)
constexpr int sliced_threads(int W) { return W <= 2 ? 512 : 384; }
constexpr uint32_t SLICED_CHUNK_BYTES = 96u << 10;  // table bytes staged at a time
#ifndef SLICED_UNROLL
#define SLICED_UNROLL 4
#endif

template <int W>
struct Sliced {
    static constexpr int LP = 8 * W;                         // positions, padded to whole packed words
    static constexpr int NB = SlicedAdder<LP>::PLANES;       // counter planes: distances 0 .. LP
    static constexpr uint32_t GROUP_BYTES = LP * 16 * 4;     // one group of 32 barcodes
    static_assert((1 << NB) - 1 > LP, "the all-ones distance must be larger than any real one (it stands for 'no barcode')");
    static_assert(SLICED_CHUNK_BYTES / GROUP_BYTES <= 256, "group index inside a chunk: 8 planes");
};

// counters of one group for this thread's read: `a` = shared-window address of the group + the read's symbol offsets
template <int W>
FQ_D void sliced_counters(const MatchParams& p, const uint32_t (&off)[8 * W], uint32_t group_addr, uint32_t (&c)[Sliced<W>::NB]) {
    constexpr int LP = Sliced<W>::LP;
    uint32_t x[LP];
#pragma unroll
    for (int i = 0; i < LP; i++)  // the add is a multiply-add (fma pipe): the alu pipe is what bounds this kernel
        x[i] = lds32_ro(off[i] * p.ck_one + group_addr + (uint32_t)i * 64u);
    SlicedAdder<LP>::add(x, c);
}

// bit-sliced a < b over 32 slots: the borrow out of a - b, one three-input logic op per plane
template <int NB>
FQ_D uint32_t sliced_less(const uint32_t (&a)[NB], const uint32_t (&b)[NB], uint32_t one) {
    uint32_t br = ~a[0] & b[0];
#pragma unroll
    for (int pl = 1; pl < NB; pl++)  // (~a & b) | (~(a ^ b) & br) — spelled as the one LOP3 it is (ptxas splits the C form)
        asm("lop3.b32 %0, %1, %2, %0, 0x8E;" : "+r"(br) : "r"(a[pl]), "r"(b[pl]));
    return br * one;  // one = p.ck_one: a multiply on the fma pipe keeps ptxas from folding the last step into each of its
                      // users (7 alu-pipe LOP3 for the top plane instead of 3)
}

// smallest value among the slots of `pool` (non-empty): returns it, narrows `pool` to the slots that hold it
template <int NB>
FQ_D uint32_t sliced_min(const uint32_t (&v)[NB], uint32_t& pool) {
    uint32_t d = 0u;
#pragma unroll
    for (int pl = NB - 1; pl >= 0; pl--) {
        const uint32_t t = pool & ~v[pl];
        d |= t ? 0u : (1u << pl);
        pool = t ? t : pool;
    }
    return d;
}

// Interim state of a read between two chunk launches, kept in its result word: idx | best << 16 | second best << 24
// (distances <= 32, a multi-chunk panel has more than one barcode).
FQ_D uint32_t sliced_pack_state(uint32_t k1, uint32_t k2) { return (k1 & 0xFFFFu) | ((k1 >> 16) << 16) | ((k2 >> 16) << 24); }

// One launch = one chunk of the panel (groups [chunk * chunk_groups, ...)) against every read; the chunk is staged once per
// CTA.  chunk 0 runs the slot form and reduces it; later chunks (panels of thousands of barcodes) only compare each group
// with the second-best distance T so far — a barcode below it is rare by then (about 64 / (barcodes seen) per group and
// read) — and extract those; the last chunk decides, writes the result words and counts.
template <int W, bool ASCII, int GBITS, bool FIRST>  // GBITS: planes of the group index inside chunk 0 (2^GBITS >= its groups);
__global__ void __launch_bounds__(sliced_threads(W), 2) k_brute_sliced(const MatchParams p, const ReadSource src,
                                                                 uint32_t* __restrict__ results, uint32_t chunk_groups,
                                                                 uint32_t chunk) {
    constexpr int LP = Sliced<W>::LP, NB = Sliced<W>::NB;
    constexpr uint32_t GB = Sliced<W>::GROUP_BYTES;
    extern __shared__ uint4 s_dyn[];
    __shared__ uint8_t s_lut[256];
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_dyn);
    uint32_t* s_hist = s_tab + (size_t)chunk_groups * (GB / 4);
    const uint32_t G = (p.S + 31u) / 32u;
    const uint32_t g0 = chunk * chunk_groups, ng = min(chunk_groups, G - g0);
    const bool last = g0 + ng == G;
    if constexpr (ASCII) init_lut(s_lut);
    Counter cnt;
    cnt.init(s_hist, p);
    {
        const uint4* src4 = reinterpret_cast<const uint4*>(p.sliced + (size_t)g0 * (GB / 4));
        uint4* dst4 = reinterpret_cast<uint4*>(s_tab);
        for (uint32_t t = threadIdx.x; t < ng * (GB / 16); t += blockDim.x) dst4[t] = __ldg(src4 + t);
    }
    __syncthreads();
    const uint32_t a_tab = smem_addr(s_tab);
    const uint32_t tail_mask = (p.S & 31u) ? ~((1u << (p.S & 31u)) - 1u) : 0u;  // slots of the last group past the panel
    const uint64_t batch = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src.n; i += batch) {
        uint32_t w[W];
        bool row_ok = true;
        if constexpr (ASCII)
            row_ok = load_ascii<W>(src, i, p.L, s_lut, w);
        else
            load_packed<W>(src.packed, i, w);
        uint32_t off[LP];  // byte offset of the read's symbol inside a position's 16 words
#pragma unroll
        for (int k = 0; k < LP; k++) {
            const int sh = 4 * (k & 7) - 2;
            off[k] = (sh >= 0 ? (w[k >> 3] >> sh) : (w[k >> 3] << 2)) & 0x3Cu;
        }
        uint32_t k1, k2;  // best / second best as (distance << 16 | index)
        if constexpr (FIRST) {  // chunk 0
            // per slot: smallest / second smallest distance so far (all ones = none yet), group of the smallest
            uint32_t mn[NB], ls[NB], gi[GBITS];
#pragma unroll
            for (int pl = 0; pl < NB; pl++) mn[pl] = ls[pl] = 0xFFFFFFFFu;
#pragma unroll
            for (int b = 0; b < GBITS; b++) gi[b] = 0u;
            // SLICED_UNROLL groups per trip: the low planes of the group index are compile-time constants and the masks of
            // the others are built once per trip (uniform datapath, but they take issue slots)
            constexpr int UB = SLICED_UNROLL == 4 ? 2 : 1;
            for (uint32_t gl = 0; gl < ng; gl += (uint32_t)SLICED_UNROLL) {
                const uint32_t gq = gl >> UB;
#pragma unroll
                for (int r = 0; r < SLICED_UNROLL; r++) {
                    if (gl + (uint32_t)r >= ng) break;  // (warp-uniform)
                    uint32_t c[NB];
                    sliced_counters<W>(p, off, a_tab + (gl + (uint32_t)r) * GB, c);
                    if (gl + (uint32_t)r + 1u == G) {  // (warp-uniform) slots past the end of the panel never win
#pragma unroll
                        for (int pl = 0; pl < NB; pl++) c[pl] |= tail_mask;
                    }
                    const uint32_t lt = sliced_less<NB>(c, mn, p.ck_one);  // slots whose new barcode is STRICTLY closer: ties keep the earlier group
                    uint32_t lost[NB];
#pragma unroll
                    for (int pl = 0; pl < NB; pl++) {
                        lost[pl] = (mn[pl] & lt) | (c[pl] & ~lt);  // the larger of the two
                        mn[pl] = (c[pl] & lt) | (mn[pl] & ~lt);
                    }
                    const uint32_t lt2 = sliced_less<NB>(lost, ls, p.ck_one);
#pragma unroll
                    for (int pl = 0; pl < NB; pl++) ls[pl] = (lost[pl] & lt2) | (ls[pl] & ~lt2);
#pragma unroll
                    for (int b = 0; b < UB; b++) gi[b] = ((r >> b) & 1) ? (gi[b] | lt) : (gi[b] & ~lt);
#pragma unroll
                    for (int b = UB; b < GBITS; b++) gi[b] = (gi[b] & ~lt) | (lt & (0u - ((gq >> (b - UB)) & 1u)));
                }
            }
            // reduce the 32 slots: best = smallest mn, among equals the smallest group, then the smallest slot
            uint32_t pool = 0xFFFFFFFFu;
            const uint32_t d1 = sliced_min<NB>(mn, pool);
            const uint32_t g1 = sliced_min<GBITS>(gi, pool);
            const uint32_t j1 = (uint32_t)__ffs(pool) - 1u;
            k1 = (d1 << 16) | (g1 * 32u + j1);
            // second smallest of the multiset: every other slot's smallest, the winning slot's second smallest
            const uint32_t win = 1u << j1;
            uint32_t q[NB];
#pragma unroll
            for (int pl = 0; pl < NB; pl++) q[pl] = (ls[pl] & win) | (mn[pl] & ~win);
            pool = 0xFFFFFFFFu;
            k2 = (p.S < 2u) ? EMPTY_KEY : (sliced_min<NB>(q, pool) << 16);  // one-sample panel: next_best stays 255 (:122)
        } else {
            const uint32_t st = results[i];
            k1 = ((st >> 16) & 0xFFu) << 16 | (st & 0xFFFFu);
            k2 = (st >> 24) << 16;
            uint32_t thr[NB];  // planes of T
#pragma unroll
            for (int pl = 0; pl < NB; pl++) thr[pl] = 0u - (((k2 >> 16) >> pl) & 1u);
            for (uint32_t gl = 0; gl < ng; gl++) {
                const uint32_t g = g0 + gl;
                uint32_t c[NB];
                sliced_counters<W>(p, off, a_tab + gl * GB, c);
                if (g + 1u == G) {
#pragma unroll
                    for (int pl = 0; pl < NB; pl++) c[pl] |= tail_mask;
                }
                uint32_t lt = sliced_less<NB>(c, thr, p.ck_one);
                if (lt == 0u) continue;
                do {
                    const uint32_t j = (uint32_t)__ffs(lt) - 1u;
                    lt &= lt - 1u;
                    uint32_t d = 0u;
#pragma unroll
                    for (int pl = 0; pl < NB; pl++) d |= ((c[pl] >> j) & 1u) << pl;
                    track2(k1, k2, (d << 16) | (g * 32u + j));
                } while (lt);
#pragma unroll
                for (int pl = 0; pl < NB; pl++) thr[pl] = 0u - (((k2 >> 16) >> pl) & 1u);
            }
        }
        if (last) {
            const uint32_t res = row_ok ? decide(k1, k2, p.max_mm, p.min_delta) : NONE;
            results[i] = res;
            cnt.add(res);
        } else {
            results[i] = sliced_pack_state(k1, k2);
        }
    }
    if (last) cnt.flush();
}

// L > 32 (W = 5 .. 32 packed words; rare — barcodes this long are unusual): nibble-word form straight from the packed
// layout.  The read's words stay in registers (WMAX is a compile-time bound, every loop over them is fully unrolled and
// predicated on the real W), the panel's ~expected words are staged in shared memory, padded to a multiple of four words
// per barcode so that a (warp-uniform, broadcast) LDS.128 fetches four at a time; when the panel does not fit it is read
// through the read-only cache instead.  Per word: x = r & ~exp; a nibble of x is non-zero iff the symbol mismatches
// (bitenc.rs:441-452); the per-nibble flags of four words share one POPC.  Rows whose length differs from L (only when
// the caller passed lengths) are None (barcode_matching.rs:167-169; longer rows were vetted by the host).
template <int WMAX, bool PANEL_SMEM>
__global__ void __launch_bounds__(256) k_brute_long(const MatchParams p, const ReadSource src,
                                                    uint32_t* __restrict__ results) {
    static_assert(WMAX % 4 == 0, "four words per LDS.128");
    extern __shared__ uint4 s_dyn[];
    const uint32_t W = p.W, Wq = (W + 3u) / 4u;  // Wq = uint4 entries per barcode
    uint4* s_ne = s_dyn;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_dyn + (PANEL_SMEM ? (size_t)p.S * Wq : 0));
    if constexpr (PANEL_SMEM) {
        uint32_t* flat = reinterpret_cast<uint32_t*>(s_ne);
        for (uint32_t t = threadIdx.x; t < p.S * Wq * 4u; t += blockDim.x) {
            const uint32_t j = t / (Wq * 4u), k = t % (Wq * 4u);
            flat[t] = k < W ? __ldg(p.not_exp + (size_t)j * W + k) : 0u;
        }
    }
    Counter cnt;
    cnt.init(s_hist, p);
    __syncthreads();
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src.n; i += total) {
        uint32_t res = NONE;
        if (src.lengths == nullptr || __ldg(src.lengths + i) == p.L) {
            uint32_t r[WMAX];
#pragma unroll
            for (int k = 0; k < WMAX; k++) r[k] = (uint32_t)k < W ? __ldg(src.packed + i * W + k) : 0u;
            uint32_t k1 = EMPTY_KEY, k2 = EMPTY_KEY;
            for (uint32_t j = 0; j < p.S; j++) {
                uint32_t d = 0;
#pragma unroll
                for (int q = 0; q < WMAX / 4; q++) {
                    if ((uint32_t)q < Wq) {
                        uint4 ne;
                        if constexpr (PANEL_SMEM) {
                            ne = s_ne[j * Wq + q];
                        } else {
                            const uint32_t* g = p.not_exp + (size_t)j * W + 4 * q;
                            ne.x = __ldg(g);
                            ne.y = 4u * q + 1u < W ? __ldg(g + 1) : 0u;
                            ne.z = 4u * q + 2u < W ? __ldg(g + 2) : 0u;
                            ne.w = 4u * q + 3u < W ? __ldg(g + 3) : 0u;
                        }
                        // bit 3 of every nibble of f* = "this nibble of (r & ~exp) is non-zero"
                        const uint32_t x0 = r[4 * q] & ne.x, x1 = r[4 * q + 1] & ne.y, x2 = r[4 * q + 2] & ne.z,
                                       x3 = r[4 * q + 3] & ne.w;
                        const uint32_t f0 = (((x0 & 0x77777777u) + 0x77777777u) | x0) & 0x88888888u;
                        const uint32_t f1 = (((x1 & 0x77777777u) + 0x77777777u) | x1) & 0x88888888u;
                        const uint32_t f2 = (((x2 & 0x77777777u) + 0x77777777u) | x2) & 0x88888888u;
                        const uint32_t f3 = (((x3 & 0x77777777u) + 0x77777777u) | x3) & 0x88888888u;
                        d += (uint32_t)__popc(f0 | (f1 >> 1) | (f2 >> 2) | (f3 >> 3));
                    }
                }
                track2(k1, k2, (d << 16) | j);
            }
            res = decide(k1, k2, p.max_mm, p.min_delta);
        }
        results[i] = res;
        cnt.add(res);
    }
    cnt.flush();
}

// ------------------------------------------------------------------------------------------------------
// memo-table kernels
// ------------------------------------------------------------------------------------------------------
template <int W>
FQ_D uint32_t warp_brute_one(const MatchParams& p, const uint32_t (&w)[W], uint32_t lane) {
    uint32_t pl[4];
    planes_from_words<W>(w, pl);
    uint32_t k1 = EMPTY_KEY, k2 = EMPTY_KEY;
    for (uint32_t j = lane; j < p.S; j += 32u) {
        const uint4 nb = __ldg(p.planes + j);
        const uint32_t m = (pl[0] & nb.x) | (pl[1] & nb.y) | (pl[2] & nb.z) | (pl[3] & nb.w);
        track2(k1, k2, ((uint32_t)__popc(m) << 16) | j);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, off);
        const uint32_t o2 = __shfl_xor_sync(0xFFFFFFFFu, k2, off);
        merge2(k1, k2, o1, o2);
    }
    return decide(k1, k2, p.max_mm, p.min_delta);
}

// Memo-table slots (kernels.h): W <= 3 -> 16 bytes {k0, k1, k2, value}, one LDG.128 per probe;
// W = 4 -> 32 bytes {k0..k3, value, pad}, one 256-bit load (LDG.E.ENL2.256 on sm_100a) per probe.
FQ_D void load_slot32(const uint32_t* __restrict__ table, uint32_t slot, uint32_t (&e)[8]) {
    const uint32_t* q = table + (size_t)slot * 8;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]), "=r"(e[4]), "=r"(e[5]), "=r"(e[6]), "=r"(e[7])
                 : "l"(q));
}

// One memo-table lookup: linear probing, single-exit loop; an empty slot (value NONE) ends the probe.
// Returns true on a hit (res = stored Some(..) word); res is NONE otherwise.
template <int W>
FQ_D bool table_lookup(const MatchParams& p, const uint32_t (&w)[W], uint32_t h, uint32_t& res) {
    uint32_t b = bucket_of_hash(h, p.n_buckets);
    for (;;) {
        uint32_t v;  // the probed slot's value
        if constexpr (W <= 3) {
            const uint4 e = __ldg(reinterpret_cast<const uint4*>(p.table) + b);
            const uint32_t k1 = W > 1 ? w[W > 1 ? 1 : 0] : 0u, k2 = W > 2 ? w[W > 2 ? 2 : 0] : 0u;
            v = e.w;
            res = (e.x == w[0] && e.y == k1 && e.z == k2) ? v : NONE;  // an empty slot's all-ones key may equal an
        } else {                                                        // all-N read: its value is NONE anyway
            uint32_t e[8];
            load_slot32(p.table, b, e);
            v = e[4];
            res = (e[0] == w[0] && e[1] == w[1] && e[2] == w[2] && e[3] == w[3]) ? v : NONE;
        }
        if (res != NONE || v == NONE) break;
        b = (b + 1u == p.n_buckets) ? 0u : b + 1u;
    }
    return res != NONE;
}

// k_probe: one read per thread; the ASCII (host-batch / e2e) route and the fallback for unaligned device buffers.
template <int W, bool ASCII>
__global__ void __launch_bounds__(PROBE_THREADS) k_probe(const MatchParams p, const ReadSource src,
                                                         uint32_t* __restrict__ results) {
    extern __shared__ uint4 s_dyn[];
    __shared__ uint8_t s_lut[256];
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_dyn);
    if constexpr (ASCII) init_lut(s_lut);
    Counter cnt;
    cnt.init(s_hist, p);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    // warp-uniform trip count so the whole warp is present for the cooperative slow path
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < src.n; base += total) {
        const uint64_t i = base + lane;
        const bool valid = i < src.n;
        uint32_t w[W];
#pragma unroll
        for (int k = 0; k < W; k++) w[k] = 0u;
        bool row_ok = true;
        if (valid) {
            if constexpr (ASCII)
                row_ok = load_ascii<W>(src, i, p.L, s_lut, w);
            else
                load_packed<W>(src.packed, i, w);
        }
        uint32_t res = NONE;
        bool slow = false;
        if (valid && row_ok) {
            const bool hit = table_lookup<W>(p, w, hash_key<W>(w), res);
            slow = !hit && !read_in_table_alphabet<W>(w, p.last_pad);
        }
        uint32_t pending = __ballot_sync(0xFFFFFFFFu, slow);
        while (pending) {
            const int src_lane = __ffs(pending) - 1;
            pending &= pending - 1u;
            uint32_t bw[W];
#pragma unroll
            for (int k = 0; k < W; k++) bw[k] = __shfl_sync(0xFFFFFFFFu, w[k], src_lane);
            const uint32_t out = warp_brute_one<W>(p, bw, lane);
            if ((int)lane == src_lane) res = out;
        }
        if (valid) {
            results[i] = res;
            cnt.add(res);
        }
    }
    cnt.flush();
}

// L2 eviction-priority policies: k_probe4's table must survive the read / result stream that flows through L2 next to it
FQ_D uint64_t l2_policy_keep() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
FQ_D uint64_t l2_policy_stream() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
FQ_D uint64_t l2_policy_normal() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
FQ_D uint4 ldg_hint(const uint4* ptr, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(ptr), "l"(pol));
    return v;
}
FQ_D void stg_hint(uint4* ptr, const uint4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol)
                 : "memory");
}

// k_probe2: the HBM-resident packed route.  A warp owns tiles of 128 consecutive reads, four per lane
// (W x LDG.128 in, one STG.128 out), and resolves them through three tiers:
//   1. hot tier in SHARED memory: the table entries whose best distance is 0 (the barcodes themselves and their
//      expansions), 2-choice cuckoo -> two independent LDS per read, no loop.  ~80 % of real reads end here.
//   2. the tile's remaining reads (all four slots) are compacted into a per-warp shared-memory queue so that the
//      global memo-table probe, the alphabet test and the result write-back run once per ~32 such reads instead of
//      once per slot with mostly idle lanes.
//   3. reads outside the table's alphabet: warp-cooperative brute force (same as k_probe).
// The < 128-read tail of a batch is finished by one warp with the one-read-per-lane code of k_probe.
constexpr int PROBE2_R = 4;
constexpr int PROBE2_TILE = 32 * PROBE2_R;   // reads per warp tile
constexpr int PROBE2_QUEUE = PROBE2_TILE;    // queue entries per warp: every read of a tile may be unresolved
constexpr uint32_t PROBE2_HIST_BYTES = 25u << 10;  // shared memory the replicated histogram may take

// histogram: hist[(result >> 16) * hrep + lane % hrep] += 1 unless the result is NONE (predicated, no branch).
// `hist_lane_addr` already includes the lane's replica offset; `hshift` = log2(hrep * 4 bytes).
FQ_D void red_hist(uint32_t hist_lane_addr, uint32_t hshift, uint32_t result) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 a;\n\t"
        "setp.ne.u32 p, %2, 0xffffffff;\n\t"
        "shr.u32 a, %2, 16;\n\tshl.b32 a, a, %1;\n\tadd.u32 a, a, %0;\n\t"
        "@p red.shared.add.u32 [a], 1;\n\t}"
        ::"r"(hist_lane_addr), "r"(hshift), "r"(result)
        : "memory");
}

template <int W>
__global__ void __launch_bounds__(PROBE2_THREADS, 1) k_probe2(const MatchParams p, const ReadSource src,
                                                           uint32_t* __restrict__ results) {
    constexpr int R = PROBE2_R;
    constexpr int TE = W <= 2 ? 2 : 4;              // tier entry words (kernels.h)
    constexpr bool SEPV = (W == 2 || W == 4);       // values in their own (unreplicated) array
    extern __shared__ uint4 s_dyn[];
    // layout: tier replicas | (W = 4: tier values) | Bloom words | per-warp queues | histogram replicas
    uint32_t* s_tier = reinterpret_cast<uint32_t*>(s_dyn);
    const uint32_t rep = p.tier_rep;
    uint32_t* s_tvals = s_tier + (size_t)p.tier_slots * rep * TE;
    uint32_t* s_bloom = s_tvals + (SEPV ? p.tier_slots : 0u);
    uint32_t* s_queue = s_bloom + p.bloom_words;
    const uint32_t n_warps = blockDim.x >> 5;
    uint32_t* s_hist = s_queue + (size_t)n_warps * PROBE2_QUEUE * W;

    // entry e, replica r lives at word (e * rep + r) * TE: with rep = 8 (16 with 8-byte entries) replica r owns one
    // fixed group of banks, so the 32 lanes of a probe hit every bank group exactly 4 (2) times — no conflicts
    for (uint32_t t = threadIdx.x; t < p.tier_slots * rep * TE; t += blockDim.x) {
        const uint32_t e = t / (rep * TE), k = t % TE;
        s_tier[t] = __ldg(p.tier_entries + e * TE + k);
    }
    if constexpr (SEPV)
        for (uint32_t t = threadIdx.x; t < p.tier_slots; t += blockDim.x)
            s_tvals[t] = __ldg(p.tier_entries + (size_t)p.tier_slots * TE + t);
    for (uint32_t t = threadIdx.x; t < p.bloom_words; t += blockDim.x) s_bloom[t] = __ldg(p.bloom + t);
    // histogram replicas: bin b of replica r at word b * hrep + r, lane l adds into replica l % hrep, so the lanes of
    // one RED spread over distinct banks; the replicas are summed when the CTA flushes
    const uint32_t hrep = p.hist_rep;
    for (uint32_t t = threadIdx.x; t < (p.S + 1u) * hrep; t += blockDim.x) s_hist[t] = 0u;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lane_lt = (1u << lane) - 1u;
    const uint32_t warp_in_cta = threadIdx.x >> 5;
    // 32-bit shared-window addresses, computed once
    uint32_t a_tier = smem_addr(s_tier) + (lane & (rep - 1u)) * (TE * 4u);  // this lane's replica
    const uint32_t tier_step = rep * TE * 4u;                                // bytes between slots of one replica
    uint32_t a_qk = smem_addr(s_queue) + warp_in_cta * (PROBE2_QUEUE * W * 4u);  // entry q: key in, result out (word 0)
    uint32_t a_hist = smem_addr(s_hist) + (lane & (hrep - 1u)) * 4u;
    const uint32_t hshift = 2u + (uint32_t)__popc(hrep - 1u);
    const uint32_t a_tvals = smem_addr(s_tvals), a_bloom = smem_addr(s_bloom);
    asm volatile("" : "+r"(a_tier), "+r"(a_qk), "+r"(a_hist));  // keep them in registers
    const uint32_t tshift = p.tier_shift;  // 32 - log2(tier_slots)
    const bool has_tier = tshift < 32u;
    uint32_t none_count = 0;

    const uint32_t n_tiles = (uint32_t)(src.n / (uint64_t)PROBE2_TILE);
    const uint32_t warp_stride = gridDim.x * n_warps;
    uint32_t tile = blockIdx.x * n_warps + warp_in_cta;

    // software pipeline: the next tile's keys are in flight while this one is resolved
    uint32_t nxt[R * W];
    if (tile < n_tiles) {
        const uint4* in = reinterpret_cast<const uint4*>(src.packed) + (size_t)(tile * 32u + lane) * W;
#pragma unroll
        for (int v = 0; v < W; v++) {
            const uint4 q = __ldg(in + v);
            nxt[4 * v + 0] = q.x;
            nxt[4 * v + 1] = q.y;
            nxt[4 * v + 2] = q.z;
            nxt[4 * v + 3] = q.w;
        }
    }
    for (; tile < n_tiles; tile += warp_stride) {
        const uint32_t g = tile * 32u + lane;  // this lane's group of 4 consecutive reads
        uint32_t w[R][W];
        uint32_t res[R];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int k = 0; k < W; k++) w[r][k] = nxt[r * W + k];
        if (tile + warp_stride < n_tiles) {
            const uint4* in = reinterpret_cast<const uint4*>(src.packed) + (size_t)((tile + warp_stride) * 32u + lane) * W;
#pragma unroll
            for (int v = 0; v < W; v++) {
                const uint4 q = __ldg(in + v);
                nxt[4 * v + 0] = q.x;
                nxt[4 * v + 1] = q.y;
                nxt[4 * v + 2] = q.z;
                nxt[4 * v + 3] = q.w;
            }
        }

        // ---- tier 1: shared-memory cuckoo probe (a NONE value = empty slot = not found here) ----
#pragma unroll
        for (int r = 0; r < R; r++) {
            res[r] = NONE;
            if (has_tier) {
                const uint32_t s1 = tier_hash1<W>(w[r]) >> tshift, s2 = tier_hash2<W>(w[r]) >> tshift;
                const uint32_t a1 = a_tier + s1 * tier_step, a2 = a_tier + s2 * tier_step;
                if constexpr (W == 1) {
                    const uint2 a = lds64_ro(a1), b = lds64_ro(a2);
                    res[r] = (a.x == w[r][0]) ? a.y : ((b.x == w[r][0]) ? b.y : NONE);
                } else if constexpr (W == 2) {
                    const uint2 a = lds64_ro(a1), b = lds64_ro(a2);
                    const bool m1 = a.x == w[r][0] && a.y == w[r][W > 1 ? 1 : 0];
                    const bool m2 = b.x == w[r][0] && b.y == w[r][W > 1 ? 1 : 0];
                    if (m1 || m2) res[r] = lds32_ro(a_tvals + (m1 ? s1 : s2) * 4u);
                } else {
                    const uint4 a = lds128_ro(a1), b = lds128_ro(a2);
                    if constexpr (W == 3) {
                        const bool m1 = a.x == w[r][0] && a.y == w[r][1] && a.z == w[r][W > 2 ? 2 : 0];
                        const bool m2 = b.x == w[r][0] && b.y == w[r][1] && b.z == w[r][W > 2 ? 2 : 0];
                        res[r] = m1 ? a.w : (m2 ? b.w : NONE);
                    } else {
                        const bool m1 = a.x == w[r][0] && a.y == w[r][1] && a.z == w[r][W > 2 ? 2 : 0] && a.w == w[r][W > 3 ? 3 : 0];
                        const bool m2 = b.x == w[r][0] && b.y == w[r][1] && b.z == w[r][W > 2 ? 2 : 0] && b.w == w[r][W > 3 ? 3 : 0];
                        if (m1 || m2) res[r] = lds32_ro(a_tvals + (m1 ? s1 : s2) * 4u);
                    }
                }
            }
        }

        // ---- tier 2: compact the tile's unresolved reads into the warp's queue ----
        uint32_t qcount = 0;  // warp-uniform
        uint32_t qi[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const bool pend = res[r] == NONE;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pend);
            qi[r] = qcount + __popc(bal & lane_lt);
            if (pend) sts_key<W>(a_qk + qi[r] * (W * 4u), w[r]);
            qcount += __popc(bal);
        }
        __syncwarp();
        for (uint32_t qb = 0; qb < qcount; qb += 32u) {
            const uint32_t q = qb + lane;
            const bool active = q < qcount;
            uint32_t kw[W];
#pragma unroll
            for (int k = 0; k < W; k++) kw[k] = 0u;
            uint32_t out = NONE;
            bool slow = false;
            if (active) {
                lds_key<W>(a_qk + q * (W * 4u), kw);
                const uint32_t h = hash_key<W>(kw);
                bool maybe = true;  // the Bloom filter (when present) rules out most reads that are in no table entry
                if (p.bloom_words) {
                    const uint32_t bm = bloom_mask(h);
                    maybe = (lds32_ro(a_bloom + (h >> p.bloom_shift) * 4u) & bm) == bm;
                }
                if (maybe) table_lookup<W>(p, kw, h, out);
                slow = out == NONE && !read_in_table_alphabet<W>(kw, p.last_pad);
            }
            uint32_t pending = __ballot_sync(0xFFFFFFFFu, slow);
            while (pending) {
                const int src_lane = __ffs(pending) - 1;
                pending &= pending - 1u;
                uint32_t bw[W];
#pragma unroll
                for (int k = 0; k < W; k++) bw[k] = __shfl_sync(0xFFFFFFFFu, kw[k], src_lane);
                const uint32_t o = warp_brute_one<W>(p, bw, lane);
                if ((int)lane == src_lane) out = o;
            }
            if (active) sts32(a_qk + q * (W * 4u), out);  // the entry's key has been consumed: reuse it for the result
        }
        __syncwarp();
        // every read that went through the queue had res == NONE; a queue result of NONE leaves it NONE
#pragma unroll
        for (int r = 0; r < R; r++)
            if (res[r] == NONE) res[r] = lds32(a_qk + qi[r] * (W * 4u));
        __syncwarp();  // the queue is reused by the next tile

        reinterpret_cast<uint4*>(results)[g] = make_uint4(res[0], res[1], res[2], res[3]);
#pragma unroll
        for (int r = 0; r < R; r++) {
            red_hist(a_hist, hshift, res[r]);
            none_count += (res[r] == NONE);
        }
    }

    // ---- tail: fewer than 128 reads, one per lane, first warp of the grid ----
    if (blockIdx.x == 0 && threadIdx.x < 32u) {
        for (uint64_t base = (uint64_t)n_tiles * PROBE2_TILE; base < src.n; base += 32u) {
            const uint64_t i = base + lane;
            const bool valid = i < src.n;
            uint32_t kw[W];
#pragma unroll
            for (int k = 0; k < W; k++) kw[k] = valid ? __ldg(src.packed + i * W + k) : 0u;
            uint32_t out = NONE;
            bool slow = false;
            if (valid) {
                const bool hit = table_lookup<W>(p, kw, hash_key<W>(kw), out);
                slow = !hit && !read_in_table_alphabet<W>(kw, p.last_pad);
            }
            uint32_t pending = __ballot_sync(0xFFFFFFFFu, slow);
            while (pending) {
                const int src_lane = __ffs(pending) - 1;
                pending &= pending - 1u;
                uint32_t bw[W];
#pragma unroll
                for (int k = 0; k < W; k++) bw[k] = __shfl_sync(0xFFFFFFFFu, kw[k], src_lane);
                const uint32_t o = warp_brute_one<W>(p, bw, lane);
                if ((int)lane == src_lane) out = o;
            }
            if (valid) {
                results[i] = out;
                red_hist(a_hist, hshift, out);
                none_count += (out == NONE);
            }
        }
    }
    // flush: unmatched from registers (one RED per warp), then the replicated bins
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) none_count += __shfl_xor_sync(0xFFFFFFFFu, none_count, off);
    if (lane == 0u && none_count) atomicAdd(&s_hist[p.S * hrep], none_count);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b <= p.S; b += blockDim.x) {
        uint32_t c = 0;
        for (uint32_t r = 0; r < hrep; r++) c += s_hist[b * hrep + r];
        if (c) atomicAdd(&p.counts[b], (unsigned long long)c);
    }
}

// ------------------------------------------------------------------------------------------------------
// k_probe3: the HBM-resident packed route when every pure-A/C/G/T memo-table entry fits in shared memory (L <= 16,
// e.g. cfg 2 / cfg 3).  No global-memory probe, no queue, no Bloom filter on the common path:
//   * the read's one-hot nibbles are compressed to a 32-bit key (2 bits per base, common.cuh acgt_key) and tested
//     for validity (every nibble exactly one bit) in ~10 integer ops;
//   * the key is looked up in an NP-ary cuckoo table of 4-byte QUOTIENT entries in shared memory: sub-table i is
//     indexed by the top sb_i bits of k * ck_mul(i) and the entry holds the low 32 - sb_i bits of that product above
//     a value code, so one LDS.32 and one multiply-add (entry - remainder) both verify the key exactly and deliver
//     the code; the NP candidates are merged with a min (at most one can match; an empty slot yields a code >= limit);
//   * a valid read that is not in the table is farther than max_mismatches from every barcode -> None;
//   * reads with any other symbol (no-calls, IUPAC, junk; ~3 % of real reads) are written as None and counted as
//     unmatched, parked in a small per-warp stash in shared memory (one ballot per pass: each lane parks its first
//     such read) and resolved a warp-full at a time through the global memo table (which also holds the N-containing
//     neighbours) or, outside its alphabet, the warp-cooperative scan; the result word and the counts are patched.
// Per-sample counts: lane-private (conflict-free) shared-memory histogram of packed 16-bit counters, one unconditional
// atomic per read, flushed to the global u64 table under a barrier every PROBE3_FLUSH_READS reads per lane.
// The kernel is bound by instruction issue and the SM's integer pipes, so adds and address computations on the hot
// path are written as multiply-adds with an operand read from the kernel parameters (IMAD, fma pipe) and only the
// logic ops, shifts, compares and selects stay on the alu pipe (both issue one warp instruction per 2 clocks).
// A warp owns tiles of 32 * R consecutive reads, R per lane (R * W / 4 x LDG.128 in, R / 4 x STG.128 out), with the
// next tile's words in flight in a second register buffer while this one is resolved.
// ------------------------------------------------------------------------------------------------------
struct Probe3Ctx {
    uint32_t base[3];  // shared-window address of sub-table i
    uint32_t a_hist;   // ... of this lane's histogram replica
    uint32_t a_stash;  // ... of this warp's stash: ck_stash_cap keys (8 bytes each), then as many read indices
};
// p.ck_one (= 1) and p.ck_four (= 4) come from the kernel parameters so that ptxas cannot fold x * one + c back into
// an alu-pipe add / LEA: those adds and scalings stay multiply-adds on the fma pipe.

FQ_D uint32_t imad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
FQ_HD uint32_t probe3_unmatched_bin(uint32_t S) { return S | 1u; }  // smallest odd index >= S (see hist_inc)
FQ_HD uint32_t probe3_hist_words(uint32_t S) { return ((probe3_unmatched_bin(S) + 1u) / 2u) * 32u; }  // the columns

// Compressed key + validity of one read: acgt_key (common.cuh), with the subtractions and the left shift on the fma pipe.
template <int W, bool PAD>
FQ_D uint32_t ck_key(const MatchParams& p, const uint32_t (&w)[W], bool& valid) {
    const uint32_t pad = PAD ? p.last_pad : 0u;
    const uint32_t x0 = (W == 1) ? (w[0] | pad) : w[0];
    uint32_t t = imad(x0, p.ck_one, 0xEEEEEEEFu) & (x0 | 0x88888888u);
    uint32_t k = (x0 ^ (x0 >> 1)) & 0x33333333u;
    if constexpr (W == 2) {
        const uint32_t x1 = w[W - 1] | pad;
        t |= imad(x1, p.ck_one, 0xEEEEEEEFu) & (x1 | 0x88888888u);
        k |= (x1 ^ imad(x1, p.ck_one, x1)) & 0xCCCCCCCCu;
    }
    valid = t == 0u;
    return k;
}

// One key through the shared-memory table: the result word (NONE for unmatched AND for reads that are not pure
// A/C/G/T) and the histogram bin (the unmatched bin when NONE).
template <int NP>
FQ_D uint32_t ck_find(const MatchParams& p, const Probe3Ctx& c, uint32_t k, bool valid, uint32_t& bin) {
    uint32_t u = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const uint32_t slot = (k * ck_mul(i)) >> p.ck_shift[i];
        const uint32_t e = lds32_ro(imad(slot, p.ck_four, c.base[i]));
        u = min(u, imad(k, p.ck_negmulb[i], e));  // entry - remainder: the value code iff this slot holds the key
    }
    const bool found = valid && u < p.ck_limit;
    // value code = idx << lb | best << nb | (next - next_min)
    const uint32_t idx = u >> p.ck_lb;
    const uint32_t low = ((u << p.ck_bsh) & p.ck_bmask8) | (u & p.ck_nmask);
    const uint32_t word = imad(idx, 65536u, imad(low, p.ck_one, p.ck_next_min));
    bin = found ? idx : probe3_unmatched_bin(p.S);
    return found ? word : NONE;
}

// memo-table / warp-cooperative resolution of one read per lane (`act` lanes only); all 32 lanes must call it
template <int W>
FQ_D uint32_t slow_resolve(const MatchParams& p, const uint32_t (&kw)[W], bool act, uint32_t lane) {
    uint32_t out = NONE;
    bool slow = false;
    if (act) {
        const bool hit = table_lookup<W>(p, kw, hash_key<W>(kw), out);
        slow = !hit && !read_in_table_alphabet<W>(kw, p.last_pad);
    }
    uint32_t pending = __ballot_sync(0xFFFFFFFFu, slow);
    while (pending) {
        const int src_lane = __ffs(pending) - 1;
        pending &= pending - 1u;
        uint32_t bw[W];
#pragma unroll
        for (int k = 0; k < W; k++) bw[k] = __shfl_sync(0xFFFFFFFFu, kw[k], src_lane);
        const uint32_t o = warp_brute_one<W>(p, bw, lane);
        if ((int)lane == src_lane) out = o;
    }
    return out;
}

// Packed histogram: 32 lane-private columns (bank = lane: no conflicts), two 16-bit counters per word; bin b is the
// (b & 1) half of word b >> 1.  Fields only ever grow: a parked read that turns out to match is counted in its
// sample's bin when it is resolved and tallied in one extra word (`fix`, right after the columns), which the flush
// takes off the unmatched total.  The CTA flushes the counters to the global u64 table every PROBE3_FLUSH_READS
// reads per lane, under a barrier, long before a 16-bit field can wrap: a field gains at most (32 warps) x (reads per
// lane, counted once and re-counted at most once) = 64 per read per lane.
constexpr uint32_t PROBE3_FLUSH_READS = 512;  // 512 * 64 = 32768 < 65536
FQ_D void hist_inc(const MatchParams& p, const Probe3Ctx& c, uint32_t bin) {
    const uint32_t addr = imad(bin & ~1u, p.ck_four * 16u, c.a_hist);  // (bin >> 1) * 128
    const uint32_t val = imad(bin & 1u, 65535u, p.ck_one);             // 1 or 65536
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(val) : "memory");
}
FQ_D void hist_unmatched_dec(const MatchParams& p, const Probe3Ctx& c) {  // cold path
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t addr = c.a_hist - lane * 4u + probe3_hist_words(p.S) * 4u;  // the `fix` word
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

// all threads of the CTA: add the packed counters to the global table and zero them
FQ_D void probe3_flush_hist(const MatchParams& p, uint32_t* s_hist) {
    __syncthreads();
    const uint32_t U = probe3_unmatched_bin(p.S);
    const uint32_t n_words = (U + 1u) / 2u;
    uint32_t* s_fix = s_hist + n_words * 32u;
    for (uint32_t wd = threadIdx.x; wd < n_words; wd += blockDim.x) {
        uint32_t lo = 0, hi = 0;
        for (uint32_t r = 0; r < 32u; r++) {
            const uint32_t col = (r + wd) & 31u;  // rotate the start column: no bank conflicts across the warp
            const uint32_t v = s_hist[wd * 32u + col];
            s_hist[wd * 32u + col] = 0u;
            lo += v & 0xFFFFu;
            hi += v >> 16;
        }
        const uint32_t b0 = 2u * wd, b1 = 2u * wd + 1u;
        if (b1 >= p.S) {  // the word of the unmatched bin: take off the parked reads that were matched after all
            hi -= *s_fix;
            *s_fix = 0u;
        }
        if (lo && b0 < p.S) atomicAdd(&p.counts[b0], (unsigned long long)lo);
        if (hi) atomicAdd(&p.counts[b1 < p.S ? b1 : p.S], (unsigned long long)hi);
    }
    __syncthreads();
}

constexpr uint32_t PROBE3_INLINE_LANES = 16;  // this many lanes with an odd read in one pass: not worth parking
constexpr uint32_t PROBE3_STASH_MIN = 16, PROBE3_STASH_MAX = 64;  // stash entries per warp (>= PROBE3_INLINE_LANES)

// Resolve one read per `act` lane now: patch its result word and the counts (it was written as None / unmatched).
template <int W>
__device__ __noinline__ void probe3_resolve_now(const MatchParams& p, const Probe3Ctx c, const uint32_t (&k2)[2], bool act,
                                                uint32_t idx, uint32_t* __restrict__ results, uint32_t lane) {
    uint32_t kw[W];
    kw[0] = k2[0];
    if constexpr (W == 2) kw[W - 1] = k2[1];
    const uint32_t out = slow_resolve<W>(p, kw, act, lane);
    if (act && out != NONE) {
        results[idx] = out;
        hist_inc(p, c, out >> 16);
        hist_unmatched_dec(p, c);
    }
}

// Resolve the first `cnt` stashed reads of the warp (`cnt` is warp-uniform): patch result words and counts.
template <int W>
__device__ __noinline__ void probe3_drain(const MatchParams& p, const Probe3Ctx c, uint32_t cnt,
                                          uint32_t* __restrict__ results, uint32_t lane) {
    __syncwarp();  // the parked entries, and the None words written for them, are ordered before this point
    for (uint32_t b = 0; b < cnt; b += 32u) {
        const uint32_t e = b + lane;
        const bool act = e < cnt;
        uint32_t kw[W];
        uint32_t idx = 0;
#pragma unroll
        for (int k = 0; k < W; k++) kw[k] = 0u;
        if (act) {
            uint32_t k2[2];
            lds_key<2>(c.a_stash + e * 8u, k2);
            kw[0] = k2[0];
            if constexpr (W == 2) kw[W - 1] = k2[1];
            idx = lds32(c.a_stash + p.ck_stash_cap * 8u + e * 4u);
        }
        const uint32_t out = slow_resolve<W>(p, kw, act, lane);
        if (act && out != NONE) {  // it was written as None and counted as unmatched when it was parked
            results[idx] = out;
            hist_inc(p, c, out >> 16);
            hist_unmatched_dec(p, c);
        }
    }
    __syncwarp();  // the stash is free again
}

// Ask L2 for a tile (32 * R reads) two tiles ahead of its use: every lane prefetches its own 32 bytes of it, no
// registers held.  Measured on B200 (cfg 3, same box): no prefetch 1.33 ms; one-lane TMA bulk prefetch
// (cp.async.bulk.prefetch.L2, ~14 instructions of uniform-datapath plumbing per tile) 1.27 ms; this per-lane
// prefetch.global.L2 (4 instructions) 1.21 ms.  2 and 4 tiles further ahead: no further gain.
template <int W, int R>
FQ_D void probe3_prefetch_l2(const uint32_t* __restrict__ packed, uint32_t tile, uint32_t lane) {
    const uint32_t* ptr = packed + (size_t)(tile * 32u + lane) * (R * W);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

template <int W, int R>
FQ_D void probe3_load(const uint32_t* __restrict__ packed, uint32_t tile, uint32_t lane, uint32_t (&w)[R][W]) {
    static_assert((W == 1 || W == 2) && (R == 4 || R == 8), "k_probe3 covers L <= 16, 4 or 8 reads per lane");
    constexpr int NV = R * W / 4;  // 16-byte vectors per lane
    const uint4* in = reinterpret_cast<const uint4*>(packed) + (size_t)(tile * 32u + lane) * NV;
    uint32_t flat[R * W];
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const uint4 q = __ldg(in + v);
        flat[4 * v + 0] = q.x; flat[4 * v + 1] = q.y; flat[4 * v + 2] = q.z; flat[4 * v + 3] = q.w;
    }
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int k = 0; k < W; k++) w[r][k] = flat[r * W + k];
}

template <int W, int NP, bool PAD, int R>
FQ_D void probe3_tile(const MatchParams& p, const Probe3Ctx& c, const uint32_t (&w)[R][W],
                      uint32_t* __restrict__ results, uint32_t tile, uint32_t lane, uint32_t& cnt) {
    uint32_t res[R], bin[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; r++) res[r] = ck_find<NP>(p, c, ck_key<W, PAD>(p, w[r], valid[r]), valid[r], bin[r]);
    const uint32_t g = tile * 32u + lane;  // this lane's group of R consecutive reads
    uint4* out4 = reinterpret_cast<uint4*>(results) + (size_t)g * (R / 4);
#pragma unroll
    for (int v = 0; v < R / 4; v++) out4[v] = make_uint4(res[4 * v], res[4 * v + 1], res[4 * v + 2], res[4 * v + 3]);
    // counts (unmatched reads, and for now the parked ones, go to the unmatched bin)
#pragma unroll
    for (int r = 0; r < R; r++) hist_inc(p, c, bin[r]);
    bool all_valid = true;
#pragma unroll
    for (int r = 0; r < R; r++) all_valid = all_valid && valid[r];
    if (!__any_sync(0xFFFFFFFFu, !all_valid)) return;
    // every pass takes each lane's first read that is not pure A/C/G/T (lanes rarely have two): a handful of them
    // are parked in the warp's stash, half a warp or more is resolved on the spot
    uint32_t bad = 0u;
#pragma unroll
    for (int r = 0; r < R; r++) bad |= valid[r] ? 0u : (1u << r);
    const uint32_t lane_lt = (1u << lane) - 1u;
    do {
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bad != 0u);
        const uint32_t n_new = (uint32_t)__popc(bal);
        const uint32_t r = (uint32_t)__ffs(bad) - 1u;
        uint32_t k2[2] = {w[0][0], w[0][W - 1]};
#pragma unroll
        for (int q = 1; q < R; q++) {
            k2[0] = (r == (uint32_t)q) ? w[q][0] : k2[0];
            k2[1] = (r == (uint32_t)q) ? w[q][W - 1] : k2[1];
        }
        if (n_new >= PROBE3_INLINE_LANES) {
            probe3_resolve_now<W>(p, c, k2, bad != 0u, g * R + r, results, lane);
        } else {
            if (cnt + n_new > p.ck_stash_cap) {  // no room: resolve what is parked first
                probe3_drain<W>(p, c, cnt, results, lane);
                cnt = 0u;
            }
            if (bad) {
                const uint32_t pos = cnt + (uint32_t)__popc(bal & lane_lt);
                sts_key<2>(c.a_stash + pos * 8u, k2);
                sts32(c.a_stash + p.ck_stash_cap * 8u + pos * 4u, g * R + r);
            }
            cnt += n_new;
        }
        bad &= bad - 1u;
    } while (__any_sync(0xFFFFFFFFu, bad != 0u));
}

template <int W, int NP, bool PAD, int R, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_probe3(const __grid_constant__ MatchParams p, const ReadSource src,
                                                       uint32_t* __restrict__ results) {
    extern __shared__ uint4 s_dyn[];
    // layout: cuckoo entries | packed histogram (32 lane columns) | per-warp stashes
    uint32_t* s_ck = reinterpret_cast<uint32_t*>(s_dyn);
    uint32_t* s_hist = s_ck + p.ck_words;
    const uint32_t n_hist_words = probe3_hist_words(p.S);
    {
        const uint4* g4 = reinterpret_cast<const uint4*>(p.ck_entries);  // ck_words is a multiple of 4
        uint4* s4 = reinterpret_cast<uint4*>(s_ck);
        for (uint32_t t = threadIdx.x; t < p.ck_words / 4u; t += blockDim.x) s4[t] = __ldg(g4 + t);
    }
    for (uint32_t t = threadIdx.x; t < n_hist_words + 4u; t += blockDim.x) s_hist[t] = 0u;  // columns + the fix word
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = blockDim.x >> 5, warp_in_cta = threadIdx.x >> 5;
    Probe3Ctx c;
#pragma unroll
    for (int i = 0; i < 3; i++) c.base[i] = smem_addr(s_ck) + p.ck_off[i < NP ? i : 0] * 4u;
    c.a_hist = smem_addr(s_hist) + lane * 4u;
    c.a_stash = smem_addr(s_hist + n_hist_words + 4u) + warp_in_cta * (p.ck_stash_cap * 12u);
    asm volatile("" : "+r"(c.base[0]), "+r"(c.base[1]), "+r"(c.base[2]), "+r"(c.a_hist));
    uint32_t cnt = 0;  // reads parked in the warp's stash (warp-uniform)

    constexpr uint32_t TILE = 32u * R;
    const uint32_t n_tiles = (uint32_t)(src.n / (uint64_t)TILE);
    const uint32_t stride = gridDim.x * n_warps;
    uint32_t tile = blockIdx.x * n_warps + warp_in_cta;
    // every warp of the CTA runs the same number of rounds (two tiles each) so that the flush barrier is legal
    const uint32_t first = blockIdx.x * n_warps;
    const uint32_t cta_tiles = first < n_tiles ? (n_tiles - first + stride - 1u) / stride : 0u;  // of the CTA's first warp
    const uint32_t rounds = (cta_tiles + 1u) / 2u;

    // two register buffers, alternating: the next tile's words are in flight while this one is resolved
    uint32_t wa[R][W], wb[R][W];
    if (tile < n_tiles) probe3_load<W, R>(src.packed, tile, lane, wa);
    for (uint32_t round = 0; round < rounds; round++) {
        if (tile < n_tiles) {
            uint32_t nt = tile + stride;
            if (nt < n_tiles) probe3_load<W, R>(src.packed, nt, lane, wb);
            if (nt + stride < n_tiles) probe3_prefetch_l2<W, R>(src.packed, nt + stride, lane);
            probe3_tile<W, NP, PAD, R>(p, c, wa, results, tile, lane, cnt);
            tile = nt;
            if (tile < n_tiles) {
                nt = tile + stride;
                if (nt < n_tiles) probe3_load<W, R>(src.packed, nt, lane, wa);
                if (nt + stride < n_tiles) probe3_prefetch_l2<W, R>(src.packed, nt + stride, lane);
                probe3_tile<W, NP, PAD, R>(p, c, wb, results, tile, lane, cnt);
                tile = nt;
            }
        }
        if ((round + 1u) % (PROBE3_FLUSH_READS / (2u * R)) == 0u) {
            probe3_drain<W>(p, c, cnt, results, lane);
            cnt = 0u;
            probe3_flush_hist(p, s_hist);
        }
    }
    probe3_drain<W>(p, c, cnt, results, lane);

    // ---- tail: fewer than a tile of reads, one per lane, first warp of the grid ----
    if (blockIdx.x == 0 && threadIdx.x < 32u) {
        for (uint64_t b0 = (uint64_t)n_tiles * TILE; b0 < src.n; b0 += 32u) {
            const uint64_t i = b0 + lane;
            const bool live = i < src.n;
            uint32_t w1[W];
#pragma unroll
            for (int k = 0; k < W; k++) w1[k] = live ? __ldg(src.packed + i * W + k) : 0u;
            bool valid;
            uint32_t bin;
            const uint32_t key = ck_key<W, PAD>(p, w1, valid);
            uint32_t out = ck_find<NP>(p, c, key, valid, bin);
            const uint32_t slow = slow_resolve<W>(p, w1, live && !valid, lane);
            if (live) {
                if (!valid) {
                    out = slow;
                    bin = (out == NONE) ? probe3_unmatched_bin(p.S) : (out >> 16);
                }
                results[i] = out;
                hist_inc(p, c, bin);
            }
        }
    }
    probe3_flush_hist(p, s_hist);
}

// ------------------------------------------------------------------------------------------------------
// k_probe4: the HBM-resident packed route for panels whose pure-A/C/G/T memo entries do not fit in shared memory
// (cfg 5: 6 144 IUPAC samples of 20 bases, 2.5 M entries).  Every pure-A/C/G/T read makes ONE random 32-byte load from
// a global FINGERPRINT table (layout and exactness argument: kernels.h) that is small enough (4 bytes per candidate
// string) to stay L2-resident while the reads stream through — table loads carry an L2 evict_last policy, the read /
// result stream evict_first — and a fingerprint match is verified against the barcode's ~expected nibble words in
// shared memory.  What bounds the kernel is the SM's L1 request path: a warp-wide load of 32 random lines costs 32 clocks
// whatever its width (tools/microbench_gather.cu: 284 G lookups/s chip-wide), so the design is one load instruction per
// read; the second bucket of the rare overflow chain is loaded only by the lanes that need it.
// Reads that are not pure A/C/G/T (no-calls, IUPAC codes, junk) and reads whose fingerprint match fails verification are
// written as None / counted as unmatched, parked in a per-warp stash and resolved a warp-full at a time through the
// exact memo table (or, outside its alphabet, the warp-cooperative scan), as in k_probe3.
// ------------------------------------------------------------------------------------------------------
constexpr int PROBE4_THREADS = 1024;  // the largest launch shape (shared-memory sizing)
constexpr int PROBE4_R = 4;

struct Probe4Ctx {
    uint32_t a_hist;   // shared-window address of this lane's histogram replica
    uint32_t a_stash;  // ... of this warp's stash: g4_stash_cap keys (W words each), then as many read indices
    uint32_t a_ne;     // ... of the panel's ~expected nibble words (probe4_ne_stride(W) words per barcode)
};

FQ_D void hist4_add(const MatchParams& p, const Probe4Ctx& c, uint32_t bin, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(imad(bin, p.g4_hist_rep * 4u, c.a_hist)), "r"(v) : "memory");
}

struct G4Bucket {
    uint32_t e[8];
};
FQ_D G4Bucket g4_load(const MatchParams& p, uint32_t bucket, uint64_t pol) {
    G4Bucket b;
    const uint32_t* q = p.g4_table + (size_t)bucket * 8;
    asm volatile("ld.global.nc.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(b.e[0]), "=r"(b.e[1]), "=r"(b.e[2]), "=r"(b.e[3]), "=r"(b.e[4]), "=r"(b.e[5]), "=r"(b.e[6]),
                   "=r"(b.e[7])
                 : "l"(q), "l"(pol));
    return b;
}
// the entry of the bucket whose fingerprint equals the read's, with the fingerprint field cleared (idx << cb | code),
// or a value >= 2^fp_shift when there is none (two matches: the smaller entry; the builder replays the same rule).
// `fp_hash` is the whole second hash: its top fp_bits are the fingerprint, `fp_mask` selects them (one LOP3 per entry).
FQ_D uint32_t g4_match(const G4Bucket& b, uint32_t fp_hash, uint32_t fp_mask) {
    uint32_t t[8];
#pragma unroll
    for (int j = 0; j < 8; j++) t[j] = b.e[j] ^ (fp_hash & fp_mask);
    return min(min(min(t[0], t[1]), min(t[2], t[3])), min(min(t[4], t[5]), min(t[6], t[7])));
}
template <int W>
FQ_D void lds_ne(uint32_t a, uint32_t (&ne)[W]) {  // one barcode's ~expected words
    if constexpr (W == 1) {
        ne[0] = lds32_ro(a);
    } else if constexpr (W == 2) {
        const uint2 v = lds64_ro(a);
        ne[0] = v.x;
        ne[W - 1] = v.y;
    } else {
        const uint4 v = lds128_ro(a);
        ne[0] = v.x;
        ne[1] = v.y;
        ne[2] = v.z;
        if constexpr (W == 4) ne[W - 1] = v.w;
    }
}

// Lookup rules (all derived constants of the entry layout come precomputed in MatchParams):
//   read with a symbol outside A/C/G/T/N   -> slow path, whatever the table says: such a read is no table key, and it can
//                                             pass the verification of another key's entry (junk / IUPAC symbols match
//                                             more than a base does), so exactness must not rest on the fingerprint
//   fingerprint match + verified          -> the entry's value (Some or None): the read IS that table key, or the builder's
//                                            replay would have re-seeded the hash
//   fingerprint match, not verified       -> slow path (another key's entry)
//   no match                              -> None: the table holds every A/C/G/T/N string within max_mm of a barcode
// `t` = g4_match of the bucket that ended the walk.  Returns the result word; `park` = the read takes the slow path.
template <int W>
FQ_D uint32_t g4_decide(const MatchParams& p, uint32_t t, const uint32_t (&w)[W], const uint32_t (&ne)[W], uint32_t idx,
                        uint32_t pad, bool& park) {
    const bool in_alphabet = acgtn_only<W>(w, pad);
    const bool hit = t < p.g4_lim;
    const bool ok = nibble_distance<W>(w, ne) <= p.max_mm;
    park = !in_alphabet || (hit && !ok);
    const uint32_t code = t & p.g4_cmask;
    const uint32_t best = code >> p.g4_nb, next = (code & p.g4_nmask) + p.g4_next_min;
    const bool some = in_alphabet && hit && ok && code < p.g4_code_none;
    return some ? ((idx << 16) | (best << 8) | next) : NONE;
}

// One read, start to end (the tail of a batch): all 32 lanes call it together.
template <int W>
FQ_D uint32_t g4_lookup_one(const MatchParams& p, const Probe4Ctx& c, const uint32_t (&w)[W], uint32_t pad, uint64_t pol,
                            bool& park) {
    uint32_t bucket, fph;
    g4_hashes<W>(w, p.g4_seed, p.g4_buckets, bucket, fph);
    G4Bucket b = g4_load(p, bucket, pol);
    uint32_t t = g4_match(b, fph, p.g4_fp_mask);
    bool more = t >= p.g4_lim && b.e[7] != 0xFFFFFFFFu;
    while (__any_sync(0xFFFFFFFFu, more)) {
        if (more) {
            bucket = (bucket + 1u == p.g4_buckets) ? 0u : bucket + 1u;
            b = g4_load(p, bucket, pol);
            t = g4_match(b, fph, p.g4_fp_mask);
            more = t >= p.g4_lim && b.e[7] != 0xFFFFFFFFu;
        }
    }
    const uint32_t idx = min(t >> p.g4_cb, p.S - 1u);  // (the clamp only matters for the lanes without a hit)
    uint32_t ne[W];
    lds_ne<W>(imad(idx, probe4_ne_stride(W) * 4u, c.a_ne), ne);
    return g4_decide<W>(p, t, w, ne, idx, pad, park);
}

// (the extra template parameters only make the copy private to one kernel: ptxas 12.9 crashes on a shared one)
template <int W, bool PAD, int THREADS, int IF>
__device__ __noinline__ void probe4_drain(const MatchParams& p, const Probe4Ctx c, uint32_t cnt,
                                          uint32_t* __restrict__ results, uint32_t lane) {
    __syncwarp();
    const uint32_t a_idx = c.a_stash + p.g4_stash_cap * (W * 4u);
    for (uint32_t b = 0; b < cnt; b += 32u) {
        const uint32_t e = b + lane;
        const bool act = e < cnt;
        uint32_t kw[W];
        uint32_t idx = 0;
#pragma unroll
        for (int k = 0; k < W; k++) kw[k] = act ? lds32(c.a_stash + (e * W + k) * 4u) : 0u;
        if (act) idx = lds32(a_idx + e * 4u);
        const uint32_t out = slow_resolve<W>(p, kw, act, lane);
        if (act && out != NONE) {  // it was written as None and counted as unmatched when it was parked
            results[idx] = out;
            hist4_add(p, c, out >> 16, 1u);
            hist4_add(p, c, p.S, 0xFFFFFFFFu);
        }
    }
    __syncwarp();
}

template <int W, bool PAD, int THREADS, int IF>  // IF = home buckets in flight per lane (2 or 4)
__global__ void __launch_bounds__(THREADS, 1) k_probe4(const __grid_constant__ MatchParams p, const ReadSource src,
                                                      uint32_t* __restrict__ results) {
    constexpr int R = PROBE4_R;
    static_assert(IF == 2 || IF == 4, "two or four lookups in flight");
    constexpr uint32_t NES = W == 3 ? 4 : W;  // probe4_ne_stride
    extern __shared__ uint4 s_dyn[];
    // layout: ~expected words of the panel | histogram replicas | per-warp stashes
    uint32_t* s_ne = reinterpret_cast<uint32_t*>(s_dyn);
    uint32_t* s_hist = s_ne + (size_t)p.S * NES;
    const uint32_t hrep = p.g4_hist_rep;
    for (uint32_t t = threadIdx.x; t < p.S * NES; t += blockDim.x) {
        const uint32_t j = t / NES, k = t % NES;
        s_ne[t] = k < (uint32_t)W ? __ldg(p.not_exp + (size_t)j * W + k) : 0u;
    }
    for (uint32_t t = threadIdx.x; t < (p.S + 1u) * hrep; t += blockDim.x) s_hist[t] = 0u;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = blockDim.x >> 5, warp_in_cta = threadIdx.x >> 5;
    Probe4Ctx c;
    c.a_ne = smem_addr(s_ne);
    c.a_hist = smem_addr(s_hist) + (lane & (hrep - 1u)) * 4u;
    c.a_stash = smem_addr(s_hist + (p.S + 1u) * hrep) + warp_in_cta * (p.g4_stash_cap * (W * 4u + 4u));
    const uint32_t pad = PAD ? p.last_pad : 0u;
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint32_t cnt = 0;  // reads parked in the warp's stash (warp-uniform)
    const uint64_t pol_keep = (p.g4_flags & 1u) ? l2_policy_keep() : l2_policy_normal();
    const uint64_t pol_stream = (p.g4_flags & 2u) ? l2_policy_stream() : l2_policy_normal();

    constexpr uint32_t TILE = 32u * R;
    constexpr int NV = R * W / 4;  // 16-byte vectors per lane per tile
    const uint32_t n_tiles = (uint32_t)(src.n / (uint64_t)TILE);
    const uint32_t stride = gridDim.x * n_warps;
    // ---- the three stages of a tile ----
    auto load_words = [&](uint32_t tile, uint32_t (&flat)[R * W]) {
        const uint4* in = reinterpret_cast<const uint4*>(src.packed) + (size_t)(tile * 32u + lane) * NV;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const uint4 q = ldg_hint(in + v, pol_stream);
            flat[4 * v + 0] = q.x; flat[4 * v + 1] = q.y; flat[4 * v + 2] = q.z; flat[4 * v + 3] = q.w;
        }
    };
    // hashes + home-bucket loads of reads [h0, h0 + N) of the lane
    auto issue = [&](const uint32_t (&flat)[R * W], uint32_t (&bucket)[R], uint32_t (&fph)[R], G4Bucket* e, int h0, int n) {
#pragma unroll
        for (int q = 0; q < R; q++) {
            if (q >= h0 && q < h0 + n) {
                uint32_t kw[W];
#pragma unroll
                for (int k = 0; k < W; k++) kw[k] = flat[q * W + k];
                g4_hashes<W>(kw, p.g4_seed, p.g4_buckets, bucket[q], fph[q]);
                e[q - h0] = g4_load(p, bucket[q], pol_keep);
            }
        }
    };
    auto match = [&](const G4Bucket* e, const uint32_t (&fph)[R], uint32_t (&t)[R], uint32_t& more, int h0, int n) {
#pragma unroll
        for (int q = 0; q < R; q++) {
            if (q >= h0 && q < h0 + n) {
                t[q] = g4_match(e[q - h0], fph[q], p.g4_fp_mask);
                more |= (t[q] >= p.g4_lim && e[q - h0].e[7] != 0xFFFFFFFFu) ? (1u << q) : 0u;
            }
        }
    };
    // overflow walk, verification, decision, result store, counts, parking
    auto finish = [&](uint32_t tile, const uint32_t (&flat)[R * W], uint32_t (&bucket)[R], const uint32_t (&fph)[R],
                      uint32_t (&t)[R], uint32_t more) {
        // the overflow walk (rare at the table's load factor): one more bucket per round for each lane's first read
        // that needs it, until no lane needs any
        while (__any_sync(0xFFFFFFFFu, more != 0u)) {
            if (more) {
                const uint32_t r = (uint32_t)__ffs(more) - 1u;
                uint32_t bk = bucket[0], fh = fph[0];
#pragma unroll
                for (int q = 1; q < R; q++) {
                    bk = (r == (uint32_t)q) ? bucket[q] : bk;
                    fh = (r == (uint32_t)q) ? fph[q] : fh;
                }
                bk = (bk + 1u == p.g4_buckets) ? 0u : bk + 1u;
                const G4Bucket e = g4_load(p, bk, pol_keep);
                const uint32_t tt = g4_match(e, fh, p.g4_fp_mask);
#pragma unroll
                for (int q = 0; q < R; q++) {
                    bucket[q] = (r == (uint32_t)q) ? bk : bucket[q];
                    t[q] = (r == (uint32_t)q) ? tt : t[q];
                }
                if (!(tt >= p.g4_lim && e.e[7] != 0xFFFFFFFFu)) more &= more - 1u;
            }
        }
        // the matched entries' barcodes from shared memory (all R loads in flight), then verification and decision
        uint32_t idx[R], ne[R][W];
#pragma unroll
        for (int r = 0; r < R; r++) {
            idx[r] = min(t[r] >> p.g4_cb, p.S - 1u);  // (the clamp only matters for the lanes without a hit)
            lds_ne<W>(imad(idx[r], NES * 4u, c.a_ne), ne[r]);
        }
        uint32_t res[R];
        uint32_t bad = 0u;  // bit r: read r takes the slow path
#pragma unroll
        for (int r = 0; r < R; r++) {
            uint32_t kw[W];
#pragma unroll
            for (int k = 0; k < W; k++) kw[k] = flat[r * W + k];
            bool park;
            res[r] = g4_decide<W>(p, t[r], kw, ne[r], idx[r], pad, park);
            hist4_add(p, c, res[r] != NONE ? idx[r] : p.S, 1u);
            bad |= park ? (1u << r) : 0u;
        }
        const uint32_t g = tile * 32u + lane;  // this lane's group of R consecutive reads
        uint4* out4 = reinterpret_cast<uint4*>(results) + (size_t)g * (R / 4);
#pragma unroll
        for (int v = 0; v < R / 4; v++)
            stg_hint(out4 + v, make_uint4(res[4 * v], res[4 * v + 1], res[4 * v + 2], res[4 * v + 3]), pol_stream);
        // park each lane's slow-path reads (symbols outside A/C/G/T/N, fingerprint collisions: rare): one pass per read
        // slot that has any
        uint32_t bal;
        while ((bal = __ballot_sync(0xFFFFFFFFu, bad != 0u)) != 0u) {
            const uint32_t n_new = (uint32_t)__popc(bal);
            if (cnt + n_new > p.g4_stash_cap) {  // no room (g4_stash_cap >= 32): resolve what is parked first
                probe4_drain<W, PAD, THREADS, IF>(p, c, cnt, results, lane);
                cnt = 0u;
            }
            if (bad) {
                const uint32_t r = (uint32_t)__ffs(bad) - 1u;
                const uint32_t pos = cnt + (uint32_t)__popc(bal & lane_lt);
#pragma unroll
                for (int k = 0; k < W; k++) {
                    uint32_t wk = flat[k];
#pragma unroll
                    for (int q = 1; q < R; q++) wk = (r == (uint32_t)q) ? flat[q * W + k] : wk;
                    sts32(c.a_stash + (pos * W + k) * 4u, wk);
                }
                sts32(c.a_stash + p.g4_stash_cap * (W * 4u) + pos * 4u, g * R + r);
                bad &= bad - 1u;
            }
            cnt += n_new;
        }
    };

    const uint32_t tile0 = blockIdx.x * n_warps + warp_in_cta;
    if constexpr (THREADS <= 512) {
        // Software pipeline across tiles (128 registers per thread): the NEXT tile's home buckets are requested before this
        // tile is verified and stored, the words of the tile after that before the next match — a warp has R lookups in
        // flight nearly all the time, which is what the L1 request path needs to stay busy with only 16 warps per SM.
        static_assert(IF == R, "the pipelined form keeps all R lookups of a tile in flight");
        uint32_t wa[R * W], wb[R * W], bucket_a[R], fph_a[R], bucket_b[R], fph_b[R];
        G4Bucket e[R];
        if (tile0 < n_tiles) {
            load_words(tile0, wa);
            issue(wa, bucket_a, fph_a, e, 0, R);
            if (tile0 + stride < n_tiles) load_words(tile0 + stride, wb);
        }
        for (uint32_t tile = tile0; tile < n_tiles; tile += stride) {
            uint32_t t[R], more = 0u;
            match(e, fph_a, t, more, 0, R);
            const uint32_t next = tile + stride;
            if (next < n_tiles) issue(wb, bucket_b, fph_b, e, 0, R);
            finish(tile, wa, bucket_a, fph_a, t, more);
#pragma unroll
            for (int k = 0; k < R * W; k++) wa[k] = wb[k];
#pragma unroll
            for (int q = 0; q < R; q++) {
                bucket_a[q] = bucket_b[q];
                fph_a[q] = fph_b[q];
            }
            if (next + stride < n_tiles) load_words(next + stride, wb);
        }
    } else {
        // Plain grid-stride loop: phases over the R reads of the lane, so that a warp waits once per kind of memory
        // round trip (words -> home buckets -> barcode words), not once per read.
        for (uint32_t tile = tile0; tile < n_tiles; tile += stride) {
            uint32_t flat[R * W];
            load_words(tile, flat);
            if ((p.g4_flags & 4u) && lane < (uint32_t)(NV * 4) && tile + stride < n_tiles) {  // ask L2 for the next tile
                const uint32_t* nxt = src.packed + (size_t)(tile + stride) * (TILE * W) + lane * 32u;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
            }
            uint32_t bucket[R], fph[R], t[R];
            uint32_t more = 0u;  // bit r: read r's bucket is full and has no match: walk on
#pragma unroll
            for (int h = 0; h < R; h += IF) {
                G4Bucket e[IF];
                issue(flat, bucket, fph, e, h, IF);
                match(e, fph, t, more, h, IF);
            }
            finish(tile, flat, bucket, fph, t, more);
        }
    }
    probe4_drain<W, PAD, THREADS, IF>(p, c, cnt, results, lane);

    // ---- tail: fewer than a tile of reads, one per lane, first warp of the grid ----
    if (blockIdx.x == 0 && threadIdx.x < 32u) {
        for (uint64_t b0 = (uint64_t)n_tiles * TILE; b0 < src.n; b0 += 32u) {
            const uint64_t i = b0 + lane;
            const bool live = i < src.n;
            uint32_t w1[W];
#pragma unroll
            for (int k = 0; k < W; k++) w1[k] = live ? __ldg(src.packed + i * W + k) : 0u;
            bool park;
            uint32_t out = g4_lookup_one<W>(p, c, w1, pad, pol_keep, park);
            const uint32_t slow = slow_resolve<W>(p, w1, live && park, lane);
            if (live) {
                if (park) out = slow;
                results[i] = out;
                hist4_add(p, c, out == NONE ? p.S : (out >> 16), 1u);
            }
        }
    }
    // ---- flush the replicated bins (bin S = unmatched) ----
    __syncthreads();
    for (uint32_t b = threadIdx.x; b <= p.S; b += blockDim.x) {
        uint32_t v = 0;
        for (uint32_t r = 0; r < hrep; r++) v += s_hist[b * hrep + r];
        if (v) atomicAdd(&p.counts[b], (unsigned long long)v);
    }
}

// ------------------------------------------------------------------------------------------------------
// k_probe5: the HBM-resident packed route for L <= 16 panels whose full neighbourhood does not fit in shared memory but
// whose EXACT entries do (cfg 4: 1 536 samples at 2 mismatches — 1.7 M pure-A/C/G/T entries, 1 536 of them exact).
// k_probe3's front end on a shared-memory cuckoo table that holds only the best-distance-0 entries (two LDS.32 per read, the
// ~80 % of a real stream that is an exact barcode ends there), and for the rest of the tile — pure reads that are not
// exact, no-call reads, anything else — k_probe2's trick: compact them into a per-warp queue and send them, ~26 busy
// lanes at a time, through k_probe4's L2-resident fingerprint table (ONE 256-bit load per read, verified against the
// barcode's ~expected words in shared memory; exactness argument in kernels.h).  What the fingerprint table cannot
// answer exactly (symbols outside A/C/G/T/N, fingerprint collisions) goes through the memo table / warp-cooperative scan
// right there.  Counts: k_probe3's lane-private packed histogram, one atomic per read once its result is known.
// ------------------------------------------------------------------------------------------------------
constexpr int PROBE5_THREADS = 1024;
constexpr int PROBE5_R = 4;

// the exact slow path as a real call (keeps its code out of the queue loop); the extra template parameters only make the
// copy private to one kernel (ptxas 12.9 crashes on shared noinline copies)
template <int W, int NP, bool PAD>
__device__ __noinline__ uint32_t probe5_slow(const MatchParams& p, uint32_t k0, uint32_t k1, bool act, uint32_t lane) {
    uint32_t kw[W];
    kw[0] = k0;
    if constexpr (W == 2) kw[W - 1] = k1;
    return slow_resolve<W>(p, kw, act, lane);
}

template <int W, int NP, bool PAD>
__global__ void __launch_bounds__(PROBE5_THREADS, 1) k_probe5(const __grid_constant__ MatchParams p, const ReadSource src,
                                                             uint32_t* __restrict__ results) {
    constexpr int R = PROBE5_R;
    constexpr uint32_t TILE = 32u * R;
    extern __shared__ uint4 s_dyn[];
    // layout: cuckoo entries (exact keys) | histogram replicas (32-bit counters: a CTA sees < 2^32 reads) | ~expected words
    //         | warp queues
    uint32_t* s_ck = reinterpret_cast<uint32_t*>(s_dyn);
    uint32_t* s_hist = s_ck + p.ck_words;
    const uint32_t hrep = p.g4_hist_rep;
    const uint32_t n_hist_words = (p.S + 1u) * hrep;
    uint32_t* s_ne = s_hist + n_hist_words;
    uint32_t* s_queue = s_ne + (size_t)p.S * W;
    {
        const uint4* g4 = reinterpret_cast<const uint4*>(p.ck_entries);  // ck_words is a multiple of 4
        uint4* s4 = reinterpret_cast<uint4*>(s_ck);
        for (uint32_t t = threadIdx.x; t < p.ck_words / 4u; t += blockDim.x) s4[t] = __ldg(g4 + t);
    }
    for (uint32_t t = threadIdx.x; t < n_hist_words; t += blockDim.x) s_hist[t] = 0u;
    for (uint32_t t = threadIdx.x; t < p.S * W; t += blockDim.x) s_ne[t] = __ldg(p.not_exp + t);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u, lane_lt = (1u << lane) - 1u;
    const uint32_t n_warps = blockDim.x >> 5, warp_in_cta = threadIdx.x >> 5;
    Probe3Ctx c;
#pragma unroll
    for (int i = 0; i < 3; i++) c.base[i] = smem_addr(s_ck) + p.ck_off[i < NP ? i : 0] * 4u;
    c.a_hist = smem_addr(s_hist) + (lane & (hrep - 1u)) * 4u;
    c.a_stash = 0u;
    auto count = [&](uint32_t bin) {  // bin S = unmatched
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(imad(bin, hrep * 4u, c.a_hist)) : "memory");
    };
    const uint32_t a_ne = smem_addr(s_ne);
    const uint32_t a_q = smem_addr(s_queue) + warp_in_cta * (TILE * W * 4u);  // entry q: key in, result out (word 0)
    const uint32_t pad = PAD ? p.last_pad : 0u;
    const uint64_t pol_keep = l2_policy_keep();

    const uint32_t n_tiles = (uint32_t)(src.n / (uint64_t)TILE);
    const uint32_t stride = gridDim.x * n_warps;
    uint32_t tile = blockIdx.x * n_warps + warp_in_cta;

    // the fingerprint-table lookup of one read per `act` lane (all 32 lanes call it together)
    auto lookup = [&](const uint32_t (&kw)[W], bool act) -> uint32_t {
        uint32_t bucket, fph;
        g4_hashes<W>(kw, p.g4_seed, p.g4_buckets, bucket, fph);
        G4Bucket b;
#pragma unroll
        for (int j = 0; j < 8; j++) b.e[j] = 0xFFFFFFFFu;
        if (act) b = g4_load(p, bucket, pol_keep);
        uint32_t t = g4_match(b, fph, p.g4_fp_mask);
        bool more = act && t >= p.g4_lim && b.e[7] != 0xFFFFFFFFu;
        while (__any_sync(0xFFFFFFFFu, more)) {
            if (more) {
                bucket = (bucket + 1u == p.g4_buckets) ? 0u : bucket + 1u;
                b = g4_load(p, bucket, pol_keep);
                t = g4_match(b, fph, p.g4_fp_mask);
                more = t >= p.g4_lim && b.e[7] != 0xFFFFFFFFu;
            }
        }
        const uint32_t idx = min(t >> p.g4_cb, p.S - 1u);
        uint32_t ne[W];
        lds_ne<W>(imad(idx, W * 4u, a_ne), ne);
        bool park;
        uint32_t out = g4_decide<W>(p, t, kw, ne, idx, pad, park);
        park = park && act;
        if (__any_sync(0xFFFFFFFFu, park)) {  // rare: symbols outside A/C/G/T/N, fingerprint collisions
            const uint32_t slow = probe5_slow<W, NP, PAD>(p, kw[0], kw[W - 1], park, lane);  // memo table, else the scan
            if (park) out = slow;
        }
        return act ? out : NONE;
    };

    auto do_tile = [&](const uint32_t (&w)[R][W], uint32_t t_idx) {
        uint32_t res[R], bin[R];
        bool valid[R];
#pragma unroll
        for (int r = 0; r < R; r++) res[r] = ck_find<NP>(p, c, ck_key<W, PAD>(p, w[r], valid[r]), valid[r], bin[r]);
        // everything the exact table did not answer goes to the warp's queue
        uint32_t qcount = 0;  // warp-uniform
        uint32_t qi[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const bool pend = res[r] == NONE;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pend);
            qi[r] = qcount + (uint32_t)__popc(bal & lane_lt);
            if (pend) sts_key<W>(a_q + qi[r] * (W * 4u), w[r]);
            qcount += (uint32_t)__popc(bal);
        }
        __syncwarp();
        for (uint32_t qb = 0; qb < qcount; qb += 32u) {
            const uint32_t q = qb + lane;
            const bool act = q < qcount;
            uint32_t kw[W];
#pragma unroll
            for (int k = 0; k < W; k++) kw[k] = 0u;
            if (act) lds_key<W>(a_q + q * (W * 4u), kw);
            const uint32_t out = lookup(kw, act);
            if (act) sts32(a_q + q * (W * 4u), out);  // the entry's key has been consumed: reuse it for the result
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (res[r] == NONE) {
                res[r] = lds32(a_q + qi[r] * (W * 4u));
                bin[r] = res[r] == NONE ? p.S : (res[r] >> 16);
            }
        }
        __syncwarp();  // the queue is reused by the next tile
        const uint32_t g = t_idx * 32u + lane;
        reinterpret_cast<uint4*>(results)[g] = make_uint4(res[0], res[1], res[2], res[3]);
#pragma unroll
        for (int r = 0; r < R; r++) count(min(bin[r], p.S));  // (ck_find's own unmatched bin is S | 1)
    };

    uint32_t wa[R][W], wb[R][W];
    if (tile < n_tiles) probe3_load<W, R>(src.packed, tile, lane, wa);
    while (tile < n_tiles) {
        {
            uint32_t nt = tile + stride;
            if (nt < n_tiles) probe3_load<W, R>(src.packed, nt, lane, wb);
            if (nt + stride < n_tiles) probe3_prefetch_l2<W, R>(src.packed, nt + stride, lane);
            do_tile(wa, tile);
            tile = nt;
            if (tile < n_tiles) {
                nt = tile + stride;
                if (nt < n_tiles) probe3_load<W, R>(src.packed, nt, lane, wa);
                if (nt + stride < n_tiles) probe3_prefetch_l2<W, R>(src.packed, nt + stride, lane);
                do_tile(wb, tile);
                tile = nt;
            }
        }
    }
    // ---- tail: fewer than a tile of reads, one per lane, first warp of the grid ----
    if (blockIdx.x == 0 && threadIdx.x < 32u) {
        for (uint64_t b0 = (uint64_t)n_tiles * TILE; b0 < src.n; b0 += 32u) {
            const uint64_t i = b0 + lane;
            const bool live = i < src.n;
            uint32_t w1[W];
#pragma unroll
            for (int k = 0; k < W; k++) w1[k] = live ? __ldg(src.packed + i * W + k) : 0u;
            const uint32_t out = lookup(w1, live);
            if (live) {
                results[i] = out;
                count(out == NONE ? p.S : (out >> 16));
            }
        }
    }
    // ---- flush the replicated bins (bin S = unmatched) ----
    __syncthreads();
    for (uint32_t b = threadIdx.x; b <= p.S; b += blockDim.x) {
        uint32_t v = 0;
        for (uint32_t r = 0; r < hrep; r++) v += s_hist[b * hrep + r];
        if (v) atomicAdd(&p.counts[b], (unsigned long long)v);
    }
}

// ------------------------------------------------------------------------------------------------------
// k_pack: encode() for a batch (mod.rs:49-61), any L
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const uint8_t* __restrict__ ascii, uint64_t n, uint32_t L,
                                              uint64_t stride, uint32_t* __restrict__ packed) {
    __shared__ uint8_t s_lut[256];
    init_lut(s_lut);
    __syncthreads();
    const uint32_t W = words_for_len(L);
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += total) {
        const uint8_t* row = ascii + i * stride;
        for (uint32_t wi = 0; wi < W; wi++) {
            uint32_t acc = 0u;
            for (uint32_t b = 0; b < 8u; b++) {
                const uint32_t k = wi * 8u + b;
                if (k < L) acc |= (uint32_t)s_lut[__ldg(row + k)] << (4u * b);
            }
            packed[i * W + wi] = acc;
        }
    }
}

// k_pack_segments: sample_barcode_sequence() (demux.rs:121-123) + encode() (mod.rs:49-61) in one pass: the B segments
// of a read are gathered from their sources in order and packed straight into the BitEnc word layout.
__global__ void __launch_bounds__(256) k_pack_segments(const SegmentSource seg, uint64_t n, uint32_t L,
                                                       uint32_t* __restrict__ packed) {
    __shared__ uint8_t s_lut[256];
    init_lut(s_lut);
    __syncthreads();
    const uint32_t W = words_for_len(L);
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += total) {
        uint32_t acc = 0u, pos = 0u;  // pos = symbols emitted so far
        for (uint32_t s = 0; s < seg.n_segments; s++) {
            const uint8_t* src = seg.base[s] + i * seg.stride[s] + seg.offset[s];
            for (uint32_t k = 0; k < seg.length[s]; k++, pos++) {
                acc |= (uint32_t)s_lut[__ldg(src + k)] << (4u * (pos & 7u));
                if ((pos & 7u) == 7u) {
                    packed[i * W + (pos >> 3)] = acc;
                    acc = 0u;
                }
            }
        }
        if (pos & 7u) packed[i * W + (pos >> 3)] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------
static uint32_t hist_bytes(const MatchParams& p) {
    return (p.S + 1u <= HIST_SMEM_BINS) ? (p.S + 1u) * 4u : 0u;
}

template <typename K>
static int grid_for(K kernel, int threads, size_t smem, const LaunchGeometry& g, uint64_t work_items) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
    uint64_t want = (work_items + threads - 1) / threads;
    uint64_t cap = (uint64_t)g.sm_count * occ;  // persistent: one wave of resident CTAs, grid-stride inside
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <int W, bool ASCII, int PK>
static cudaError_t launch_brute_wp(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                   const LaunchGeometry& g, cudaStream_t stream) {
    const uint32_t n_pairs = (p.S + 1u) / 2u;
    const size_t panel_bytes = (size_t)(PK == 2 ? n_pairs : 2u * n_pairs) * sizeof(uint4);
    const size_t hb = hist_bytes(p);
    const bool panel_smem = panel_bytes + hb + 1024 <= (size_t)g.max_smem_optin;
    if (panel_smem) {
        auto k = k_brute<W, ASCII, true, PK>;
        const size_t smem = panel_bytes + hb;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int grid = grid_for(k, BRUTE_THREADS, smem, g, src.n);
        k<<<grid, BRUTE_THREADS, smem, stream>>>(p, src, d_results);
    } else {
        auto k = k_brute<W, ASCII, false, PK>;
        const int grid = grid_for(k, BRUTE_THREADS, hb, g, src.n);
        k<<<grid, BRUTE_THREADS, hb, stream>>>(p, src, d_results);
    }
    count_launch();
    return cudaGetLastError();
}

template <int W, bool ASCII>
static cudaError_t launch_brute_w(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                  const LaunchGeometry& g, cudaStream_t stream) {
    if constexpr (W <= 2) {
        return launch_brute_wp<W, ASCII, 2>(p, src, d_results, g, stream);  // L <= 16: two barcodes per plane word
    } else {
        return launch_brute_wp<W, ASCII, 1>(p, src, d_results, g, stream);
    }
}

template <int WMAX>
static cudaError_t launch_brute_long_w(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                       const LaunchGeometry& g, cudaStream_t stream) {
    const size_t hb = hist_bytes(p);
    const size_t panel_bytes = (size_t)p.S * ((p.W + 3u) / 4u) * sizeof(uint4);
    if (panel_bytes + hb + 1024 <= (size_t)g.max_smem_optin) {
        auto k = k_brute_long<WMAX, true>;
        const size_t smem = panel_bytes + hb;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int grid = grid_for(k, 256, smem, g, src.n);
        k<<<grid, 256, smem, stream>>>(p, src, d_results);
    } else {
        auto k = k_brute_long<WMAX, false>;
        const int grid = grid_for(k, 256, hb, g, src.n);
        k<<<grid, 256, hb, stream>>>(p, src, d_results);
    }
    count_launch();
    return cudaGetLastError();
}

template <int W, bool ASCII, int GBITS>
static cudaError_t launch_brute_sliced_g(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                         const LaunchGeometry& g, cudaStream_t stream, uint32_t chunk_groups) {
    const uint32_t G = (p.S + 31u) / 32u;
    const size_t smem = (size_t)chunk_groups * Sliced<W>::GROUP_BYTES + hist_bytes(p);
    auto k0 = k_brute_sliced<W, ASCII, GBITS, true>;
    auto k1 = k_brute_sliced<W, ASCII, 1, false>;  // later chunks: no slot state
    cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (uint32_t chunk = 0; chunk * chunk_groups < G; chunk++) {  // stream order carries the interim state in d_results
        if (chunk == 0u)
            k0<<<grid_for(k0, sliced_threads(W), smem, g, src.n), sliced_threads(W), smem, stream>>>(p, src, d_results, chunk_groups, chunk);
        else
            k1<<<grid_for(k1, sliced_threads(W), smem, g, src.n), sliced_threads(W), smem, stream>>>(p, src, d_results, chunk_groups, chunk);
        count_launch();
    }
    return cudaGetLastError();
}
template <int W, bool ASCII>
static cudaError_t launch_brute_sliced(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                       const LaunchGeometry& g, cudaStream_t stream) {
    const uint32_t G = (p.S + 31u) / 32u;
    const uint32_t chunk_groups = std::max<uint32_t>(1u, std::min<uint32_t>(G, SLICED_CHUNK_BYTES / Sliced<W>::GROUP_BYTES));
    if (chunk_groups <= 16u) return launch_brute_sliced_g<W, ASCII, 4>(p, src, d_results, g, stream, chunk_groups);
    return launch_brute_sliced_g<W, ASCII, 8>(p, src, d_results, g, stream, chunk_groups);  // a chunk is <= 192 groups
}

static bool brute_v1() {  // FQTK_B200_BRUTE_V1=1 (A/B timing): the barcode-pair kernel of round 1 instead of the bit-sliced one
    static const bool v = [] {
        const char* e = getenv("FQTK_B200_BRUTE_V1");
        return e && atoi(e) != 0;
    }();
    return v;
}

cudaError_t launch_brute(const MatchParams& p, const ReadSource& src, uint32_t* d_results, const LaunchGeometry& g,
                         cudaStream_t stream) {
    if (src.n == 0) return cudaSuccess;
    const bool ascii = src.ascii != nullptr;
    if (p.sliced != nullptr && p.W <= (uint32_t)MAX_FAST_WORDS && !brute_v1()) {
        switch (p.W) {
            case 1: return ascii ? launch_brute_sliced<1, true>(p, src, d_results, g, stream) : launch_brute_sliced<1, false>(p, src, d_results, g, stream);
            case 2: return ascii ? launch_brute_sliced<2, true>(p, src, d_results, g, stream) : launch_brute_sliced<2, false>(p, src, d_results, g, stream);
            case 3: return ascii ? launch_brute_sliced<3, true>(p, src, d_results, g, stream) : launch_brute_sliced<3, false>(p, src, d_results, g, stream);
            default: return ascii ? launch_brute_sliced<4, true>(p, src, d_results, g, stream) : launch_brute_sliced<4, false>(p, src, d_results, g, stream);
        }
    }
    if (p.W > (uint32_t)MAX_FAST_WORDS) {
        if (ascii) return cudaErrorInvalidValue;  // caller packs first
        return p.W <= 8u ? launch_brute_long_w<8>(p, src, d_results, g, stream)
                         : (p.W <= 16u ? launch_brute_long_w<16>(p, src, d_results, g, stream)
                                       : launch_brute_long_w<32>(p, src, d_results, g, stream));
    }
    switch (p.W) {
        case 1: return ascii ? launch_brute_w<1, true>(p, src, d_results, g, stream) : launch_brute_w<1, false>(p, src, d_results, g, stream);
        case 2: return ascii ? launch_brute_w<2, true>(p, src, d_results, g, stream) : launch_brute_w<2, false>(p, src, d_results, g, stream);
        case 3: return ascii ? launch_brute_w<3, true>(p, src, d_results, g, stream) : launch_brute_w<3, false>(p, src, d_results, g, stream);
        default: return ascii ? launch_brute_w<4, true>(p, src, d_results, g, stream) : launch_brute_w<4, false>(p, src, d_results, g, stream);
    }
}

template <int W, bool ASCII>
static cudaError_t launch_probe_w(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                  const LaunchGeometry& g, cudaStream_t stream) {
    auto k = k_probe<W, ASCII>;
    const size_t hb = hist_bytes(p);
    const int grid = grid_for(k, PROBE_THREADS, hb, g, src.n);
    k<<<grid, PROBE_THREADS, hb, stream>>>(p, src, d_results);
    count_launch();
    return cudaGetLastError();
}

uint32_t probe2_hist_rep(uint32_t S) {  // histogram replicas (power of two <= 32) within PROBE2_HIST_BYTES
    uint32_t rep = 32;
    while (rep > 1 && (size_t)(S + 1u) * rep * 4u > PROBE2_HIST_BYTES) rep >>= 1;
    return rep;
}

size_t probe2_fixed_smem_bytes(uint32_t W, uint32_t S, int threads) {  // queues + replicated histogram
    return (size_t)(threads / 32) * PROBE2_QUEUE * W * 4 + (size_t)(S + 1u) * probe2_hist_rep(S) * 4u;
}

size_t probe2_smem_bytes(const MatchParams& p, int threads) {
    return (size_t)p.tier_slots * p.tier_rep * tier_entry_words((int)p.W) * 4 +
           (tier_separate_values((int)p.W) ? (size_t)p.tier_slots * 4 : 0) +
           (size_t)p.bloom_words * 4 + probe2_fixed_smem_bytes(p.W, p.S, threads);
}

int probe2_threads() { return PROBE2_THREADS; }

template <int W>
static cudaError_t launch_probe2_w(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                   const LaunchGeometry& g, cudaStream_t stream) {
    auto k = k_probe2<W>;
    const size_t smem = probe2_smem_bytes(p, PROBE2_THREADS);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t n_groups = (src.n + PROBE2_R - 1) / PROBE2_R;
    const int grid = grid_for(k, PROBE2_THREADS, smem, g, n_groups);
    k<<<grid, PROBE2_THREADS, smem, stream>>>(p, src, d_results);
    count_launch();
    return cudaGetLastError();
}

constexpr uint32_t PROBE3_MAX_WARPS = 32;
size_t probe3_smem_bytes(uint32_t ck_words, uint32_t S, uint32_t stash_cap) {
    return (size_t)ck_words * 4 + (size_t)probe3_hist_words(S) * 4 + 16 +
           (size_t)PROBE3_MAX_WARPS * stash_cap * 12;
}

// launch shape of k_probe3: 1024 threads (the kernel needs all 32 resident warps: 768 / 512 threads measured 10 - 30 %
// slower); 4 reads per lane for two-word keys, 8 for one-word keys (same register footprint; cfg 2: 0.204 -> 0.193 ms).
// FQTK_B200_P3_SHAPE=4x1024 (A/B timing) forces 4 reads per lane.
static int probe3_shape() {
    static const int v = [] {
        const char* e = getenv("FQTK_B200_P3_SHAPE");
        return (e && !strcmp(e, "4x1024")) ? 1 : 0;
    }();
    return v;
}

template <int W, int NP, bool PAD, int R, int THREADS>
static cudaError_t launch_probe3_shape(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                       const LaunchGeometry& g, cudaStream_t stream) {
    auto k = k_probe3<W, NP, PAD, R, THREADS>;
    const size_t smem = probe3_smem_bytes(p.ck_words, p.S, p.ck_stash_cap);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t n_warp_tiles = (src.n + 32 * R - 1) / (32 * R);
    const uint64_t want = (n_warp_tiles + THREADS / 32 - 1) / (THREADS / 32);
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)g.sm_count));
    k<<<grid, THREADS, smem, stream>>>(p, src, d_results);
    count_launch();
    return cudaGetLastError();
}

template <int W, int NP, bool PAD>
static cudaError_t launch_probe3_wnp(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                     const LaunchGeometry& g, cudaStream_t stream) {
    if constexpr (W == 1) {  // same register footprint as two-word keys at 4 reads per lane
        if (probe3_shape() == 0) return launch_probe3_shape<W, NP, PAD, 8, 1024>(p, src, d_results, g, stream);
    }
    return launch_probe3_shape<W, NP, PAD, 4, 1024>(p, src, d_results, g, stream);
}

size_t probe4_smem_bytes(uint32_t W, uint32_t S, uint32_t hist_rep, uint32_t stash_cap) {
    return (size_t)S * probe4_ne_stride(W) * 4 + (size_t)(S + 1u) * hist_rep * 4 +
           (size_t)(PROBE4_THREADS / 32) * stash_cap * (W * 4 + 4) + 16;
}

// launch shape of k_probe4: threads per CTA (one CTA per SM) x home buckets in flight per lane; g4_flags bits 4-5 (A/B
// timing): 0 = 1024 x 2, 1 = 1024 x 4, 2 = 768 x 4, 3 = 512 x 4
template <int W, bool PAD, int THREADS, int IF>
static cudaError_t launch_probe4_shape(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                       const LaunchGeometry& g, cudaStream_t stream) {
    const size_t smem = probe4_smem_bytes(W, p.S, p.g4_hist_rep, p.g4_stash_cap);
    const uint64_t n_warp_tiles = (src.n + 32 * PROBE4_R - 1) / (32 * PROBE4_R);
    const uint64_t want = (n_warp_tiles + THREADS / 32 - 1) / (THREADS / 32);
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)g.sm_count));
    auto k = k_probe4<W, PAD, THREADS, IF>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, THREADS, smem, stream>>>(p, src, d_results);
    count_launch();
    return cudaGetLastError();
}
template <int W, bool PAD>
static cudaError_t launch_probe4_wp(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                    const LaunchGeometry& g, cudaStream_t stream) {
    switch ((p.g4_flags >> 4) & 3u) {
        case 0: return launch_probe4_shape<W, PAD, 1024, 2>(p, src, d_results, g, stream);
        case 1: return launch_probe4_shape<W, PAD, 1024, 4>(p, src, d_results, g, stream);
        case 2: return launch_probe4_shape<W, PAD, 768, 4>(p, src, d_results, g, stream);
        default: return launch_probe4_shape<W, PAD, 512, 4>(p, src, d_results, g, stream);
    }
}
template <int W>
static cudaError_t launch_probe4_w(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                   const LaunchGeometry& g, cudaStream_t stream) {
    return p.last_pad ? launch_probe4_wp<W, true>(p, src, d_results, g, stream)
                      : launch_probe4_wp<W, false>(p, src, d_results, g, stream);
}

size_t probe5_smem_bytes(uint32_t ck_words, uint32_t S, uint32_t W, uint32_t hist_rep) {
    return (size_t)ck_words * 4 + (size_t)(S + 1u) * hist_rep * 4 + (size_t)S * W * 4 +
           (size_t)(PROBE5_THREADS / 32) * (32 * PROBE5_R) * W * 4 + 16;
}

template <int W, int NP>
static cudaError_t launch_probe5_wn(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                    const LaunchGeometry& g, cudaStream_t stream) {
    const size_t smem = probe5_smem_bytes(p.ck_words, p.S, W, p.g4_hist_rep);
    const uint64_t n_warp_tiles = (src.n + 32 * PROBE5_R - 1) / (32 * PROBE5_R);
    const uint64_t want = (n_warp_tiles + PROBE5_THREADS / 32 - 1) / (PROBE5_THREADS / 32);
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)g.sm_count));
    if (p.last_pad) {
        auto k = k_probe5<W, NP, true>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, PROBE5_THREADS, smem, stream>>>(p, src, d_results);
    } else {
        auto k = k_probe5<W, NP, false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, PROBE5_THREADS, smem, stream>>>(p, src, d_results);
    }
    count_launch();
    return cudaGetLastError();
}

template <int W, int NP>
static cudaError_t launch_probe3_wn(const MatchParams& p, const ReadSource& src, uint32_t* d_results,
                                    const LaunchGeometry& g, cudaStream_t stream) {
    return p.last_pad ? launch_probe3_wnp<W, NP, true>(p, src, d_results, g, stream)
                      : launch_probe3_wnp<W, NP, false>(p, src, d_results, g, stream);
}

cudaError_t launch_probe(const MatchParams& p, const ReadSource& src, uint32_t* d_results, const LaunchGeometry& g,
                         cudaStream_t stream) {
    if (src.n == 0) return cudaSuccess;
    if (p.table == nullptr || p.W > (uint32_t)MAX_FAST_WORDS) return cudaErrorInvalidValue;
    const bool ascii = src.ascii != nullptr;
    if (!ascii && p.ck_np && p.ck_exact_only && p.g4_table && p.W <= 2u &&
        ((reinterpret_cast<uintptr_t>(src.packed) | reinterpret_cast<uintptr_t>(d_results)) & 15u) == 0u) {
        if (p.W == 1) return p.ck_np == 2 ? launch_probe5_wn<1, 2>(p, src, d_results, g, stream)
                                          : launch_probe5_wn<1, 3>(p, src, d_results, g, stream);
        return p.ck_np == 2 ? launch_probe5_wn<2, 2>(p, src, d_results, g, stream)
                            : launch_probe5_wn<2, 3>(p, src, d_results, g, stream);
    }
    if (!ascii && p.ck_np && !p.ck_exact_only && p.W <= 2u &&
        ((reinterpret_cast<uintptr_t>(src.packed) | reinterpret_cast<uintptr_t>(d_results)) & 15u) == 0u) {
        if (p.W == 1) return p.ck_np == 2 ? launch_probe3_wn<1, 2>(p, src, d_results, g, stream)
                                          : launch_probe3_wn<1, 3>(p, src, d_results, g, stream);
        return p.ck_np == 2 ? launch_probe3_wn<2, 2>(p, src, d_results, g, stream)
                            : launch_probe3_wn<2, 3>(p, src, d_results, g, stream);
    }
    if (!ascii && p.g4_table && ((reinterpret_cast<uintptr_t>(src.packed) | reinterpret_cast<uintptr_t>(d_results)) & 15u) == 0u) {
        switch (p.W) {
            case 1: return launch_probe4_w<1>(p, src, d_results, g, stream);
            case 2: return launch_probe4_w<2>(p, src, d_results, g, stream);
            case 3: return launch_probe4_w<3>(p, src, d_results, g, stream);
            default: return launch_probe4_w<4>(p, src, d_results, g, stream);
        }
    }
    if (ascii) {
        switch (p.W) {
            case 1: return launch_probe_w<1, true>(p, src, d_results, g, stream);
            case 2: return launch_probe_w<2, true>(p, src, d_results, g, stream);
            case 3: return launch_probe_w<3, true>(p, src, d_results, g, stream);
            default: return launch_probe_w<4, true>(p, src, d_results, g, stream);
        }
    }
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(src.packed) | reinterpret_cast<uintptr_t>(d_results)) & 15u) == 0u &&
                        p.S + 1u <= HIST_SMEM_BINS &&
                        probe2_smem_bytes(p, PROBE2_THREADS) + 1024 <= (size_t)g.max_smem_optin;
    if (vec_ok) {
        switch (p.W) {
            case 1: return launch_probe2_w<1>(p, src, d_results, g, stream);
            case 2: return launch_probe2_w<2>(p, src, d_results, g, stream);
            case 3: return launch_probe2_w<3>(p, src, d_results, g, stream);
            default: return launch_probe2_w<4>(p, src, d_results, g, stream);
        }
    }
    switch (p.W) {
        case 1: return launch_probe_w<1, false>(p, src, d_results, g, stream);
        case 2: return launch_probe_w<2, false>(p, src, d_results, g, stream);
        case 3: return launch_probe_w<3, false>(p, src, d_results, g, stream);
        default: return launch_probe_w<4, false>(p, src, d_results, g, stream);
    }
}

cudaError_t launch_pack(const uint8_t* d_ascii, uint64_t n, uint32_t L, uint64_t stride, uint32_t* d_packed,
                        const LaunchGeometry& g, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(k_pack, 256, 0, g, n);
    k_pack<<<grid, 256, 0, stream>>>(d_ascii, n, L, stride, d_packed);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_pack_segments(const SegmentSource& seg, uint64_t n, uint32_t L, uint32_t* d_packed,
                                 const LaunchGeometry& g, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(k_pack_segments, 256, 0, g, n);
    k_pack_segments<<<grid, 256, 0, stream>>>(seg, n, L, d_packed);
    count_launch();
    return cudaGetLastError();
}

// k_pack_offsets: ReadSetIterator::next's segment extraction (demux.rs:288-342) + sample_barcode_sequence (:121-123) +
// encode() (mod.rs:49-61) for reads that still sit in raw FASTQ text: the B segments are gathered by per-read offsets of
// the sequence lines (fqtk_b200_fastq_scan) and packed straight into BitEnc words.  lengths[i] = bases gathered for read i
// (it differs from L only with a trailing `+B` segment); symbols beyond L are dropped, missing ones stay 0.
__global__ void __launch_bounds__(256) k_pack_offsets(const OffsetSource seg, uint64_t n, uint32_t L,
                                                      uint32_t* __restrict__ packed, uint32_t* __restrict__ lengths) {
    __shared__ uint8_t s_lut[256];
    init_lut(s_lut);
    __syncthreads();
    const uint32_t W = words_for_len(L);
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += total) {
        uint32_t acc = 0u, pos = 0u;  // pos = symbols gathered so far
        for (uint32_t s = 0; s < seg.n_segments; s++) {
            const uint32_t so = seg.source_of[s];
            const uint8_t* src = seg.base[so] + __ldg(seg.seq_offsets[so] + i) + seg.offset[s];
            uint32_t len = seg.length[s];
            if (len == OffsetSource::REST) {
                const uint32_t sl = __ldg(seg.seq_lengths[so] + i);
                len = sl > seg.offset[s] ? sl - seg.offset[s] : 0u;
            }
            for (uint32_t k = 0; k < len; k++, pos++) {
                if (pos < L) {
                    acc |= (uint32_t)s_lut[__ldg(src + k)] << (4u * (pos & 7u));
                    if ((pos & 7u) == 7u) {
                        packed[i * W + (pos >> 3)] = acc;
                        acc = 0u;
                    }
                }
            }
        }
        if (pos < L) {  // a short read: zero the rest of its row (its result is None whatever the row holds)
            for (uint32_t wd = pos >> 3; wd < W; wd++) {
                packed[i * W + wd] = acc;
                acc = 0u;
            }
        } else if (L & 7u) {
            packed[i * W + (W - 1u)] = acc;
        }
        if (lengths) lengths[i] = pos;
    }
}

// reads whose gathered barcode length differs from L are None (barcode_matching.rs:167-169; longer ones were vetted by
// the host): undo whatever the matching kernel made of their row, counts included
__global__ void __launch_bounds__(256) k_fix_lengths(uint32_t* __restrict__ results, const uint32_t* __restrict__ lengths,
                                                     uint64_t n, uint32_t L, uint32_t S, unsigned long long* __restrict__ counts) {
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += total) {
        if (__ldg(lengths + i) != L) {
            const uint32_t r = results[i];
            if (r != NONE) {
                results[i] = NONE;
                atomicAdd(&counts[r >> 16], ~0ull);  // - 1
                atomicAdd(&counts[S], 1ull);
            }
        }
    }
}

cudaError_t launch_pack_offsets(const OffsetSource& seg, uint64_t n, uint32_t L, uint32_t* d_packed, uint32_t* d_lengths,
                                const LaunchGeometry& g, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(k_pack_offsets, 256, 0, g, n);
    k_pack_offsets<<<grid, 256, 0, stream>>>(seg, n, L, d_packed, d_lengths);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_fix_lengths(uint32_t* d_results, const uint32_t* d_lengths, uint64_t n, uint32_t L, uint32_t S,
                               unsigned long long* d_counts, const LaunchGeometry& g, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(k_fix_lengths, 256, 0, g, n);
    k_fix_lengths<<<grid, 256, 0, stream>>>(d_results, d_lengths, n, L, S, d_counts);
    count_launch();
    return cudaGetLastError();
}

// result words -> u16 sample indices (0xFFFF = None): the 2-byte-per-read return format of the packed host call
__global__ void __launch_bounds__(256) k_narrow_u16(const uint32_t* __restrict__ results, uint64_t n,
                                                    uint16_t* __restrict__ out) {
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += total) {
        const uint32_t r = __ldg(results + i);
        out[i] = r == NONE ? (uint16_t)0xFFFFu : (uint16_t)(r >> 16);
    }
}

cudaError_t launch_narrow_u16(const uint32_t* d_results, uint64_t n, uint16_t* d_out, const LaunchGeometry& g,
                              cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(k_narrow_u16, 256, 0, g, n);
    k_narrow_u16<<<grid, 256, 0, stream>>>(d_results, n, d_out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t prepare_kernels(const LaunchGeometry&) { return cudaSuccess; }

}  // namespace fq
