// bgzf.cu — BGZF (blocked gzip) output compression on the GPU: SURVEY.md 8f "next" #4, the tail of the reference's hot loop.
//
// The reference hands every output record to `pooled_writer` (src/bin/commands/demux.rs:755-798: `PoolBuilder::<_,
// BgzfCompressor>` with `compression_level`, default 5, demux.rs:641-643), which buffers 65 280 bytes per writer
// (bgzf crate `BGZF_BLOCK_SIZE`), deflates every full buffer as ONE independent gzip member with the BGZF `BC` extra
// field (SAM spec 4.1), and ends each file with the 28-byte EOF block.  The members are independent of each other, so a
// batch of output text is an embarrassingly parallel set of 64 KB compression jobs: one CTA per BGZF block here.
//
// Per block (16 warps, everything in shared memory; two blocks per SM):
//   1  LZ77 parse: warp w owns one sixteenth of the block.  32 positions per step: a 4-byte hash finds the most recent
//      earlier position of the warp's part with the same hash (512-entry table per warp; the step's own positions are
//      inserted with "largest position wins", which makes the table — and the output — deterministic), plus the run
//      candidate at distance 1; match lengths by 4-byte compares; the greedy parse of the 32 positions (which of them start
//      a token, given how far the previous step's last match reaches) is the set reachable from the first uncovered
//      position along "next token" pointers: five rounds of pointer doubling with `redux.sync.or`, no serial walk.
//      Tokens go to a global scratch ring (L2), symbol counts to a per-warp histogram.
//   2  length-limited canonical Huffman codes for the literal/length and distance alphabets: parallel rank sort of the
//      used symbols, then the in-place Moffat-Katajainen construction and the Kraft repair of over-long codes by one
//      thread per alphabet; the code-length alphabet (RFC 1951 3.2.7, run-length symbols 16 / 17 / 18) the same way.
//   3  bit emission: every warp knows its bit offset from its own histogram; 32 tokens per step, warp scan of the bit
//      lengths, bits OR-ed into a warp-private shared window, whole words stored coalesced.
//   4  CRC-32 of the block: 512 partial CRCs, each advanced over the bytes behind it by multiplication with
//      x^(8 n) mod P (zlib's crc32_combine arithmetic), XOR-reduced.
// A block that does not shrink is written as a stored deflate block (BGZF guarantees the 64 KB bound that way).
// The bytes are not libdeflate's — no two deflate implementations agree — so parity is what the reference's own tests
// check: the members inflate to the input, block boundaries every 65 280 bytes, valid BGZF framing, EOF block.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fqtk_b200.h"
#include "kernels.h"

namespace fq {
void set_last_error(const std::string& msg);
}

namespace {

constexpr uint32_t BZ_IN = 65280;        // input bytes per BGZF block (bgzf crate BGZF_BLOCK_SIZE)
constexpr uint32_t BZ_HDR = 18, BZ_TRL = 8;
constexpr uint32_t BZ_SLOT = 65536 + 64;  // output slot stride per block (multiple of 16; block image at +2)
constexpr int BZ_THREADS = 512, BZ_WARPS = 16;
constexpr uint32_t BZ_HASH_BITS = 9, BZ_HASH = 1u << BZ_HASH_BITS;  // 16 KB of heads per CTA: two CTAs (32 warps) fit an SM
constexpr uint32_t BZ_TOK_PER_WARP = 4096;  // >= ceil(65280 / 16) tokens
constexpr uint32_t NLIT = 286, NDIST = 30, DOFF = 288, NSYM = 320;  // histogram layout: lit/len 0..285, dist 288..317
constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr uint32_t BZ_WIN = 56;  // words of a warp's bit window: 31 + 32 tokens x 48 bits = 50 words at most

__device__ uint32_t d_crc_table[256];
__device__ uint32_t d_x2n[32];

// ---------------------------------------------------------------------------------------------------- symbols
__device__ __forceinline__ void len_symbol(uint32_t y /* length - 3 */, uint32_t& sym, uint32_t& ebits, uint32_t& eval) {
    if (y < 8u) {
        sym = 257u + y; ebits = 0; eval = 0;
    } else if (y == 255u) {
        sym = 285u; ebits = 0; eval = 0;
    } else {
        const uint32_t nb = 31u - __clz(y);
        sym = 257u + 4u * (nb - 1u) + ((y >> (nb - 2u)) & 3u);
        ebits = nb - 2u;
        eval = y & ((1u << ebits) - 1u);
    }
}
__device__ __forceinline__ void dist_symbol(uint32_t x /* distance - 1 */, uint32_t& sym, uint32_t& ebits, uint32_t& eval) {
    if (x < 4u) {
        sym = x; ebits = 0; eval = 0;
    } else {
        const uint32_t nb = 31u - __clz(x);
        sym = 2u * nb + ((x >> (nb - 1u)) & 1u);
        ebits = nb - 1u;
        eval = x & ((1u << ebits) - 1u);
    }
}
__device__ __forceinline__ uint32_t len_extra_bits(uint32_t sym) {  // sym 257..285
    const uint32_t i = sym - 257u;
    return (i < 8u || i == 28u) ? 0u : (i - 4u) >> 2;
}
__device__ __forceinline__ uint32_t dist_extra_bits(uint32_t sym) { return sym < 4u ? 0u : (sym - 2u) >> 1; }

// ---------------------------------------------------------------------------------------------------- CRC-32
__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {  // a(x) * b(x) mod P, reflected (zlib crc32.c)
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1u)) == 0u) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
__device__ __forceinline__ uint32_t x2nmodp(uint32_t n, uint32_t k) {  // x^(n * 2^k) mod P
    uint32_t p = 1u << 31;
    while (n) {
        if (n & 1u) p = multmodp(d_x2n[k & 31u], p);
        n >>= 1;
        k++;
    }
    return p;
}

// ---------------------------------------------------------------------------------------------------- shared memory
struct BzPost {                              // what the phases after the parse need (aliases the hash heads, dead by then)
    uint32_t freq[NSYM];                     // block totals
    uint32_t win[BZ_WARPS][BZ_WIN];          // bit windows of the emitters
    uint32_t hdr[160];                       // the dynamic-block header bits
    uint32_t sortA[2][NLIT + 2];             // Huffman scratch (lit/len, dist): frequencies in ascending order -> depths
    uint16_t order[2][NLIT + 2];             //   symbol at every sorted position
    uint8_t rle_sym[NLIT + NDIST + 4], rle_ext[NLIT + NDIST + 4];
    uint32_t num[3][36];                     // huff_lengths scratch (lit/len, dist, code-length alphabet)
    uint32_t cnt[3][16];                     // codes per length -> first code per length
    uint32_t clf[20], clA[20];               // code-length alphabet: frequencies, sorted frequencies -> depths
    uint16_t clord[20], clc[20];             //   sorted symbols, codes
    uint8_t cll[20];                         //   lengths
};
struct BzShared {
    uint32_t in[(65536 + 64) / 4];           // the block, zero-padded
    union {
        uint32_t head[BZ_WARPS][BZ_HASH];    // per-warp hash heads: 1 + position inside the warp's part, 0 = none (32-bit: atomicMax)
        BzPost post;
    };
    uint32_t hist[BZ_WARPS][NSYM / 2];       // per-warp symbol counts, two 16-bit counters per word (a part has < 65 536 tokens)
    uint32_t crc_tab[256];
    uint16_t code[NSYM];                     // bit-reversed canonical codes
    uint8_t clen[NSYM];                      // code lengths
    uint32_t used[2];                        // used symbols per alphabet
    uint32_t ntok[BZ_WARPS];
    uint32_t start_bit[BZ_WARPS + 2];        // bit offset of the header (0), of every warp's tokens, and the end
    uint32_t hdr_bits;
    uint32_t crc_part[BZ_WARPS];
};
static_assert(sizeof(BzPost) <= sizeof(uint32_t) * BZ_WARPS * BZ_HASH, "the post-parse scratch must fit in the hash heads");
static_assert(2 * (sizeof(BzShared) + 1024) <= 233472, "two blocks per SM");

__device__ __forceinline__ void hist_inc(uint32_t* h, uint32_t sym) { atomicAdd(&h[sym >> 1], 1u << ((sym & 1u) * 16u)); }
__device__ __forceinline__ uint32_t hist_get(const uint32_t* h, uint32_t sym) { return (h[sym >> 1] >> ((sym & 1u) * 16u)) & 0xFFFFu; }

__device__ __forceinline__ uint32_t read4(const uint32_t* in, uint32_t pos) {  // unaligned little-endian 4 bytes
    const uint32_t lo = in[pos >> 2], hi = in[(pos >> 2) + 1u];
    return __funnelshift_r(lo, hi, (pos & 3u) * 8u);
}
__device__ __forceinline__ uint32_t read1(const uint32_t* in, uint32_t pos) {
    return (in[pos >> 2] >> ((pos & 3u) * 8u)) & 0xFFu;
}

// ---------------------------------------------------------------------------------------------------- Huffman
// One thread.  A[0..m) = frequencies ascending (m >= 2) -> A[i] = code length of sorted position i (Moffat & Katajainen,
// "In-place calculation of minimum-redundancy codes"), then lengths above `limit` are repaired the way miniz does
// (count per length, fold the over-long ones into `limit`, restore the Kraft sum) and handed out again in sorted order:
// the most frequent symbol (last sorted position) gets the shortest code.
__device__ void huff_lengths(uint32_t* A, const uint16_t* order, uint32_t m, uint32_t limit, uint8_t* out_len, uint32_t* num /* 33 words of scratch */) {
    if (m == 2u) {
        out_len[order[0]] = 1;
        out_len[order[1]] = 1;
        return;
    }
    A[0] += A[1];
    uint32_t root = 0, leaf = 2;
    for (uint32_t next = 1; next < m - 1u; next++) {
        if (leaf >= m || A[root] < A[leaf]) { A[next] = A[root]; A[root++] = next; } else A[next] = A[leaf++];
        if (leaf >= m || (root < next && A[root] < A[leaf])) { A[next] += A[root]; A[root++] = next; } else A[next] += A[leaf++];
    }
    A[m - 2u] = 0;
    for (int next = (int)m - 3; next >= 0; next--) A[next] = A[A[next]] + 1u;
    int avbl = 1, used = 0, dpth = 0, rt = (int)m - 2, nx = (int)m - 1;
    while (avbl > 0) {
        while (rt >= 0 && (int)A[rt] == dpth) { used++; rt--; }
        while (avbl > used) { A[nx--] = (uint32_t)dpth; avbl--; }
        avbl = 2 * used; dpth++; used = 0;
    }
    // A[i] = depth, non-increasing in i.  Length limit.
    for (uint32_t i = 0; i <= 32u; i++) num[i] = 0;
    for (uint32_t i = 0; i < m; i++) num[min(A[i], 32u)]++;
    if (A[0] > limit) {
        for (uint32_t i = limit + 1u; i <= 32u; i++) { num[limit] += num[i]; num[i] = 0; }
        uint32_t total = 0;
        for (uint32_t i = limit; i > 0; i--) total += num[i] << (limit - i);
        while (total != (1u << limit)) {
            num[limit]--;
            for (uint32_t i = limit - 1u; i > 0; i--)
                if (num[i]) { num[i]--; num[i + 1u] += 2u; break; }
            total--;
        }
    }
    uint32_t j = m;
    for (uint32_t l = 1; l <= limit; l++)
        for (uint32_t c = num[l]; c > 0; c--) out_len[order[--j]] = (uint8_t)l;
}

// Canonical codes (RFC 1951 3.2.2), stored bit-reversed (deflate sends codes MSB first), in three small steps around two
// barriers: every thread counts its symbols' lengths into cnt[16] (shared atomics), one thread turns the counts into the
// first code of every length, then the k-th symbol of a length, in symbol order, gets first_code + k.
__device__ __forceinline__ void huff_count(const uint8_t* len, uint32_t n, uint32_t* cnt, uint32_t t, uint32_t nthreads) {
    for (uint32_t s = t; s < n; s += nthreads)
        if (len[s]) atomicAdd(&cnt[len[s]], 1u);
}
__device__ __forceinline__ void huff_first(uint32_t* cnt /* in: counts, out: first code per length */) {
    uint32_t c = 0, prev = 0;
    for (int b = 1; b < 16; b++) {
        c = (c + prev) << 1;
        prev = cnt[b];
        cnt[b] = c;
    }
}
__device__ void huff_assign(const uint8_t* len, uint32_t n, const uint32_t* first, uint16_t* code, uint32_t t, uint32_t nthreads) {
    for (uint32_t s = t; s < n; s += nthreads) {
        const uint32_t l = len[s];
        if (!l) { code[s] = 0; continue; }
        uint32_t k = 0;
        for (uint32_t u = 0; u < s; u++) k += len[u] == l;
        code[s] = (uint16_t)(__brev(first[l] + k) >> (32u - l));
    }
}

// All threads: the used symbols of freq[0..n) in ascending (frequency, symbol) order -> A / order, count -> *m_out.
__device__ void rank_sort(const uint32_t* freq, uint32_t n, uint32_t* A, uint16_t* order, uint32_t* m_out) {
    for (uint32_t s = threadIdx.x; s < n; s += blockDim.x) {
        const uint32_t f = freq[s];
        if (!f) continue;
        uint32_t r = 0;
        for (uint32_t t = 0; t < n; t++) {
            const uint32_t g = freq[t];
            r += (g != 0u && (g < f || (g == f && t < s))) ? 1u : 0u;
        }
        A[r] = f;
        order[r] = (uint16_t)s;
        atomicAdd(m_out, 1u);
    }
}

struct BitWriter {  // one thread, into shared words
    uint32_t* w;
    uint32_t pos;
    __device__ void put(uint32_t v, uint32_t n) {
        if (!n) return;
        const uint32_t i = pos >> 5, sh = pos & 31u;
        w[i] |= v << sh;
        if (sh + n > 32u) w[i + 1u] |= v >> (32u - sh);
        pos += n;
    }
};

struct BzBlockDesc {  // one BGZF block of a segmented input
    unsigned long long offset;
    uint32_t bytes, pad;
};

// ---------------------------------------------------------------------------------------------------- the kernel
// grid-stride over BGZF blocks; tokens: BZ_WARPS * BZ_TOK_PER_WARP words per CTA.
__global__ void __launch_bounds__(BZ_THREADS, 2)
    k_bgzf_deflate(const uint8_t* __restrict__ in, uint64_t n_bytes, uint32_t n_blocks, int level, uint8_t* __restrict__ slots,
                   uint32_t* __restrict__ sizes, uint32_t* __restrict__ tokens_all, const BzBlockDesc* __restrict__ desc) {
    extern __shared__ uint4 s_raw[];
    BzShared& S = *reinterpret_cast<BzShared*>(s_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5, lane_lt = (1u << lane) - 1u;
    uint32_t* tokens = tokens_all + ((size_t)blockIdx.x * BZ_WARPS + w) * BZ_TOK_PER_WARP;
    for (uint32_t t = tid; t < 256u; t += BZ_THREADS) S.crc_tab[t] = d_crc_table[t];

    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        // uniform cut every 65 280 bytes, or (segments: several writers' texts in one buffer) the block list given
        const uint64_t in_off = desc ? desc[blk].offset : (uint64_t)blk * BZ_IN;
        const uint32_t n = desc ? desc[blk].bytes : (uint32_t)min((uint64_t)BZ_IN, n_bytes - in_off);
        uint8_t* slot = slots + (size_t)blk * BZ_SLOT + 2;  // block image: 18-byte header, then the deflate data 4-byte aligned
        uint32_t* out_words = reinterpret_cast<uint32_t*>(slot + BZ_HDR);
        __syncthreads();  // the previous block's shared state is dead
        // ---- load (16-byte vectors when the source allows), zero padding behind
        {
            const uint8_t* src = in + in_off;
            const uint32_t n16 = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) ? (n >> 4) : 0u;
            const uint4* src4 = reinterpret_cast<const uint4*>(src);
            uint4* dst4 = reinterpret_cast<uint4*>(S.in);
            for (uint32_t t = tid; t < n16; t += BZ_THREADS) dst4[t] = __ldg(src4 + t);
            uint8_t* dstb = reinterpret_cast<uint8_t*>(S.in);
            for (uint32_t t = n16 * 16u + tid; t < n; t += BZ_THREADS) dstb[t] = __ldg(src + t);
            for (uint32_t t = n + tid; t < ((n + 3u) & ~3u) + 64u; t += BZ_THREADS) dstb[t] = 0;
        }
        for (uint32_t t = tid; t < BZ_WARPS * BZ_HASH; t += BZ_THREADS) (&S.head[0][0])[t] = 0u;
        for (uint32_t t = tid; t < BZ_WARPS * NSYM / 2u; t += BZ_THREADS) (&S.hist[0][0])[t] = 0u;
        for (uint32_t t = tid; t < NSYM; t += BZ_THREADS) { S.clen[t] = 0; S.code[t] = 0; }
        if (tid < 2u) S.used[tid] = 0u;
        __syncthreads();

        const bool try_deflate = level != 0 && n > 0u;
        uint32_t ntok = 0;
        if (try_deflate) {
            // ---- 1: LZ77 parse of this warp's part
            const uint32_t sub = (n + BZ_WARPS - 1u) / BZ_WARPS;
            const uint32_t s0 = min(w * sub, n), s1 = min(s0 + sub, n);
            uint32_t* head = S.head[w];
            uint32_t* hist = S.hist[w];
            uint32_t carry = 0;  // positions of the next step already covered by the last match
            for (uint32_t p = s0; p < s1; p += 32u) {
                const uint32_t pos = p + lane;
                const bool valid = pos < s1;
                const uint32_t v = read4(S.in, pos);
                const uint32_t h = (v * 2654435761u) >> (32u - BZ_HASH_BITS);
                const uint32_t rel = pos - s0;
                const uint32_t cand = valid ? head[h] : 0u;  // 1 + position, 0 = none
                __syncwarp();
                // the largest position of the step wins its hash slot (every earlier step's positions are smaller): one
                // atomic max, deterministic whatever order the lanes arrive in.  Runs (quality strings, poly-N) make
                // neighbouring lanes hash alike: a lane whose upper neighbour has the same hash leaves the slot to it.
                const uint32_t h_up = __shfl_down_sync(0xFFFFFFFFu, h, 1);
                if (valid && !(lane < 31u && pos + 1u < s1 && h_up == h)) atomicMax(&head[h], rel + 1u);
                __syncwarp();
                uint32_t best_len = 0, best_dist = 0;
                if (carry < 32u) {  // (a step that lies inside the previous match only feeds the hash table)
                    const uint32_t maxlen = valid ? min(258u, s1 - pos) : 0u;
                    if (cand != 0u && maxlen >= 4u) {
                        const uint32_t cpos = s0 + cand - 1u;
                        uint32_t k = 0;
                        while (k < maxlen) {
                            const uint32_t x = read4(S.in, cpos + k) ^ read4(S.in, pos + k);
                            if (x) { k += (uint32_t)(__ffs(x) - 1) >> 3; break; }
                            k += 4u;
                        }
                        k = min(k, maxlen);
                        if (k >= 4u) { best_len = k; best_dist = pos - cpos; }
                    }
                    if (pos > s0 && maxlen >= 3u) {  // run of the previous byte
                        const uint32_t b = read1(S.in, pos - 1u) * 0x01010101u;
                        if (((v ^ b) & 0x00FFFFFFu) == 0u) {
                            uint32_t k = 0;
                            while (k < maxlen) {
                                const uint32_t x = read4(S.in, pos + k) ^ b;
                                if (x) { k += (uint32_t)(__ffs(x) - 1) >> 3; break; }
                                k += 4u;
                            }
                            k = min(k, maxlen);
                            if (k >= 3u && k >= best_len) { best_len = k; best_dist = 1u; }
                        }
                    }
                }
                if (carry >= 32u) {
                    carry -= 32u;
                    continue;
                }
                // lazy matching (zlib's rule, free here because every position's match is already known): a match gives
                // way to a literal when the next position starts a longer one
                {
                    const uint32_t len_up = __shfl_down_sync(0xFFFFFFFFu, best_len, 1);
                    if (lane < 31u && best_len && len_up > best_len) best_len = 0;
                }
                // token starts = positions reachable from `carry` along next-token pointers
                const uint32_t nxt = lane + (best_len ? best_len : 1u);
                uint32_t jump = min(nxt, 32u);
                uint32_t M = 1u << carry;
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    const uint32_t add = __reduce_or_sync(0xFFFFFFFFu, (((M >> lane) & 1u) && jump < 32u) ? (1u << jump) : 0u);
                    M |= add;
                    const uint32_t j2 = __shfl_sync(0xFFFFFFFFu, jump, jump & 31u);
                    jump = jump < 32u ? j2 : 32u;
                }
                M &= __ballot_sync(0xFFFFFFFFu, valid);
                if (M) {
                    const uint32_t last = 31u - __clz(M);
                    const uint32_t end_off = __shfl_sync(0xFFFFFFFFu, nxt, last);
                    carry = max(end_off, 32u) - 32u;
                    if ((M >> lane) & 1u) {
                        const uint32_t idx = ntok + __popc(M & lane_lt);
                        if (best_len) {
                            uint32_t sym, eb, ev;
                            len_symbol(best_len - 3u, sym, eb, ev);
                            hist_inc(hist, sym);
                            dist_symbol(best_dist - 1u, sym, eb, ev);
                            hist_inc(hist, DOFF + sym);
                            tokens[idx] = 0x80000000u | ((best_len - 3u) << 16) | (best_dist - 1u);
                        } else {
                            hist_inc(hist, v & 0xFFu);
                            tokens[idx] = v & 0xFFu;
                        }
                    }
                    ntok += __popc(M);
                } else {
                    carry = 0;
                }
            }
            if (lane == 0) S.ntok[w] = ntok;
        }
        __syncthreads();
        // ---- 2: block totals, Huffman codes
        bool stored = !try_deflate;
        if (try_deflate) {
            // (the hash heads are dead: their memory is the scratch of the phases below)
            for (uint32_t t = tid; t < 160u; t += BZ_THREADS) S.post.hdr[t] = 0u;
            for (uint32_t t = tid; t < BZ_WARPS * BZ_WIN; t += BZ_THREADS) (&S.post.win[0][0])[t] = 0u;
            for (uint32_t s = tid; s < NSYM; s += BZ_THREADS) {
                uint32_t f = 0;
#pragma unroll
                for (int ww = 0; ww < BZ_WARPS; ww++) f += hist_get(S.hist[ww], s);
                if (s == 256u) f = 1u;                       // end of block
                S.post.freq[s] = f;
            }
            __syncthreads();
            if (tid == 0) {  // at least two used distance codes (zlib does the same: some inflaters insist); the literal /
                             // length alphabet always has two (end of block + the first token)
                uint32_t nd = 0;
                for (uint32_t s = 0; s < NDIST; s++) nd += S.post.freq[DOFF + s] != 0u;
                for (uint32_t s = 0; s < 3u && nd < 2u; s++)
                    if (S.post.freq[DOFF + s] == 0u) { S.post.freq[DOFF + s] = 1u; nd++; }
            }
            __syncthreads();
            rank_sort(S.post.freq, NLIT, S.post.sortA[0], S.post.order[0], &S.used[0]);
            rank_sort(S.post.freq + DOFF, NDIST, S.post.sortA[1], S.post.order[1], &S.used[1]);
            __syncthreads();
            if (tid == 0) huff_lengths(S.post.sortA[0], S.post.order[0], S.used[0], 15u, S.clen, S.post.num[0]);
            else if (tid == 32) huff_lengths(S.post.sortA[1], S.post.order[1], S.used[1], 15u, S.clen + DOFF, S.post.num[1]);
            else if (tid >= 64 && tid < 64 + 48) (&S.post.cnt[0][0])[tid - 64] = 0u;
            __syncthreads();
            huff_count(S.clen, NLIT, S.post.cnt[0], tid, BZ_THREADS);
            huff_count(S.clen + DOFF, NDIST, S.post.cnt[1], tid, BZ_THREADS);
            __syncthreads();
            if (tid == 32) huff_first(S.post.cnt[0]);
            else if (tid == 64) huff_first(S.post.cnt[1]);
            // (thread 0 does not need the codes for the header: it goes straight on; the barrier behind the header orders the rest)
            // ---- 3a: header (one thread): BFINAL = 1, BTYPE = 2, HLIT, HDIST, HCLEN, code-length codes, run-length coded lengths
            if (tid == 0) {
                uint32_t hlit = NLIT, hdist = NDIST;
                while (hlit > 257u && S.clen[hlit - 1u] == 0) hlit--;
                while (hdist > 1u && S.clen[DOFF + hdist - 1u] == 0) hdist--;
                const uint32_t total = hlit + hdist;
                auto L = [&](uint32_t i) -> uint32_t { return i < hlit ? S.clen[i] : S.clen[DOFF + i - hlit]; };
                uint32_t nr = 0, i = 0;
                uint32_t* clf = S.post.clf;
                for (int k = 0; k < 19; k++) clf[k] = 0;
                while (i < total) {
                    const uint32_t l = L(i);
                    uint32_t run = 1;
                    while (i + run < total && L(i + run) == l) run++;
                    if (l == 0u && run >= 3u) {
                        const uint32_t r = min(run, 138u);
                        if (r <= 10u) { S.post.rle_sym[nr] = 17; S.post.rle_ext[nr] = (uint8_t)(r - 3u); }
                        else { S.post.rle_sym[nr] = 18; S.post.rle_ext[nr] = (uint8_t)(r - 11u); }
                        clf[S.post.rle_sym[nr]]++; nr++; i += r;
                    } else if (l != 0u && run >= 4u) {  // the length itself, then "repeat previous" 3..6 times
                        S.post.rle_sym[nr] = (uint8_t)l; S.post.rle_ext[nr] = 0; clf[l]++; nr++; i++;
                        const uint32_t r = min(run - 1u, 6u);
                        S.post.rle_sym[nr] = 16; S.post.rle_ext[nr] = (uint8_t)(r - 3u); clf[16]++; nr++; i += r;
                    } else {
                        S.post.rle_sym[nr] = (uint8_t)l; S.post.rle_ext[nr] = 0; clf[l]++; nr++; i++;
                    }
                }
                // code-length alphabet: at least two used symbols, lengths <= 7
                uint32_t nused = 0;
                for (int k = 0; k < 19; k++) nused += clf[k] != 0u;
                for (int k = 0; k < 19 && nused < 2u; k++) if (!clf[k]) { clf[k] = 1; nused++; }
                uint32_t* A = S.post.clA;
                uint16_t* ord = S.post.clord;
                uint8_t* cll = S.post.cll;
                uint16_t* clc = S.post.clc;
                uint32_t m = 0;
                for (int k = 0; k < 19; k++) cll[k] = 0;
                for (int k = 0; k < 19; k++) if (clf[k]) { A[m] = clf[k]; ord[m] = (uint16_t)k; m++; }
                for (uint32_t a = 1; a < m; a++) {  // insertion sort by (frequency, symbol)
                    const uint32_t fa = A[a]; const uint16_t oa = ord[a];
                    int b = (int)a - 1;
                    while (b >= 0 && (A[b] > fa || (A[b] == fa && ord[b] > oa))) { A[b + 1] = A[b]; ord[b + 1] = ord[b]; b--; }
                    A[b + 1] = fa; ord[b + 1] = oa;
                }
                huff_lengths(A, ord, m, 7u, cll, S.post.num[2]);
                {
                    uint32_t* cc = S.post.cnt[2];
                    for (int k = 0; k < 16; k++) cc[k] = 0;
                    for (int k = 0; k < 19; k++) if (cll[k]) cc[cll[k]]++;
                    huff_first(cc);
                    huff_assign(cll, 19u, cc, clc, 0u, 1u);
                }
                static const uint8_t perm[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint32_t hclen = 19;
                while (hclen > 4u && cll[perm[hclen - 1u]] == 0) hclen--;
                BitWriter bw{S.post.hdr, 0u};
                bw.put(1u, 1u); bw.put(2u, 2u);
                bw.put(hlit - 257u, 5u); bw.put(hdist - 1u, 5u); bw.put(hclen - 4u, 4u);
                for (uint32_t k = 0; k < hclen; k++) bw.put(cll[perm[k]], 3u);
                for (uint32_t k = 0; k < nr; k++) {
                    const uint32_t s = S.post.rle_sym[k];
                    bw.put(clc[s], cll[s]);
                    if (s == 16u) bw.put(S.post.rle_ext[k], 2u);
                    else if (s == 17u) bw.put(S.post.rle_ext[k], 3u);
                    else if (s == 18u) bw.put(S.post.rle_ext[k], 7u);
                }
                S.hdr_bits = bw.pos;
            }
            __syncthreads();
            huff_assign(S.clen, NLIT, S.post.cnt[0], S.code, tid, BZ_THREADS);
            if (tid < 32u) huff_assign(S.clen + DOFF, NDIST, S.post.cnt[1], S.code + DOFF, tid, 32u);
            __syncthreads();
            // ---- 3b: bit offsets of the warps' token streams from their own histograms
            {
                uint32_t bits = 0;
                for (uint32_t s = lane; s < NSYM; s += 32u) {
                    const uint32_t c = hist_get(S.hist[w], s);
                    if (!c) continue;
                    uint32_t eb = 0;
                    if (s >= 257u && s < NLIT) eb = len_extra_bits(s);
                    else if (s >= DOFF) eb = dist_extra_bits(s - DOFF);
                    bits += c * (S.clen[s] + eb);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xFFFFFFFFu, bits, o);
                if (lane == 0) S.start_bit[w + 1u] = bits;  // (lengths for now)
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t run = S.hdr_bits;
                S.start_bit[0] = 0u;
                for (int ww = 0; ww < BZ_WARPS; ww++) { const uint32_t b = S.start_bit[ww + 1]; S.start_bit[ww + 1] = run; run += b; }
                S.start_bit[BZ_WARPS + 1] = run + S.clen[256];  // after the end-of-block code
            }
            __syncthreads();
            const uint32_t total_bits = S.start_bit[BZ_WARPS + 1];
            const uint32_t cbytes = (total_bits + 7u) >> 3;
            stored = cbytes >= n + 5u;
            if (!stored) {
                // zero the words shared by two emitters, then everybody ORs into them
                if (tid <= BZ_WARPS + 1u) out_words[S.start_bit[tid] >> 5] = 0u;
                __syncthreads();
                if (w == 0) {  // header words
                    const uint32_t hb = S.hdr_bits, nw = (hb + 31u) >> 5;
                    for (uint32_t k = lane; k < nw; k += 32u) {
                        const bool edge = k == 0u || k == (hb >> 5);
                        if (edge) atomicOr(&out_words[k], S.post.hdr[k]); else out_words[k] = S.post.hdr[k];
                    }
                }
                // ---- 3c: the warp's tokens
                {
                    uint32_t* win = S.post.win[w];
                    uint32_t bitpos = S.start_bit[w + 1u];
                    const uint32_t first_word = bitpos >> 5;
                    const uint32_t nt = S.ntok[w] + (w == BZ_WARPS - 1u ? 1u : 0u);  // the last warp appends end-of-block
                    for (uint32_t t0 = 0; t0 < nt; t0 += 32u) {
                        const uint32_t t = t0 + lane;
                        unsigned long long val = 0;
                        uint32_t nb = 0;
                        if (t < nt) {
                            if (t == S.ntok[w]) {  // (only the last warp gets here)
                                val = S.code[256]; nb = S.clen[256];
                            } else {
                                const uint32_t tok = tokens[t];
                                if (tok & 0x80000000u) {
                                    uint32_t sym, eb, ev;
                                    len_symbol((tok >> 16) & 0xFFu, sym, eb, ev);
                                    val = S.code[sym]; nb = S.clen[sym];
                                    val |= (unsigned long long)ev << nb; nb += eb;
                                    dist_symbol(tok & 0xFFFFu, sym, eb, ev);
                                    val |= (unsigned long long)S.code[DOFF + sym] << nb; nb += S.clen[DOFF + sym];
                                    val |= (unsigned long long)ev << nb; nb += eb;
                                } else {
                                    val = S.code[tok]; nb = S.clen[tok];
                                }
                            }
                        }
                        uint32_t off = nb;  // inclusive scan
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, off, o);
                            if (lane >= (uint32_t)o) off += u;
                        }
                        const uint32_t total = __shfl_sync(0xFFFFFFFFu, off, 31);
                        const uint32_t my = (bitpos & 31u) + off - nb;
                        if (nb) {
                            const uint32_t wi = my >> 5, sh = my & 31u;
                            const unsigned long long lo = val << sh;
                            atomicOr(&win[wi], (uint32_t)lo);
                            const uint32_t w1 = (uint32_t)(lo >> 32);
                            if (w1) atomicOr(&win[wi + 1u], w1);
                            if (sh > 16u) {
                                const uint32_t w2 = (uint32_t)(val >> (64u - sh));
                                if (w2) atomicOr(&win[wi + 2u], w2);
                            }
                        }
                        __syncwarp();
                        const uint32_t filled = (bitpos & 31u) + total, full = filled >> 5;
                        const uint32_t base_word = bitpos >> 5;
                        for (uint32_t k = lane; k < full; k += 32u) {
                            if (base_word + k == first_word) atomicOr(&out_words[first_word], win[k]);
                            else out_words[base_word + k] = win[k];
                        }
                        const uint32_t rem = win[full];
                        __syncwarp();
                        for (uint32_t k = lane; k <= full + 1u && k < BZ_WIN; k += 32u) win[k] = 0u;
                        __syncwarp();
                        if (lane == 0) win[0] = rem;
                        __syncwarp();
                        bitpos += total;
                    }
                    if (lane == 0 && (bitpos & 31u)) atomicOr(&out_words[bitpos >> 5], win[0]);
                    __syncwarp();
                    if (lane == 0) win[0] = 0u;
                }
            }
        }
        // ---- 4: CRC-32 of the input: partial CRCs advanced over what follows them
        {
            const uint32_t per = (n + BZ_THREADS - 1u) / BZ_THREADS;
            const uint32_t c0 = min(tid * per, n), c1 = min(c0 + per, n);
            uint32_t crc = 0u;
            if (c1 > c0) {
                crc = 0xFFFFFFFFu;
                for (uint32_t i = c0; i < c1; i++) crc = S.crc_tab[(crc ^ read1(S.in, i)) & 0xFFu] ^ (crc >> 8);
                crc ^= 0xFFFFFFFFu;
                const uint32_t after = n - c1;
                if (after) crc = multmodp(x2nmodp(after, 3u), crc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) crc ^= __shfl_xor_sync(0xFFFFFFFFu, crc, o);
            if (lane == 0) S.crc_part[w] = crc;
        }
        __syncthreads();
        uint32_t cbytes;
        if (stored) {  // BFINAL = 1, BTYPE = 0, LEN, NLEN, bytes
            uint8_t* c = slot + BZ_HDR;
            if (tid == 0) {
                c[0] = 1; c[1] = (uint8_t)(n & 0xFFu); c[2] = (uint8_t)(n >> 8); c[3] = (uint8_t)(~n & 0xFFu); c[4] = (uint8_t)((~n >> 8) & 0xFFu);
            }
            const uint8_t* sb = reinterpret_cast<const uint8_t*>(S.in);
            for (uint32_t t = tid; t < n; t += BZ_THREADS) c[5u + t] = sb[t];
            cbytes = n + 5u;
        } else {
            cbytes = (S.start_bit[BZ_WARPS + 1] + 7u) >> 3;
        }
        if (tid == 0) {
            uint32_t crc = 0;
            for (int ww = 0; ww < BZ_WARPS; ww++) crc ^= S.crc_part[ww];
            const uint32_t bsize = BZ_HDR + cbytes + BZ_TRL - 1u;
            const uint8_t xfl = level >= 9 ? 2 : (level == 1 ? 4 : 0);
            const uint8_t hdr[BZ_HDR] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, xfl, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)(bsize & 0xFFu), (uint8_t)(bsize >> 8)};
            for (uint32_t k = 0; k < BZ_HDR; k++) slot[k] = hdr[k];
            uint8_t* t = slot + BZ_HDR + cbytes;
            for (int k = 0; k < 4; k++) { t[k] = (uint8_t)(crc >> (8 * k)); t[4 + k] = (uint8_t)(n >> (8 * k)); }
            sizes[blk] = BZ_HDR + cbytes + BZ_TRL;
        }
    }
}

// exclusive scan of the block sizes (one CTA), offsets[n_blocks] = total
__global__ void __launch_bounds__(1024) k_bgzf_scan(const uint32_t* __restrict__ sizes, uint32_t n_blocks,
                                                     unsigned long long* __restrict__ offsets) {
    __shared__ unsigned long long part[1024];
    const uint32_t per = (n_blocks + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, n_blocks), hi = min(lo + per, n_blocks);
    unsigned long long sum = 0;
    for (uint32_t k = lo; k < hi; k++) sum += sizes[k];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) {
        const unsigned long long v = threadIdx.x >= off ? part[threadIdx.x - off] : 0ull;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = part[threadIdx.x] - sum;
    for (uint32_t k = lo; k < hi; k++) { offsets[k] = run; run += sizes[k]; }
    if (threadIdx.x == 1023u) offsets[n_blocks] = part[1023];
}

// the block images, back to back
__global__ void __launch_bounds__(256) k_bgzf_gather(const uint8_t* __restrict__ slots, const uint32_t* __restrict__ sizes,
                                                     const unsigned long long* __restrict__ offsets, uint32_t n_blocks,
                                                     uint8_t* __restrict__ out, uint64_t out_cap) {
    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const uint8_t* src = slots + (size_t)blk * BZ_SLOT + 2;
        const uint32_t sz = sizes[blk];
        const unsigned long long off = offsets[blk];
        if (off + sz > out_cap) continue;  // (the host reports the overflow from offsets[n_blocks])
        uint8_t* dst = out + off;
        // head bytes up to the destination's 4-byte boundary, then words assembled from the (differently aligned) source
        const uint32_t headb = min(sz, (uint32_t)((4u - (reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u));
        if (threadIdx.x < headb) dst[threadIdx.x] = src[threadIdx.x];
        const uint32_t nwords = (sz - headb) >> 2;
        uint32_t* dw = reinterpret_cast<uint32_t*>(dst + headb);
        const uint8_t* sp = src + headb;
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 3u);
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(sp - mis);
        for (uint32_t k = threadIdx.x; k < nwords; k += blockDim.x)
            dw[k] = mis ? __funnelshift_r(sw[k], sw[k + 1u], mis * 8u) : sw[k];
        for (uint32_t k = headb + nwords * 4u + threadIdx.x; k < sz; k += blockDim.x) dst[k] = src[k];
    }
}

// ---------------------------------------------------------------------------------------------------- host
cudaError_t upload_tables() {
    uint32_t crc[256], x2n[32];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
        crc[i] = c;
    }
    auto mult = [](uint32_t a, uint32_t b) {
        uint32_t m = 1u << 31, p = 0;
        for (;;) {
            if (a & m) { p ^= b; if ((a & (m - 1u)) == 0u) break; }
            m >>= 1;
            b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
        }
        return p;
    };
    uint32_t p = 1u << 30;  // x^1
    x2n[0] = p;
    for (int k = 1; k < 32; k++) x2n[k] = p = mult(p, p);
    cudaError_t e = cudaMemcpyToSymbol(d_crc_table, crc, sizeof(crc));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(d_x2n, x2n, sizeof(x2n));
}

int bz_fail(int code, const std::string& msg) {
    fq::set_last_error(msg);
    return code;
}
#define BZ_CU(call)                                                                                         \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) return bz_fail(FQTK_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                              0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

}  // namespace

struct fqtk_b200_bgzf {
    int device = 0;
    int sm_count = 0;
    uint64_t chunk_bytes = 0;   // input bytes per device pass (a multiple of 65 280)
    uint32_t chunk_blocks = 0;
    uint32_t grid = 0;
    static constexpr int NPIPE = 2;
    uint8_t* d_in[NPIPE] = {};
    uint8_t* d_slots[NPIPE] = {};
    uint8_t* d_out[NPIPE] = {};
    uint32_t* d_sizes[NPIPE] = {};
    unsigned long long* d_offsets[NPIPE] = {};
    unsigned long long* h_total[NPIPE] = {};  // pinned
    uint32_t* d_tokens = nullptr;
    cudaStream_t stream[NPIPE] = {};
    cudaEvent_t done[NPIPE] = {};
    uint64_t out_cap = 0;  // bytes of d_out[k]
};

static int bgzf_launch(fqtk_b200_bgzf* z, int k, const uint8_t* d_in, uint64_t n, int level, uint8_t* d_out, uint64_t out_cap,
                       unsigned long long* d_total, cudaStream_t st) {
    const uint32_t n_blocks = (uint32_t)((n + BZ_IN - 1) / BZ_IN);
    if (n_blocks == 0) return FQTK_B200_OK;
    const uint32_t grid = std::min<uint32_t>(z->grid, n_blocks);
    k_bgzf_deflate<<<grid, BZ_THREADS, sizeof(BzShared), st>>>(d_in, n, n_blocks, level, z->d_slots[k], z->d_sizes[k], z->d_tokens + (size_t)k * z->grid * BZ_WARPS * BZ_TOK_PER_WARP, nullptr);
    fq::count_launch();
    k_bgzf_scan<<<1, 1024, 0, st>>>(z->d_sizes[k], n_blocks, z->d_offsets[k]);
    fq::count_launch();
    k_bgzf_gather<<<std::min<uint32_t>(n_blocks, (uint32_t)z->sm_count * 8u), 256, 0, st>>>(z->d_slots[k], z->d_sizes[k], z->d_offsets[k], n_blocks, d_out, out_cap);
    fq::count_launch();
    if (d_total) BZ_CU(cudaMemcpyAsync(d_total, z->d_offsets[k] + n_blocks, 8, cudaMemcpyDeviceToDevice, st));
    BZ_CU(cudaGetLastError());
    return FQTK_B200_OK;
}

extern "C" {

uint64_t fqtk_b200_bgzf_bound(uint64_t n_bytes) {
    const uint64_t blocks = (n_bytes + BZ_IN - 1) / BZ_IN;
    return n_bytes + blocks * (BZ_HDR + BZ_TRL + 5) + sizeof(BGZF_EOF);
}

int fqtk_b200_bgzf_create(int device, uint64_t chunk_bytes, fqtk_b200_bgzf** out) {
    if (!out) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_create: null output");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return bz_fail(FQTK_B200_ERR_CUDA, "bgzf_create: no such CUDA device (there is no CPU compressor in this library)");
    BZ_CU(cudaSetDevice(device));
    auto* z = new fqtk_b200_bgzf();
    z->device = device;
    BZ_CU(cudaDeviceGetAttribute(&z->sm_count, cudaDevAttrMultiProcessorCount, device));
    if (chunk_bytes == 0) chunk_bytes = 64ull << 20;
    z->chunk_blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(chunk_bytes / BZ_IN, 1u << 20));
    z->chunk_bytes = (uint64_t)z->chunk_blocks * BZ_IN;
    z->grid = 2u * (uint32_t)z->sm_count;  // two resident CTAs per SM
    z->out_cap = z->chunk_bytes + (uint64_t)z->chunk_blocks * (BZ_HDR + BZ_TRL + 5);
    BZ_CU(upload_tables());
    BZ_CU(cudaFuncSetAttribute(k_bgzf_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BzShared)));
    for (int k = 0; k < fqtk_b200_bgzf::NPIPE; k++) {
        BZ_CU(cudaMalloc(&z->d_in[k], z->chunk_bytes + 64));
        BZ_CU(cudaMalloc(&z->d_slots[k], (size_t)z->chunk_blocks * BZ_SLOT));
        BZ_CU(cudaMalloc(&z->d_out[k], z->out_cap));
        BZ_CU(cudaMalloc(&z->d_sizes[k], (size_t)z->chunk_blocks * 4));
        BZ_CU(cudaMalloc(&z->d_offsets[k], ((size_t)z->chunk_blocks + 1) * 8));
        BZ_CU(cudaHostAlloc(&z->h_total[k], 8, cudaHostAllocDefault));
        BZ_CU(cudaStreamCreateWithFlags(&z->stream[k], cudaStreamNonBlocking));
        BZ_CU(cudaEventCreateWithFlags(&z->done[k], cudaEventDisableTiming));
    }
    BZ_CU(cudaMalloc(&z->d_tokens, (size_t)fqtk_b200_bgzf::NPIPE * z->grid * BZ_WARPS * BZ_TOK_PER_WARP * 4));
    *out = z;
    return FQTK_B200_OK;
}

void fqtk_b200_bgzf_destroy(fqtk_b200_bgzf* z) {
    if (!z) return;
    cudaSetDevice(z->device);
    for (int k = 0; k < fqtk_b200_bgzf::NPIPE; k++) {
        if (z->stream[k]) cudaStreamSynchronize(z->stream[k]);
        cudaFree(z->d_in[k]); cudaFree(z->d_slots[k]); cudaFree(z->d_out[k]); cudaFree(z->d_sizes[k]); cudaFree(z->d_offsets[k]);
        if (z->h_total[k]) cudaFreeHost(z->h_total[k]);
        if (z->stream[k]) cudaStreamDestroy(z->stream[k]);
        if (z->done[k]) cudaEventDestroy(z->done[k]);
    }
    cudaFree(z->d_tokens);
    delete z;
}

/* device buffers in, device buffer out (the measured core): n_bytes <= the handle's chunk size */
int fqtk_b200_bgzf_compress_device(fqtk_b200_bgzf* z, const uint8_t* d_in, uint64_t n_bytes, int level, uint8_t* d_out,
                                   uint64_t out_capacity, uint64_t* d_out_bytes, void* stream) {
    if (!z || (!d_in && n_bytes) || !d_out || !d_out_bytes) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress_device: null argument");
    if (n_bytes > z->chunk_bytes) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress_device: more bytes than the handle's chunk size");
    if (level < 0 || level > 12) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress_device: compression level must be 0..12");
    BZ_CU(cudaSetDevice(z->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n_bytes == 0) {
        BZ_CU(cudaMemsetAsync(d_out_bytes, 0, 8, st));
        return FQTK_B200_OK;
    }
    return bgzf_launch(z, 0, d_in, n_bytes, level, d_out, out_capacity, reinterpret_cast<unsigned long long*>(d_out_bytes), st);
}

/* host buffers: chunks of the handle's size, H2D / kernels / D2H of consecutive chunks overlap on two streams */
int fqtk_b200_bgzf_compress(fqtk_b200_bgzf* z, const uint8_t* in, uint64_t n_bytes, int level, int write_eof, uint8_t* out,
                            uint64_t out_capacity, uint64_t* out_bytes) {
    if (!z || (!in && n_bytes) || !out || !out_bytes) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress: null argument");
    if (level < 0 || level > 12) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress: compression level must be 0..12");
    BZ_CU(cudaSetDevice(z->device));
    *out_bytes = 0;
    uint64_t written = 0;
    constexpr int NP = fqtk_b200_bgzf::NPIPE;
    bool pend[NP] = {};
    auto collect = [&](int k) -> int {  // wait for slot k's kernels, copy its compressed bytes out
        if (!pend[k]) return FQTK_B200_OK;
        BZ_CU(cudaEventSynchronize(z->done[k]));
        const uint64_t total = *z->h_total[k];
        if (written + total > out_capacity) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress: output buffer too small (see fqtk_b200_bgzf_bound)");
        BZ_CU(cudaMemcpyAsync(out + written, z->d_out[k], total, cudaMemcpyDeviceToHost, z->stream[k]));
        written += total;
        pend[k] = false;
        return FQTK_B200_OK;
    };
    int rc = FQTK_B200_OK;
    uint64_t off = 0;
    for (int k = 0; off < n_bytes && rc == FQTK_B200_OK; k = (k + 1) % NP) {
        // slot k's previous chunk must have left the device before its buffers are reused (its D2H is on stream k too)
        rc = collect(k);
        if (rc != FQTK_B200_OK) break;
        const uint64_t n = std::min<uint64_t>(z->chunk_bytes, n_bytes - off);
        cudaStream_t st = z->stream[k];
        BZ_CU(cudaMemcpyAsync(z->d_in[k], in + off, n, cudaMemcpyHostToDevice, st));
        rc = bgzf_launch(z, k, z->d_in[k], n, level, z->d_out[k], z->out_cap, nullptr, st);
        if (rc != FQTK_B200_OK) break;
        const uint32_t nb = (uint32_t)((n + BZ_IN - 1) / BZ_IN);
        BZ_CU(cudaMemcpyAsync(z->h_total[k], z->d_offsets[k] + nb, 8, cudaMemcpyDeviceToHost, st));
        BZ_CU(cudaEventRecord(z->done[k], st));
        pend[k] = true;
        off += n;
        // the chunks leave in input order: the OTHER slot holds the older chunk
        const int other = (k + 1) % NP;
        if (pend[other]) rc = collect(other);
    }
    if (rc == FQTK_B200_OK) {
        // whatever is still pending, oldest first
        uint64_t left = 0;
        for (int k = 0; k < NP; k++) left += pend[k];
        if (left == 2) return bz_fail(FQTK_B200_ERR_CUDA, "bgzf_compress: internal pipeline order error");
        for (int k = 0; k < NP && rc == FQTK_B200_OK; k++) rc = collect(k);
    }
    for (int k = 0; k < NP; k++) cudaStreamSynchronize(z->stream[k]);
    if (rc != FQTK_B200_OK) return rc;
    if (write_eof) {
        if (written + sizeof(BGZF_EOF) > out_capacity) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress: output buffer too small for the EOF block");
        std::memcpy(out + written, BGZF_EOF, sizeof(BGZF_EOF));
        written += sizeof(BGZF_EOF);
    }
    *out_bytes = written;
    return FQTK_B200_OK;
}

/* several writers' texts in one device buffer: segment k = d_in[seg_offsets[k], seg_offsets[k+1]) is cut every 65 280 bytes
 * on its own (a member never spans two writers); the members of all segments leave back to back in d_out and
 * out_seg_offsets[k] (host) says where segment k's members start (out_seg_offsets[n_segments] = total).  No EOF blocks. */
int fqtk_b200_bgzf_compress_segments_device(fqtk_b200_bgzf* z, const uint8_t* d_in, const uint64_t* seg_offsets,
                                            uint32_t n_segments, int level, uint8_t* d_out, uint64_t out_capacity,
                                            uint64_t* out_seg_offsets, void* stream) {
    if (!z || !seg_offsets || !out_seg_offsets || (n_segments && !d_out)) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress_segments_device: null argument");
    if (level < 0 || level > 12) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress_segments_device: compression level must be 0..12");
    BZ_CU(cudaSetDevice(z->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<BzBlockDesc> desc;
    std::vector<uint32_t> first_block(n_segments + 1, 0);
    for (uint32_t k = 0; k < n_segments; k++) {
        if (seg_offsets[k + 1] < seg_offsets[k]) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress_segments_device: offsets must not decrease");
        first_block[k] = (uint32_t)desc.size();
        for (uint64_t o = seg_offsets[k]; o < seg_offsets[k + 1]; o += BZ_IN)
            desc.push_back(BzBlockDesc{o, (uint32_t)std::min<uint64_t>(BZ_IN, seg_offsets[k + 1] - o), 0u});
    }
    first_block[n_segments] = (uint32_t)desc.size();
    for (uint32_t k = 0; k <= n_segments; k++) out_seg_offsets[k] = 0;
    const uint32_t nb = (uint32_t)desc.size();
    if (nb == 0) return FQTK_B200_OK;
    if (!d_in) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress_segments_device: null input");
    // scratch of this call: block list, slots, sizes, offsets (the handle's own buffers are sized for uniform chunks)
    fq::TempBuf t_desc, t_slots, t_sizes, t_offs;
    cudaError_t e = t_desc.alloc((size_t)nb * sizeof(BzBlockDesc), st);
    if (e == cudaSuccess) e = t_slots.alloc((size_t)nb * BZ_SLOT, st);
    if (e == cudaSuccess) e = t_sizes.alloc((size_t)nb * 4, st);
    if (e == cudaSuccess) e = t_offs.alloc(((size_t)nb + 1) * 8, st);
    BzBlockDesc* d_desc = t_desc.as<BzBlockDesc>();
    uint8_t* d_slots = t_slots.as<uint8_t>();
    uint32_t* d_sizes = t_sizes.as<uint32_t>();
    unsigned long long* d_offs = t_offs.as<unsigned long long>();
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_desc, desc.data(), (size_t)nb * sizeof(BzBlockDesc), cudaMemcpyHostToDevice, st);
    std::vector<unsigned long long> offs(nb + 1);
    if (e == cudaSuccess) {
        k_bgzf_deflate<<<std::min<uint32_t>(z->grid, nb), BZ_THREADS, sizeof(BzShared), st>>>(d_in, 0, nb, level, d_slots, d_sizes, z->d_tokens, d_desc);
        fq::count_launch();
        k_bgzf_scan<<<1, 1024, 0, st>>>(d_sizes, nb, d_offs);
        fq::count_launch();
        k_bgzf_gather<<<std::min<uint32_t>(nb, (uint32_t)z->sm_count * 8u), 256, 0, st>>>(d_slots, d_sizes, d_offs, nb, d_out, out_capacity);
        fq::count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(offs.data(), d_offs, ((size_t)nb + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bz_fail(FQTK_B200_ERR_CUDA, std::string("bgzf_compress_segments_device: ") + cudaGetErrorString(e));
    if (offs[nb] > out_capacity) return bz_fail(FQTK_B200_ERR_ARG, "bgzf_compress_segments_device: output buffer too small");
    for (uint32_t k = 0; k <= n_segments; k++) out_seg_offsets[k] = offs[first_block[k]];
    return FQTK_B200_OK;
}

uint64_t fqtk_b200_bgzf_chunk_bytes(const fqtk_b200_bgzf* z) { return z ? z->chunk_bytes : 0; }

}  // extern "C"
