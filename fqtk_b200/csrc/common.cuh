// common.cuh — encoding, bit-plane and hashing primitives shared by host and device code.
//
// Everything here restates *semantics* of the reference (fulcrumgenomics/fqtk @ 45dbb99), not its code:
//   encode_byte      : IUPAC_MASKS + byte_is_nocall + to_ascii_uppercase   src/lib/mod.rs:26-61,85-87
//   packed layout    : width-4 BitEnc, symbol i at bits 4*(i%8) of u32 i/8  src/lib/bitenc.rs:311-322
//   mismatch rule    : obs & !exp != 0 per symbol                           src/lib/bitenc.rs:441-452
//   decision rule    : best/next-best, max_mismatches, min_mismatch_delta   src/lib/barcode_matching.rs:119-160
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define FQ_HD __host__ __device__ __forceinline__
#define FQ_D __device__ __forceinline__

namespace fq {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t EMPTY_KEY = 0xFFFFFFFFu;  // running-min sentinel: no barcode seen yet
constexpr int MAX_FAST_WORDS = 4;            // L <= 32 keeps a read in <= 4 packed words / one 32-bit plane set

// ASCII byte -> 4-bit base set {A=1,C=2,G=4,T=8}.  No-calls 'N','n','.' -> 15; letters are upper-cased first;
// 'U' = T; IUPAC degenerate codes are unions; every other byte -> 0 (which can never mismatch).
FQ_HD uint32_t encode_byte(uint32_t b) {
    if (b == '.') return 15u;
    if (b >= 'a' && b <= 'z') b -= 32u;
    switch (b) {
        case 'A': return 1u;
        case 'C': return 2u;
        case 'G': return 4u;
        case 'T': return 8u;
        case 'U': return 8u;
        case 'M': return 3u;   // A|C
        case 'R': return 5u;   // A|G
        case 'W': return 9u;   // A|T
        case 'S': return 6u;   // C|G
        case 'Y': return 10u;  // C|T
        case 'K': return 12u;  // G|T
        case 'V': return 7u;   // A|C|G
        case 'H': return 11u;  // A|C|T
        case 'D': return 13u;  // A|G|T
        case 'B': return 14u;  // C|G|T
        case 'N': return 15u;
        default: return 0u;
    }
}

FQ_HD bool byte_is_nocall(uint32_t b) { return b == 'N' || b == 'n' || b == '.'; }

FQ_HD uint32_t words_for_len(uint32_t L) { return (L + 7u) / 8u; }
FQ_HD uint32_t planes_for_len(uint32_t L) { return (L + 31u) / 32u; }

// Bit k of each of the 8 nibbles of x, gathered into the low 8 bits (nibble i -> bit i).
FQ_HD uint32_t gather_nibble_bit(uint32_t x, int k) {
    uint32_t y = (x >> k) & 0x11111111u;
    y = (y | (y >> 3)) & 0x03030303u;
    y = (y | (y >> 6)) & 0x000F000Fu;
    y = (y | (y >> 12)) & 0x000000FFu;
    return y;
}

// Four "has base X" bit-planes (X = A,C,G,T) of up to 32 symbols held in up to four packed words.
template <int W>
FQ_HD void planes_from_words(const uint32_t (&w)[W], uint32_t (&pl)[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < W; i++) v |= gather_nibble_bit(w[i], k) << (8 * i);
        pl[k] = v;
    }
}

// True when every symbol of the read is one of A,C,G,T,N (masks 1,2,4,8,15) — the alphabet the memo table
// enumerates.  `pad_or` has 0x1 in every nibble beyond the barcode length so padding never trips the test.
// Per nibble: bit 0 of the result is set iff the nibble's popcount is 1 or 4 (other result bits are garbage).
// With a,b,c,d = the nibble's four bits: parity3 = a^b^c, maj3 = majority(a,b,c);
//   exactly one of four = (parity3 & ~maj3 & ~d) | (~parity3 & ~maj3 & d);  all four = parity3 & maj3 & d.
FQ_HD uint32_t nibble_ok_bits(uint32_t x) {
    const uint32_t b = x >> 1, c = x >> 2, d = x >> 3;
    const uint32_t par = x ^ b ^ c;
    const uint32_t maj = (x & b) | (x & c) | (b & c);
    return (par & ~maj & ~d) | (~par & ~maj & d) | (par & maj & d);
}

template <int W>
FQ_HD bool read_in_table_alphabet(const uint32_t (&w)[W], uint32_t last_word_pad) {
    uint32_t ok = 0x11111111u;
#pragma unroll
    for (int i = 0; i < W; i++) ok &= nibble_ok_bits(i == W - 1 ? (w[i] | last_word_pad) : w[i]);
    return ok == 0x11111111u;
}

FQ_HD uint32_t last_word_pad_for_len(uint32_t L) {
    const uint32_t used = L % 8u;
    return used == 0u ? 0u : (0x11111111u << (4u * used));
}

// ---- compressed A/C/G/T keys (k_probe3) -------------------------------------------------------------------
// A read of L <= 16 symbols that are all one of A,C,G,T (one-hot nibbles 1,2,4,8) is mapped injectively onto 32
// bits, two per symbol, with no table: word 0 keeps (n0^n1, n1^n2) in bits {0,1} of every nibble, word 1 keeps
// (n1^n2, n2^n3) in bits {2,3} — a Gray code of the bit position.  `valid` is false for every other read (any
// nibble that is not exactly one bit: no-calls, IUPAC codes, junk), which must not use the key.
//   one-hot test per word: d = x - 0x11111111; every nibble is one-hot  <=>  d & (x | 0x88888888) == 0
//   (no zero nibble -> no borrows -> (n-1) & n == 0 per nibble; the lowest zero nibble always leaves its bit 3 set).
// `pad` (last_word_pad_for_len) puts an 'A' in the unused nibbles of the last word so they pass the test.
template <int W>
FQ_HD uint32_t acgt_key(const uint32_t (&w)[W], uint32_t pad, bool& valid) {
    static_assert(W == 1 || W == 2, "compressed keys cover L <= 16");
    const uint32_t x0 = (W == 1) ? (w[0] | pad) : w[0];
    uint32_t bad = (x0 - 0x11111111u) & (x0 | 0x88888888u);
    uint32_t k = (x0 ^ (x0 >> 1)) & 0x33333333u;
    if constexpr (W == 2) {
        const uint32_t x1 = w[W > 1 ? 1 : 0] | pad;
        bad |= (x1 - 0x11111111u) & (x1 | 0x88888888u);
        k |= (x1 ^ (x1 << 1)) & 0xCCCCCCCCu;
    }
    valid = bad == 0u;
    return k;
}

// ---- k_probe4's fingerprint table (kernels.h) ----------------------------------------------------------------
// One-hot test of a whole read: every nibble of every word exactly one bit (pure A/C/G/T).  `pad` = last_word_pad_for_len.
template <int W>
FQ_HD bool acgt_only(const uint32_t (&w)[W], uint32_t pad) {
    uint32_t bad = 0u;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const uint32_t x = (i == W - 1) ? (w[i] | pad) : w[i];
        bad |= (x - 0x11111111u) & (x | 0x88888888u);
    }
    return bad == 0u;
}
// True when every symbol of the read is one of A,C,G,T,N (masks 1,2,4,8,15) — the same predicate as
// read_in_table_alphabet, in the form that costs the fewest instructions on the device: the all-ones nibbles are found
// (x & x>>1 & x>>2 & x>>3), rewritten to a one-hot nibble, and the word then takes the one-hot test of acgt_only.
// `pad` (last_word_pad_for_len) makes the unused nibbles of the last word pass.
template <int W>
FQ_HD bool acgtn_only(const uint32_t (&w)[W], uint32_t pad) {
    uint32_t bad = 0u;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const uint32_t x = (i == W - 1) ? (w[i] | pad) : w[i];
        const uint32_t t = x & (x >> 1);
        const uint32_t f = t & (t >> 2) & 0x11111111u;  // 1 in every nibble that is 0xF
        const uint32_t y = x ^ (f * 14u);                // 0xF -> 0x1
        bad |= (y - 0x11111111u) & (y | 0x88888888u);
    }
    return bad == 0u;
}

// Home bucket and fingerprint hash of a read's packed words: two independent 32-bit mixes, so that the fingerprint bits
// are not a function of the bucket.  Every word is first folded onto itself (w ^ w >> 15): a plain multiply-add over the
// words is LINEAR in their top nibbles (w << 28 times an odd constant keeps 4 bits), which made reads that differ only
// in symbols 7 / 15 / 23 / 31 collide in bucket AND fingerprint systematically; after the fold such a difference reaches
// 19 bits of the product.  (Exactness never rests on the hash: see kernels.h; a weak hash only costs re-seeds.)
// Identical on host (table build) and device (probe).  The fingerprint of an entry is the TOP fp_bits of `fp_hash`.
template <int W>
FQ_HD void g4_hashes(const uint32_t (&w)[W], uint32_t seed, uint32_t n_buckets, uint32_t& bucket, uint32_t& fp_hash) {
    uint32_t f[W];
#pragma unroll
    for (int i = 0; i < W; i++) f[i] = w[i] ^ (w[i] >> 15);
    uint32_t h1 = f[0] * 0x9E3779B1u + seed, h2 = f[0] * 0x165667B1u + seed;
    if constexpr (W > 1) { h1 += f[W > 1 ? 1 : 0] * 0x85EBCA77u; h2 += f[W > 1 ? 1 : 0] * 0xD3A2646Du; }
    if constexpr (W > 2) { h1 += f[W > 2 ? 2 : 0] * 0xC2B2AE3Du; h2 += f[W > 2 ? 2 : 0] * 0xFD7046C5u; }
    if constexpr (W > 3) { h1 += f[W > 3 ? 3 : 0] * 0x27D4EB2Fu; h2 += f[W > 3 ? 3 : 0] * 0xB55A4F09u; }
    h1 ^= h1 >> 15;
    h1 *= 0x2C1B3C6Du;
    h1 ^= h1 >> 12;
    h2 ^= h2 >> 16;
    h2 *= 0x45D9F3B5u;
    bucket = (uint32_t)(((uint64_t)h1 * (uint64_t)n_buckets) >> 32);
    fp_hash = h2;
}
// Mismatches between a read and one barcode from the barcode's ~expected nibble words (bitenc.rs:441-452): a nibble of
// r & ~exp is non-zero iff the symbol mismatches; the flags of up to four words share one popcount.
template <int W>
FQ_HD uint32_t nibble_distance(const uint32_t (&w)[W], const uint32_t (&ne)[W]) {
    static_assert(W <= 4, "one popcount covers four words");
    uint32_t acc = 0u;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const uint32_t x = w[i] & ne[i];
        const uint32_t f = (((x & 0x77777777u) + 0x77777777u) | x) & 0x88888888u;  // bit 3 of every non-zero nibble
        acc |= f >> (3 - i);
    }
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(acc);
#else
    return (uint32_t)__builtin_popcount(acc);
#endif
}

// Fixed odd multipliers of the (up to three) cuckoo sub-tables: slot_i = (k * CK_MUL[i]) >> (32 - sb_i), and the low
// 32 - sb_i bits of the same product are the remainder stored in the entry (multiplication by an odd constant is a
// bijection of 32-bit keys, so slot + remainder identify the key exactly).
constexpr uint32_t CK_MUL0 = 0x9E3779B1u, CK_MUL1 = 0x85EBCA77u, CK_MUL2 = 0xC2B2AE3Du;
FQ_HD uint32_t ck_mul(int i) { return i == 0 ? CK_MUL0 : (i == 1 ? CK_MUL1 : CK_MUL2); }

// 32-bit mix of a W-word key; identical on host (table build) and device (probe): multiply-add over the words
// (FMA pipe), then one xorshift-multiply round so that every output bit depends on every input nibble.
template <int W>
FQ_HD uint32_t hash_key(const uint32_t (&w)[W]) {
    uint32_t h = w[0] * 0x9E3779B9u;
    if constexpr (W > 1) h += w[W > 1 ? 1 : 0] * 0x85EBCA6Bu;
    if constexpr (W > 2) h += w[W > 2 ? 2 : 0] * 0xC2B2AE35u;
    if constexpr (W > 3) h += w[W > 3 ? 3 : 0] * 0x27D4EB2Fu;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

// Memo-table bucket (global memory) of a key, from its 32-bit hash.
FQ_HD uint32_t bucket_of_hash(uint32_t h, uint32_t n_buckets) {                    // fast range: [0, n_buckets)
    return (uint32_t)(((uint64_t)(h * 0x9E3779B1u) * (uint64_t)n_buckets) >> 32);
}
// Hot-tier slots: two cheap multiply-add hashes, top bits (IMAD runs on the FMA pipe, off the ALU pipe the
// compares live on).  slot = hash >> tier_shift.
template <int W>
FQ_HD uint32_t tier_hash(const uint32_t (&w)[W], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t h = w[0] * c0;
    if constexpr (W > 1) h += w[W > 1 ? 1 : 0] * c1;
    if constexpr (W > 2) h += w[W > 2 ? 2 : 0] * c2;
    if constexpr (W > 3) h += w[W > 3 ? 3 : 0] * c3;
    return h;
}
template <int W>
FQ_HD uint32_t tier_hash1(const uint32_t (&w)[W]) { return tier_hash<W>(w, 0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu); }
template <int W>
FQ_HD uint32_t tier_hash2(const uint32_t (&w)[W]) { return tier_hash<W>(w, 0x165667B1u, 0xD3A2646Du, 0xFD7046C5u, 0xB55A4F09u); }


// Blocked Bloom filter: one 32-bit word per key (top hash bits), three bit positions from the low hash bits.
FQ_HD uint32_t bloom_mask(uint32_t h) { return (1u << (h & 31u)) | (1u << ((h >> 5) & 31u)) | (1u << ((h >> 10) & 31u)); }

// Running best / second-best on keys (distance << 16 | sample index): keys are unique per sample, so plain
// min / second-min on the key gives min distance, FIRST index among ties (strict '<' at
// barcode_matching.rs:132) and the second-smallest distance with multiplicity (:140).
FQ_HD void track2(uint32_t& k1, uint32_t& k2, uint32_t key) {
    const uint32_t hi = k1 > key ? k1 : key;
    k1 = k1 < key ? k1 : key;
    k2 = k2 < hi ? k2 : hi;
}

FQ_HD void merge2(uint32_t& k1, uint32_t& k2, uint32_t o1, uint32_t o2) {
    const uint32_t hi = k1 > o1 ? k1 : o1;
    const uint32_t lo2 = k2 < o2 ? k2 : o2;
    k1 = k1 < o1 ? k1 : o1;
    k2 = hi < lo2 ? hi : lo2;
}

// barcode_matching.rs:149-159.  k2 == EMPTY_KEY means a single-sample panel: next_best stays 255 (:122).
FQ_HD uint32_t decide(uint32_t k1, uint32_t k2, uint32_t max_mismatches, uint32_t min_delta) {
    const uint32_t best = k1 >> 16, idx = k1 & 0xFFFFu;
    const uint32_t next = (k2 == EMPTY_KEY) ? 255u : (k2 >> 16);
    if (best > max_mismatches || (next - best) < min_delta) return NONE;
    return (idx << 16) | (best << 8) | next;
}

}  // namespace fq
