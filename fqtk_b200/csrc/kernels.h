// kernels.h — host-side launch interface of the sm_100a kernels (internal; the public ABI is include/fqtk_b200.h).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fq {

// Device-resident description of a matcher's panel and decision parameters.
struct MatchParams {
    const uint4* planes;       // [S * P] "forbidden base" bit-planes per barcode: .x/.y/.z/.w bit i set iff position
                               //  32*p + i of the barcode does NOT admit A/C/G/T
    const uint4* planes2;      // L <= 16 only: [ceil(S/2)] the planes of barcodes 2q (low 16 bits) and 2q+1 (high 16 bits)
                               //  packed into one word each; planes[] itself is padded to an even number of entries
    const uint32_t* not_exp;   // [S * W] ~expected nibble words (only read by the L > 32 kernel)
    const uint32_t* table;     // memo table slots in global memory (nullptr in brute mode)
    const uint32_t* tier_entries;  // hot tier (table entries whose best distance is 0), staged into shared memory by
                                   //   k_probe2: 2-choice cuckoo, tier_slots entries of tier_entry_words(W) words
                                   //   (W = 2, 4: tier_slots key entries, then tier_slots value words)
    const uint32_t* bloom;         // blocked Bloom filter over every memo-table key (one 32-bit word, 3 bits per key)
    unsigned long long* counts;  // [S + 1] per-sample counts, last = unmatched
    uint32_t S, L, W, P;
    uint32_t max_mm, min_delta;
    uint32_t last_pad;         // 0x1 in every padding nibble of the last packed word
    uint32_t n_buckets;        // memo-table slots
    uint32_t tier_slots;       // power of two, 0 = no hot tier
    uint32_t tier_shift;       // 32 - log2(tier_slots); 32 = no hot tier
    uint32_t tier_rep;         // shared-memory replicas of the hot tier (power of two): lane l reads replica l % rep, so
                               //   lanes of a warp spread over distinct banks whatever slots they probe
    uint32_t hist_rep;         // shared-memory replicas of the k_probe2 histogram (power of two <= 32)
    uint32_t bloom_words;      // power of two, 0 = no filter
    uint32_t bloom_shift;      // 32 - log2(bloom_words)
    // k_probe3 (L <= 16): every memo-table entry whose key is pure A/C/G/T, as a 2- or 3-ary cuckoo table of 4-byte
    // quotient entries staged into shared memory (layout: CuckooLayout below)
    const uint32_t* ck_entries;  // ck_words words: sub-table 0 | sub-table 1 | (sub-table 2); nullptr = none
    uint32_t ck_np;              // sub-tables (= probes per read): 0 = no cuckoo table, 2 or 3
    uint32_t ck_words;           // total entries
    uint32_t ck_off[3];          // first entry of sub-table i
    uint32_t ck_shift[3];        // 32 - sb_i  (sub-table i has 2^sb_i slots)
    uint32_t ck_negmulb[3];      // -(ck_mul(i) << sb_i): entry + k * ck_negmulb[i] = entry - (the key's remainder << sb_i)
    uint32_t ck_limit;           // 2^cb - 1: a probe found its key iff entry - remainder < ck_limit (then it is the code)
    uint32_t ck_lb;              // value code = idx << lb | best << nb | (next - ck_next_min), lb = bb + nb
    uint32_t ck_bsh, ck_bmask8;  // (code << ck_bsh) & ck_bmask8 = best << 8   (ck_bsh = 8 - nb)
    uint32_t ck_nmask;           // code & ck_nmask = next - ck_next_min
    uint32_t ck_next_min;
    uint32_t ck_stash_cap;       // stash entries per warp of k_probe3 (16 .. 64)
    // k_probe4 (L <= 24): every pure-A/C/G/T memo entry under its compressed key in 8-byte slots {key lo, hi word},
    // grouped in 32-byte BUCKETS of four (one L2 sector per probe), bucketised linear probing in GLOBAL memory at load
    // <= 0.6, small enough to stay L2-resident.  hi word = the result word itself for L <= 16, (key hi << g4_vb) |
    // value code (k_probe3's layout, ck_lb .. ck_next_min) for L > 16; an empty slot is all ones; the slots of a
    // bucket fill in order, so a bucket is full iff its last slot is taken.
    const uint2* g4_table;       // nullptr = none
    uint32_t g4_buckets;
    uint32_t g4_himask;          // low 2 * (L - 16) bits (0 for L <= 16)
    uint32_t g4_vb;              // value-code bits (L > 16)
    uint32_t g4_limit;           // 2^g4_vb - 1: the reserved (largest) code
    uint32_t g4_hist_rep;        // histogram replicas of k_probe4 (power of two <= 16)
    uint32_t g4_stash_cap;       // stash entries per warp of k_probe4
    uint32_t g4_kernel;          // 1: the packed route runs k_probe4; 0: k_probe2, whose queue phase probes this table
    uint32_t ck_one, ck_four;    // 1 and 4 (see Probe3Ctx in match_kernels.cu)
};

// Cuckoo entry (4 bytes) of sub-table i: (remainder << sb_i) | code, where remainder = low (32 - sb_i) bits of
// k * ck_mul(i), code < 2^cb - 1 <= 2^sb_i - 1.  Empty = 0xFFFFFFFF (its code field is all ones, which no real entry
// uses, so an empty slot can never be taken for a key).  entry - (remainder' << sb_i) equals the code when the
// remainders agree and is >= 2^sb_i otherwise.

// Where a batch of reads lives on the device.
struct ReadSource {
    const uint32_t* packed;    // n * W words, or nullptr
    const uint8_t* ascii;      // n rows of `stride` bytes, or nullptr
    const uint32_t* lengths;   // optional per-row lengths for the ASCII form
    uint64_t stride;
    uint64_t n;                // < 2^32
};

// Fixed-length barcode segments scattered over several row sources (demux.rs:121-123 concatenation order).
struct SegmentSource {
    const uint8_t* base[8];
    uint64_t stride[8];
    uint32_t offset[8];
    uint32_t length[8];
    uint32_t n_segments;
};

struct LaunchGeometry {
    int sm_count;
    int max_smem_optin;
};

// Memo-table geometry: open addressing, linear probing, one slot per probe.
//   W <= 3: 16-byte slots {k0, k1, k2, value} (missing key words are 0), one LDG.128 per probe;
//   W == 4: 32-byte slots {k0..k3, value, 0, 0, 0}, one 256-bit load per probe.
// Empty slots are all-ones (value NONE).  `n_buckets` is the slot count.
inline __host__ __device__ int table_slot_words(int W) { return W <= 3 ? 4 : 8; }
inline __host__ __device__ int table_value_index(int W) { return W <= 3 ? 3 : 4; }
// Hot-tier entries: W = 1: {k0, value} (LDS.64); W = 2: {k0, k1} (LDS.64) with the values in a separate array;
//   W = 3: {k0, k1, k2, value} (LDS.128); W = 4: {k0, k1, k2, k3} (LDS.128) with the values in a separate array.
//   Measured on B200 (tools/microbench_lds.cu): a conflict-free LDS.64 costs 2.4 clk per warp, an LDS.128 8.0 clk, a
//   random LDS.32 2.5 clk — so 8-byte probes plus one value fetch (7.1 clk) beat two 16-byte probes (16 clk).
//   Empty slots are all-ones (value NONE).
inline __host__ __device__ int tier_entry_words(int W) { return W <= 2 ? 2 : 4; }
inline __host__ __device__ bool tier_separate_values(int W) { return W == 2 || W == 4; }
inline __host__ __device__ int tier_value_index(int W) { return W == 1 ? 1 : 3; }  // fused layouts only (W = 1, 3)
inline __host__ __device__ int tier_max_rep(int W) { return W <= 2 ? 16 : 8; }  // replicas that tile all 32 banks once

cudaError_t launch_brute(const MatchParams& p, const ReadSource& src, uint32_t* d_results, const LaunchGeometry& g,
                         cudaStream_t stream);
cudaError_t launch_probe(const MatchParams& p, const ReadSource& src, uint32_t* d_results, const LaunchGeometry& g,
                         cudaStream_t stream);
cudaError_t launch_pack(const uint8_t* d_ascii, uint64_t n, uint32_t L, uint64_t stride, uint32_t* d_packed,
                        const LaunchGeometry& g, cudaStream_t stream);
cudaError_t launch_pack_segments(const SegmentSource& seg, uint64_t n, uint32_t L, uint32_t* d_packed,
                                 const LaunchGeometry& g, cudaStream_t stream);
size_t route_workspace_bytes(uint64_t n, uint32_t S, const LaunchGeometry& g);
bool route_supported(uint32_t S, const LaunchGeometry& g);
cudaError_t launch_route(const uint32_t* d_results, uint64_t n, uint32_t S, uint32_t* d_order,
                         unsigned long long* d_offsets, void* d_workspace, const LaunchGeometry& g, cudaStream_t stream);
cudaError_t prepare_kernels(const LaunchGeometry& g);  // opt-in shared memory attributes, once per device

size_t probe3_smem_bytes(uint32_t ck_words, uint32_t S, uint32_t stash_cap);
size_t probe4_smem_bytes(uint32_t W, uint32_t S, uint32_t hist_rep, uint32_t stash_cap);
size_t probe2_fixed_smem_bytes(uint32_t W, uint32_t S, int threads);  // k_probe2 shared memory besides tier + Bloom
int probe2_threads();
uint32_t probe2_hist_rep(uint32_t S);

uint64_t kernel_launches();
void count_launch();

}  // namespace fq
