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
    const uint32_t* not_exp;   // [S * W] ~expected nibble words (the L > 32 kernel; k_probe4's verification)
    const uint32_t* sliced;    // k_brute_sliced (L <= 32): [G][8W][16] words, G = ceil(S/32) groups of 32 barcodes: bit j of
                               //   word (g, i, v) is set iff barcode 32g + j mismatches a read symbol with 4-bit mask v at
                               //   position i (positions >= L: 0); nullptr = not built
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
    uint32_t ck_exact_only;      // 1: the cuckoo table holds only the best-distance-0 entries (k_probe5: the rest of the
                                 //    neighbourhood is in the fingerprint table below)
    // k_probe4 (L <= 32): FINGERPRINT table in global memory of every candidate string (every A/C/G/T/N string within
    // max_mm of some barcode, whether its result is Some or None), keyed by the read's packed words: 4-byte entries
    //     fingerprint << (ib + cb) | sample index << cb | value code,
    // eight to a 32-byte bucket (ONE 256-bit load per probe: a random lookup costs the SM's L1 one clock per lane whatever
    // its width, tools/microbench_gather.cu), bucketised linear probing, slots of a bucket filled in order, empty = all
    // ones.  Half the bytes per entry of an exact-key table, so the whole table stays L2-resident next to the stream.
    // Exactness does not rest on the hash: (i) only reads over A/C/G/T/N use the table at all, every other read takes the
    // exact slow path; (ii) a fingerprint match is VERIFIED: the read must be within max_mm of the entry's barcode (its
    // ~expected nibble words are staged in shared memory) — an A/C/G/T/N read that is NOT a table key is farther than
    // max_mm from every barcode, so it passes for no entry; (iii) a table key that meets another key's entry either fails
    // the verification (slow path) or is caught by the builder, which replays the kernel's lookup for EVERY key and
    // re-seeds the hash until each one yields its own value or the slow path.
    // value code = best << nb | (next - next_min); 2^cb - 2 = "candidate whose result is None"; 2^cb - 1 only in empty slots.
    const uint32_t* g4_table;    // g4_buckets * 8 entries; nullptr = none
    uint32_t g4_buckets;
    uint32_t g4_seed;
    uint32_t g4_fp_bits;         // fingerprint bits (>= 8)
    uint32_t g4_fp_shift;        // ib + cb
    uint32_t g4_cb;              // value-code bits
    uint32_t g4_nb;              // bits of (next - next_min)
    uint32_t g4_fp_mask, g4_lim; // derived: ~(2^fp_shift - 1), 2^fp_shift
    uint32_t g4_cmask, g4_nmask; // derived: 2^cb - 1, 2^nb - 1
    uint32_t g4_code_none;       // derived: 2^cb - 2
    uint32_t g4_next_min;
    uint32_t g4_hist_rep;        // histogram replicas of k_probe4 (power of two <= 16)
    uint32_t g4_stash_cap;       // stash entries per warp of k_probe4
    uint32_t g4_flags;           // A/B switches: 1 = table loads evict_last, 2 = stream evict_first, 4 = L2 prefetch of the next tile, bits 4-5 = launch shape
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

// Barcode segments taken straight out of raw FASTQ chunks (demux.rs:288-342 + :121-123): read i's sequence line starts at
// base[s] + seq_offsets[s][i] in source s; segment k of that source = bytes [offset, offset + length) of the line, or
// [offset, end of line) when length == REST (a trailing `+B`), which needs seq_lengths[s].
struct OffsetSource {
    static constexpr uint32_t REST = 0xFFFFFFFFu;
    const uint8_t* base[8];
    const uint64_t* seq_offsets[8];
    const uint32_t* seq_lengths[8];  // may be nullptr when no segment of the source is REST
    uint32_t source_of[8];           // segment -> source
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
cudaError_t launch_pack_offsets(const OffsetSource& seg, uint64_t n, uint32_t L, uint32_t* d_packed, uint32_t* d_lengths,
                                const LaunchGeometry& g, cudaStream_t stream);
// Stream-ordered temporaries (cudaMallocAsync from the device's default pool, which is told once to keep what it is given
// back instead of returning it to the driver at every synchronisation): the per-call scratch of the scanner, the record
// writer and the segmented BGZF call costs microseconds, not a cudaMalloc / cudaFree pair per buffer.
inline cudaError_t temp_alloc(void** p, size_t bytes, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured[dev] = true;
    }
    return cudaMallocAsync(p, bytes ? bytes : 1, stream);
}
struct TempBuf {  // freed in stream order when it goes out of scope
    void* p = nullptr;
    cudaStream_t stream = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t st) {
        stream = st;
        return temp_alloc(&p, bytes, st);
    }
    template <typename T>
    T* as() const { return static_cast<T*>(p); }
    ~TempBuf() {
        if (p) cudaFreeAsync(p, stream);
    }
};

// device FASTQ scanner (ingest_kernels.cu)
uint32_t fastq_scan_tiles(uint64_t bytes);
cudaError_t launch_nl_count(const uint8_t* d_chunk, uint64_t bytes, uint32_t* d_tile_counts, unsigned long long* d_prefix,
                            cudaStream_t stream);
cudaError_t launch_fq_records(const uint8_t* d_chunk, uint64_t bytes, const unsigned long long* d_prefix, uint64_t max_nl,
                              unsigned long long* d_nl, uint64_t n_records, unsigned long long* d_head_offsets,
                              unsigned long long* d_seq_offsets, uint32_t* d_seq_lengths, unsigned long long* d_err,
                              const LaunchGeometry& g, cudaStream_t stream);
cudaError_t launch_min_len(const uint32_t* d_seq_lengths, uint64_t n, uint32_t min_len, unsigned long long* d_err,
                           const LaunchGeometry& g, cudaStream_t stream);
cudaError_t launch_fq_vet(const OffsetSource& os, uint64_t n, uint32_t L, uint32_t max_nocalls, unsigned long long* d_err2,
                          const LaunchGeometry& g, cudaStream_t stream);
cudaError_t launch_fix_lengths(uint32_t* d_results, const uint32_t* d_lengths, uint64_t n, uint32_t L, uint32_t S,
                               unsigned long long* d_counts, const LaunchGeometry& g, cudaStream_t stream);
cudaError_t launch_narrow_u16(const uint32_t* d_results, uint64_t n, uint16_t* d_out, const LaunchGeometry& g,
                              cudaStream_t stream);
cudaError_t prepare_kernels(const LaunchGeometry& g);  // opt-in shared memory attributes, once per device

size_t probe3_smem_bytes(uint32_t ck_words, uint32_t S, uint32_t stash_cap);
size_t probe4_smem_bytes(uint32_t W, uint32_t S, uint32_t hist_rep, uint32_t stash_cap);
size_t probe5_smem_bytes(uint32_t ck_words, uint32_t S, uint32_t W, uint32_t hist_rep);
inline __host__ __device__ uint32_t probe4_ne_stride(uint32_t W) { return W == 3u ? 4u : W; }  // words per barcode in shared memory
size_t probe2_fixed_smem_bytes(uint32_t W, uint32_t S, int threads);  // k_probe2 shared memory besides tier + Bloom
int probe2_threads();
uint32_t probe2_hist_rep(uint32_t S);

uint64_t kernel_launches();
void count_launch();

}  // namespace fq
