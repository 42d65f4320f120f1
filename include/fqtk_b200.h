/*
 * fqtk_b200.h — C ABI of the B200-native (sm_100a) replacement for fqtk's `demux` barcode-matcher hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, never unwinds.  It replaces
 * the public Rust API of `fqtk_lib::barcode_matching` (reference paths relative to the fqtk repo root):
 *
 *   BarcodeMatcher::new(samples, max_mismatches, min_mismatch_delta, use_cache)   src/lib/barcode_matching.rs:55-86
 *   BarcodeMatcher::assign(&mut self, read_bases) -> Option<BarcodeMatch>          src/lib/barcode_matching.rs:165-186
 *   struct BarcodeMatch { best_match, best_mismatches, next_best_mismatches }      src/lib/barcode_matching.rs:16-25
 *   the caller's count rule (templates += 1 per Some / per None)                  src/bin/commands/demux.rs:968-975
 *   encode() / IUPAC_MASKS / byte_is_nocall                                        src/lib/mod.rs:26-61,85-87
 *
 * The only change asked of the caller (demux.rs:945-977) is to collect K read-sets, make ONE batch call,
 * and route the results in index order.  INTEGRATION.md shows the Rust `extern "C"` block and safe wrapper.
 *
 * Threading: a handle is NOT thread-safe (same contract as `&mut self`); one handle per device.
 * Errors: every function returns FQTK_B200_OK or a negative code; fqtk_b200_last_error() (thread-local)
 * holds the text, which for the reference's panics is the reference's own panic message.
 * There is no CPU fallback: every entry point that matches reads needs a CUDA device.
 */
#ifndef FQTK_B200_H
#define FQTK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FQTK_B200_API
#else
#define FQTK_B200_API __attribute__((visibility("default")))
#endif

/* ---- result word: one u32 per read (replaces Option<BarcodeMatch>, barcode_matching.rs:16-25) ----
 *   FQTK_B200_NONE                                   Option::None
 *   (best_match << 16) | (best_mismatches << 8) | next_best_mismatches      Some(..)
 * best_match < S <= 65535; next_best_mismatches keeps the reference's 255 sentinel for a 1-sample panel. */
#define FQTK_B200_NONE 0xFFFFFFFFu
#define FQTK_B200_BEST_MATCH(w) ((uint32_t)(w) >> 16)
#define FQTK_B200_BEST_MISMATCHES(w) (((uint32_t)(w) >> 8) & 0xFFu)
#define FQTK_B200_NEXT_BEST_MISMATCHES(w) ((uint32_t)(w) & 0xFFu)

#define FQTK_B200_MAX_SAMPLES 65535u
#define FQTK_B200_MAX_BARCODE_LEN 254u /* u8 distances with a 255 sentinel, as in the reference */

/* ---- status codes ---- */
#define FQTK_B200_OK 0
#define FQTK_B200_ERR_EMPTY_PANEL (-1)   /* panic "Must provide at least one sample"       barcode_matching.rs:61    */
#define FQTK_B200_ERR_EMPTY_BARCODE (-2) /* panic "Sample barcode cannot be empty string"  barcode_matching.rs:62-65 */
#define FQTK_B200_ERR_LENGTH (-3)        /* panic "Read barcode (..) length (..) differs.." barcode_matching.rs:95-106 */
#define FQTK_B200_ERR_ARG (-4)
#define FQTK_B200_ERR_CUDA (-5)
#define FQTK_B200_ERR_UNSUPPORTED (-6)   /* S > 65535 or L > 254 */

typedef struct fqtk_b200_matcher fqtk_b200_matcher;

/* which kernel family a matcher runs */
#define FQTK_B200_MODE_BRUTE 1 /* every read x every barcode, bit-plane popcount, thread per read */
#define FQTK_B200_MODE_TABLE 2 /* pre-computed memo table of the <= max_mismatches neighbourhood, warp-cooperative
                                  brute force for reads outside the table's alphabet */

typedef struct {
    uint32_t n_samples;          /* S */
    uint32_t barcode_len;        /* L */
    uint32_t words_per_read;     /* W = ceil(L/8): u32 blocks of the 4-bit BitEnc layout (bitenc.rs:311-322) */
    uint32_t max_ns_in_barcodes; /* barcode_matching.rs:73-74 */
    uint32_t mode;               /* FQTK_B200_MODE_* */
    uint32_t device;
    uint64_t table_entries;      /* memo-table entries (0 in brute mode) */
    uint64_t table_slots;
    uint64_t table_bytes;
    uint64_t table_candidates;   /* neighbourhood strings enumerated before filtering to Some(..) */
    uint64_t tier_entries;       /* memo-table entries (best distance 0) also cached in the shared-memory hot tier */
    uint64_t tier_slots;
    uint64_t cuckoo_entries;     /* pure-A/C/G/T memo entries held in the shared-memory cuckoo table (0 = none) */
    uint32_t cuckoo_probes;      /* sub-tables = probes per read (2 or 3) */
    uint32_t cuckoo_slots;       /* 4-byte slots over all sub-tables */
    uint64_t l2_table_entries;   /* pure-A/C/G/T candidate strings in the L2-resident fingerprint table of k_probe4 (0 = none) */
    uint64_t l2_table_bytes;
    uint64_t l2_table_slow_keys; /* of those, keys the kernel sends to the exact slow path (fingerprint collisions) */
} fqtk_b200_matcher_info;

/* ---- lifecycle -------------------------------------------------------------------------------------
 * BarcodeMatcher::new (barcode_matching.rs:55-86).  `panel_ascii` = S rows of L bytes, row-major, in sample-sheet
 * order (first-index tie-break depends on it).  Bytes are upper-cased (:71) and encoded exactly as the reference's
 * encode() does; the panel is copied, the caller may free it on return.
 * `use_cache`: the reference's memo-cache switch (demux.rs:925 passes true).  Non-zero builds the device memo table
 * when the panel's <= max_mismatches neighbourhood fits `fqtk_b200_set_table_budget` (default 32 Mi candidates),
 * zero (or an over-budget panel, or L > 32) selects the brute-force kernels.  Results are identical either way. */
FQTK_B200_API int fqtk_b200_matcher_create(const uint8_t* panel_ascii, uint32_t n_samples, uint32_t barcode_len,
                                           uint8_t max_mismatches, uint8_t min_mismatch_delta, int use_cache,
                                           int device, fqtk_b200_matcher** out);
/* The same with per-handle options instead of the defaults (two host threads may create matchers concurrently:
 * nothing about a handle depends on process-global state).  Initialise with fqtk_b200_options_init, then change fields. */
#define FQTK_B200_KERNEL_AUTO (-1) /* k_probe3 when the pure-A/C/G/T memo entries fit in shared memory, else k_probe5 (L <= 16)
                                      or k_probe4 (L <= 32), else k_probe2 */
typedef struct {
    uint32_t struct_size;       /* sizeof(fqtk_b200_options) as the caller compiled it (lets the struct grow) */
    int32_t kernel;             /* packed-route kernel: FQTK_B200_KERNEL_AUTO, 0 = k_probe2 only, 1 = k_probe4 (the L2-resident
                                   fingerprint table only), 2 / 3 = k_probe3 with that many sub-tables, 5 = k_probe5 (exact
                                   entries in shared memory + the fingerprint table).  Results are identical for every value */
    uint64_t table_budget;      /* largest neighbourhood (candidate strings) a memo table is built for; 0 = 32 Mi */
    uint64_t chunk_bytes;       /* input bytes per chunk of the host-buffer pipeline; 0 = 32 MiB */
    uint32_t l2_table_load_pct; /* load factor of the L2-resident table in percent, 0 = automatic */
    uint32_t reserved;
} fqtk_b200_options;
FQTK_B200_API void fqtk_b200_options_init(fqtk_b200_options* opts);
FQTK_B200_API int fqtk_b200_matcher_create_ex(const uint8_t* panel_ascii, uint32_t n_samples, uint32_t barcode_len,
                                              uint8_t max_mismatches, uint8_t min_mismatch_delta, int use_cache,
                                              int device, const fqtk_b200_options* opts, fqtk_b200_matcher** out);
FQTK_B200_API void fqtk_b200_matcher_destroy(fqtk_b200_matcher* m);
FQTK_B200_API int fqtk_b200_matcher_get_info(const fqtk_b200_matcher* m, fqtk_b200_matcher_info* info);
/* Deprecated convenience (kept for the tests and A/B timing): defaults of fqtk_b200_matcher_create for matchers created
 * afterwards BY THE CALLING THREAD (thread-local, so concurrent creates on other threads are unaffected). */
FQTK_B200_API void fqtk_b200_set_table_budget(uint64_t max_candidates);
/* Which kernel the HBM-resident packed route runs (= options.kernel):
 * -1 = automatic (default): k_probe3 (shared-memory cuckoo table of the pure-A/C/G/T memo entries) when it fits
 *      (L <= 16), else k_probe4 (the same entries in an L2-resident table of 8-byte slots, L <= 24), else k_probe2;
 *  0 = k_probe2 only;  1 = k_probe4's table only;  2 or 3 = k_probe3 with that many sub-tables.
 * Results are identical for every setting. */
FQTK_B200_API void fqtk_b200_set_cuckoo_arity(int arity);

/* ---- reference-facing calls: HOST buffers ----------------------------------------------------------
 * BarcodeMatcher::assign (barcode_matching.rs:165-186) for one read, any length: len < L -> NONE (:167-169);
 * len > L -> FQTK_B200_ERR_LENGTH (:95-106) unless the no-call pre-filter (:170-172) already made it NONE.
 * Adds 1 to counts[best_match] or counts[S] like the caller does (demux.rs:970-974). */
FQTK_B200_API int fqtk_b200_matcher_assign(fqtk_b200_matcher* m, const uint8_t* read_bases, size_t len,
                                           uint32_t* result);

/* The batched form of the same call: `n_reads` rows, row i at barcodes_ascii + i*row_stride.
 * `lengths` NULL  -> every row holds exactly L bases (row_stride >= L).
 * `lengths` given -> row i holds lengths[i] <= row_stride bases with the single-read length rules above;
 *                    the whole call fails with FQTK_B200_ERR_LENGTH (nothing counted) if any row would panic.
 * Synchronous.  Internally chunked and double-buffered (H2D, kernel, D2H overlap); pass memory from
 * fqtk_b200_host_alloc (pinned) for full PCIe rate.  results[i] is written for every row; counts accumulate. */
FQTK_B200_API int fqtk_b200_matcher_assign_batch(fqtk_b200_matcher* m, const uint8_t* barcodes_ascii,
                                                 uint64_t n_reads, uint64_t row_stride, const uint32_t* lengths,
                                                 uint32_t* results);

/* The same batch call for a host that already holds the reads in the reference's own BitEnc form (what encode(),
 * mod.rs:49-61, returns — fqtk_b200_pack_host produces it): `packed` = n_reads * W u32 words in HOST memory, layout as for
 * the device calls below.  Half the bytes of ASCII rows over PCIe (8 instead of 16 at dual 8+8 bp).  Results come back as
 * result words (`results`, 4 bytes per read) and / or as bare sample indices (`sample_index`, 2 bytes per read, 0xFFFF =
 * None: all the caller's routing needs, demux.rs:970-975); either pointer may be NULL, not both.  Synchronous, chunked. */
FQTK_B200_API int fqtk_b200_matcher_assign_batch_packed(fqtk_b200_matcher* m, const uint32_t* packed, uint64_t n_reads,
                                                        uint32_t* results, uint16_t* sample_index);
/* encode() of n_reads host rows (mod.rs:49-61): row i at rows + i*row_stride, barcode_len symbols each -> W words each.
 * Pure encoding on the host, no GPU involved, `threads` host threads (0 = all): rows back to back with barcode_len a multiple
 * of 8 are one stream of symbols and take an AVX2 form (32 symbols per step; csrc/host_pack.cpp), any other layout two symbols
 * per table lookup; the result is encode()'s for every byte value either way. */
FQTK_B200_API int fqtk_b200_pack_host(const uint8_t* rows, uint64_t n_reads, uint32_t barcode_len, uint64_t row_stride,
                                      uint32_t* out_packed, int threads);

/* The same call for barcodes that arrive in pieces (SURVEY 8f "next" #1: B-segment extraction fused on the GPU).
 * ReadSet::sample_barcode_sequence (demux.rs:121-123) concatenates the sample-barcode segments of ALL inputs, in input
 * order then in-read order; here each fixed-length B segment is described once and gathered + encoded on the device:
 * read i's barcode = seg[0].base[i*row_stride + offset .. +length) ++ seg[1]... (e.g. I1 rows ++ I2 rows for a
 * dual index, or the first 16 bytes of R1 for an inline barcode).  The segment lengths must add up to L exactly.
 * Up to FQTK_B200_MAX_SEGMENTS segments.  Host buffers, synchronous, chunked like assign_batch. */
#define FQTK_B200_MAX_SEGMENTS 8
typedef struct {
    const uint8_t* base; /* row 0 of the source holding this segment */
    uint64_t row_stride; /* bytes between consecutive reads in that source */
    uint32_t offset;     /* first byte of the segment inside a row */
    uint32_t length;     /* bases in the segment (> 0) */
} fqtk_b200_segment;
FQTK_B200_API int fqtk_b200_matcher_assign_segments(fqtk_b200_matcher* m, const fqtk_b200_segment* segments,
                                                    uint32_t n_segments, uint64_t n_reads, uint32_t* results);
/* Device-buffer form (segment bases are device pointers), asynchronous on `stream`. */
FQTK_B200_API int fqtk_b200_matcher_assign_segments_device(fqtk_b200_matcher* m, const fqtk_b200_segment* segments,
                                                           uint32_t n_segments, uint64_t n_reads,
                                                           uint32_t* d_results, void* stream);

/* ---- barcodes straight out of raw FASTQ text (SURVEY 8f "next" #1 / #2: ReadSetIterator::next, demux.rs:288-342) ----
 * fqtk_b200_fastq_scan finds the records of an in-memory, uncompressed 4-line FASTQ chunk: for record r the offset of
 * its header line ('@' included; head_offsets may be NULL), the offset and length of its sequence line.  Only records
 * that lie completely inside the chunk are reported; *consumed = the byte after the last of them (the caller carries the
 * rest over to the next chunk).  '\r' before '\n' is not part of a line.  A record whose first line does not start with
 * '@', whose third does not start with '+', or whose quality line differs in length from its sequence is an error
 * (FQTK_B200_ERR_ARG, with the record number in the message).  Host only, no GPU involved. */
FQTK_B200_API int fqtk_b200_fastq_scan(const uint8_t* chunk, uint64_t chunk_bytes, uint64_t max_records,
                                       uint64_t* head_offsets, uint64_t* seq_offsets, uint32_t* seq_lengths,
                                       uint64_t* n_records, uint64_t* consumed);
/* One input FASTQ of the read set (demux.rs:945-966 walks the inputs in lock step): the chunk and the per-read tables
 * fqtk_b200_fastq_scan filled for it.  All pointers are HOST pointers for fqtk_b200_matcher_assign_fastq and DEVICE
 * pointers for ..._assign_fastq_device. */
typedef struct {
    const uint8_t* chunk;
    uint64_t chunk_bytes;
    const uint64_t* seq_offsets; /* n_reads */
    const uint32_t* seq_lengths; /* n_reads; may be NULL for the device call when no segment of this source is REST */
} fqtk_b200_fastq_source;
/* One sample-barcode segment of a read structure: bytes [offset, offset + length) of the sequence line of `source`, or
 * everything from `offset` on when length == FQTK_B200_SEGMENT_REST (a trailing `+B`).  Segments are concatenated in the
 * order given — inputs in order, then segments in read order (ReadSet::sample_barcode_sequence, demux.rs:121-123). */
#define FQTK_B200_SEGMENT_REST 0xFFFFFFFFu
typedef struct {
    uint32_t source;
    uint32_t offset;
    uint32_t length;
} fqtk_b200_fastq_segment;
/* The batch call on raw FASTQ text: the B segments are gathered by offset and encoded ON THE DEVICE (no dense barcode
 * rows are built on the host), then matched like any batch.  Host form: ships every chunk and its tables, synchronous;
 * a read that is too short for a segment fails the call like the reference's "too few bases" panic (demux.rs:309-315;
 * filter such records out first to get --skip-reasons too-few-bases); barcodes whose gathered length differs from L
 * follow BarcodeMatcher::assign (shorter: None; longer: FQTK_B200_ERR_LENGTH unless the no-call pre-filter fires).
 * Device form: asynchronous on `stream`, no vetting (too-short / too-long barcodes come back None). */
FQTK_B200_API int fqtk_b200_matcher_assign_fastq(fqtk_b200_matcher* m, const fqtk_b200_fastq_source* sources,
                                                 uint32_t n_sources, const fqtk_b200_fastq_segment* segments,
                                                 uint32_t n_segments, uint64_t n_reads, uint32_t* results);
FQTK_B200_API int fqtk_b200_matcher_assign_fastq_device(fqtk_b200_matcher* m, const fqtk_b200_fastq_source* sources,
                                                        uint32_t n_sources, const fqtk_b200_fastq_segment* segments,
                                                        uint32_t n_segments, uint64_t n_reads, uint32_t* d_results,
                                                        void* stream);

/* The scanner on the DEVICE (the chunk crosses PCIe anyway): same tables, same rules and error texts as
 * fqtk_b200_fastq_scan for a chunk that is already in device memory; every pointer but n_records / consumed is a device
 * pointer (d_head_offsets may be NULL).  Synchronous on `stream` (the two counts come back to the host). */
FQTK_B200_API int fqtk_b200_fastq_scan_device(int device, const uint8_t* d_chunk, uint64_t chunk_bytes, uint64_t max_records,
                                              uint64_t* d_head_offsets, uint64_t* d_seq_offsets, uint32_t* d_seq_lengths,
                                              uint64_t* n_records, uint64_t* consumed, void* stream);
/* The whole ingest step in one call: raw FASTQ chunks of the inputs (HOST memory, one per input, in lock step) -> result
 * words.  The chunks are copied to the device, scanned THERE, the reference's per-read rules (too few bases,
 * demux.rs:298-315; barcode longer than the panel's, barcode_matching.rs:95-106 after the no-call pre-filter :170-172) are
 * checked there too, then gather + encode + match as in fqtk_b200_matcher_assign_fastq.  *n_reads = records matched
 * (complete in every chunk, at most max_reads), consumed[s] = bytes of chunk s they cover (the caller carries the rest
 * over).  Errors are the table form's, for the first offending read in input order. */
typedef struct {
    const uint8_t* data;
    uint64_t bytes;
} fqtk_b200_fastq_chunk;
FQTK_B200_API int fqtk_b200_matcher_assign_fastq_chunks(fqtk_b200_matcher* m, const fqtk_b200_fastq_chunk* chunks,
                                                        uint32_t n_sources, const fqtk_b200_fastq_segment* segments,
                                                        uint32_t n_segments, uint64_t max_reads, uint32_t* results,
                                                        uint64_t* n_reads, uint64_t* consumed);

/* ---- HBM-resident calls: DEVICE buffers, asynchronous on `stream` (a cudaStream_t, NULL = default) ----
 * `d_packed`: n_reads * W u32 words, read i at words [i*W, (i+1)*W), symbol k of a read in bits 4*(k%8) of its
 * word k/8 — the reference's own BitEnc layout (mod.rs:49-61, bitenc.rs:311-322).  `d_results`: n_reads u32.
 * n_reads < 2^32 per call.  Counts accumulate on the device; read them with fqtk_b200_matcher_counts.
 * Device calls on one handle share its scratch buffers: issue them on ONE stream (or serialise them yourself). */
FQTK_B200_API int fqtk_b200_matcher_assign_packed_device(fqtk_b200_matcher* m, const uint32_t* d_packed,
                                                         uint64_t n_reads, uint32_t* d_results, void* stream);
/* Same with ASCII rows on the device (encode fused into the kernel); d_lengths may be NULL. */
FQTK_B200_API int fqtk_b200_matcher_assign_ascii_device(fqtk_b200_matcher* m, const uint8_t* d_ascii,
                                                        uint64_t n_reads, uint64_t row_stride,
                                                        const uint32_t* d_lengths, uint32_t* d_results, void* stream);
/* encode() on the device: ASCII rows -> packed words (mod.rs:49-61). */
FQTK_B200_API int fqtk_b200_pack_device(const uint8_t* d_ascii, uint64_t n_reads, uint32_t barcode_len,
                                        uint64_t row_stride, uint32_t* d_packed, void* stream);
/* encode() on the host for one sequence: out = ceil(len/8) u32 blocks.  Pure encoding, no matching. */
FQTK_B200_API int fqtk_b200_encode_host(const uint8_t* bases, size_t len, uint32_t* out_blocks);

/* ---- per-sample routing of a batch (SURVEY 8f "next" #3) ----------------------------------------------
 * The reference hands every read to its sample's writer in input order (demux.rs:970-975; order pinned by the tests
 * at :1505-1523).  For a batch that is a STABLE partition of the read indices by assignment:
 *   order[offsets[j] .. offsets[j+1])  = indices of the reads assigned to sample j, ascending (input order), j < S
 *   order[offsets[S] .. offsets[S+1])  = indices of the unmatched reads, ascending;  offsets[S+1] = n_reads
 * so the host appends ONE contiguous run per sample per batch.  `d_results` are the batch's result words.
 * d_order: n_reads u32; d_offsets: S + 2 u64.  Asynchronous on `stream`; n_reads < 2^32; S + 1 <= ~50 000. */
FQTK_B200_API int fqtk_b200_matcher_route_device(fqtk_b200_matcher* m, const uint32_t* d_results, uint64_t n_reads,
                                                 uint32_t* d_order, uint64_t* d_offsets, void* stream);
/* Host-buffer convenience: ships result words in, brings order + offsets back.  Synchronous. */
FQTK_B200_API int fqtk_b200_matcher_route(fqtk_b200_matcher* m, const uint32_t* results, uint64_t n_reads,
                                          uint32_t* order, uint64_t* offsets);

/* ---- per-sample counts (DemuxMetric.templates, demux.rs:458,971,974): S+1 u64, last = unmatched ---- */
FQTK_B200_API int fqtk_b200_matcher_counts(fqtk_b200_matcher* m, uint64_t* out_counts); /* syncs the device */
FQTK_B200_API int fqtk_b200_matcher_counts_device(fqtk_b200_matcher* m, uint64_t** d_counts); /* for ncclAllReduce */
FQTK_B200_API int fqtk_b200_matcher_reset_counts(fqtk_b200_matcher* m);

/* ---- several GPUs of one box in ONE process ----------------------------------------------------
 * The reference owns one matcher and one count table (demux.rs:921-926, 970-975).  A group is the same thing over G
 * devices: the panel's tables are built on every device (concurrently, one host thread each), a batch is split into G
 * contiguous shards (fqtk_b200_group_shard: shard k = reads [first, first + count), in input order, so results[] is
 * filled exactly as by a single matcher), each shard runs on its device from its own host thread, and the G count tables
 * are summed on the first device — peers are read over NVLink through peer-mapped pointers by one small kernel (staged
 * copies where peer access is not available).  `devices` NULL / n_devices 0 = every visible device.  A host that
 * already runs NCCL can instead all-reduce the per-device tables itself: fqtk_b200_group_matcher + ..._counts_device. */
typedef struct fqtk_b200_group fqtk_b200_group;
FQTK_B200_API int fqtk_b200_group_create(const uint8_t* panel_ascii, uint32_t n_samples, uint32_t barcode_len,
                                         uint8_t max_mismatches, uint8_t min_mismatch_delta, int use_cache,
                                         const int* devices, uint32_t n_devices, const fqtk_b200_options* opts,
                                         fqtk_b200_group** out);
FQTK_B200_API void fqtk_b200_group_destroy(fqtk_b200_group* g);
FQTK_B200_API uint32_t fqtk_b200_group_size(const fqtk_b200_group* g);
FQTK_B200_API int fqtk_b200_group_device(const fqtk_b200_group* g, uint32_t k);                /* device ordinal of shard k */
FQTK_B200_API fqtk_b200_matcher* fqtk_b200_group_matcher(fqtk_b200_group* g, uint32_t k);     /* borrowed, not owned */
FQTK_B200_API void fqtk_b200_group_shard(const fqtk_b200_group* g, uint64_t n_reads, uint32_t k, uint64_t* first,
                                         uint64_t* count);
/* fqtk_b200_matcher_assign_batch / _assign_batch_packed over the group: host buffers, synchronous. */
FQTK_B200_API int fqtk_b200_group_assign_batch(fqtk_b200_group* g, const uint8_t* barcodes_ascii, uint64_t n_reads,
                                               uint64_t row_stride, const uint32_t* lengths, uint32_t* results);
FQTK_B200_API int fqtk_b200_group_assign_batch_packed(fqtk_b200_group* g, const uint32_t* packed, uint64_t n_reads,
                                                      uint32_t* results, uint16_t* sample_index);
/* HBM-resident shards: d_packed[k] / d_results[k] live on device k (n_reads[k] reads); asynchronous on streams[k]
 * (streams NULL = default streams). */
FQTK_B200_API int fqtk_b200_group_assign_packed_device(fqtk_b200_group* g, const uint32_t* const* d_packed,
                                                       const uint64_t* n_reads, uint32_t* const* d_results,
                                                       void* const* streams);
/* The ONE count table: S + 1 u64 summed over the devices (synchronises them).  */
FQTK_B200_API int fqtk_b200_group_counts(fqtk_b200_group* g, uint64_t* out_counts);
FQTK_B200_API int fqtk_b200_group_reset_counts(fqtk_b200_group* g);

/* ---- measurement aid: what the platform's pinned-memory copies alone can do (no kernels) ----
 * Moves in_bytes host->device and out_bytes device->host in chunks over three streams, like the host-buffer calls do,
 * `reps` times after one warm-up; *seconds_per_rep = mean wall time.  The ceiling any host-buffer call is held against. */
FQTK_B200_API int fqtk_b200_copy_ceiling(int device, uint64_t in_bytes, uint64_t out_bytes, uint64_t chunk_in_bytes,
                                         int reps, double* seconds_per_rep);

/* ---- control / introspection ---- */
FQTK_B200_API int fqtk_b200_matcher_set_mode(fqtk_b200_matcher* m, int mode); /* TABLE needs a built table */
/* Opt-in for fqtk_b200_matcher_assign_batch: let `threads` host threads (-1 = the CPUs the caller may run on, at most 16;
 * 0 = off, the default) do encode() (mod.rs:49-61) while the batch is in flight, so that the reference's BitEnc words —
 * half the bytes of the ASCII rows — cross PCIe.  Taken when the rows lie back to back (row_stride == barcode_len), the
 * barcode length is a multiple of 8, no per-row lengths are given and the batch has >= 2^20 reads; results are the plain
 * route's.  Worth it when the call is PCIe-bound and host cores are idle (one GPU per 8+ cores); it costs host memory
 * bandwidth (the rows are read and the words written once more), so not when many GPUs share one host's memory system. */
FQTK_B200_API int fqtk_b200_matcher_set_host_pack(fqtk_b200_matcher* m, int threads);
FQTK_B200_API uint64_t fqtk_b200_kernel_launches(void); /* kernels launched by this library in this process */
FQTK_B200_API const char* fqtk_b200_last_error(void);
FQTK_B200_API int fqtk_b200_device_count(void);

/* ---- pinned host memory for the host-buffer calls ---- */
FQTK_B200_API int fqtk_b200_host_alloc(void** ptr, size_t bytes);
FQTK_B200_API int fqtk_b200_host_free(void* ptr);
/* asynchronous copies on `stream` between host memory (pinned: full PCIe rate) and device memory */
FQTK_B200_API int fqtk_b200_copy_to_device(void* d_dst, const void* src, uint64_t bytes, void* stream);
FQTK_B200_API int fqtk_b200_copy_to_host(void* dst, const void* d_src, uint64_t bytes, void* stream);

/* ---- BGZF output compression (SURVEY 8f "next" #4; replaces the reference's pooled BGZF writers) ----
 * src/bin/commands/demux.rs:755-798 builds `PoolBuilder::<_, BgzfCompressor>` (pooled-writer 0.4.0 -> bgzf crate ->
 * libdeflate) with `compression_level` (demux.rs:641-643, default 5): every writer buffers 65 280 bytes (the bgzf crate's
 * BGZF_BLOCK_SIZE), each full buffer becomes ONE gzip member with the `BC` extra field (SAM spec 4.1), the file ends with
 * the 28-byte EOF block.  This entry does the same for one writer's bytes: input cut every 65 280 bytes, one member per
 * piece (LZ77 + dynamic Huffman on the device, a stored block when that does not shrink the piece), members back to back
 * in input order, optionally the EOF block.  level 0 = stored blocks only (as libdeflate), 1..12 = compressed (one
 * strategy; the level only sets the header's XFL byte).  The deflate bytes are not libdeflate's (no two implementations
 * agree); every member inflates to its piece and the CRC32 / ISIZE / BSIZE fields are exact.
 * A handle owns the device buffers for chunks of `chunk_bytes` input bytes (0 = 64 MiB); not thread-safe. */
typedef struct fqtk_b200_bgzf fqtk_b200_bgzf;
FQTK_B200_API int fqtk_b200_bgzf_create(int device, uint64_t chunk_bytes, fqtk_b200_bgzf** out);
FQTK_B200_API void fqtk_b200_bgzf_destroy(fqtk_b200_bgzf* z);
FQTK_B200_API uint64_t fqtk_b200_bgzf_chunk_bytes(const fqtk_b200_bgzf* z);
FQTK_B200_API uint64_t fqtk_b200_bgzf_bound(uint64_t n_bytes); /* output bytes that always suffice, EOF block included */
/* host buffers (pinned memory makes the copies asynchronous): H2D, kernels and D2H of consecutive chunks overlap */
FQTK_B200_API int fqtk_b200_bgzf_compress(fqtk_b200_bgzf* z, const uint8_t* in, uint64_t n_bytes, int level, int write_eof,
                                          uint8_t* out, uint64_t out_capacity, uint64_t* out_bytes);
/* device buffers, one chunk (n_bytes <= chunk_bytes), enqueued on `stream`; *d_out_bytes (device u64) = bytes written */
FQTK_B200_API int fqtk_b200_bgzf_compress_device(fqtk_b200_bgzf* z, const uint8_t* d_in, uint64_t n_bytes, int level,
                                                 uint8_t* d_out, uint64_t out_capacity, uint64_t* d_out_bytes, void* stream);

/* several writers' texts in ONE device buffer (what fqtk_b200_demux_emit_device leaves): segment k =
 * d_in[seg_offsets[k], seg_offsets[k+1]) is cut every 65 280 bytes on its own, so no member spans two writers; all members
 * leave back to back in d_out, out_seg_offsets[k] (HOST, n_segments + 1 entries) = where segment k's members start.  No EOF
 * blocks (append the 28-byte EOF member when a file is closed).  Synchronous on `stream`. */
FQTK_B200_API int fqtk_b200_bgzf_compress_segments_device(fqtk_b200_bgzf* z, const uint8_t* d_in, const uint64_t* seg_offsets,
                                                          uint32_t n_segments, int level, uint8_t* d_out,
                                                          uint64_t out_capacity, uint64_t* out_seg_offsets, void* stream);

/* ---- the routed records written out on the device (SURVEY 8f "next" #3 / #4) ----
 * SampleWriters::write (demux.rs:396-415) + ReadSet::write_header_internal (:171-267) for a whole batch whose FASTQ chunks are
 * in device memory: for every output stream — the T, B, M, C segments selected by `output_kinds`, in the order the
 * reference walks its writers, e.g. R1 R2 I1 — one text buffer region in which every sample's records
 * (`@<rewritten header>\n<bases>\n+\n<quals>\n`) form one contiguous run in input order.  The header of a record is
 * the FIRST input's header rewritten: read number, UMI (M) segments appended to the name, sample-barcode (B) segments
 * appended to the comment's index field.
 *   sources[s]     chunk + the tables of fqtk_b200_fastq_scan_device (d_head_offsets is needed for source 0 only)
 *   segments       EVERY segment of every read structure, inputs in order, segments in read order
 *   d_order, d_offsets   fqtk_b200_matcher_route_device's output for the batch (n_buckets = S + 1)
 *   file_offsets   HOST out, [n_streams][n_buckets + 1]: (stream t, bucket b) = d_text[file_offsets[t][b], file_offsets[t][b+1])
 * Header rule violations (demux.rs:190-194,229-233) fail the call with the reference's text and the read's index. */
typedef struct {
    const uint8_t* d_chunk;
    uint64_t chunk_bytes;
    const uint64_t* d_head_offsets;
    const uint64_t* d_seq_offsets;
    const uint32_t* d_seq_lengths;
} fqtk_b200_emit_source;
typedef struct {
    uint32_t source; /* which input */
    uint32_t kind;   /* 'T', 'B', 'M', 'S' or 'C' */
    uint32_t offset; /* first base of the segment in the read */
    uint32_t length; /* bases, or FQTK_B200_SEGMENT_REST */
} fqtk_b200_read_segment;
/* which streams `output_kinds` selects: kind letter and 1-based number of each (file name <prefix>.<R|I|U|C><number>.fq.gz) */
FQTK_B200_API int fqtk_b200_emit_streams(const fqtk_b200_read_segment* segments, uint32_t n_segments, const char* output_kinds,
                                         uint32_t* n_streams, char* stream_kinds, uint32_t* stream_numbers);
FQTK_B200_API int fqtk_b200_demux_emit_device(int device, const fqtk_b200_emit_source* sources, uint32_t n_sources,
                                              const fqtk_b200_read_segment* segments, uint32_t n_segments,
                                              const char* output_kinds, const uint32_t* d_order, const uint64_t* d_offsets,
                                              uint32_t n_buckets, uint64_t n_reads, uint8_t* d_text, uint64_t text_capacity,
                                              uint64_t* file_offsets, uint64_t* text_bytes, void* stream);

/* ---- one call per batch: FASTQ chunks in, per-sample BGZF members out ----
 * The reference's main loop body (demux.rs:945-977) for a batch, entirely on the device: chunks of the inputs (HOST memory,
 * lock step) -> records found and vetted (too few bases for the read structure: the reference's error text, demux.rs:309-315)
 * -> B segments matched -> routed -> every selected output segment written with its rewritten header -> every (stream,
 * sample) run deflated into its own BGZF members -> `out` (HOST).  Run k = stream t * (S + 1) + bucket b (bucket S =
 * unmatched) occupies out[out_offsets[k], out_offsets[k+1]): append it to that file (and the 28-byte EOF member at close).
 * batch_counts (S + 1, may be NULL) = this batch's templates per bucket; the matcher's own counters accumulate as usual.
 * *n_reads / consumed[s] as in fqtk_b200_matcher_assign_fastq_chunks.  Streams: fqtk_b200_emit_streams. */
FQTK_B200_API int fqtk_b200_demux_chunks(fqtk_b200_matcher* m, fqtk_b200_bgzf* z, const fqtk_b200_fastq_chunk* chunks,
                                         uint32_t n_sources, const fqtk_b200_read_segment* segments, uint32_t n_segments,
                                         const char* output_kinds, int level, uint64_t max_reads, uint8_t* out,
                                         uint64_t out_capacity, uint64_t* out_offsets, uint32_t* n_streams,
                                         uint64_t* batch_counts, uint64_t* n_reads, uint64_t* consumed);

/* ---- deterministic synthetic workload; counter-based, identical on host and device ----
 * Panel: S barcodes of length L over ACGT with pairwise Hamming distance >= min_distance (greedy, rejection);
 * `n_degenerate` positions per barcode are then rewritten to IUPAC degenerate codes (cfg 5).
 * Reads: read i depends only on (seed, i): 90 % true (uniform sample, degenerate positions resolved, per-base
 * 0.5 % substitution, 0.2 % no-call), 8 % near-miss (plus 2-3 forced substitutions), 2 % uniform random L-mers. */
FQTK_B200_API int fqtk_b200_synth_panel(uint64_t seed, uint32_t n_samples, uint32_t barcode_len,
                                        uint32_t min_distance, uint32_t n_degenerate, uint8_t* out_panel_ascii);
FQTK_B200_API int fqtk_b200_synth_reads_host(const uint8_t* panel_ascii, uint32_t n_samples, uint32_t barcode_len,
                                             uint64_t seed, uint64_t first_read, uint64_t n_reads,
                                             uint8_t* out_ascii /* n_reads * L */);
FQTK_B200_API int fqtk_b200_synth_reads_device(const uint8_t* panel_ascii, uint32_t n_samples, uint32_t barcode_len,
                                               uint64_t seed, uint64_t first_read, uint64_t n_reads,
                                               uint8_t* d_ascii /* or NULL */, uint32_t* d_packed /* or NULL */,
                                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FQTK_B200_H */
