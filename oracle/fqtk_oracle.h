/*
 * fqtk_oracle.h — CPU ORACLE for the fqtk `demux` barcode-matcher hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fqtk_b200/ (the product) may include, link, import or
 * execute this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the CPU baseline.
 *
 * It is a plain-C restatement of the reference's algorithm (fulcrumgenomics/fqtk @ 45dbb99); every
 * function cites the reference file:line it follows (paths relative to /root/reference).  The
 * reference is Rust and no Rust toolchain exists in this image, so it cannot be compiled here
 * (no oracle/_ref).  Parity is PINNED by replaying every known-answer test the reference holds for
 * this path (tests/golden/reference_kats.json, transcribed from src/lib/barcode_matching.rs:251-447,
 * src/lib/bitenc.rs:558-580, src/lib/mod.rs:94-169 and src/bin/commands/demux.rs:1421-1611).
 *
 * Two independent restatements live here:
 *   (a) LITERAL    — width-4 BitEnc blocks, capped early-exit hamming, the sequential
 *                    best/next/cap scan, the no-call pre-filter and the memo cache;
 *   (b) CLOSED FORM — exact distances to every barcode, lexicographic (d, j) min and second-min
 *                    (SURVEY.md Appendix A.2).  This is what the GPU computes.
 * tests/ fuzz (a) == (b).
 */
#ifndef FQTK_ORACLE_H
#define FQTK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Result word shared with the product ABI (include/fqtk_b200.h):
 *   FQO_NONE (0xFFFFFFFF)                       -> Option::None
 *   (best_match << 16) | (best << 8) | next     -> Some(BarcodeMatch{..})  (barcode_matching.rs:16-25) */
#define FQO_NONE 0xFFFFFFFFu
#define FQO_MAX_SYMBOLS 256u
#define FQO_MAX_BLOCKS (FQO_MAX_SYMBOLS / 8u)

/* error codes standing in for the reference's panics */
#define FQO_OK 0
#define FQO_ERR_EMPTY_PANEL (-1)   /* "Must provide at least one sample"      barcode_matching.rs:61   */
#define FQO_ERR_EMPTY_BARCODE (-2) /* "Sample barcode cannot be empty string" barcode_matching.rs:62-65 */
#define FQO_ERR_LENGTH (-3)        /* "Read barcode (..) length (..) differs" barcode_matching.rs:95-106 */
#define FQO_ERR_ARG (-4)

/* width-4 BitEnc (bitenc.rs:37-43): symbol i lives at bits 4*(i%8) of block i/8 (bitenc.rs:311-322) */
typedef struct {
    uint32_t blk[FQO_MAX_BLOCKS];
    uint32_t len; /* nr_symbols, bitenc.rs:396 */
} fqo_bitenc;

typedef struct fqo_matcher fqo_matcher;

/* src/lib/mod.rs:26-46 (IUPAC_MASKS), :85-87 (byte_is_nocall), :90-92 (is_valid_iupac) */
uint8_t fqo_iupac_mask(uint8_t byte);
int fqo_byte_is_nocall(uint8_t byte);
int fqo_is_valid_iupac(uint8_t byte);

/* src/lib/mod.rs:49-61 (encode) and :68-82 (decode; returns -1 on an undecodable nibble = the panic) */
int fqo_encode(const uint8_t* bases, size_t len, fqo_bitenc* out);
int fqo_decode(const fqo_bitenc* enc, char* out /* len+1 bytes */);

/* src/lib/bitenc.rs:432-459.  Returns UINT32_MAX when lengths differ (the assert at :433). */
uint32_t fqo_hamming(const fqo_bitenc* self, const fqo_bitenc* other, uint32_t max_mismatches);

/* BarcodeMatcher::count_mismatches, barcode_matching.rs:89-110.  Returns mismatches (0..255) or
 * FQO_ERR_LENGTH; on error `msg` (if non-NULL, >= 512 bytes) receives the reference's panic text. */
int fqo_count_mismatches(const uint8_t* observed, size_t obs_len, const uint8_t* expected, size_t exp_len,
                         const char* sample_id, uint8_t max_mismatches, char* msg);

/* BarcodeMatcher::new, barcode_matching.rs:55-86.  panel = S rows of L ASCII bytes. */
int fqo_matcher_new(const uint8_t* panel, uint32_t S, uint32_t L, uint8_t max_mismatches,
                    uint8_t min_mismatch_delta, int use_cache, fqo_matcher** out);
void fqo_matcher_free(fqo_matcher* m);
uint32_t fqo_matcher_max_ns(const fqo_matcher* m); /* max_ns_in_barcodes, :73-74 */
uint64_t fqo_matcher_cache_len(const fqo_matcher* m);

/* BarcodeMatcher::assign, barcode_matching.rs:165-186 (LITERAL).  *result = FQO_NONE or packed Some.
 * Returns FQO_OK or FQO_ERR_LENGTH (len > L reaching count_mismatches = the reference panics). */
int fqo_assign(fqo_matcher* m, const uint8_t* read_bases, size_t len, uint32_t* result);

/* BarcodeMatcher::assign_internal, barcode_matching.rs:119-160 (LITERAL, no pre-filter, no cache). */
int fqo_assign_internal(const fqo_matcher* m, const uint8_t* read_bases, size_t len, uint32_t* result);

/* CLOSED FORM: exact d_j for all j, first-index min, second-min with multiplicity. */
int fqo_assign_closed(const fqo_matcher* m, const uint8_t* read_bases, size_t len, uint32_t* result);

/* The caller's loop body, demux.rs:967-975: N dense rows of L bytes -> result word per read and
 * counts[S+1] (+= ; counts[S] = unmatched).  mode: 0 literal assign (with cache if enabled),
 * 1 closed form.  Single thread — this is the reference's threading model for the matcher. */
int fqo_assign_batch(fqo_matcher* m, const uint8_t* reads, uint64_t N, uint32_t* results /* may be NULL */,
                     uint64_t* counts /* S+1, may be NULL */, int mode);

/* NOT something the reference does: reads sharded over `threads` OpenMP threads, one private
 * matcher + memo cache per thread.  Upper bound for the CPU baseline only.  Returns threads used. */
int fqo_assign_batch_mt(const uint8_t* panel, uint32_t S, uint32_t L, uint8_t max_mismatches,
                        uint8_t min_mismatch_delta, int use_cache, const uint8_t* reads, uint64_t N,
                        uint32_t* results, uint64_t* counts, int threads);

int fqo_max_threads(void);

/* ---- synthetic workload (fqtk_synth.c): the benchmark's deterministic read stream and panels, byte for byte what
 * fqtk_b200/csrc/synth.cu generates — here so that the CPU reference arm of bench.py loads nothing of the product.
 * Workload generation only: nothing below matches reads. ---- */
void fqo_synth_reads(const uint8_t* panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t first, uint64_t n,
                     uint8_t* out /* n * L */);
int fqo_synth_panel(uint64_t seed, uint32_t S, uint32_t L, uint32_t min_distance, uint32_t n_degenerate,
                    uint8_t* out /* S * L */);

#ifdef __cplusplus
}
#endif
#endif
