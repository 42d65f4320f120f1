/*
 * fqtk_oracle.c — CPU ORACLE (test infrastructure only; see fqtk_oracle.h for the rules).
 *
 * Plain-C restatement of the fqtk `demux` matcher path.  Reference citations are file:line under
 * /root/reference (fulcrumgenomics/fqtk @ 45dbb99).  Pinned against the reference's own
 * known-answer tests by tests/test_oracle_kats.py.
 */
#include "fqtk_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * src/lib/mod.rs:26-46  IUPAC_MASKS — only uppercase keys are populated, everything else is 0.
 * ---------------------------------------------------------------------------------------------- */
static uint8_t g_masks[256];
static int g_masks_ready = 0;

static void init_masks(void) {
    if (g_masks_ready) return;
    const uint8_t a = 1, c = 2, g = 4, t = 8;
    memset(g_masks, 0, sizeof g_masks);
    g_masks['A'] = a;
    g_masks['C'] = c;
    g_masks['G'] = g;
    g_masks['T'] = t;
    g_masks['U'] = t;
    g_masks['M'] = a | c;
    g_masks['R'] = a | g;
    g_masks['W'] = a | t;
    g_masks['S'] = c | g;
    g_masks['Y'] = c | t;
    g_masks['K'] = g | t;
    g_masks['V'] = a | c | g;
    g_masks['H'] = a | c | t;
    g_masks['D'] = a | g | t;
    g_masks['B'] = c | g | t;
    g_masks['N'] = a | c | g | t;
    g_masks_ready = 1;
}

uint8_t fqo_iupac_mask(uint8_t byte) {
    init_masks();
    return g_masks[byte];
}

/* src/lib/mod.rs:85-87 */
int fqo_byte_is_nocall(uint8_t b) { return b == 'N' || b == 'n' || b == '.'; }

/* src/lib/mod.rs:90-92 */
int fqo_is_valid_iupac(uint8_t b) { return fqo_iupac_mask(b) != 0 || fqo_byte_is_nocall(b); }

static uint8_t ascii_upper(uint8_t b) { return (b >= 'a' && b <= 'z') ? (uint8_t)(b - 32) : b; }

/* ------------------------------------------------------------------------------------------------
 * src/lib/bitenc.rs:114-121 (push), :311-316 (set_by_addr), :319-322 (addr) at width 4:
 * usable_bits_per_block = 32, so value i goes to block i/8, bit offset 4*(i%8).
 * ---------------------------------------------------------------------------------------------- */
static void bitenc_push(fqo_bitenc* e, uint8_t value) {
    const uint32_t k = e->len * 4u;
    const uint32_t block = k / 32u, bit = k % 32u;
    if (bit == 0) e->blk[block] = 0; /* storage.push(0) */
    const uint32_t mask = 0xFu << bit;
    e->blk[block] |= mask;
    e->blk[block] ^= mask;
    e->blk[block] |= ((uint32_t)value & 0xFu) << bit;
    e->len += 1;
}

static uint32_t bitenc_nr_blocks(const fqo_bitenc* e) { return (e->len + 7u) / 8u; }

/* src/lib/mod.rs:49-61 */
int fqo_encode(const uint8_t* bases, size_t len, fqo_bitenc* out) {
    init_masks();
    if (len > FQO_MAX_SYMBOLS) return FQO_ERR_ARG;
    out->len = 0;
    for (size_t i = 0; i < len; i++) {
        const uint8_t base = bases[i];
        uint8_t bit;
        if (fqo_byte_is_nocall(base)) {
            bit = g_masks['N'];
        } else {
            bit = g_masks[ascii_upper(base)];
        }
        bitenc_push(out, bit);
    }
    return FQO_OK;
}

/* src/lib/mod.rs:68-82 — first IUPAC_BASES entry ("ACGTMRWSYKVHDBN", mod.rs:8) whose mask equals the nibble */
int fqo_decode(const fqo_bitenc* enc, char* out) {
    static const char IUPAC_BASES[] = "ACGTMRWSYKVHDBN";
    init_masks();
    for (uint32_t i = 0; i < enc->len; i++) {
        const uint8_t v = (uint8_t)((enc->blk[i / 8u] >> (4u * (i % 8u))) & 0xFu);
        int found = 0;
        for (int b = 0; b < 15; b++) {
            if (g_masks[(uint8_t)IUPAC_BASES[b]] == v) {
                out[i] = IUPAC_BASES[b];
                found = 1;
                break;
            }
        }
        if (!found) return -1; /* "Invalid bit mask for base" */
    }
    out[enc->len] = 0;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * src/lib/bitenc.rs:432-459  BitEnc::hamming — literal, including the block-granular early exit.
 * ---------------------------------------------------------------------------------------------- */
uint32_t fqo_hamming(const fqo_bitenc* self, const fqo_bitenc* other, uint32_t max_mismatches) {
    if (self->len != other->len) return UINT32_MAX; /* assert at :433 */
    uint32_t count = 0;
    const uint32_t values_per_block = 32u / 4u;
    const uint32_t nblocks = bitenc_nr_blocks(self);
    for (uint32_t block_index = 0; block_index < nblocks; block_index++) {
        const uint32_t block_diff = self->blk[block_index] & ~other->blk[block_index];
        if (block_diff != 0) {
            uint32_t shift_i = 0;
            for (uint32_t v = 0; v < values_per_block; v++) {
                const uint32_t block_diff_sub = (block_diff >> shift_i) & 0xFu;
                if (block_diff_sub != 0) count += 1;
                shift_i += 4;
            }
            if (count >= max_mismatches) return max_mismatches;
        }
    }
    return count;
}

/* ------------------------------------------------------------------------------------------------
 * Memo cache: stands in for AHashMap<Vec<u8>, BarcodeMatch> (barcode_matching.rs:44,84,173-182).
 * Open addressing, linear probing, keys of exactly L bytes held in an arena.  Result-neutral.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t tag;  /* high hash bits | 1 (0 = empty) */
    uint32_t idx;  /* entry index */
} cache_slot;

typedef struct {
    cache_slot* slots;
    uint64_t nslots; /* power of two */
    uint8_t* keys;   /* n * L */
    uint32_t* vals;
    uint64_t n, cap;
    uint32_t L;
} memo_cache;

#define STARTING_CACHE_SIZE 1000000ull /* barcode_matching.rs:12 */

static inline uint64_t mum(uint64_t a, uint64_t b) {
    __uint128_t r = (__uint128_t)a * b;
    return (uint64_t)r ^ (uint64_t)(r >> 64);
}

static inline uint64_t hash_bytes(const uint8_t* p, uint32_t len) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ len;
    while (len >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = mum(h ^ w, 0xA0761D6478BD642Full);
        p += 8;
        len -= 8;
    }
    if (len) {
        uint64_t w = 0;
        memcpy(&w, p, len);
        h = mum(h ^ w, 0xE7037ED1A0B428DBull);
    }
    return mum(h, 0x8EBC6AF09C88C6E3ull);
}

static int cache_init(memo_cache* c, uint32_t L, uint64_t capacity) {
    memset(c, 0, sizeof *c);
    c->L = L;
    uint64_t nslots = 1;
    while (nslots < capacity * 2) nslots <<= 1; /* ~hashbrown: 1M capacity -> 2^21 buckets */
    c->nslots = nslots;
    c->slots = (cache_slot*)calloc(nslots, sizeof(cache_slot));
    c->cap = capacity;
    c->keys = (uint8_t*)malloc((size_t)c->cap * (L ? L : 1));
    c->vals = (uint32_t*)malloc((size_t)c->cap * sizeof(uint32_t));
    return (c->slots && c->keys && c->vals) ? 0 : -1;
}

static void cache_free(memo_cache* c) {
    free(c->slots);
    free(c->keys);
    free(c->vals);
    memset(c, 0, sizeof *c);
}

static void cache_grow(memo_cache* c) {
    const uint64_t nslots = c->nslots * 2;
    cache_slot* slots = (cache_slot*)calloc(nslots, sizeof(cache_slot));
    for (uint64_t i = 0; i < c->n; i++) {
        const uint64_t h = hash_bytes(c->keys + i * c->L, c->L);
        uint64_t p = h & (nslots - 1);
        while (slots[p].tag) p = (p + 1) & (nslots - 1);
        slots[p].tag = (uint32_t)(h >> 32) | 1u;
        slots[p].idx = (uint32_t)i;
    }
    free(c->slots);
    c->slots = slots;
    c->nslots = nslots;
}

static inline int cache_get(const memo_cache* c, const uint8_t* key, uint64_t h, uint32_t* val) {
    const uint32_t tag = (uint32_t)(h >> 32) | 1u;
    uint64_t p = h & (c->nslots - 1);
    for (;;) {
        const cache_slot s = c->slots[p];
        if (!s.tag) return 0;
        if (s.tag == tag && memcmp(c->keys + (uint64_t)s.idx * c->L, key, c->L) == 0) {
            *val = c->vals[s.idx];
            return 1;
        }
        p = (p + 1) & (c->nslots - 1);
    }
}

static void cache_insert(memo_cache* c, const uint8_t* key, uint64_t h, uint32_t val) {
    if (c->n == c->cap) {
        c->cap *= 2;
        c->keys = (uint8_t*)realloc(c->keys, (size_t)c->cap * c->L);
        c->vals = (uint32_t*)realloc(c->vals, (size_t)c->cap * sizeof(uint32_t));
    }
    if ((c->n + 1) * 2 > c->nslots) cache_grow(c);
    memcpy(c->keys + c->n * c->L, key, c->L);
    c->vals[c->n] = val;
    uint64_t p = h & (c->nslots - 1);
    while (c->slots[p].tag) p = (p + 1) & (c->nslots - 1);
    c->slots[p].tag = (uint32_t)(h >> 32) | 1u;
    c->slots[p].idx = (uint32_t)c->n;
    c->n += 1;
}

/* ------------------------------------------------------------------------------------------------
 * struct BarcodeMatcher, barcode_matching.rs:29-45
 * ---------------------------------------------------------------------------------------------- */
struct fqo_matcher {
    uint32_t S, L;
    uint8_t* barcodes;           /* samples[i].barcode, upper-cased (:71) */
    fqo_bitenc* sample_barcodes; /* :75 */
    uint32_t max_ns_in_barcodes; /* :73-74 */
    uint8_t max_mismatches, min_mismatch_delta;
    int use_cache;
    memo_cache cache;
};

/* barcode_matching.rs:55-86 */
int fqo_matcher_new(const uint8_t* panel, uint32_t S, uint32_t L, uint8_t max_mismatches,
                    uint8_t min_mismatch_delta, int use_cache, fqo_matcher** out) {
    init_masks();
    if (!out) return FQO_ERR_ARG;
    *out = NULL;
    if (S == 0) return FQO_ERR_EMPTY_PANEL;   /* :61 */
    if (L == 0) return FQO_ERR_EMPTY_BARCODE; /* :62-65 */
    if (L > FQO_MAX_SYMBOLS || !panel) return FQO_ERR_ARG;
    fqo_matcher* m = (fqo_matcher*)calloc(1, sizeof *m);
    if (!m) return FQO_ERR_ARG;
    m->S = S;
    m->L = L;
    m->barcodes = (uint8_t*)malloc((size_t)S * L);
    m->sample_barcodes = (fqo_bitenc*)malloc((size_t)S * sizeof(fqo_bitenc));
    m->max_mismatches = max_mismatches;
    m->min_mismatch_delta = min_mismatch_delta;
    m->use_cache = use_cache;
    uint32_t max_ns = 0;
    for (uint32_t j = 0; j < S; j++) {
        uint8_t* bc = m->barcodes + (size_t)j * L;
        uint32_t num_ns = 0;
        for (uint32_t i = 0; i < L; i++) {
            bc[i] = ascii_upper(panel[(size_t)j * L + i]); /* :71 */
            if (fqo_byte_is_nocall(bc[i])) num_ns++;       /* :73 */
        }
        if (num_ns > max_ns) max_ns = num_ns; /* :74 */
        fqo_encode(bc, L, &m->sample_barcodes[j]); /* :75 */
    }
    m->max_ns_in_barcodes = max_ns;
    /* The reference always allocates the map (:84); only a cache-enabled matcher ever touches it. */
    if (use_cache && cache_init(&m->cache, L, STARTING_CACHE_SIZE) != 0) {
        fqo_matcher_free(m);
        return FQO_ERR_ARG;
    }
    *out = m;
    return FQO_OK;
}

void fqo_matcher_free(fqo_matcher* m) {
    if (!m) return;
    if (m->use_cache) cache_free(&m->cache);
    free(m->barcodes);
    free(m->sample_barcodes);
    free(m);
}

uint32_t fqo_matcher_max_ns(const fqo_matcher* m) { return m->max_ns_in_barcodes; }
uint64_t fqo_matcher_cache_len(const fqo_matcher* m) { return m->use_cache ? m->cache.n : 0; }

/* barcode_matching.rs:89-110 */
int fqo_count_mismatches(const uint8_t* observed, size_t obs_len, const uint8_t* expected, size_t exp_len,
                         const char* sample_id, uint8_t max_mismatches, char* msg) {
    fqo_bitenc obs, exp;
    if (fqo_encode(observed, obs_len, &obs) || fqo_encode(expected, exp_len, &exp)) return FQO_ERR_ARG;
    if (obs.len != exp.len) {
        if (msg) {
            char dec[FQO_MAX_SYMBOLS + 1];
            if (fqo_decode(&obs, dec) != 0) strcpy(dec, "?");
            snprintf(msg, 512,
                     "Read barcode (%s) length (%u) differs from expected barcode (%.*s) length (%u) for sample %s",
                     dec, obs.len, (int)exp_len, (const char*)expected, exp.len, sample_id ? sample_id : "");
        }
        return FQO_ERR_LENGTH;
    }
    const uint32_t count = fqo_hamming(&obs, &exp, (uint32_t)max_mismatches);
    return (int)(uint8_t)count; /* u8::try_from, :109 (count <= max <= 255) */
}

static inline uint32_t pack_result(uint32_t idx, uint8_t best, uint8_t next) {
    return (idx << 16) | ((uint32_t)best << 8) | (uint32_t)next;
}

/* barcode_matching.rs:119-160 — literal */
int fqo_assign_internal(const fqo_matcher* m, const uint8_t* read_bases, size_t len, uint32_t* result) {
    uint32_t best_barcode_index = m->S;
    uint8_t best_mismatches = 255, next_best_mismatches = 255, max_mismatches = 255; /* :120-123 */
    fqo_bitenc obs;
    if (fqo_encode(read_bases, len, &obs)) return FQO_ERR_ARG; /* :124 */
    for (uint32_t index = 0; index < m->S; index++) {           /* :125 */
        const fqo_bitenc* exp = &m->sample_barcodes[index];
        if (obs.len != exp->len) return FQO_ERR_LENGTH; /* count_mismatches :95-106 */
        const uint8_t mismatches = (uint8_t)fqo_hamming(&obs, exp, (uint32_t)max_mismatches); /* :108-109 */
        if (mismatches < best_mismatches) { /* :132 */
            next_best_mismatches = best_mismatches;
            best_mismatches = mismatches;
            best_barcode_index = index;
            if (next_best_mismatches < 255 - m->min_mismatch_delta) { /* :136 */
                const uint8_t c = (uint8_t)(next_best_mismatches + m->min_mismatch_delta);
                if (c < max_mismatches) max_mismatches = c;
            }
        } else if (mismatches < next_best_mismatches) { /* :140 */
            next_best_mismatches = mismatches;
            if (next_best_mismatches < 255 - m->min_mismatch_delta) {
                const uint8_t c = (uint8_t)(next_best_mismatches + m->min_mismatch_delta);
                if (c < max_mismatches) max_mismatches = c;
            }
        }
    }
    if (best_mismatches > m->max_mismatches ||
        (uint8_t)(next_best_mismatches - best_mismatches) < m->min_mismatch_delta) { /* :149-151 */
        *result = FQO_NONE;
    } else {
        *result = pack_result(best_barcode_index, best_mismatches, next_best_mismatches); /* :153-158 */
    }
    return FQO_OK;
}

/* barcode_matching.rs:165-186 — literal */
int fqo_assign(fqo_matcher* m, const uint8_t* read_bases, size_t len, uint32_t* result) {
    if (len < m->L) { /* :167-169 */
        *result = FQO_NONE;
        return FQO_OK;
    }
    size_t num_no_calls = 0;
    for (size_t i = 0; i < len; i++) num_no_calls += (size_t)fqo_byte_is_nocall(read_bases[i]);
    if (num_no_calls > (size_t)m->max_mismatches + m->max_ns_in_barcodes) { /* :170-172 */
        *result = FQO_NONE;
        return FQO_OK;
    }
    if (len != m->L) return FQO_ERR_LENGTH; /* a longer key can never be cached: every path panics at :95 */
    if (m->use_cache) {
        const uint64_t h = hash_bytes(read_bases, m->L);
        if (cache_get(&m->cache, read_bases, h, result)) return FQO_OK; /* :174-175 */
        const int rc = fqo_assign_internal(m, read_bases, len, result);   /* :177 */
        if (rc != FQO_OK) return rc;
        if (*result != FQO_NONE) cache_insert(&m->cache, read_bases, h, *result); /* :178-180 */
        return FQO_OK;
    }
    return fqo_assign_internal(m, read_bases, len, result); /* :184 */
}

/* SURVEY.md Appendix A.2 — closed form; independent of the BitEnc layout on purpose. */
int fqo_assign_closed(const fqo_matcher* m, const uint8_t* read_bases, size_t len, uint32_t* result) {
    init_masks();
    if (len < m->L) {
        *result = FQO_NONE;
        return FQO_OK;
    }
    if (len != m->L) {
        /* reference: None if the no-call filter fires, panic otherwise */
        size_t nc = 0;
        for (size_t i = 0; i < len; i++) nc += (size_t)fqo_byte_is_nocall(read_bases[i]);
        if (nc > (size_t)m->max_mismatches + m->max_ns_in_barcodes) {
            *result = FQO_NONE;
            return FQO_OK;
        }
        return FQO_ERR_LENGTH;
    }
    uint8_t obs[FQO_MAX_SYMBOLS];
    for (uint32_t i = 0; i < m->L; i++) {
        const uint8_t b = read_bases[i];
        obs[i] = fqo_byte_is_nocall(b) ? 15 : g_masks[ascii_upper(b)];
    }
    uint32_t best = 0xFFFFFFFFu, next = 0xFFFFFFFFu, best_idx = m->S;
    for (uint32_t j = 0; j < m->S; j++) {
        const uint8_t* bc = m->barcodes + (size_t)j * m->L;
        uint32_t d = 0;
        for (uint32_t i = 0; i < m->L; i++) {
            const uint8_t e = fqo_byte_is_nocall(bc[i]) ? 15 : g_masks[bc[i]];
            d += (obs[i] & (uint8_t)~e & 0xF) != 0;
        }
        if (d < best) {
            next = best;
            best = d;
            best_idx = j;
        } else if (d < next) {
            next = d;
        }
    }
    const uint32_t next8 = (next == 0xFFFFFFFFu) ? 255u : next;
    if (best > 255u || best > m->max_mismatches || next8 - best < m->min_mismatch_delta) {
        *result = FQO_NONE;
    } else {
        *result = pack_result(best_idx, (uint8_t)best, (uint8_t)next8);
    }
    return FQO_OK;
}

/* demux.rs:967-975 */
int fqo_assign_batch(fqo_matcher* m, const uint8_t* reads, uint64_t N, uint32_t* results, uint64_t* counts,
                     int mode) {
    for (uint64_t i = 0; i < N; i++) {
        uint32_t r;
        const uint8_t* rb = reads + i * m->L;
        const int rc = mode == 1 ? fqo_assign_closed(m, rb, m->L, &r) : fqo_assign(m, rb, m->L, &r);
        if (rc != FQO_OK) return rc;
        if (results) results[i] = r;
        if (counts) counts[r == FQO_NONE ? m->S : (r >> 16)] += 1; /* :970-974 */
    }
    return FQO_OK;
}

int fqo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int fqo_assign_batch_mt(const uint8_t* panel, uint32_t S, uint32_t L, uint8_t max_mismatches,
                        uint8_t min_mismatch_delta, int use_cache, const uint8_t* reads, uint64_t N,
                        uint32_t* results, uint64_t* counts, int threads) {
    if (threads <= 0) threads = fqo_max_threads();
    int used = 1, err = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
        int tid = 0, nt = 1;
#ifdef _OPENMP
        tid = omp_get_thread_num();
        nt = omp_get_num_threads();
#endif
        fqo_matcher* m = NULL;
        uint64_t* local = (uint64_t*)calloc((size_t)S + 1, sizeof(uint64_t));
        if (fqo_matcher_new(panel, S, L, max_mismatches, min_mismatch_delta, use_cache, &m) != FQO_OK || !local) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            err = 1;
        } else {
            const uint64_t lo = N * (uint64_t)tid / (uint64_t)nt, hi = N * (uint64_t)(tid + 1) / (uint64_t)nt;
            if (fqo_assign_batch(m, reads + lo * L, hi - lo, results ? results + lo : NULL, local, 0) != FQO_OK) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
                err = 1;
            }
            if (counts) {
#ifdef _OPENMP
#pragma omp critical
#endif
                for (uint32_t j = 0; j <= S; j++) counts[j] += local[j];
            }
        }
        if (tid == 0) used = nt;
        fqo_matcher_free(m);
        free(local);
    }
    return err ? FQO_ERR_ARG : used;
}
