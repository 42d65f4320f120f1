/* fqtk_synth.c — TEST INFRASTRUCTURE ONLY: the deterministic synthetic workload of the benchmark, as plain
 * C on the oracle's side of the fence, so that `bench.py --impl reference` (the CPU arm) loads nothing of the product.
 *
 * It generates exactly the stream of fqtk_b200/csrc/synth.cu (read i is a pure function of (seed, i, panel); panel =
 * greedy rejection on pairwise Hamming distance + optional degenerate rewriting); tests/test_oracle_kats.py checks the two
 * byte for byte.  Nothing here matches reads or decides assignments — it is a workload generator.  The only piece of the
 * reference it needs is the IUPAC mask table (src/lib/mod.rs:26-46), taken from fqtk_oracle.c (fqo_iupac_mask).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "fqtk_oracle.h"

static uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
typedef struct { uint64_t s; } rng_t;
static uint64_t rng_next(rng_t* g) {
    g->s += 0x9E3779B97F4A7C15ull;
    return mix64(g->s);
}
static uint32_t bounded(uint64_t r, uint32_t n) { return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32); }
static uint32_t popc4(uint32_t m) { return (m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u); }
static uint32_t nth_base(uint32_t m, uint32_t t) { /* index (0..3 = A,C,G,T) of the t-th set bit of a base set */
    for (uint32_t b = 0; b < 4u; b++) {
        if ((m >> b) & 1u) {
            if (t == 0u) return b;
            t--;
        }
    }
    return 0u;
}
/* the generator's own byte -> base-set map: the reference's masks, with no-calls = all four and lower case folded */
static uint32_t base_set(uint8_t b) {
    if (b == '.') return 15u;
    if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);
    return fqo_iupac_mask(b);
}

static void synth_read(const uint8_t* panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t i, uint8_t* out) {
    static const char BASES[4] = {'A', 'C', 'G', 'T'};
    rng_t g = {mix64(seed ^ (i * 0xD1B54A32D192ED03ull))};
    const uint32_t kind = bounded(rng_next(&g), 100u);
    if (kind >= 98u) { /* 2 %: uniform random L-mer */
        uint64_t r = 0;
        for (uint32_t k = 0; k < L; k++) {
            if ((k & 31u) == 0u) r = rng_next(&g);
            out[k] = (uint8_t)BASES[r & 3u];
            r >>= 2;
        }
        return;
    }
    const uint32_t j = bounded(rng_next(&g), S);
    for (uint32_t k = 0; k < L; k++) {
        uint32_t m = base_set(panel[(size_t)j * L + k]);
        if (m == 0u) m = 15u;
        const uint64_t r = rng_next(&g);
        const uint32_t n = popc4(m);
        uint32_t b = nth_base(m, (uint32_t)(((r >> 40) * (uint64_t)n) >> 24)); /* resolve degenerate positions */
        if ((r & 0xFFFFu) < 328u) b = (b + 1u + (uint32_t)((r >> 16) & 0xFFu) % 3u) & 3u; /* 0.5 % substitution */
        out[k] = (uint8_t)BASES[b];
        if (((r >> 24) & 0xFFFFu) < 131u) out[k] = 'N'; /* 0.2 % no-call */
    }
    if (kind >= 90u) { /* 8 %: near-miss, 2-3 forced substitutions at distinct positions */
        uint32_t nsub = 2u + (uint32_t)(rng_next(&g) & 1u);
        if (nsub > L) nsub = L;
        uint32_t pos[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        for (uint32_t s = 0; s < nsub; s++) {
            uint32_t pk;
            uint64_t r;
            do {
                r = rng_next(&g);
                pk = bounded(r, L);
            } while (pk == pos[0] || pk == pos[1] || pk == pos[2]);
            pos[s] = pk;
            const uint8_t c = out[pk];
            const uint32_t cur = c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
            out[pk] = (uint8_t)BASES[(cur + 1u + (uint32_t)(r & 0xFFFFu) % 3u) & 3u];
        }
    }
}

void fqo_synth_reads(const uint8_t* panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t first, uint64_t n,
                     uint8_t* out) {
    for (uint64_t t = 0; t < n; t++) synth_read(panel, S, L, seed, first + t, out + t * L);
}

int fqo_synth_panel(uint64_t seed, uint32_t S, uint32_t L, uint32_t min_distance, uint32_t n_degenerate, uint8_t* out) {
    static const char BASES[4] = {'A', 'C', 'G', 'T'};
    static const char CODES[] = "RYSWKMBDHVN";
    uint8_t* acc = (uint8_t*)malloc((size_t)S * L); /* accepted, 1 byte per base (0..3) */
    uint8_t* cand = (uint8_t*)malloc(L);
    uint8_t* orig = (uint8_t*)malloc(L);
    if (!acc || !cand || !orig) {
        free(acc); free(cand); free(orig);
        return -1;
    }
    uint64_t counter = 0, rejected = 0;
    uint32_t have = 0;
    int rc = 0;
    while (have < S) {
        rng_t g = {mix64(seed + 0x5851F42D4C957F2Dull * (++counter))};
        uint64_t r = 0;
        for (uint32_t k = 0; k < L; k++) {
            if ((k & 31u) == 0u) r = rng_next(&g);
            cand[k] = (uint8_t)(r & 3u);
            r >>= 2;
        }
        int ok = 1;
        for (uint32_t j = 0; j < have && ok; j++) {
            uint32_t d = 0;
            const uint8_t* o = acc + (size_t)j * L;
            for (uint32_t k = 0; k < L && d < min_distance; k++) d += (o[k] != cand[k]);
            ok = d >= min_distance;
        }
        if (!ok) {
            if (++rejected > 50000000ull) { rc = -1; goto done; } /* infeasible for this (S, L, min_distance) */
            continue;
        }
        memcpy(acc + (size_t)have * L, cand, L);
        have++;
    }
    for (size_t t = 0; t < (size_t)S * L; t++) out[t] = (uint8_t)BASES[acc[t]];
    if (n_degenerate > 0) {
        if (n_degenerate > L) n_degenerate = L;
        uint32_t used[256];
        for (uint32_t j = 0; j < S; j++) {
            uint8_t* bc = out + (size_t)j * L;
            memcpy(orig, bc, L);
            for (uint64_t attempt = 0;; attempt++) {
                memcpy(bc, orig, L);
                rng_t g = {mix64(seed ^ (0xA24BAED4963EE407ull * (j + 1)) ^ (attempt << 40))};
                uint32_t n_used = 0;
                while (n_used < n_degenerate) {
                    const uint64_t r = rng_next(&g);
                    const uint32_t pk = bounded(r, L);
                    int dup = 0;
                    for (uint32_t u = 0; u < n_used; u++) dup = dup || (used[u] == pk);
                    if (dup) continue;
                    const uint32_t have_mask = base_set(orig[pk]);
                    uint32_t n_ok = 0; /* uniformly among the degenerate codes that still admit the original base */
                    for (int c = 0; c < 11; c++) n_ok += (base_set((uint8_t)CODES[c]) & have_mask) != 0u;
                    uint32_t pick = (uint32_t)((r & 0xFFFFu) % n_ok);
                    for (int c = 0; c < 11; c++) {
                        if ((base_set((uint8_t)CODES[c]) & have_mask) != 0u) {
                            if (pick == 0u) {
                                bc[pk] = (uint8_t)CODES[c];
                                break;
                            }
                            pick--;
                        }
                    }
                    used[n_used++] = pk;
                }
                int unique = 1; /* samples.rs:112-115: barcodes must be unique as strings */
                for (uint32_t o = 0; o < j && unique; o++) unique = memcmp(out + (size_t)o * L, bc, L) != 0;
                if (unique) break;
                if (attempt > 1000) { rc = -1; goto done; }
            }
        }
    }
done:
    free(acc); free(cand); free(orig);
    return rc;
}
