// capi.cu — the extern "C" boundary (include/fqtk_b200.h) over the sm_100a kernels.
//
// Host-side responsibilities that mirror the reference (fulcrumgenomics/fqtk @ 45dbb99):
//   * BarcodeMatcher::new        src/lib/barcode_matching.rs:55-86   -> fqtk_b200_matcher_create (panel upper-casing,
//                                                                       max_ns_in_barcodes, encoding, memo table)
//   * BarcodeMatcher::assign     src/lib/barcode_matching.rs:165-186 -> length rules + no-call pre-filter for rows
//                                                                       whose length differs from L
//   * demux.rs:968-975           -> counts[S+1]
// No CPU matching path exists here: distances, decisions and the memo-table contents all come from the kernels.
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/fqtk_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace fq {
bool have_avx2();                                                                         // host_pack.cpp
void pack_stream(const uint8_t* in, uint64_t n_bytes, uint8_t* out, const uint8_t* lut);  // host_pack.cpp
}  // namespace fq

namespace fq {
cudaError_t synth_reads_device(const uint8_t* d_panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t first,
                               uint64_t n, uint8_t* d_ascii, uint32_t* d_packed, int sm_count, cudaStream_t stream);
void synth_reads_host(const uint8_t* panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t first, uint64_t n,
                      uint8_t* out);
int synth_panel_host(uint64_t seed, uint32_t S, uint32_t L, uint32_t min_distance, uint32_t n_degenerate,
                     uint8_t* out);
}  // namespace fq

namespace {
thread_local std::string g_err;
}
namespace fq {
void set_last_error(const std::string& msg) { g_err = msg; }  // for the other translation units of the library
}
namespace {

constexpr size_t TIER_BUDGET_BYTES = 132u << 10;  // shared memory the hot tier may take per CTA

// Defaults of fqtk_b200_matcher_create (the form without an options struct).  Thread-local: two host threads that each
// create a matcher for their own device cannot race on them.  The environment variables are read once
// (thread-safe static initialisation) and only seed these defaults; they exist for A/B timing.
int env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    return e ? atoi(e) : fallback;
}
fqtk_b200_options env_defaults() {
    static const fqtk_b200_options v = [] {
        fqtk_b200_options o{};
        o.struct_size = (uint32_t)sizeof(fqtk_b200_options);
        o.kernel = env_int("FQTK_B200_CK_NP", FQTK_B200_KERNEL_AUTO);
        o.table_budget = 32ull << 20;
        const int mb = env_int("FQTK_B200_CHUNK_MB", 0);
        o.chunk_bytes = (uint64_t)(mb > 0 ? mb : 32) << 20;
        const int load = env_int("FQTK_B200_G4_LOAD", 0);
        o.l2_table_load_pct = (load >= 5 && load <= 90) ? (uint32_t)load : 0u;
        return o;
    }();
    return v;
}
thread_local fqtk_b200_options t_defaults = env_defaults();

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(FQTK_B200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                         \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// decode() of one 4-bit mask (mod.rs:66-83): first IUPAC letter with that mask; the reference panics on mask 0
char decode_mask(uint32_t mask) {
    static const char T[17] = "?ACMGRSVTWYHKDBN";
    return T[mask & 15u];
}

constexpr int N_PIPE = 3;                       // chunks in flight for the host-buffer call
constexpr int N_STAGE = 4;                      // pinned staging slots of the host-pack route

}  // namespace

struct fqtk_b200_matcher {
    int device = 0;
    fqtk_b200_options opt{};  // this handle's options (create_ex), fixed at creation
    fq::LaunchGeometry geo{};
    uint32_t S = 0, L = 0, W = 0, P = 0;
    uint8_t max_mm = 0, min_delta = 0;
    uint32_t max_ns = 0;
    std::vector<uint8_t> panel;  // upper-cased ASCII, S x L
    uint4* d_planes = nullptr;
    uint4* d_planes2 = nullptr;
    uint32_t* d_not_exp = nullptr;
    uint32_t* d_sliced = nullptr;
    uint32_t* d_table = nullptr;
    uint32_t* d_tier = nullptr;
    uint32_t* d_bloom = nullptr;
    uint32_t* d_cuckoo = nullptr;
    uint32_t* d_g4 = nullptr;
    uint64_t g4_entries = 0, g4_slow_keys = 0;
    std::vector<uint32_t> h_not_exp;  // host copy of the panel's ~expected nibble words (fingerprint-table build)
    uint64_t cuckoo_entries = 0;
    uint64_t tier_entries = 0;
    unsigned long long* d_counts = nullptr;
    fq::MatchParams params{};
    int mode = FQTK_B200_MODE_BRUTE;
    uint64_t table_entries = 0, table_slots = 0, table_bytes = 0, table_candidates = 0;
    // host-buffer pipeline
    cudaStream_t streams[N_PIPE] = {};
    // host-pack route of assign_batch (fqtk_b200_matcher_set_host_pack): pinned staging ring + the events of its copies
    int host_pack_threads = 0;                 // 0 = off
    uint8_t* h_stage[N_STAGE] = {};
    size_t stage_cap = 0;
    cudaEvent_t stage_ev[N_STAGE] = {};
    uint8_t* d_in[N_PIPE] = {};
    uint32_t* d_out[N_PIPE] = {};
    uint32_t* d_len[N_PIPE] = {};
    size_t in_cap = 0, out_cap = 0;  // bytes / reads per pipeline slot
    void* d_route_ws = nullptr;      // routing workspace (per-warp histograms)
    size_t route_ws_bytes = 0;
    uint32_t* d_scratch = nullptr;   // packed scratch for the L > 32 ASCII route and the device segment gather
    size_t scratch_words = 0;
    uint32_t* d_seg_packed[N_PIPE] = {};  // packed scratch per pipeline slot for the host segment gather
    void* d_fq[3 * FQTK_B200_MAX_SEGMENTS] = {};  // raw FASTQ chunks + their offset / length tables (assign_fastq)
    size_t fq_cap[3 * FQTK_B200_MAX_SEGMENTS] = {};
    void* d_fq_head[FQTK_B200_MAX_SEGMENTS] = {};  // header offsets of the records (the whole-batch path)
    size_t fq_head_cap[FQTK_B200_MAX_SEGMENTS] = {};
    uint32_t* d_fq_len = nullptr;  // gathered barcode lengths (REST segments)
    size_t fq_len_cap = 0;
    size_t seg_packed_words[N_PIPE] = {};
};

namespace {

// ------------------------------------------------------------------------------------------------------
// memo-table construction: enumerate every A/C/G/T/N string within max_mm mismatches of some barcode, let the
// brute-force KERNEL compute each one's result, keep the Some(..) ones in an open-addressing table.
// ------------------------------------------------------------------------------------------------------
const uint32_t ALPHABET[5] = {1u, 2u, 4u, 8u, 15u};

// number of neighbourhood strings (with duplicates across barcodes), saturating at limit + 1
uint64_t neighbourhood_size(const std::vector<uint8_t>& masks, uint32_t S, uint32_t L, uint32_t mm, uint64_t limit) {
    const uint32_t K = std::min(mm, L);
    long double total = 0;
    std::vector<long double> ways(K + 1), nxt(K + 1);
    for (uint32_t j = 0; j < S; j++) {
        std::fill(ways.begin(), ways.end(), 0.0L);
        ways[0] = 1;
        for (uint32_t i = 0; i < L; i++) {
            const uint32_t e = masks[(size_t)j * L + i];
            uint32_t n_match = 0;
            for (uint32_t o : ALPHABET) n_match += (o & ~e & 0xFu) == 0u;
            const uint32_t n_mis = 5u - n_match;
            for (uint32_t k = 0; k <= K; k++) nxt[k] = ways[k] * n_match + (k ? ways[k - 1] * n_mis : 0.0L);
            ways.swap(nxt);
        }
        for (uint32_t k = 0; k <= K; k++) total += ways[k];
        if (total > (long double)limit) return limit + 1;
    }
    return (uint64_t)total;
}

struct Enumerator {
    const uint8_t* e;  // L masks of one barcode
    uint32_t L, W, K;
    std::vector<uint32_t>* out;
    uint32_t cur[fq::MAX_FAST_WORDS];
    std::vector<uint32_t>* owner = nullptr;  // optional: the barcode each string was enumerated from
    uint32_t barcode = 0;
    void rec(uint32_t i, uint32_t used) {
        if (i == L) {
            out->insert(out->end(), cur, cur + W);
            if (owner) owner->push_back(barcode);
            return;
        }
        const uint32_t word = i >> 3, sh = 4u * (i & 7u);
        for (uint32_t o : ALPHABET) {
            const bool mis = (o & ~(uint32_t)e[i] & 0xFu) != 0u;
            if (mis && used == K) continue;
            cur[word] = (cur[word] & ~(0xFu << sh)) | (o << sh);
            rec(i + 1, used + (mis ? 1u : 0u));
        }
        cur[word] &= ~(0xFu << sh);
    }
};

// Insert into the memo table (kernels.h: one slot per probe, linear probing).
template <int W>
void host_insert(std::vector<uint32_t>& table, uint32_t n_slots, const uint32_t* key, uint32_t val,
                 uint64_t& inserted) {
    uint32_t kw[W];
    for (int k = 0; k < W; k++) kw[k] = key[k];
    uint32_t b = fq::bucket_of_hash(fq::hash_key<W>(kw), n_slots);
    const int SW = fq::table_slot_words(W), VI = fq::table_value_index(W);
    for (;;) {
        uint32_t* ent = table.data() + (size_t)b * SW;
        if (ent[VI] == fq::NONE) {
            for (int k = 0; k < SW; k++) ent[k] = 0u;
            for (int k = 0; k < W; k++) ent[k] = kw[k];
            ent[VI] = val;
            inserted++;
            return;
        }
        bool same = true;
        for (int k = 0; k < W; k++) same = same && ent[k] == kw[k];
        if (same) return;  // the same string reached from two barcodes: identical value by construction
        b = (b + 1u == n_slots) ? 0u : b + 1u;
    }
}

// Hot tier: 2-choice cuckoo over the entries whose best distance is 0 (layout: kernels.h).  It is only a cache of
// the memo table, so a key that cannot be placed (or does not fit the shared-memory budget) is simply left out.
template <int W>
void build_tier(const std::vector<uint32_t>& keys, const std::vector<uint32_t>& res, uint64_t n, size_t budget_bytes,
                std::vector<uint32_t>& tier, uint32_t& slots_out, uint64_t& placed) {
    const int TE = fq::tier_entry_words(W);
    uint64_t n_hot = 0;
    for (uint64_t t = 0; t < n; t++) n_hot += (res[t] != fq::NONE && ((res[t] >> 8) & 0xFFu) == 0u);
    uint32_t slots = 16;
    while ((uint64_t)slots * 2 < n_hot * 5 && slots < (1u << 20)) slots <<= 1;            // load <= 0.4
    const int SEPV = fq::tier_separate_values(W) ? 1 : 0;
    while (slots > 16 && (size_t)slots * (TE + SEPV) * 4 > budget_bytes) slots >>= 1;  // shared-memory budget
    placed = 0;
    slots_out = 0;
    if (n_hot == 0 || (size_t)slots * (TE + SEPV) * 4 > budget_bytes) return;
    uint32_t shift = 32;
    while ((1u << (32 - shift)) < slots) shift--;
    std::vector<uint32_t> tk((size_t)slots * W, 0xFFFFFFFFu), tv(slots, fq::NONE);
    const uint64_t cap = (uint64_t)slots * 45 / 100;
    for (uint64_t t = 0; t < n && placed < cap; t++) {
        if (res[t] == fq::NONE || ((res[t] >> 8) & 0xFFu) != 0u) continue;
        uint32_t ck[W], cv = res[t];
        for (int k = 0; k < W; k++) ck[k] = keys[t * W + k];
        bool done = false;
        for (int kick = 0; kick < 256 && !done; kick++) {
            const uint32_t s[2] = {fq::tier_hash1<W>(ck) >> shift, fq::tier_hash2<W>(ck) >> shift};
            for (int c = 0; c < 2 && !done; c++) {  // already there (duplicate candidate)?
                bool same = tv[s[c]] != fq::NONE;
                for (int k = 0; k < W; k++) same = same && tk[(size_t)s[c] * W + k] == ck[k];
                if (same) done = true;
            }
            if (done) break;
            for (int c = 0; c < 2 && !done; c++) {
                if (tv[s[c]] == fq::NONE) {
                    for (int k = 0; k < W; k++) tk[(size_t)s[c] * W + k] = ck[k];
                    tv[s[c]] = cv;
                    placed++;
                    done = true;
                }
            }
            if (done) break;
            const uint32_t victim = s[(kick ^ (ck[0] >> 9)) & 1u];  // evict a resident and carry it on
            for (int k = 0; k < W; k++) std::swap(ck[k], tk[(size_t)victim * W + k]);
            std::swap(cv, tv[victim]);
        }
        // not placed after 256 kicks: whichever key is in hand stays out of the tier (it is still in the memo table)
    }
    // serialise into the device layout
    if (SEPV) {  // key entries, then the value array
        tier.assign((size_t)slots * (TE + 1), 0xFFFFFFFFu);
        for (uint32_t e = 0; e < slots; e++) {
            for (int k = 0; k < TE; k++) tier[(size_t)e * TE + k] = tk[(size_t)e * W + (k < W ? k : 0)];
            tier[(size_t)slots * TE + e] = tv[e];
        }
    } else {
        const int TV = fq::tier_value_index(W);
        tier.assign((size_t)slots * TE, 0u);
        for (uint32_t e = 0; e < slots; e++) {
            uint32_t* ent = tier.data() + (size_t)e * TE;
            for (int k = 0; k < W; k++) ent[k] = tk[(size_t)e * W + k];
            ent[TV] = tv[e];
            if (tv[e] == fq::NONE)
                for (int k = 0; k < TE; k++) ent[k] = 0xFFFFFFFFu;
        }
    }
    slots_out = slots;
}

// k_probe3's shared-memory table (kernels.h): every Some(..) memo entry whose key is pure A/C/G/T, keyed by the
// compressed 32-bit key, as an NP-ary cuckoo table of 4-byte quotient entries.  Returns 0 when built, 1 when the
// panel does not qualify (the matcher then stays on k_probe2), < 0 on a CUDA error.
// m->opt.kernel: -1 auto, 0 no cuckoo table (k_probe2 runs instead), 1 the L2-resident table only, 2 | 3 force the arity

template <int W>
int build_cuckoo_w(fqtk_b200_matcher* m, const std::vector<uint32_t>& keys, const std::vector<uint32_t>& res,
                   uint64_t n, bool exact_only) {
    // exact_only (k_probe5): only the entries with best distance 0; the rest of the neighbourhood is in the fingerprint table
    const uint32_t S = m->S, pad = m->params.last_pad;
    if (S + 1u > 8192u) return 1;
    // pure-A/C/G/T entries, de-duplicated (the same string is enumerated once per barcode it is close to)
    std::vector<std::pair<uint32_t, uint32_t>> ent;  // (compressed key, result word)
    for (uint64_t t = 0; t < n; t++) {
        if (res[t] == fq::NONE) continue;
        if (exact_only && ((res[t] >> 8) & 0xFFu) != 0u) continue;
        uint32_t kw[W];
        for (int k = 0; k < W; k++) kw[k] = keys[t * W + k];
        bool valid;
        const uint32_t k = fq::acgt_key<W>(kw, pad, valid);
        if (valid) ent.emplace_back(k, res[t]);
    }
    std::sort(ent.begin(), ent.end());
    ent.erase(std::unique(ent.begin(), ent.end()), ent.end());
    for (size_t t = 1; t < ent.size(); t++)
        if (ent[t].first == ent[t - 1].first) return 1;  // cannot happen: one string, one result
    if (ent.empty()) return 1;
    // value code = idx << (bb + nb) | best << nb | (next - next_min)
    uint32_t max_idx = 0, max_best = 0, min_next = 255, max_next = 0;
    for (auto& e : ent) {
        max_idx = std::max(max_idx, e.second >> 16);
        max_best = std::max(max_best, (e.second >> 8) & 0xFFu);
        min_next = std::min(min_next, e.second & 0xFFu);
        max_next = std::max(max_next, e.second & 0xFFu);
    }
    auto bits_for = [](uint32_t v) { uint32_t b = 0; while (b < 32 && (v >> b)) b++; return b; };
    const uint32_t ib = bits_for(max_idx), bb = bits_for(max_best), nb = bits_for(max_next - min_next);
    uint32_t cb = std::max(1u, ib + bb + nb);
    auto code_of = [&](uint32_t r) {
        return ((r >> 16) << (bb + nb)) | (((r >> 8) & 0xFFu) << nb) | ((r & 0xFFu) - min_next);
    };
    for (auto& e : ent)
        if (code_of(e.second) == (1u << cb) - 1u) { cb++; break; }  // the all-ones code means "empty slot"
    if (cb > 16) return 1;

    const uint32_t lb = bb + nb;
    // geometry: the smallest 2-ary layout at load <= 0.40, else the smallest 3-ary one at load <= 0.80, that leaves
    // room for >= 4 histogram replicas
    const size_t smem_max = (size_t)m->geo.max_smem_optin - 1024;
    struct Geo { uint32_t np, sb[3]; };
    std::vector<Geo> options;
    const int force = m->opt.kernel;
    if (force == 0 || force == 1) return 1;  // 0: k_probe2 only, 1: k_probe4's table only
    if (exact_only != (force == 5) && force != -1) return 1;  // 5: k_probe5 (exact-only table); 2 / 3: k_probe3 (full table)
    for (uint32_t s = std::max(cb, 4u); s <= 16; s++) {
        if (force != 3) {
            options.push_back({2, {s, s, 0}});
            options.push_back({2, {s + 1, s, 0}});
        }
        if (force != 2) options.push_back({3, {s, s, s}});
    }
    auto slots_of = [](const Geo& g) { uint64_t t = 0; for (uint32_t i = 0; i < g.np; i++) t += 1ull << g.sb[i]; return t; };
    std::stable_sort(options.begin(), options.end(), [&](const Geo& a, const Geo& b) {
        if (force < 0 && a.np != b.np) return a.np < b.np;  // prefer two probes per read whenever they fit
        return slots_of(a) < slots_of(b);
    });
    for (const Geo& g : options) {
        const uint64_t slots = slots_of(g);
        const double max_load = g.np == 2 ? 0.40 : 0.80;
        if ((double)ent.size() > max_load * (double)slots) continue;
        if ((exact_only ? fq::probe5_smem_bytes((uint32_t)slots, S, W, 1) : fq::probe3_smem_bytes((uint32_t)slots, S, 16)) > smem_max)
            continue;  // table + histogram + minimal stashes (k_probe5: + ~expected words + warp queues)
        uint32_t off[3] = {0, 0, 0};
        for (uint32_t i = 1; i < g.np; i++) off[i] = off[i - 1] + (1u << g.sb[i - 1]);
        std::vector<uint32_t> slot_key(slots, 0u), slot_code(slots, 0xFFFFFFFFu);  // code 0xFFFFFFFF = empty
        uint64_t rng = 0x9E3779B97F4A7C15ull;
        bool ok = true;
        for (auto& e : ent) {
            uint32_t ck = e.first, cc = code_of(e.second);
            uint32_t avoid = 0xFFFFFFFFu;
            bool placed = false;
            for (int kick = 0; kick < 2000 && !placed; kick++) {
                uint32_t pos[3];
                for (uint32_t i = 0; i < g.np; i++) pos[i] = off[i] + ((ck * fq::ck_mul((int)i)) >> (32u - g.sb[i]));
                for (uint32_t i = 0; i < g.np && !placed; i++)
                    if (slot_code[pos[i]] == 0xFFFFFFFFu) {
                        slot_key[pos[i]] = ck;
                        slot_code[pos[i]] = cc;
                        placed = true;
                    }
                if (placed) break;
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                uint32_t v = (uint32_t)(rng >> 33) % g.np;
                if (pos[v] == avoid) v = (v + 1u) % g.np;  // do not bounce straight back
                std::swap(ck, slot_key[pos[v]]);
                std::swap(cc, slot_code[pos[v]]);
                avoid = pos[v];
            }
            if (!placed) { ok = false; break; }
        }
        if (!ok) continue;
        // serialise + verify with the kernel's own arithmetic
        std::vector<uint32_t> words(slots, 0xFFFFFFFFu);
        for (uint32_t i = 0; i < g.np; i++)
            for (uint32_t q = 0; q < (1u << g.sb[i]); q++) {
                const uint32_t at = off[i] + q;
                if (slot_code[at] == 0xFFFFFFFFu) continue;
                words[at] = ((slot_key[at] * fq::ck_mul((int)i)) << g.sb[i]) | slot_code[at];
            }
        const uint32_t limit = (1u << cb) - 1u;
        const uint32_t bsh = 8u - nb, bmask8 = ((1u << bb) - 1u) << 8, nmask = (1u << nb) - 1u;
        uint32_t negmulb[3] = {0, 0, 0};
        for (uint32_t i = 0; i < g.np; i++) negmulb[i] = 0u - (fq::ck_mul((int)i) << g.sb[i]);
        for (auto& e : ent) {
            uint32_t u = 0xFFFFFFFFu;
            for (uint32_t i = 0; i < g.np; i++) {
                const uint32_t sl = (e.first * fq::ck_mul((int)i)) >> (32u - g.sb[i]);
                u = std::min(u, words[off[i] + sl] + e.first * negmulb[i]);
            }
            const uint32_t word = ((u >> lb) << 16) + (((u << bsh) & bmask8) | (u & nmask)) + min_next;
            if (!(u < limit) || word != e.second) return fail(FQTK_B200_ERR_CUDA, "cuckoo table self-check failed");
        }
        CU(cudaMalloc(&m->d_cuckoo, words.size() * 4));
        CU(cudaMemcpy(m->d_cuckoo, words.data(), words.size() * 4, cudaMemcpyHostToDevice));
        fq::MatchParams& p = m->params;
        p.ck_entries = m->d_cuckoo;
        p.ck_np = g.np;
        p.ck_words = (uint32_t)slots;
        for (uint32_t i = 0; i < 3; i++) {
            const uint32_t j = i < g.np ? i : 0;
            p.ck_off[i] = off[j];
            p.ck_shift[i] = 32u - g.sb[j];
            p.ck_negmulb[i] = negmulb[j];
        }
        p.ck_limit = limit;
        p.ck_lb = lb;
        p.ck_bsh = bsh;
        p.ck_bmask8 = bmask8;
        p.ck_nmask = nmask;
        p.ck_next_min = min_next;
        uint32_t cap = 16;
        while (cap < 64 && fq::probe3_smem_bytes((uint32_t)slots, S, cap + 1) <= smem_max) cap++;
        p.ck_stash_cap = cap;
        p.ck_exact_only = exact_only ? 1u : 0u;
        m->cuckoo_entries = ent.size();
        return 0;
    }
    return 1;
}

// k_probe4's global fingerprint table (kernels.h): every candidate string — every A/C/G/T/N string within max_mm of
// some barcode, whether its result is Some or None — as a 4-byte entry fingerprint | sample index | value code, eight to a
// 32-byte bucket, bucketised linear probing.  The kernel's lookup is replayed here for every key: a key must come back
// with its own value or be sent to the (exact) slow path; if any key comes back with a WRONG value — a fingerprint
// collision with an entry of another barcode that the read is also within max_mm of — the hash is re-seeded.
// Returns 0 when built, 1 when the panel does not qualify, < 0 on error.
template <int W>
int build_g4_w(fqtk_b200_matcher* m, const std::vector<uint32_t>& keys, const std::vector<uint32_t>& res,
               const std::vector<uint32_t>& owners, uint64_t n) {
    const uint32_t S = m->S;
    if (S + 1u > 8192u || owners.size() != n) return 1;
    struct Ent { uint32_t w[W]; uint32_t word, owner; };
    std::vector<Ent> ent(n);
    for (uint64_t t = 0; t < n; t++) {
        for (int k = 0; k < W; k++) ent[t].w[k] = keys[t * W + k];
        ent[t].word = res[t];
        // a Some entry is verified against its best match, a None candidate against the barcode that enumerated it
        ent[t].owner = res[t] != fq::NONE ? (res[t] >> 16) : owners[t];
    }
    auto key_less = [](const Ent& a, const Ent& b) {
        for (int k = W - 1; k >= 0; k--)
            if (a.w[k] != b.w[k]) return a.w[k] < b.w[k];
        return false;
    };
    auto key_eq = [](const Ent& a, const Ent& b) {
        for (int k = 0; k < W; k++)
            if (a.w[k] != b.w[k]) return false;
        return true;
    };
    std::sort(ent.begin(), ent.end(), key_less);
    for (size_t t = 1; t < ent.size(); t++)
        if (key_eq(ent[t], ent[t - 1]) && ent[t].word != ent[t - 1].word) return 1;  // cannot happen: one string, one result
    ent.erase(std::unique(ent.begin(), ent.end(), key_eq), ent.end());
    if (ent.empty()) return 1;
    auto distance_to = [&](const Ent& e, uint32_t j) {
        uint32_t ne[W];
        for (int k = 0; k < W; k++) ne[k] = m->h_not_exp[(size_t)j * W + k];
        return fq::nibble_distance<W>(e.w, ne);
    };
    std::vector<uint32_t> owner(ent.size());
    for (size_t t = 0; t < ent.size(); t++) {
        owner[t] = ent[t].owner;
        if (distance_to(ent[t], owner[t]) > m->max_mm)
            return fail(FQTK_B200_ERR_CUDA, "fingerprint table: a candidate is farther than max_mm from its barcode");
    }
    // value code = best << nb | (next - next_min)
    uint32_t max_best = 0, min_next = 255, max_next = 0;
    for (auto& e : ent) {
        if (e.word == fq::NONE) continue;
        max_best = std::max(max_best, (e.word >> 8) & 0xFFu);
        min_next = std::min(min_next, e.word & 0xFFu);
        max_next = std::max(max_next, e.word & 0xFFu);
    }
    if (min_next > max_next) min_next = max_next = 0;  // no Some entry at all
    auto bits_for = [](uint32_t v) { uint32_t b = 0; while (b < 32 && (v >> b)) b++; return b; };
    const uint32_t ib = std::max(1u, bits_for(S - 1u)), bb = bits_for(max_best), nb = bits_for(max_next - min_next);
    uint32_t cb = std::max(1u, bb + nb);
    const uint32_t max_code = (max_best << nb) | (max_next - min_next);
    while (max_code >= (1u << cb) - 2u) cb++;  // 2^cb - 2 = None candidate, 2^cb - 1 = empty slot
    if (ib + cb > 24u) return 1;               // fewer than 8 fingerprint bits
    const uint32_t fp_bits = 32u - ib - cb, fp_shift = ib + cb;
    auto entry_of = [&](size_t t, uint32_t fp_in_place) {
        const uint32_t w = ent[t].word;
        const uint32_t code = w == fq::NONE ? (1u << cb) - 2u : ((((w >> 8) & 0xFFu) << nb) | ((w & 0xFFu) - min_next));
        return fp_in_place | (owner[t] << cb) | code;
    };
    // shared memory of the kernel: ~expected words + histogram replicas + stashes
    const size_t smem_max = (size_t)m->geo.max_smem_optin - 1024;
    uint32_t rep = 16;
    while (rep > 1 && fq::probe4_smem_bytes(W, S, rep, 32) > smem_max) rep >>= 1;
    if (fq::probe4_smem_bytes(W, S, rep, 32) > smem_max) return 1;
    if (m->params.ck_exact_only)  // k_probe5 shares the replica count: it must also fit next to its cuckoo table
        while (rep > 1 && fq::probe5_smem_bytes(m->params.ck_words, S, W, rep) > smem_max) rep >>= 1;
    uint32_t cap = 32;
    while (cap < 64 && fq::probe4_smem_bytes(W, S, rep, cap + 1) <= smem_max) cap++;
    // load factor: measured on B200 (profiles/README.md); options.l2_table_load_pct overrides (A/B timing)
    // Measured on B200 (1 B reads): cfg 5 (k_probe4) 12.2 ms at 0.4, 10.75 at 0.5, 10.66 at 0.6, 11.4 at 0.75, 16.3 at 0.85;
    // cfg 4 (k_probe5) 6.19 ms at 0.3, 5.55 at 0.4, 4.73 at 0.55, 4.66 at 0.65, 5.18 at 0.75, 8.46 at 0.85 — a smaller table
    // stays L2-resident against the stream (the ~half of the probes that find nothing land on uniformly random buckets),
    // a fuller one walks overflow chains
    const double load = m->opt.l2_table_load_pct ? m->opt.l2_table_load_pct / 100.0 : 0.6;
    const uint64_t buckets64 = std::max<uint64_t>(16, (uint64_t)((double)ent.size() / (8 * load)) + 1);
    if (buckets64 >= (1ull << 28)) return 1;
    const uint32_t n_buckets = (uint32_t)buckets64;
    // insertion order: the closer a key is to its barcode, the more reads carry it (exact barcodes first), so the keys
    // that end up in an overflow bucket are the rare ones; None candidates last
    std::vector<uint32_t> insert_order(ent.size());
    for (size_t t = 0; t < ent.size(); t++) insert_order[t] = (uint32_t)t;
    std::stable_sort(insert_order.begin(), insert_order.end(), [&](uint32_t a, uint32_t b) {
        const uint32_t ka = ent[a].word == fq::NONE ? 256u : ((ent[a].word >> 8) & 0xFFu);
        const uint32_t kb = ent[b].word == fq::NONE ? 256u : ((ent[b].word >> 8) & 0xFFu);
        return ka < kb;
    });
    std::vector<uint32_t> table;
    uint32_t seed = 0;
    uint64_t slow_keys = 0;
    bool built = false;
    for (uint32_t attempt = 0; attempt < 8 && !built; attempt++) {
        seed = attempt * 0x632BE5ABu;
        table.assign((size_t)n_buckets * 8, 0xFFFFFFFFu);
        for (size_t ti = 0; ti < ent.size(); ti++) {
            const size_t t = insert_order[ti];
            uint32_t b, fpp;
            fq::g4_hashes<W>(ent[t].w, seed, n_buckets, b, fpp);
            fpp &= 0u - (1u << fp_shift);  // the fingerprint: top fp_bits of the second hash, in place
            while (table[(size_t)b * 8 + 7] != 0xFFFFFFFFu) b = (b + 1u == n_buckets) ? 0u : b + 1u;  // full: next bucket
            uint32_t j = 0;
            while (table[(size_t)b * 8 + j] != 0xFFFFFFFFu) j++;
            table[(size_t)b * 8 + j] = entry_of(t, fpp);
        }
        // replay the kernel's lookup (g4_lookup in match_kernels.cu) for every key
        built = true;
        slow_keys = 0;
        for (size_t t = 0; t < ent.size() && built; t++) {
            uint32_t b, fpp;
            fq::g4_hashes<W>(ent[t].w, seed, n_buckets, b, fpp);
            fpp &= 0u - (1u << fp_shift);  // the fingerprint: top fp_bits of the second hash, in place
            uint32_t tmin;
            for (;;) {
                tmin = 0xFFFFFFFFu;
                for (int j = 0; j < 8; j++) tmin = std::min(tmin, table[(size_t)b * 8 + j] ^ fpp);
                if (tmin < (1u << fp_shift) || table[(size_t)b * 8 + 7] == 0xFFFFFFFFu) break;
                b = (b + 1u == n_buckets) ? 0u : b + 1u;
            }
            if (tmin >= (1u << fp_shift)) return fail(FQTK_B200_ERR_CUDA, "fingerprint table self-check failed (key not found)");
            const uint32_t idx = std::min(tmin >> cb, S - 1u), code = tmin & ((1u << cb) - 1u);
            if (distance_to(ent[t], idx) > m->max_mm) {  // another key's entry: the kernel parks the read (exact slow path)
                slow_keys++;
                continue;
            }
            uint32_t word = fq::NONE;
            if (code < (1u << cb) - 2u) word = (idx << 16) | ((code >> nb) << 8) | ((code & ((1u << nb) - 1u)) + min_next);
            if (word != ent[t].word) built = false;  // collision that passes verification with another value: re-seed
        }
    }
    if (!built) return 1;
    if (getenv("FQTK_B200_DEBUG")) {
        uint64_t full = 0, used = 0;
        for (uint32_t b = 0; b < n_buckets; b++) {
            full += table[(size_t)b * 8 + 7] != 0xFFFFFFFFu;
            for (int j = 0; j < 8; j++) used += table[(size_t)b * 8 + j] != 0xFFFFFFFFu;
        }
        // walk length of an unsuccessful search from every bucket
        uint64_t walks = 0, longest = 0;
        for (uint32_t b = 0; b < n_buckets; b++) {
            uint32_t c = b, len = 0;
            while (table[(size_t)c * 8 + 7] != 0xFFFFFFFFu && len < 1000) { c = (c + 1u == n_buckets) ? 0u : c + 1u; len++; }
            walks += len;
            longest = std::max<uint64_t>(longest, len);
        }
        std::fprintf(stderr, "[fqtk_b200] fingerprint table: %zu keys, %u buckets, load %.3f, full buckets %.2f %%, mean miss walk %.3f, longest %llu, slow keys %llu, seed %u\n",
                     ent.size(), n_buckets, (double)used / (8.0 * n_buckets), 100.0 * full / n_buckets, (double)walks / n_buckets,
                     (unsigned long long)longest, (unsigned long long)slow_keys, seed);
    }
    CU(cudaMalloc(&m->d_g4, table.size() * 4));
    CU(cudaMemcpy(m->d_g4, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
    fq::MatchParams& p = m->params;
    p.g4_table = m->d_g4;
    p.g4_buckets = n_buckets;
    p.g4_seed = seed;
    p.g4_fp_bits = fp_bits;
    p.g4_fp_shift = fp_shift;
    p.g4_cb = cb;
    p.g4_nb = nb;
    p.g4_fp_mask = 0u - (1u << fp_shift);
    p.g4_lim = 1u << fp_shift;
    p.g4_cmask = (1u << cb) - 1u;
    p.g4_nmask = (1u << nb) - 1u;
    p.g4_code_none = (1u << cb) - 2u;
    p.g4_next_min = min_next;
    p.g4_hist_rep = rep;
    p.g4_stash_cap = cap;
    p.g4_flags = (uint32_t)env_int("FQTK_B200_G4_FLAGS", 3);
    m->g4_entries = ent.size();
    m->g4_slow_keys = slow_keys;
    return 0;
}

int build_table(fqtk_b200_matcher* m) {
    const uint32_t S = m->S, L = m->L, W = m->W;
    std::vector<uint8_t> masks((size_t)S * L);
    for (size_t t = 0; t < masks.size(); t++) masks[t] = (uint8_t)fq::encode_byte(m->panel[t]);
    const uint64_t cand = neighbourhood_size(masks, S, L, m->max_mm, m->opt.table_budget);
    if (cand > m->opt.table_budget || cand >= (1ull << 31)) return 1;  // over budget: stay in brute mode

    std::vector<uint32_t> keys;
    keys.reserve((size_t)cand * W);
    std::vector<uint32_t> owners;  // the barcode every candidate was enumerated from (it is within max_mm of it)
    owners.reserve((size_t)cand);
    Enumerator en{nullptr, L, W, std::min<uint32_t>(m->max_mm, L), &keys, {0, 0, 0, 0}};
    en.owner = &owners;
    for (uint32_t j = 0; j < S; j++) {
        en.barcode = j;
        en.e = masks.data() + (size_t)j * L;
        std::memset(en.cur, 0, sizeof en.cur);
        en.rec(0, 0);
    }
    const uint64_t n = keys.size() / W;
    m->table_candidates = n;

    // evaluate every candidate with the brute-force kernel
    struct DeviceBuffer {  // freed on every exit path of this function
        uint32_t* p = nullptr;
        ~DeviceBuffer() { if (p) cudaFree(p); }
    } d_keys, d_res;
    CU(cudaMalloc(&d_keys.p, std::max<size_t>(16, keys.size() * 4)));
    CU(cudaMalloc(&d_res.p, std::max<size_t>(16, n * 4)));
    CU(cudaMemcpy(d_keys.p, keys.data(), keys.size() * 4, cudaMemcpyHostToDevice));
    fq::ReadSource src{d_keys.p, nullptr, nullptr, 0, n};
    CU(fq::launch_brute(m->params, src, d_res.p, m->geo, m->streams[0]));
    std::vector<uint32_t> res(n);
    CU(cudaMemcpyAsync(res.data(), d_res.p, n * 4, cudaMemcpyDeviceToHost, m->streams[0]));
    CU(cudaStreamSynchronize(m->streams[0]));
    CU(cudaMemsetAsync(m->d_counts, 0, (size_t)(S + 1) * 8, m->streams[0]));  // the build pass is not a batch

    uint64_t n_some = 0;
    for (uint64_t t = 0; t < n; t++) n_some += res[t] != fq::NONE;
    // load factor: 0.2 while the table stays small (a probe almost never goes past its first slot), 0.4 beyond 64 MB
    const int SW = fq::table_slot_words((int)W);
    const uint64_t pct = (n_some * SW * 4 * 100 / 20 <= (64ull << 20)) ? 20 : 40;
    uint64_t slots64 = (n_some * 100 + pct - 1) / pct;
    if (slots64 < 64) slots64 = 64;
    if (slots64 >= (1ull << 31)) return 1;
    const uint32_t n_buckets = (uint32_t)slots64;
    std::vector<uint32_t> table((size_t)n_buckets * SW, 0xFFFFFFFFu);
    uint64_t inserted = 0;
    for (uint64_t t = 0; t < n; t++) {
        if (res[t] == fq::NONE) continue;
        const uint32_t* key = keys.data() + t * W;
        switch (W) {
            case 1: host_insert<1>(table, n_buckets, key, res[t], inserted); break;
            case 2: host_insert<2>(table, n_buckets, key, res[t], inserted); break;
            case 3: host_insert<3>(table, n_buckets, key, res[t], inserted); break;
            default: host_insert<4>(table, n_buckets, key, res[t], inserted); break;
        }
    }
    CU(cudaMalloc(&m->d_table, table.size() * 4));
    CU(cudaMemcpy(m->d_table, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
    m->table_entries = inserted;
    m->table_slots = n_buckets;
    m->table_bytes = table.size() * 4;
    m->params.table = m->d_table;
    m->params.n_buckets = n_buckets;

    // hot tier + Bloom filter for shared memory (k_probe2)
    const size_t smem_max = (size_t)m->geo.max_smem_optin - 2048;
    const size_t fixed = fq::probe2_fixed_smem_bytes(W, S, fq::probe2_threads());
    std::vector<uint32_t> tier;
    uint32_t tslots = 0;
    uint64_t placed = 0;
    const size_t budget = std::min(TIER_BUDGET_BYTES, smem_max > fixed ? smem_max - fixed : 0);
    switch (W) {
        case 1: build_tier<1>(keys, res, n, budget, tier, tslots, placed); break;
        case 2: build_tier<2>(keys, res, n, budget, tier, tslots, placed); break;
        case 3: build_tier<3>(keys, res, n, budget, tier, tslots, placed); break;
        default: build_tier<4>(keys, res, n, budget, tier, tslots, placed); break;
    }
    size_t tier_bytes = 0;
    if (tslots) {
        CU(cudaMalloc(&m->d_tier, tier.size() * 4));
        CU(cudaMemcpy(m->d_tier, tier.data(), tier.size() * 4, cudaMemcpyHostToDevice));
        m->params.tier_entries = m->d_tier;
        m->params.tier_slots = tslots;
        uint32_t shift = 32;
        while ((1u << (32 - shift)) < tslots) shift--;
        m->params.tier_shift = shift;
        m->tier_entries = placed;
        // replicas: as many as tile the banks once, while the tier stays within its budget
        const size_t entry_bytes = (size_t)fq::tier_entry_words((int)W) * 4;
        uint32_t rep = (uint32_t)fq::tier_max_rep((int)W);
        const size_t vals_bytes = fq::tier_separate_values((int)W) ? (size_t)tslots * 4 : 0;
        while (rep > 1 && (size_t)tslots * rep * entry_bytes + vals_bytes > budget) rep >>= 1;
        m->params.tier_rep = rep;
        tier_bytes = (size_t)tslots * rep * entry_bytes + vals_bytes;
    }
    // Bloom filter over every table key: >= 8 bits per key, at most 32 KB, only if it still fits
    {
        uint32_t words = 64;
        while ((uint64_t)words * 32 < inserted * 12 && words < 8192u) words <<= 1;
        if ((uint64_t)words * 32 >= inserted * 8 && fixed + tier_bytes + (size_t)words * 4 <= smem_max) {
            std::vector<uint32_t> bloom(words, 0u);
            uint32_t bshift = 32;
            while ((1u << (32 - bshift)) < words) bshift--;
            for (uint64_t t = 0; t < n; t++) {
                if (res[t] == fq::NONE) continue;
                uint32_t h;
                const uint32_t* key = keys.data() + t * W;
                switch (W) {
                    case 1: { uint32_t kw[1] = {key[0]}; h = fq::hash_key<1>(kw); break; }
                    case 2: { uint32_t kw[2] = {key[0], key[1]}; h = fq::hash_key<2>(kw); break; }
                    case 3: { uint32_t kw[3] = {key[0], key[1], key[2]}; h = fq::hash_key<3>(kw); break; }
                    default: { uint32_t kw[4] = {key[0], key[1], key[2], key[3]}; h = fq::hash_key<4>(kw); break; }
                }
                bloom[h >> bshift] |= fq::bloom_mask(h);
            }
            CU(cudaMalloc(&m->d_bloom, bloom.size() * 4));
            CU(cudaMemcpy(m->d_bloom, bloom.data(), bloom.size() * 4, cudaMemcpyHostToDevice));
            m->params.bloom = m->d_bloom;
            m->params.bloom_words = words;
            m->params.bloom_shift = bshift;
        }
    }
    // k_probe3's shared-memory cuckoo table of the pure-A/C/G/T entries (L <= 16)
    if (W <= 2 && m->opt.kernel != 5) {
        const int rc = (W == 1) ? build_cuckoo_w<1>(m, keys, res, n, false) : build_cuckoo_w<2>(m, keys, res, n, false);
        if (rc < 0) return rc;
    }
    // else, for one- and two-word panels, k_probe5: only the EXACT entries in shared memory + the fingerprint table below
    bool probe5 = false;
    if (m->params.ck_np == 0 && W <= 2 && (m->opt.kernel == -1 || m->opt.kernel == 5)) {
        const int rc = (W == 1) ? build_cuckoo_w<1>(m, keys, res, n, true) : build_cuckoo_w<2>(m, keys, res, n, true);
        if (rc < 0) return rc;
        probe5 = m->params.ck_np != 0;
    }
    // else k_probe4's L2-resident fingerprint table of every candidate.  Measured on B200: it wins where k_probe2 has no
    // useful hot tier (cfg 5, W = 3: 27 -> 10.7 ms per 1 B reads); with a hot tier that ends 80 % of the reads in shared
    // memory (cfg 4, W = 2) k_probe2 is still ahead (5.2 vs 5.7 ms), so one- and two-word panels only take it on request
    if ((m->params.ck_np == 0 || probe5) && m->opt.kernel != 0 && (W >= 3 || m->opt.kernel == 1 || probe5)) {
        const int rc = (W == 1) ? build_g4_w<1>(m, keys, res, owners, n) : (W == 2) ? build_g4_w<2>(m, keys, res, owners, n)
                     : (W == 3) ? build_g4_w<3>(m, keys, res, owners, n) : build_g4_w<4>(m, keys, res, owners, n);
        if (rc < 0) return rc;
        if (probe5 && m->params.g4_table == nullptr) {  // no fingerprint table after all: k_probe5 cannot run
            cudaFree(m->d_cuckoo);
            m->d_cuckoo = nullptr;
            m->params.ck_entries = nullptr;
            m->params.ck_np = 0;
            m->params.ck_exact_only = 0;
            m->cuckoo_entries = 0;
        }
    }
    return 0;
}

int ensure_pipeline(fqtk_b200_matcher* m, size_t in_bytes, size_t out_reads, bool want_len) {
    if (in_bytes > m->in_cap) {
        for (int s = 0; s < N_PIPE; s++) {
            if (m->d_in[s]) cudaFree(m->d_in[s]);
            m->d_in[s] = nullptr;
            CU(cudaMalloc(&m->d_in[s], in_bytes));
        }
        m->in_cap = in_bytes;
    }
    if (out_reads > m->out_cap) {
        for (int s = 0; s < N_PIPE; s++) {
            if (m->d_out[s]) cudaFree(m->d_out[s]);
            if (m->d_len[s]) cudaFree(m->d_len[s]);
            m->d_out[s] = m->d_len[s] = nullptr;
            CU(cudaMalloc(&m->d_out[s], out_reads * 4));
        }
        m->out_cap = out_reads;
    }
    if (want_len) {
        for (int s = 0; s < N_PIPE; s++)
            if (!m->d_len[s]) CU(cudaMalloc(&m->d_len[s], m->out_cap * 4));
    }
    return FQTK_B200_OK;
}

int ensure_scratch(uint32_t** slot, size_t* cap, size_t words) {
    if (words > *cap) {
        if (*slot) cudaFree(*slot);
        *slot = nullptr;
        *cap = 0;
        CU(cudaMalloc(slot, words * 4));
        *cap = words;
    }
    return FQTK_B200_OK;
}

// `pipe_slot` >= 0: called from the host-buffer pipeline, whose chunks run on N_PIPE independent streams — scratch must
// then be the slot's own (a shared buffer would be overwritten by the next chunk's pack while this chunk's kernel reads it).
int run_device(fqtk_b200_matcher* m, const fq::ReadSource& src, uint32_t* d_results, cudaStream_t st, int pipe_slot = -1) {
    if (src.n >= (1ull << 32)) return fail(FQTK_B200_ERR_ARG, "n_reads must be < 2^32 per device call");
    fq::ReadSource s = src;
    if (m->W > (uint32_t)fq::MAX_FAST_WORDS && s.ascii) {  // L > 32: pack first, then the long-barcode kernel
        uint32_t** buf = pipe_slot >= 0 ? &m->d_seg_packed[pipe_slot] : &m->d_scratch;
        size_t* cap = pipe_slot >= 0 ? &m->seg_packed_words[pipe_slot] : &m->scratch_words;
        const int rc = ensure_scratch(buf, cap, (size_t)s.n * m->W + 4);
        if (rc != FQTK_B200_OK) return rc;
        CU(fq::launch_pack(s.ascii, s.n, m->L, s.stride, *buf, m->geo, st));
        s.ascii = nullptr;   // s.lengths stays: rows whose length differs from L are None (barcode_matching.rs:167-169;
        s.packed = *buf;     // longer rows were vetted by the host), the kernel skips them
    }
    if (m->mode == FQTK_B200_MODE_TABLE)
        CU(fq::launch_probe(m->params, s, d_results, m->geo, st));
    else
        CU(fq::launch_brute(m->params, s, d_results, m->geo, st));
    return FQTK_B200_OK;
}

// ---- assign_batch with encode() done by host threads (opt-in: fqtk_b200_matcher_set_host_pack) -------------------------
// The ASCII rows of a batch cross PCIe at 16 bytes per read (L = 16) and the call runs at the copy rate.  With rows back to
// back and L a multiple of 8 the host can encode them itself at more than that rate (host_pack.cpp: AVX2, 32 symbols per
// step) and ship the reference's own BitEnc words — half the bytes — to the packed-route kernel.  T packer threads live for
// the duration of the call; each packs its slice of every chunk into a ring of N_STAGE pinned staging slots, the calling
// thread sends a chunk on as soon as all slices are in (H2D -> kernel -> D2H on the pipeline streams) and frees a slot
// when the event behind its copy has fired.  No locks: per-chunk arrival counters and one `released` counter.
// Packing costs host memory traffic (rows read, words written, words read by the DMA engine: 36 bytes per read at L = 16
// against 20 for the plain route) and leaves PCIe half idle, so `mix` of every 8 chunks go the plain way — ASCII rows
// straight over PCIe, encode() in the kernel — to load both at once (FQTK_B200_HOST_PACK_MIX, 0 .. 7).
int host_pack_mix() {
    static const int v = [] {
        const int e = env_int("FQTK_B200_HOST_PACK_MIX", 3);
        return e < 0 ? 0 : (e > 7 ? 7 : e);
    }();
    return v;
}
int assign_batch_host_pack(fqtk_b200_matcher* m, const uint8_t* rows, uint64_t n, uint32_t* results, int T) {
    static const std::vector<uint8_t> lut = [] {
        std::vector<uint8_t> t(256);
        for (uint32_t v = 0; v < 256u; v++) t[v] = (uint8_t)fq::encode_byte(v);
        return t;
    }();
    const uint32_t L = m->L;
    const uint64_t row_bytes = (uint64_t)m->W * 4u;  // == L / 2
    uint64_t chunk = std::max<uint64_t>(m->opt.chunk_bytes / (4u * row_bytes), 65536) & ~3ull;  // finer than the plain
    chunk = std::min<uint64_t>(chunk, n);                                                       // route: pack || copy
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    const int mix = host_pack_mix();
    auto plain = [mix](uint64_t c) { return (int)((c * (uint64_t)mix) % 8u) + mix >= 8; };  // `mix` of every 8 chunks, spread out
    int rc = ensure_pipeline(m, (size_t)(chunk * (mix ? (uint64_t)L : row_bytes) + 16), (size_t)chunk, false);
    if (rc != FQTK_B200_OK) return rc;
    if (chunk * row_bytes > m->stage_cap) {
        for (int k = 0; k < N_STAGE; k++) {
            if (m->h_stage[k]) cudaFreeHost(m->h_stage[k]);
            m->h_stage[k] = nullptr;
        }
        m->stage_cap = 0;
        for (int k = 0; k < N_STAGE; k++) CU(cudaHostAlloc(&m->h_stage[k], chunk * row_bytes, cudaHostAllocPortable));
        m->stage_cap = chunk * row_bytes;
    }
    for (int k = 0; k < N_STAGE; k++)
        if (!m->stage_ev[k]) CU(cudaEventCreateWithFlags(&m->stage_ev[k], cudaEventDisableTiming));

    std::unique_ptr<std::atomic<int>[]> arrived(new std::atomic<int>[n_chunks]);
    for (uint64_t c = 0; c < n_chunks; c++) arrived[c].store(plain(c) ? T : 0, std::memory_order_relaxed);
    std::atomic<uint64_t> released{0};  // chunks whose H2D copy is complete: chunk c may be packed iff c < released + N_STAGE
    std::atomic<bool> stop{false};
    auto packer = [&](int t) {
        for (uint64_t c = 0; c < n_chunks; c++) {
            if (plain(c)) continue;
            while (c >= released.load(std::memory_order_acquire) + (uint64_t)N_STAGE) {
                if (stop.load(std::memory_order_relaxed)) return;
                std::this_thread::yield();
            }
            const uint64_t c0 = c * chunk, cnt = std::min(chunk, n - c0);
            // slices start at multiples of four rows: 16-byte aligned in the staging slot (non-temporal stores)
            const uint64_t lo = (cnt * (uint64_t)t / (uint64_t)T) & ~3ull;
            const uint64_t hi = t + 1 == T ? cnt : ((cnt * (uint64_t)(t + 1) / (uint64_t)T) & ~3ull);
            if (hi > lo)
                fq::pack_stream(rows + (c0 + lo) * L, (hi - lo) * L, m->h_stage[c % N_STAGE] + lo * row_bytes, lut.data());
            arrived[c].fetch_add(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    pool.reserve((size_t)T);
    struct Joiner {  // every exit path: stop the packers, then wait for them
        std::vector<std::thread>& pool;
        std::atomic<bool>& stop;
        ~Joiner() {
            stop.store(true);
            for (auto& th : pool) th.join();
        }
    } joiner{pool, stop};
    try {
        for (int t = 0; t < T; t++) pool.emplace_back(packer, t);
    } catch (const std::exception& e) {  // (nothing may unwind across the C boundary)
        return fail(FQTK_B200_ERR_ARG, std::string("host-pack route: cannot start ") + std::to_string(T) + " threads: " + e.what());
    }

    uint64_t submitted = 0;
    auto poll = [&]() {  // staging slots whose copy has completed
        uint64_t r = released.load(std::memory_order_relaxed);
        while (r < submitted && cudaEventQuery(m->stage_ev[r % N_STAGE]) == cudaSuccess) r++;
        released.store(r, std::memory_order_release);
    };
    for (uint64_t c = 0; c < n_chunks; c++) {
        while (arrived[c].load(std::memory_order_acquire) < T) {
            poll();
            std::this_thread::yield();
        }
        const uint64_t c0 = c * chunk, cnt = std::min(chunk, n - c0);
        const int slot = (int)(c % N_PIPE);
        cudaStream_t st = m->streams[slot];
        uint32_t* d_words = reinterpret_cast<uint32_t*>(m->d_in[slot]);
        fq::ReadSource src{d_words, nullptr, nullptr, 0, cnt};
        if (plain(c)) {  // (its staging slot is not used: the event marks it free at once)
            CU(cudaMemcpyAsync(m->d_in[slot], rows + c0 * L, (size_t)(cnt * L), cudaMemcpyHostToDevice, st));
            src = fq::ReadSource{nullptr, m->d_in[slot], nullptr, L, cnt};
        } else {
            CU(cudaMemcpyAsync(d_words, m->h_stage[c % N_STAGE], (size_t)(cnt * row_bytes), cudaMemcpyHostToDevice, st));
        }
        CU(cudaEventRecord(m->stage_ev[c % N_STAGE], st));
        submitted = c + 1;
        rc = run_device(m, src, m->d_out[slot], st, slot);
        if (rc != FQTK_B200_OK) return rc;
        CU(cudaMemcpyAsync(results + c0, m->d_out[slot], cnt * 4, cudaMemcpyDeviceToHost, st));
        poll();
    }
    for (int s = 0; s < N_PIPE; s++) CU(cudaStreamSynchronize(m->streams[s]));
    return FQTK_B200_OK;
}

// The host-buffer calls are synchronous: on EVERY exit path (errors included) the pipeline streams are drained, so no
// DMA is still reading the caller's rows or writing the caller's results when the call returns.
struct PipelineDrain {
    fqtk_b200_matcher* m;
    ~PipelineDrain() {
        for (int s = 0; s < N_PIPE; s++)
            if (m->streams[s]) cudaStreamSynchronize(m->streams[s]);
    }
};

}  // namespace

// ======================================================================================================
extern "C" {

const char* fqtk_b200_last_error(void) { return g_err.c_str(); }

int fqtk_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void fqtk_b200_set_table_budget(uint64_t max_candidates) { t_defaults.table_budget = max_candidates; }

void fqtk_b200_set_cuckoo_arity(int arity) {
    t_defaults.kernel = ((arity >= 0 && arity <= 3) || arity == 5) ? arity : FQTK_B200_KERNEL_AUTO;
}

void fqtk_b200_options_init(fqtk_b200_options* opts) {
    if (opts) *opts = env_defaults();
}

uint64_t fqtk_b200_kernel_launches(void) { return fq::kernel_launches(); }

int fqtk_b200_matcher_create(const uint8_t* panel_ascii, uint32_t S, uint32_t L, uint8_t max_mm, uint8_t min_delta,
                             int use_cache, int device, fqtk_b200_matcher** out) {
    const fqtk_b200_options o = t_defaults;  // this thread's defaults (fqtk_b200_set_*), seeded from the environment
    return fqtk_b200_matcher_create_ex(panel_ascii, S, L, max_mm, min_delta, use_cache, device, &o, out);
}

int fqtk_b200_matcher_create_ex(const uint8_t* panel_ascii, uint32_t S, uint32_t L, uint8_t max_mm, uint8_t min_delta,
                                int use_cache, int device, const fqtk_b200_options* opts, fqtk_b200_matcher** out) {
    if (!out) return fail(FQTK_B200_ERR_ARG, "out is NULL");
    fqtk_b200_options opt = env_defaults();
    if (opts) {
        if (opts->struct_size < 8u || opts->struct_size > sizeof(fqtk_b200_options))
            return fail(FQTK_B200_ERR_ARG, "options.struct_size: call fqtk_b200_options_init first");
        std::memcpy(&opt, opts, opts->struct_size);  // fields an older caller does not know keep their defaults
        opt.struct_size = (uint32_t)sizeof(fqtk_b200_options);
        if (opt.kernel < -1 || opt.kernel > 5 || opt.kernel == 4) return fail(FQTK_B200_ERR_ARG, "options.kernel out of range");
        if (opt.table_budget == 0) opt.table_budget = 32ull << 20;
        if (opt.chunk_bytes == 0) opt.chunk_bytes = 32ull << 20;
        if (opt.l2_table_load_pct && (opt.l2_table_load_pct < 5 || opt.l2_table_load_pct > 90))
            return fail(FQTK_B200_ERR_ARG, "options.l2_table_load_pct must be 0 (auto) or 5..90");
    }
    *out = nullptr;
    if (S == 0) return fail(FQTK_B200_ERR_EMPTY_PANEL, "Must provide at least one sample");
    if (L == 0) return fail(FQTK_B200_ERR_EMPTY_BARCODE, "Sample barcode cannot be empty string");
    if (!panel_ascii) return fail(FQTK_B200_ERR_ARG, "panel_ascii is NULL");
    if (S > FQTK_B200_MAX_SAMPLES) return fail(FQTK_B200_ERR_UNSUPPORTED, "more than 65535 samples");
    if (L > FQTK_B200_MAX_BARCODE_LEN) return fail(FQTK_B200_ERR_UNSUPPORTED, "barcodes longer than 254 bases");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(FQTK_B200_ERR_CUDA, "no CUDA device: fqtk_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FQTK_B200_ERR_ARG, "bad device ordinal");
    CU(cudaSetDevice(device));
    if (const int mb = env_int("FQTK_B200_L2_PERSIST_MB", -1); mb >= 0)  // A/B: persisting-L2 carve-out of the device
        CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)mb << 20));

    fqtk_b200_matcher* m = new (std::nothrow) fqtk_b200_matcher();
    if (!m) return fail(FQTK_B200_ERR_ARG, "out of host memory");
    m->device = device;
    m->opt = opt;
    m->S = S;
    m->L = L;
    m->W = fq::words_for_len(L);
    m->P = fq::planes_for_len(L);
    m->max_mm = max_mm;
    m->min_delta = min_delta;
    cudaDeviceProp prop{};
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete m;
        return cuda_fail(e, "cudaGetDeviceProperties");
    }
    m->geo.sm_count = prop.multiProcessorCount;
    m->geo.max_smem_optin = (int)prop.sharedMemPerBlockOptin;

    // barcode_matching.rs:67-76: upper-case, count no-calls, encode
    m->panel.resize((size_t)S * L);
    for (uint32_t j = 0; j < S; j++) {
        uint32_t ns = 0;
        for (uint32_t i = 0; i < L; i++) {
            uint8_t b = panel_ascii[(size_t)j * L + i];
            if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);
            m->panel[(size_t)j * L + i] = b;
            ns += fq::byte_is_nocall(b);
        }
        m->max_ns = std::max(m->max_ns, ns);
    }
    const uint32_t W = m->W, P = m->P;
    std::vector<uint4> planes((size_t)(S + 1) * P, make_uint4(0, 0, 0, 0));  // one spare entry: pairs of barcodes
    std::vector<uint32_t> not_exp((size_t)S * W, 0u);
    for (uint32_t j = 0; j < S; j++) {
        for (uint32_t i = 0; i < L; i++) {
            const uint32_t emask = fq::encode_byte(m->panel[(size_t)j * L + i]);
            const uint32_t forbid = ~emask & 0xFu;
            not_exp[(size_t)j * W + (i >> 3)] |= forbid << (4u * (i & 7u));
            uint4& q = planes[(size_t)j * P + (i >> 5)];
            const uint32_t bit = 1u << (i & 31u);
            if (forbid & 1u) q.x |= bit;
            if (forbid & 2u) q.y |= bit;
            if (forbid & 4u) q.z |= bit;
            if (forbid & 8u) q.w |= bit;
        }
    }
    std::vector<uint4> planes2;
    if (L <= 16) {  // two barcodes per plane word for k_brute<.., PK = 2>
        planes2.assign((S + 1) / 2, make_uint4(0, 0, 0, 0));
        for (uint32_t j = 0; j < S; j++) {
            const uint4 q = planes[j];
            uint4& d = planes2[j / 2];
            const uint32_t sh = (j & 1u) ? 16u : 0u;
            d.x |= q.x << sh;
            d.y |= q.y << sh;
            d.z |= q.z << sh;
            d.w |= q.w << sh;
        }
    }
    int rc = FQTK_B200_OK;
    auto bail = [&](int code) {
        fqtk_b200_matcher_destroy(m);
        return code;
    };
#define CUB(call)                                      \
    do {                                               \
        cudaError_t e__ = (call);                      \
        if (e__ != cudaSuccess) return bail(cuda_fail(e__, #call)); \
    } while (0)
    CUB(cudaMalloc(&m->d_planes, planes.size() * sizeof(uint4)));
    CUB(cudaMemcpy(m->d_planes, planes.data(), planes.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    if (!planes2.empty()) {
        CUB(cudaMalloc(&m->d_planes2, planes2.size() * sizeof(uint4)));
        CUB(cudaMemcpy(m->d_planes2, planes2.data(), planes2.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    }
    CUB(cudaMalloc(&m->d_not_exp, not_exp.size() * 4));
    CUB(cudaMemcpy(m->d_not_exp, not_exp.data(), not_exp.size() * 4, cudaMemcpyHostToDevice));
    if (W <= (uint32_t)fq::MAX_FAST_WORDS && S + 1u <= 8192u) {
        // k_brute_sliced's transposed panel: word (g, i, v), bit j: barcode 32g + j mismatches a read symbol with mask v at i
        const uint32_t G = (S + 31u) / 32u, LP = 8u * W;
        std::vector<uint32_t> sliced((size_t)G * LP * 16u, 0u);
        for (uint32_t j = 0; j < S; j++)
            for (uint32_t i = 0; i < L; i++) {
                const uint32_t forbid = ~fq::encode_byte(m->panel[(size_t)j * L + i]) & 0xFu;
                for (uint32_t v = 0; v < 16u; v++)
                    if (v & forbid) sliced[((size_t)(j >> 5) * LP + i) * 16u + v] |= 1u << (j & 31u);
            }
        CUB(cudaMalloc(&m->d_sliced, sliced.size() * 4));
        CUB(cudaMemcpy(m->d_sliced, sliced.data(), sliced.size() * 4, cudaMemcpyHostToDevice));
    }
    CUB(cudaMalloc(&m->d_counts, (size_t)(S + 1) * 8));
    CUB(cudaMemset(m->d_counts, 0, (size_t)(S + 1) * 8));
    for (int s = 0; s < N_PIPE; s++) CUB(cudaStreamCreateWithFlags(&m->streams[s], cudaStreamNonBlocking));
#undef CUB
    m->params.planes = m->d_planes;
    m->params.planes2 = m->d_planes2;
    m->params.not_exp = m->d_not_exp;
    m->params.sliced = m->d_sliced;
    m->params.table = nullptr;
    m->params.counts = m->d_counts;
    m->params.S = S;
    m->params.L = L;
    m->params.W = W;
    m->params.P = P;
    m->params.max_mm = max_mm;
    m->params.min_delta = min_delta;
    m->params.last_pad = fq::last_word_pad_for_len(L);
    m->params.n_buckets = 0;
    m->params.tier_entries = nullptr;
    m->params.tier_slots = 0;
    m->params.tier_shift = 32;
    m->params.tier_rep = 1;
    m->params.hist_rep = fq::probe2_hist_rep(S);
    m->params.bloom = nullptr;
    m->params.bloom_words = 0;
    m->params.bloom_shift = 32;
    m->params.ck_entries = nullptr;
    m->params.ck_np = 0;
    m->params.ck_words = 0;
    m->params.ck_one = 1;
    m->params.ck_four = 4;
    m->params.ck_stash_cap = 32;
    m->params.ck_exact_only = 0;
    m->params.g4_table = nullptr;
    m->params.g4_buckets = 0;
    m->params.g4_hist_rep = 1;
    m->params.g4_stash_cap = 32;
    m->h_not_exp = not_exp;
    m->mode = FQTK_B200_MODE_BRUTE;
    if (use_cache && W <= (uint32_t)fq::MAX_FAST_WORDS) {
        rc = build_table(m);
        if (rc < 0) return bail(rc);
        if (rc == 0) m->mode = FQTK_B200_MODE_TABLE;
    }
    cudaError_t es = cudaDeviceSynchronize();
    if (es != cudaSuccess) return bail(cuda_fail(es, "create sync"));
    *out = m;
    return FQTK_B200_OK;
}

void fqtk_b200_matcher_destroy(fqtk_b200_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    for (int s = 0; s < N_PIPE; s++) {
        if (m->d_in[s]) cudaFree(m->d_in[s]);
        if (m->d_out[s]) cudaFree(m->d_out[s]);
        if (m->d_len[s]) cudaFree(m->d_len[s]);
        if (m->streams[s]) cudaStreamDestroy(m->streams[s]);
    }
    if (m->d_scratch) cudaFree(m->d_scratch);
    for (void* q : m->d_fq)
        if (q) cudaFree(q);
    for (void* q : m->d_fq_head)
        if (q) cudaFree(q);
    if (m->d_fq_len) cudaFree(m->d_fq_len);
    if (m->d_route_ws) cudaFree(m->d_route_ws);
    for (int k = 0; k < N_STAGE; k++) {
        if (m->h_stage[k]) cudaFreeHost(m->h_stage[k]);
        if (m->stage_ev[k]) cudaEventDestroy(m->stage_ev[k]);
    }
    for (int s = 0; s < N_PIPE; s++)
        if (m->d_seg_packed[s]) cudaFree(m->d_seg_packed[s]);
    if (m->d_planes) cudaFree(m->d_planes);
    if (m->d_planes2) cudaFree(m->d_planes2);
    if (m->d_not_exp) cudaFree(m->d_not_exp);
    if (m->d_sliced) cudaFree(m->d_sliced);
    if (m->d_table) cudaFree(m->d_table);
    if (m->d_tier) cudaFree(m->d_tier);
    if (m->d_bloom) cudaFree(m->d_bloom);
    if (m->d_cuckoo) cudaFree(m->d_cuckoo);
    if (m->d_g4) cudaFree(m->d_g4);
    if (m->d_counts) cudaFree(m->d_counts);
    delete m;
}

int fqtk_b200_matcher_get_info(const fqtk_b200_matcher* m, fqtk_b200_matcher_info* info) {
    if (!m || !info) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    info->n_samples = m->S;
    info->barcode_len = m->L;
    info->words_per_read = m->W;
    info->max_ns_in_barcodes = m->max_ns;
    info->mode = (uint32_t)m->mode;
    info->device = (uint32_t)m->device;
    info->table_entries = m->table_entries;
    info->table_slots = m->table_slots;
    info->table_bytes = m->table_bytes;
    info->table_candidates = m->table_candidates;
    info->tier_entries = m->tier_entries;
    info->tier_slots = m->params.tier_slots;
    info->cuckoo_entries = m->cuckoo_entries;
    info->cuckoo_probes = m->params.ck_np;
    info->cuckoo_slots = m->params.ck_words;
    info->l2_table_entries = m->g4_entries;
    info->l2_table_bytes = (uint64_t)m->params.g4_buckets * 32u;
    info->l2_table_slow_keys = m->g4_slow_keys;
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_set_host_pack(fqtk_b200_matcher* m, int threads) {
    if (!m) return fail(FQTK_B200_ERR_ARG, "NULL matcher");
    if (threads < 0) {  // auto: the CPUs this thread may run on, at most 16
        cpu_set_t set;
        CPU_ZERO(&set);
        threads = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        threads = std::max(1, std::min(threads, 16));
    }
    m->host_pack_threads = std::min(threads, 64);
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_set_mode(fqtk_b200_matcher* m, int mode) {
    if (!m) return fail(FQTK_B200_ERR_ARG, "NULL matcher");
    if (mode == FQTK_B200_MODE_BRUTE) {
        m->mode = mode;
        return FQTK_B200_OK;
    }
    if (mode == FQTK_B200_MODE_TABLE) {
        if (!m->d_table) return fail(FQTK_B200_ERR_ARG, "this matcher has no memo table");
        m->mode = mode;
        return FQTK_B200_OK;
    }
    return fail(FQTK_B200_ERR_ARG, "unknown mode");
}

int fqtk_b200_matcher_assign_packed_device(fqtk_b200_matcher* m, const uint32_t* d_packed, uint64_t n,
                                           uint32_t* d_results, void* stream) {
    if (!m || (n && (!d_packed || !d_results))) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    if ((reinterpret_cast<uintptr_t>(d_packed) & 15u) || (reinterpret_cast<uintptr_t>(d_results) & 3u))
        return fail(FQTK_B200_ERR_ARG, "d_packed must be 16-byte aligned, d_results 4-byte aligned");
    CU(cudaSetDevice(m->device));
    fq::ReadSource src{d_packed, nullptr, nullptr, 0, n};
    return run_device(m, src, d_results, (cudaStream_t)stream);
}

int fqtk_b200_matcher_assign_ascii_device(fqtk_b200_matcher* m, const uint8_t* d_ascii, uint64_t n, uint64_t stride,
                                          const uint32_t* d_lengths, uint32_t* d_results, void* stream) {
    if (!m || (n && (!d_ascii || !d_results))) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    if (stride < m->L && !d_lengths) return fail(FQTK_B200_ERR_ARG, "row_stride smaller than the barcode length");
    CU(cudaSetDevice(m->device));
    fq::ReadSource src{nullptr, d_ascii, d_lengths, stride, n};
    return run_device(m, src, d_results, (cudaStream_t)stream);
}

int fqtk_b200_pack_device(const uint8_t* d_ascii, uint64_t n, uint32_t L, uint64_t stride, uint32_t* d_packed,
                          void* stream) {
    if (n && (!d_ascii || !d_packed)) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    if (L == 0 || L > FQTK_B200_MAX_BARCODE_LEN || stride < L) return fail(FQTK_B200_ERR_ARG, "bad length / stride");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, dev));
    fq::LaunchGeometry g{prop.multiProcessorCount, (int)prop.sharedMemPerBlockOptin};
    CU(fq::launch_pack(d_ascii, n, L, stride, d_packed, g, (cudaStream_t)stream));
    return FQTK_B200_OK;
}

int fqtk_b200_encode_host(const uint8_t* bases, size_t len, uint32_t* out_blocks) {
    if ((len && !bases) || !out_blocks) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    const size_t W = (len + 7) / 8;
    for (size_t w = 0; w < W; w++) out_blocks[w] = 0;
    for (size_t i = 0; i < len; i++) out_blocks[i >> 3] |= fq::encode_byte(bases[i]) << (4u * (i & 7u));
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_assign_batch(fqtk_b200_matcher* m, const uint8_t* rows, uint64_t n, uint64_t stride,
                                   const uint32_t* lengths, uint32_t* results) {
    if (!m || (n && (!rows || !results))) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    if (n == 0) return FQTK_B200_OK;
    const uint32_t L = m->L;
    if (!lengths && stride < L) return fail(FQTK_B200_ERR_ARG, "row_stride smaller than the barcode length");
    if (lengths) {
        // barcode_matching.rs:165-172 for rows that would reach count_mismatches with the wrong length (:95-106)
        for (uint64_t i = 0; i < n; i++) {
            const uint32_t len = lengths[i];
            if (len > stride) return fail(FQTK_B200_ERR_ARG, "lengths[i] exceeds row_stride");
            if (len <= L) continue;
            const uint8_t* r = rows + i * stride;
            size_t nocalls = 0;
            for (uint32_t k = 0; k < len; k++) nocalls += fq::byte_is_nocall(r[k]);
            if (nocalls > (size_t)m->max_mm + m->max_ns) continue;  // pre-filter makes it None before the panic
            // the reference's panic text (barcode_matching.rs:99-105): decoded read, lengths, sample 0's barcode; the
            // sample id is not known on this side of the boundary, so the sample INDEX (always 0: the scan panics at
            // the first barcode) stands in for it — the host mirrors put the id back
            std::string msg = "Read barcode (";
            for (uint32_t k = 0; k < len; k++) msg.push_back(decode_mask(fq::encode_byte(r[k])));
            msg += ") length (" + std::to_string(len) + ") differs from expected barcode (";
            msg.append(reinterpret_cast<const char*>(m->panel.data()), L);
            msg += ") length (" + std::to_string(L) + ") for sample 0";
            return fail(FQTK_B200_ERR_LENGTH, msg);
        }
    }
    CU(cudaSetDevice(m->device));
    PipelineDrain drain{m};
    if (m->host_pack_threads > 0 && !lengths && stride == L && (L % 8u) == 0u && m->W <= (uint32_t)fq::MAX_FAST_WORDS &&
        n >= (1u << 20) && fq::have_avx2())
        return assign_batch_host_pack(m, rows, n, results, m->host_pack_threads);
    uint64_t chunk = m->opt.chunk_bytes / std::max<uint64_t>(stride, 1);
    chunk = std::max<uint64_t>(chunk, 1024);
    chunk = std::min<uint64_t>(chunk, n);
    chunk &= ~3ull;
    if (chunk == 0) chunk = n;
    int rc = ensure_pipeline(m, (size_t)(chunk * stride + 16), (size_t)chunk, lengths != nullptr);
    if (rc != FQTK_B200_OK) return rc;
    uint64_t done = 0;
    int slot = 0;
    while (done < n) {
        const uint64_t c = std::min(chunk, n - done);
        cudaStream_t st = m->streams[slot];
        // the caller's last row need not be padded out to `stride`: copy only up to its last valid byte
        const size_t copy_bytes = (size_t)((c - 1) * stride) + (lengths ? (size_t)lengths[done + c - 1] : (size_t)L);
        if (copy_bytes) CU(cudaMemcpyAsync(m->d_in[slot], rows + done * stride, copy_bytes, cudaMemcpyHostToDevice, st));
        if (lengths) CU(cudaMemcpyAsync(m->d_len[slot], lengths + done, c * 4, cudaMemcpyHostToDevice, st));
        fq::ReadSource src{nullptr, m->d_in[slot], lengths ? m->d_len[slot] : nullptr, stride, c};
        rc = run_device(m, src, m->d_out[slot], st, slot);
        if (rc != FQTK_B200_OK) return rc;
        CU(cudaMemcpyAsync(results + done, m->d_out[slot], c * 4, cudaMemcpyDeviceToHost, st));
        done += c;
        slot = (slot + 1) % N_PIPE;
    }
    for (int s = 0; s < N_PIPE; s++) CU(cudaStreamSynchronize(m->streams[s]));
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_assign_batch_packed(fqtk_b200_matcher* m, const uint32_t* packed, uint64_t n, uint32_t* results,
                                          uint16_t* sample_index) {
    if (!m || (n && !packed) || (n && !results && !sample_index)) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    if (n == 0) return FQTK_B200_OK;
    CU(cudaSetDevice(m->device));
    PipelineDrain drain{m};
    const uint64_t row_bytes = (uint64_t)m->W * 4u;
    uint64_t chunk = std::max<uint64_t>(m->opt.chunk_bytes / row_bytes, 1024);
    chunk = std::min<uint64_t>(chunk, n) & ~3ull;  // chunk starts stay 16-byte aligned on the device (vector loads)
    if (chunk == 0) chunk = n;
    int rc = ensure_pipeline(m, (size_t)(chunk * row_bytes + 16), (size_t)chunk, sample_index != nullptr);
    if (rc != FQTK_B200_OK) return rc;
    uint64_t done = 0;
    int slot = 0;
    while (done < n) {
        const uint64_t c = std::min(chunk, n - done);
        cudaStream_t st = m->streams[slot];
        uint32_t* d_words = reinterpret_cast<uint32_t*>(m->d_in[slot]);
        CU(cudaMemcpyAsync(d_words, packed + done * m->W, (size_t)(c * row_bytes), cudaMemcpyHostToDevice, st));
        fq::ReadSource src{d_words, nullptr, nullptr, 0, c};
        rc = run_device(m, src, m->d_out[slot], st, slot);
        if (rc != FQTK_B200_OK) return rc;
        if (results) CU(cudaMemcpyAsync(results + done, m->d_out[slot], c * 4, cudaMemcpyDeviceToHost, st));
        if (sample_index) {
            uint16_t* d16 = reinterpret_cast<uint16_t*>(m->d_len[slot]);
            CU(fq::launch_narrow_u16(m->d_out[slot], c, d16, m->geo, st));
            CU(cudaMemcpyAsync(sample_index + done, d16, c * 2, cudaMemcpyDeviceToHost, st));
        }
        done += c;
        slot = (slot + 1) % N_PIPE;
    }
    for (int s = 0; s < N_PIPE; s++) CU(cudaStreamSynchronize(m->streams[s]));
    return FQTK_B200_OK;
}

static int check_segments(const fqtk_b200_matcher* m, const fqtk_b200_segment* segs, uint32_t n_segs,
                          fq::SegmentSource& out) {
    if (!m || !segs || n_segs == 0 || n_segs > FQTK_B200_MAX_SEGMENTS)
        return fail(FQTK_B200_ERR_ARG, "need 1..8 segments");
    uint64_t total = 0;
    out.n_segments = n_segs;
    for (uint32_t s = 0; s < n_segs; s++) {
        if (!segs[s].base || segs[s].length == 0 || (uint64_t)segs[s].offset + segs[s].length > segs[s].row_stride)
            return fail(FQTK_B200_ERR_ARG, "bad segment (NULL base, zero length, or offset + length > row_stride)");
        out.base[s] = segs[s].base;
        out.stride[s] = segs[s].row_stride;
        out.offset[s] = segs[s].offset;
        out.length[s] = segs[s].length;
        total += segs[s].length;
    }
    if (total != m->L) {  // shorter would make every read None (:167-169), longer is the reference's panic (:95-106)
        char buf[160];
        std::snprintf(buf, sizeof buf, "Read barcode length (%llu) differs from expected barcode length (%u)",
                      (unsigned long long)total, m->L);
        return fail(FQTK_B200_ERR_LENGTH, buf);
    }
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_assign_segments_device(fqtk_b200_matcher* m, const fqtk_b200_segment* segs, uint32_t n_segs,
                                             uint64_t n, uint32_t* d_results, void* stream) {
    fq::SegmentSource ss{};
    int rc = check_segments(m, segs, n_segs, ss);
    if (rc != FQTK_B200_OK) return rc;
    if (n && !d_results) return fail(FQTK_B200_ERR_ARG, "NULL results");
    if (n >= (1ull << 32)) return fail(FQTK_B200_ERR_ARG, "n_reads must be < 2^32 per device call");
    CU(cudaSetDevice(m->device));
    rc = ensure_scratch(&m->d_scratch, &m->scratch_words, (size_t)n * m->W + 4);
    if (rc != FQTK_B200_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CU(fq::launch_pack_segments(ss, n, m->L, m->d_scratch, m->geo, st));
    fq::ReadSource src{m->d_scratch, nullptr, nullptr, 0, n};
    return run_device(m, src, d_results, st);
}

int fqtk_b200_matcher_assign_segments(fqtk_b200_matcher* m, const fqtk_b200_segment* segs, uint32_t n_segs, uint64_t n,
                                      uint32_t* results) {
    fq::SegmentSource ss{};
    int rc = check_segments(m, segs, n_segs, ss);
    if (rc != FQTK_B200_OK) return rc;
    if (n == 0) return FQTK_B200_OK;
    if (!results) return fail(FQTK_B200_ERR_ARG, "NULL results");
    CU(cudaSetDevice(m->device));
    PipelineDrain drain{m};
    // sources may overlap or alias (two segments of the same FASTQ row): ship each distinct (base, stride) once
    uint32_t n_src = 0, src_of[FQTK_B200_MAX_SEGMENTS];
    const uint8_t* sbase[FQTK_B200_MAX_SEGMENTS];
    uint64_t sstride[FQTK_B200_MAX_SEGMENTS], sspan[FQTK_B200_MAX_SEGMENTS];  // span = bytes of a row actually needed
    for (uint32_t s = 0; s < n_segs; s++) {
        uint32_t k = 0;
        while (k < n_src && !(sbase[k] == ss.base[s] && sstride[k] == ss.stride[s])) k++;
        if (k == n_src) {
            sbase[k] = ss.base[s];
            sstride[k] = ss.stride[s];
            sspan[k] = 0;
            n_src++;
        }
        src_of[s] = k;
        sspan[k] = std::max<uint64_t>(sspan[k], (uint64_t)ss.offset[s] + ss.length[s]);
    }
    uint64_t bytes_per_read = 0;
    for (uint32_t k = 0; k < n_src; k++) bytes_per_read += sstride[k];
    uint64_t chunk = std::max<uint64_t>(m->opt.chunk_bytes / std::max<uint64_t>(bytes_per_read, 1), 1024);
    chunk = std::min<uint64_t>(chunk, n) & ~3ull;
    if (chunk == 0) chunk = n;
    rc = ensure_pipeline(m, (size_t)(chunk * bytes_per_read + 64 * n_src), (size_t)chunk, false);
    if (rc != FQTK_B200_OK) return rc;
    for (int s = 0; s < N_PIPE; s++) {
        rc = ensure_scratch(&m->d_seg_packed[s], &m->seg_packed_words[s], (size_t)chunk * m->W + 4);
        if (rc != FQTK_B200_OK) return rc;
    }
    uint64_t done = 0;
    int slot = 0;
    while (done < n) {
        const uint64_t c = std::min(chunk, n - done);
        cudaStream_t st = m->streams[slot];
        fq::SegmentSource dev = ss;
        uint8_t* cursor = m->d_in[slot];
        uint8_t* dbase[FQTK_B200_MAX_SEGMENTS];
        for (uint32_t k = 0; k < n_src; k++) {
            const size_t bytes = (size_t)((c - 1) * sstride[k] + sspan[k]);  // the last row need not be padded out
            CU(cudaMemcpyAsync(cursor, sbase[k] + done * sstride[k], bytes, cudaMemcpyHostToDevice, st));
            dbase[k] = cursor;
            cursor += (c * sstride[k] + 63) & ~(size_t)63;
        }
        for (uint32_t s = 0; s < n_segs; s++) dev.base[s] = dbase[src_of[s]];
        CU(fq::launch_pack_segments(dev, c, m->L, m->d_seg_packed[slot], m->geo, st));
        fq::ReadSource src{m->d_seg_packed[slot], nullptr, nullptr, 0, c};
        rc = run_device(m, src, m->d_out[slot], st);
        if (rc != FQTK_B200_OK) return rc;
        CU(cudaMemcpyAsync(results + done, m->d_out[slot], c * 4, cudaMemcpyDeviceToHost, st));
        done += c;
        slot = (slot + 1) % N_PIPE;
    }
    for (int s = 0; s < N_PIPE; s++) CU(cudaStreamSynchronize(m->streams[s]));
    return FQTK_B200_OK;
}

static int check_fastq_args(const fqtk_b200_matcher* m, const fqtk_b200_fastq_source* sources, uint32_t n_sources,
                            const fqtk_b200_fastq_segment* segs, uint32_t n_segs, bool& any_rest, uint64_t& fixed_total) {
    if (!m || !sources || !segs || n_sources == 0 || n_sources > FQTK_B200_MAX_SEGMENTS || n_segs == 0 ||
        n_segs > FQTK_B200_MAX_SEGMENTS)
        return fail(FQTK_B200_ERR_ARG, "need 1..8 sources and 1..8 segments");
    any_rest = false;
    fixed_total = 0;
    for (uint32_t k = 0; k < n_segs; k++) {
        if (segs[k].source >= n_sources) return fail(FQTK_B200_ERR_ARG, "segment names a source that was not given");
        if (segs[k].length == 0) return fail(FQTK_B200_ERR_ARG, "zero-length segment");
        if (segs[k].length == FQTK_B200_SEGMENT_REST)
            any_rest = true;
        else
            fixed_total += segs[k].length;
    }
    for (uint32_t s = 0; s < n_sources; s++)
        if (!sources[s].chunk || !sources[s].seq_offsets) return fail(FQTK_B200_ERR_ARG, "NULL chunk / seq_offsets");
    if (!any_rest && fixed_total != m->L) {  // as for fixed-stride segments: every read would be None or panic
        char buf[160];
        std::snprintf(buf, sizeof buf, "Read barcode length (%llu) differs from expected barcode length (%u)",
                      (unsigned long long)fixed_total, m->L);
        return fail(FQTK_B200_ERR_LENGTH, buf);
    }
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_assign_fastq_device(fqtk_b200_matcher* m, const fqtk_b200_fastq_source* sources, uint32_t n_sources,
                                          const fqtk_b200_fastq_segment* segs, uint32_t n_segs, uint64_t n,
                                          uint32_t* d_results, void* stream) {
    bool any_rest;
    uint64_t fixed_total;
    int rc = check_fastq_args(m, sources, n_sources, segs, n_segs, any_rest, fixed_total);
    if (rc != FQTK_B200_OK) return rc;
    if (n && !d_results) return fail(FQTK_B200_ERR_ARG, "NULL results");
    if (n >= (1ull << 32)) return fail(FQTK_B200_ERR_ARG, "n_reads must be < 2^32 per device call");
    if (n == 0) return FQTK_B200_OK;
    CU(cudaSetDevice(m->device));
    fq::OffsetSource os{};
    os.n_segments = n_segs;
    for (uint32_t s = 0; s < n_sources; s++) {
        os.base[s] = sources[s].chunk;
        os.seq_offsets[s] = sources[s].seq_offsets;
        os.seq_lengths[s] = sources[s].seq_lengths;
    }
    for (uint32_t k = 0; k < n_segs; k++) {
        os.source_of[k] = segs[k].source;
        os.offset[k] = segs[k].offset;
        os.length[k] = segs[k].length;
        if (segs[k].length == FQTK_B200_SEGMENT_REST && !sources[segs[k].source].seq_lengths)
            return fail(FQTK_B200_ERR_ARG, "a REST segment needs the source's seq_lengths");
    }
    rc = ensure_scratch(&m->d_scratch, &m->scratch_words, (size_t)n * m->W + 4);
    if (rc != FQTK_B200_OK) return rc;
    if (any_rest && n > m->fq_len_cap) {
        if (m->d_fq_len) cudaFree(m->d_fq_len);
        m->d_fq_len = nullptr;
        m->fq_len_cap = 0;
        CU(cudaMalloc(&m->d_fq_len, (size_t)n * 4));
        m->fq_len_cap = (size_t)n;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CU(fq::launch_pack_offsets(os, n, m->L, m->d_scratch, any_rest ? m->d_fq_len : nullptr, m->geo, st));
    fq::ReadSource src{m->d_scratch, nullptr, nullptr, 0, n};
    rc = run_device(m, src, d_results, st);
    if (rc != FQTK_B200_OK) return rc;
    if (any_rest) CU(fq::launch_fix_lengths(d_results, m->d_fq_len, n, m->L, m->S, m->d_counts, m->geo, st));
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_assign_fastq(fqtk_b200_matcher* m, const fqtk_b200_fastq_source* sources, uint32_t n_sources,
                                   const fqtk_b200_fastq_segment* segs, uint32_t n_segs, uint64_t n, uint32_t* results) {
    bool any_rest;
    uint64_t fixed_total;
    int rc = check_fastq_args(m, sources, n_sources, segs, n_segs, any_rest, fixed_total);
    if (rc != FQTK_B200_OK) return rc;
    if (n == 0) return FQTK_B200_OK;
    if (!results) return fail(FQTK_B200_ERR_ARG, "NULL results");
    for (uint32_t s = 0; s < n_sources; s++)
        if (!sources[s].seq_lengths) return fail(FQTK_B200_ERR_ARG, "the host call needs seq_lengths for every source");
    // ReadSetIterator::next (demux.rs:298-315): every read must hold its segments; then BarcodeMatcher::assign's length rules
    // for barcodes that come out longer than L (only possible with a REST segment)
    for (uint64_t i = 0; i < n; i++) {
        uint64_t total = 0;
        for (uint32_t k = 0; k < n_segs; k++) {
            const fqtk_b200_fastq_source& so = sources[segs[k].source];
            const uint64_t len = so.seq_lengths[i], need = (uint64_t)segs[k].offset +
                                                           (segs[k].length == FQTK_B200_SEGMENT_REST ? 1u : segs[k].length);
            if (so.seq_offsets[i] + len > so.chunk_bytes) return fail(FQTK_B200_ERR_ARG, "a sequence line runs past its chunk");
            if (len < need) {
                char buf[200];
                std::snprintf(buf, sizeof buf, "Read %llu had too few bases to demux %llu vs. %llu needed in read structure.",
                              (unsigned long long)i, (unsigned long long)len, (unsigned long long)need);
                return fail(FQTK_B200_ERR_ARG, buf);
            }
            total += segs[k].length == FQTK_B200_SEGMENT_REST ? len - segs[k].offset : segs[k].length;
        }
        if (total > m->L) {
            std::string bc;
            for (uint32_t k = 0; k < n_segs; k++) {
                const fqtk_b200_fastq_source& so = sources[segs[k].source];
                const uint64_t len = segs[k].length == FQTK_B200_SEGMENT_REST ? so.seq_lengths[i] - segs[k].offset : segs[k].length;
                bc.append(reinterpret_cast<const char*>(so.chunk + so.seq_offsets[i] + segs[k].offset), (size_t)len);
            }
            size_t nocalls = 0;
            for (char ch : bc) nocalls += fq::byte_is_nocall((uint8_t)ch);
            if (nocalls > (size_t)m->max_mm + m->max_ns) continue;  // the pre-filter makes it None before the panic
            std::string msg = "Read barcode (";
            for (char ch : bc) msg.push_back(decode_mask(fq::encode_byte((uint8_t)ch)));
            msg += ") length (" + std::to_string(bc.size()) + ") differs from expected barcode (";
            msg.append(reinterpret_cast<const char*>(m->panel.data()), m->L);
            msg += ") length (" + std::to_string(m->L) + ") for sample 0";
            return fail(FQTK_B200_ERR_LENGTH, msg);
        }
    }
    CU(cudaSetDevice(m->device));
    PipelineDrain drain{m};
    cudaStream_t st = m->streams[0];
    auto ensure = [&](int slot, size_t bytes) -> int {
        if (bytes > m->fq_cap[slot]) {
            if (m->d_fq[slot]) cudaFree(m->d_fq[slot]);
            m->d_fq[slot] = nullptr;
            m->fq_cap[slot] = 0;
            CU(cudaMalloc(&m->d_fq[slot], bytes + 64));
            m->fq_cap[slot] = bytes;
        }
        return FQTK_B200_OK;
    };
    fqtk_b200_fastq_source dev[FQTK_B200_MAX_SEGMENTS];
    for (uint32_t s = 0; s < n_sources; s++) {
        if ((rc = ensure(3 * s, sources[s].chunk_bytes)) != FQTK_B200_OK) return rc;
        if ((rc = ensure(3 * s + 1, n * 8)) != FQTK_B200_OK) return rc;
        if ((rc = ensure(3 * s + 2, n * 4)) != FQTK_B200_OK) return rc;
        CU(cudaMemcpyAsync(m->d_fq[3 * s], sources[s].chunk, sources[s].chunk_bytes, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m->d_fq[3 * s + 1], sources[s].seq_offsets, n * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m->d_fq[3 * s + 2], sources[s].seq_lengths, n * 4, cudaMemcpyHostToDevice, st));
        dev[s].chunk = static_cast<const uint8_t*>(m->d_fq[3 * s]);
        dev[s].chunk_bytes = sources[s].chunk_bytes;
        dev[s].seq_offsets = static_cast<const uint64_t*>(m->d_fq[3 * s + 1]);
        dev[s].seq_lengths = static_cast<const uint32_t*>(m->d_fq[3 * s + 2]);
    }
    rc = ensure_pipeline(m, 16, (size_t)n, false);
    if (rc != FQTK_B200_OK) return rc;
    rc = fqtk_b200_matcher_assign_fastq_device(m, dev, n_sources, segs, n_segs, n, m->d_out[0], st);
    if (rc != FQTK_B200_OK) return rc;
    CU(cudaMemcpyAsync(results, m->d_out[0], n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FQTK_B200_OK;
}

// ---- the record scanner on the device + the batch call on whole chunks ------------------------------------------------
namespace {
struct DevScan {  // temporaries of one chunk's scan (stream-ordered: fq::temp_alloc)
    uint32_t* tile_counts = nullptr;
    unsigned long long* prefix = nullptr;  // tiles + 1
    unsigned long long* nl = nullptr;
    unsigned long long* err = nullptr;     // 1 scan error word + 2 vetting words
    cudaStream_t stream = nullptr;
    ~DevScan() {
        if (tile_counts) cudaFreeAsync(tile_counts, stream);
        if (prefix) cudaFreeAsync(prefix, stream);
        if (nl) cudaFreeAsync(nl, stream);
        if (err) cudaFreeAsync(err, stream);
    }
};
const char* const SCAN_ERR[3] = {"header line does not start with '@'", "separator line does not start with '+'",
                                 "sequence and quality lengths differ"};
}  // namespace

int fqtk_b200_fastq_scan_device(int device, const uint8_t* d_chunk, uint64_t chunk_bytes, uint64_t max_records,
                                uint64_t* d_head_offsets, uint64_t* d_seq_offsets, uint32_t* d_seq_lengths,
                                uint64_t* n_records, uint64_t* consumed, void* stream) {
    if ((chunk_bytes && !d_chunk) || !d_seq_offsets || !d_seq_lengths || !n_records || !consumed)
        return fail(FQTK_B200_ERR_ARG, "NULL argument");
    *n_records = 0;
    *consumed = 0;
    if (chunk_bytes == 0 || max_records == 0) return FQTK_B200_OK;
    CU(cudaSetDevice(device));
    fq::LaunchGeometry geo{};
    CU(cudaDeviceGetAttribute(&geo.sm_count, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t tiles = fq::fastq_scan_tiles(chunk_bytes);
    DevScan t;
    t.stream = st;
    CU(fq::temp_alloc(reinterpret_cast<void**>(&t.tile_counts), (size_t)tiles * 4, st));
    CU(fq::temp_alloc(reinterpret_cast<void**>(&t.prefix), ((size_t)tiles + 1) * 8, st));
    CU(fq::temp_alloc(reinterpret_cast<void**>(&t.err), 8, st));
    CU(cudaMemsetAsync(t.err, 0xFF, 8, st));
    CU(fq::launch_nl_count(d_chunk, chunk_bytes, t.tile_counts, t.prefix, st));
    unsigned long long total_nl = 0;
    CU(cudaMemcpyAsync(&total_nl, t.prefix + tiles, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const uint64_t n = std::min<uint64_t>(total_nl / 4, max_records);
    if (n == 0) return FQTK_B200_OK;
    CU(fq::temp_alloc(reinterpret_cast<void**>(&t.nl), (size_t)n * 4 * 8, st));
    CU(fq::launch_fq_records(d_chunk, chunk_bytes, t.prefix, n * 4, t.nl, n, reinterpret_cast<unsigned long long*>(d_head_offsets),
                             reinterpret_cast<unsigned long long*>(d_seq_offsets), d_seq_lengths, t.err, geo, st));
    unsigned long long err = 0, last_nl = 0;
    CU(cudaMemcpyAsync(&err, t.err, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&last_nl, t.nl + (n * 4 - 1), 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (err != ~0ull)
        return fail(FQTK_B200_ERR_ARG, "FASTQ record " + std::to_string(err >> 2) + ": " + SCAN_ERR[err & 3u]);
    *n_records = n;
    *consumed = last_nl + 1;
    return FQTK_B200_OK;
}

// Shared by fqtk_b200_matcher_assign_fastq_chunks and fqtk_b200_demux_chunks: chunks to the device, records found and vetted
// there, B segments matched; the result words stay in m->d_out[0] (`results` != NULL: also copied to the host), the record
// tables in m->d_fq[] (`dev_out`), header offsets in m->d_fq_head[] when `min_len` is given (the whole-batch path: every
// read must hold ALL segments of its read structure, demux.rs:298-315, not only the B ones).
static int ingest_chunks_impl(fqtk_b200_matcher* m, const fqtk_b200_fastq_chunk* chunks, uint32_t n_sources,
                              const fqtk_b200_fastq_segment* segs, uint32_t n_segs, uint64_t max_reads, uint32_t* results,
                              uint64_t* n_reads, uint64_t* consumed, const uint32_t* min_len, fqtk_b200_fastq_source* dev_out,
                              const std::string* rs_text = nullptr) {  // rs_text[s]: read structure of input s as the reference prints it
    const bool want_heads = min_len != nullptr;
    if (!m || !chunks || !segs || !n_reads || !consumed || n_sources == 0 || n_sources > FQTK_B200_MAX_SEGMENTS || n_segs == 0 ||
        n_segs > FQTK_B200_MAX_SEGMENTS)
        return fail(FQTK_B200_ERR_ARG, "need 1..8 sources and 1..8 segments");
    *n_reads = 0;
    for (uint32_t s = 0; s < n_sources; s++) {
        consumed[s] = 0;
        if (chunks[s].bytes && !chunks[s].data) return fail(FQTK_B200_ERR_ARG, "NULL chunk");
    }
    // the same argument rules as the table form, on stand-in sources (pointers are only tested for NULL there)
    fqtk_b200_fastq_source dev[FQTK_B200_MAX_SEGMENTS];
    static const uint64_t dummy = 0;
    for (uint32_t s = 0; s < n_sources; s++) dev[s] = fqtk_b200_fastq_source{reinterpret_cast<const uint8_t*>(&dummy), chunks[s].bytes, &dummy, nullptr};
    bool any_rest;
    uint64_t fixed_total;
    int rc = check_fastq_args(m, dev, n_sources, segs, n_segs, any_rest, fixed_total);
    if (rc != FQTK_B200_OK) return rc;
    if (max_reads >= (1ull << 32)) max_reads = (1ull << 32) - 1;
    CU(cudaSetDevice(m->device));
    PipelineDrain drain{m};
    cudaStream_t st = m->streams[0];
    auto ensure = [&](int slot, size_t bytes) -> int {
        if (bytes > m->fq_cap[slot]) {
            if (m->d_fq[slot]) cudaFree(m->d_fq[slot]);
            m->d_fq[slot] = nullptr;
            m->fq_cap[slot] = 0;
            CU(cudaMalloc(&m->d_fq[slot], bytes + 64));
            m->fq_cap[slot] = bytes;
        }
        return FQTK_B200_OK;
    };
    // chunks over PCIe, newline counts per chunk
    DevScan t[FQTK_B200_MAX_SEGMENTS];
    unsigned long long total_nl[FQTK_B200_MAX_SEGMENTS] = {};
    for (uint32_t s = 0; s < n_sources; s++) {
        if ((rc = ensure(3 * s, chunks[s].bytes)) != FQTK_B200_OK) return rc;
        const uint32_t tiles = fq::fastq_scan_tiles(chunks[s].bytes);
        t[s].stream = st;
        CU(fq::temp_alloc(reinterpret_cast<void**>(&t[s].tile_counts), (size_t)tiles * 4 + 4, st));
        CU(fq::temp_alloc(reinterpret_cast<void**>(&t[s].prefix), ((size_t)tiles + 1) * 8, st));
        CU(fq::temp_alloc(reinterpret_cast<void**>(&t[s].err), 40, st));  // scan | (unused) | min length | vet: few, long
        CU(cudaMemsetAsync(t[s].err, 0xFF, 40, st));
        CU(cudaMemcpyAsync(m->d_fq[3 * s], chunks[s].data, chunks[s].bytes, cudaMemcpyHostToDevice, st));
        CU(fq::launch_nl_count(static_cast<const uint8_t*>(m->d_fq[3 * s]), chunks[s].bytes, t[s].tile_counts, t[s].prefix, st));
        CU(cudaMemcpyAsync(&total_nl[s], t[s].prefix + tiles, 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    uint64_t n = max_reads;
    for (uint32_t s = 0; s < n_sources; s++) n = std::min<uint64_t>(n, total_nl[s] / 4);  // the inputs advance in lock step
    if (n == 0) return FQTK_B200_OK;
    if (!results && !want_heads) return fail(FQTK_B200_ERR_ARG, "NULL results");
    // record tables, then the reference's per-read rules, all on the device
    unsigned long long last_nl[FQTK_B200_MAX_SEGMENTS] = {}, err_scan[FQTK_B200_MAX_SEGMENTS] = {}, err_vet[2] = {};
    fq::OffsetSource os{};
    os.n_segments = n_segs;
    for (uint32_t s = 0; s < n_sources; s++) {
        if ((rc = ensure(3 * s + 1, n * 8)) != FQTK_B200_OK) return rc;
        if ((rc = ensure(3 * s + 2, n * 4)) != FQTK_B200_OK) return rc;
        CU(fq::temp_alloc(reinterpret_cast<void**>(&t[s].nl), (size_t)n * 4 * 8, st));
        const uint8_t* d_chunk = static_cast<const uint8_t*>(m->d_fq[3 * s]);
        if (want_heads && n * 8 > m->fq_head_cap[s]) {
            if (m->d_fq_head[s]) cudaFree(m->d_fq_head[s]);
            m->d_fq_head[s] = nullptr;
            m->fq_head_cap[s] = 0;
            CU(cudaMalloc(&m->d_fq_head[s], n * 8 + 64));
            m->fq_head_cap[s] = n * 8;
        }
        CU(fq::launch_fq_records(d_chunk, chunks[s].bytes, t[s].prefix, n * 4, t[s].nl, n,
                                 want_heads ? static_cast<unsigned long long*>(m->d_fq_head[s]) : nullptr,
                                 static_cast<unsigned long long*>(m->d_fq[3 * s + 1]), static_cast<uint32_t*>(m->d_fq[3 * s + 2]),
                                 t[s].err, m->geo, st));
        if (want_heads) CU(fq::launch_min_len(static_cast<const uint32_t*>(m->d_fq[3 * s + 2]), n, min_len[s], t[s].err + 2, m->geo, st));
        CU(cudaMemcpyAsync(&err_scan[s], t[s].err, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&last_nl[s], t[s].nl + (n * 4 - 1), 8, cudaMemcpyDeviceToHost, st));
        dev[s].chunk = d_chunk;
        dev[s].chunk_bytes = chunks[s].bytes;
        dev[s].seq_offsets = static_cast<const uint64_t*>(m->d_fq[3 * s + 1]);
        dev[s].seq_lengths = static_cast<const uint32_t*>(m->d_fq[3 * s + 2]);
        os.base[s] = d_chunk;
        os.seq_offsets[s] = dev[s].seq_offsets;
        os.seq_lengths[s] = dev[s].seq_lengths;
    }
    for (uint32_t k = 0; k < n_segs; k++) {
        os.source_of[k] = segs[k].source;
        os.offset[k] = segs[k].offset;
        os.length[k] = segs[k].length;
    }
    CU(fq::launch_fq_vet(os, n, m->L, (uint32_t)m->max_mm + m->max_ns, t[0].err + 3, m->geo, st));
    CU(cudaMemcpyAsync(err_vet, t[0].err + 3, 16, cudaMemcpyDeviceToHost, st));
    unsigned long long err_min[FQTK_B200_MAX_SEGMENTS];
    for (uint32_t s = 0; s < n_sources; s++) {
        err_min[s] = ~0ull;
        if (want_heads) CU(cudaMemcpyAsync(&err_min[s], t[s].err + 2, 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    for (uint32_t s = 0; s < n_sources; s++)
        if (err_scan[s] != ~0ull)
            return fail(FQTK_B200_ERR_ARG, "FASTQ record " + std::to_string(err_scan[s] >> 2) + " of input " + std::to_string(s) + ": " +
                                               SCAN_ERR[err_scan[s] & 3u]);
    if (want_heads) {  // ReadSetIterator::next: "Read {name} had too few bases to demux {len} vs. {min} needed in read structure {rs}."
        unsigned long long first = ~0ull;
        for (uint32_t s = 0; s < n_sources; s++) first = std::min(first, err_min[s]);
        if (first != ~0ull && first <= std::min(err_vet[0], err_vet[1])) {
            for (uint32_t s = 0; s < n_sources; s++) {
                uint32_t len = 0;
                CU(cudaMemcpy(&len, static_cast<const uint32_t*>(m->d_fq[3 * s + 2]) + first, 4, cudaMemcpyDeviceToHost));
                if (len >= min_len[s]) continue;
                unsigned long long h0 = 0, s0 = 0;
                CU(cudaMemcpy(&h0, static_cast<const unsigned long long*>(m->d_fq_head[0]) + first, 8, cudaMemcpyDeviceToHost));
                CU(cudaMemcpy(&s0, static_cast<const unsigned long long*>(m->d_fq[1]) + first, 8, cudaMemcpyDeviceToHost));
                std::string name(reinterpret_cast<const char*>(chunks[0].data + h0 + 1), (size_t)(s0 - h0 - 2));
                if (!name.empty() && name.back() == '\r') name.pop_back();
                return fail(FQTK_B200_ERR_ARG, "Read " + name + " had too few bases to demux " + std::to_string(len) + " vs. " +
                                                   std::to_string(min_len[s]) + " needed in read structure" +
                                                   (rs_text ? " " + rs_text[s] : std::string()) + ".");
            }
        }
    }
    if (err_vet[0] != ~0ull || err_vet[1] != ~0ull) {
        // the first offending read, as the reference's read-by-read loop would meet it: its table rows come back and the
        // message is built from the host's own chunk bytes
        const bool few = err_vet[0] <= err_vet[1];
        const uint64_t i = few ? err_vet[0] : err_vet[1];
        uint64_t off[FQTK_B200_MAX_SEGMENTS] = {};
        uint32_t len[FQTK_B200_MAX_SEGMENTS] = {};
        for (uint32_t s = 0; s < n_sources; s++) {
            CU(cudaMemcpy(&off[s], dev[s].seq_offsets + i, 8, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(&len[s], dev[s].seq_lengths + i, 4, cudaMemcpyDeviceToHost));
        }
        if (few) {
            for (uint32_t k = 0; k < n_segs; k++) {
                const uint64_t need = (uint64_t)segs[k].offset + (segs[k].length == FQTK_B200_SEGMENT_REST ? 1u : segs[k].length);
                if (len[segs[k].source] < need) {
                    char buf[200];
                    std::snprintf(buf, sizeof buf, "Read %llu had too few bases to demux %llu vs. %llu needed in read structure.",
                                  (unsigned long long)i, (unsigned long long)len[segs[k].source], (unsigned long long)need);
                    return fail(FQTK_B200_ERR_ARG, buf);
                }
            }
        }
        std::string bc;
        for (uint32_t k = 0; k < n_segs; k++) {
            const uint32_t s = segs[k].source;
            const uint64_t l = segs[k].length == FQTK_B200_SEGMENT_REST ? len[s] - segs[k].offset : segs[k].length;
            bc.append(reinterpret_cast<const char*>(chunks[s].data + off[s] + segs[k].offset), (size_t)l);
        }
        std::string msg = "Read barcode (";
        for (char ch : bc) msg.push_back(decode_mask(fq::encode_byte((uint8_t)ch)));
        msg += ") length (" + std::to_string(bc.size()) + ") differs from expected barcode (";
        msg.append(reinterpret_cast<const char*>(m->panel.data()), m->L);
        msg += ") length (" + std::to_string(m->L) + ") for sample 0";
        return fail(FQTK_B200_ERR_LENGTH, msg);
    }
    rc = ensure_pipeline(m, 16, (size_t)n, false);
    if (rc != FQTK_B200_OK) return rc;
    rc = fqtk_b200_matcher_assign_fastq_device(m, dev, n_sources, segs, n_segs, n, m->d_out[0], st);
    if (rc != FQTK_B200_OK) return rc;
    if (results) CU(cudaMemcpyAsync(results, m->d_out[0], n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_reads = n;
    for (uint32_t s = 0; s < n_sources; s++) consumed[s] = last_nl[s] + 1;
    if (dev_out)
        for (uint32_t s = 0; s < n_sources; s++) dev_out[s] = dev[s];
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_assign_fastq_chunks(fqtk_b200_matcher* m, const fqtk_b200_fastq_chunk* chunks, uint32_t n_sources,
                                          const fqtk_b200_fastq_segment* segs, uint32_t n_segs, uint64_t max_reads,
                                          uint32_t* results, uint64_t* n_reads, uint64_t* consumed) {
    if (n_reads) *n_reads = 0;
    if (!results && max_reads) {
        // (a NULL result buffer is only legal when nothing can come back)
        bool any = false;
        for (uint32_t s = 0; chunks && s < n_sources; s++) any = any || chunks[s].bytes;
        if (any) return fail(FQTK_B200_ERR_ARG, "NULL results");
    }
    return ingest_chunks_impl(m, chunks, n_sources, segs, n_segs, max_reads, results, n_reads, consumed, nullptr, nullptr);
}

// ---- one call per batch: FASTQ chunks in, per-sample BGZF members out -------------------------------------------------
int fqtk_b200_demux_chunks(fqtk_b200_matcher* m, fqtk_b200_bgzf* z, const fqtk_b200_fastq_chunk* chunks, uint32_t n_sources,
                           const fqtk_b200_read_segment* segments, uint32_t n_segments, const char* output_kinds, int level,
                           uint64_t max_reads, uint8_t* out, uint64_t out_capacity, uint64_t* out_offsets, uint32_t* n_streams_out,
                           uint64_t* batch_counts, uint64_t* n_reads, uint64_t* consumed) {
    if (!m || !z || !chunks || !segments || !output_kinds || !out_offsets || !n_streams_out || !n_reads || !consumed)
        return fail(FQTK_B200_ERR_ARG, "NULL argument");
    if (n_segments == 0 || n_segments > 32) return fail(FQTK_B200_ERR_ARG, "need 1..32 segments");
    *n_reads = 0;
    // the B segments for the matcher, the minimum read length of every input (demux.rs:298: fixed lengths, + 1 for a `+`)
    fqtk_b200_fastq_segment bsegs[FQTK_B200_MAX_SEGMENTS];
    uint32_t nb = 0, min_len[FQTK_B200_MAX_SEGMENTS] = {};
    for (uint32_t k = 0; k < n_segments; k++) {
        if (segments[k].source >= n_sources || segments[k].source >= FQTK_B200_MAX_SEGMENTS)
            return fail(FQTK_B200_ERR_ARG, "segment names a source that was not given");
        min_len[segments[k].source] = std::max(min_len[segments[k].source],
                                               segments[k].offset + (segments[k].length == FQTK_B200_SEGMENT_REST ? 1u : segments[k].length));
        if (segments[k].kind == 'B') {
            if (nb == FQTK_B200_MAX_SEGMENTS) return fail(FQTK_B200_ERR_UNSUPPORTED, "more than 8 sample-barcode segments");
            bsegs[nb++] = fqtk_b200_fastq_segment{segments[k].source, segments[k].offset, segments[k].length};
        }
    }
    if (nb == 0) return fail(FQTK_B200_ERR_ARG, "the read structures hold no sample-barcode segment");
    uint32_t ns = 0;
    int rc = fqtk_b200_emit_streams(segments, n_segments, output_kinds, &ns, nullptr, nullptr);
    if (rc != FQTK_B200_OK) return rc;
    *n_streams_out = ns;
    const uint32_t B = m->S + 1u, n_seg = ns * B;
    for (uint32_t k = 0; k <= n_seg; k++) out_offsets[k] = 0;
    if (batch_counts)
        for (uint32_t b = 0; b < B; b++) batch_counts[b] = 0;
    const bool trace = getenv("FQTK_B200_TRACE") != nullptr;  // stage times on stderr (each stage ends in a synchronisation)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_prev = now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto t = now();
        std::fprintf(stderr, "[fqtk_b200] demux_chunks %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    fqtk_b200_fastq_source dev[FQTK_B200_MAX_SEGMENTS];
    uint64_t n = 0;
    // the read structures as the reference prints them (`{length}{kind}`, `+{kind}` for the rest), for its panic text
    std::string rs_text[FQTK_B200_MAX_SEGMENTS];
    for (uint32_t k = 0; k < n_segments; k++)
        rs_text[segments[k].source] += (segments[k].length == FQTK_B200_SEGMENT_REST ? std::string("+") : std::to_string(segments[k].length)) +
                                       std::string(1, (char)segments[k].kind);
    rc = ingest_chunks_impl(m, chunks, n_sources, bsegs, nb, max_reads, nullptr, &n, consumed, min_len, dev, rs_text);
    if (rc != FQTK_B200_OK || n == 0) return rc;
    lap("H2D + scan + vet + match");
    CU(cudaSetDevice(m->device));
    cudaStream_t st = m->streams[0];
    // route
    fq::TempBuf d_order, d_off, d_text, d_out;
    CU(d_order.alloc((size_t)n * 4, st));
    CU(d_off.alloc(((size_t)B + 1) * 8, st));
    rc = fqtk_b200_matcher_route_device(m, m->d_out[0], n, d_order.as<uint32_t>(), d_off.as<uint64_t>(), st);
    if (rc != FQTK_B200_OK) return rc;
    if (batch_counts) {
        std::vector<uint64_t> off(B + 1);
        CU(cudaMemcpyAsync(off.data(), d_off.p, ((size_t)B + 1) * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (uint32_t b = 0; b < B; b++) batch_counts[b] = off[b + 1] - off[b];
    }
    if (trace) CU(cudaStreamSynchronize(st));
    lap("route");
    if (ns == 0) {
        *n_reads = n;
        return FQTK_B200_OK;
    }
    // records
    fqtk_b200_emit_source esrc[FQTK_B200_MAX_SEGMENTS];
    uint64_t in_bytes = 0, bc_bytes = 16;
    for (uint32_t s = 0; s < n_sources; s++) {
        esrc[s] = fqtk_b200_emit_source{dev[s].chunk, dev[s].chunk_bytes, static_cast<const uint64_t*>(m->d_fq_head[s]), dev[s].seq_offsets,
                                        dev[s].seq_lengths};
        in_bytes += consumed[s];
    }
    for (uint32_t k = 0; k < n_segments; k++)
        if ((segments[k].kind == 'B' || segments[k].kind == 'M') && segments[k].length != FQTK_B200_SEGMENT_REST) bc_bytes += segments[k].length + 1;
    // every stream repeats the header (+ read number, UMIs, barcodes) and holds at most the bases + qualities of one input
    const uint64_t text_cap = (uint64_t)ns * (in_bytes + n * (bc_bytes + 16)) + 64;
    CU(d_text.alloc(text_cap, st));
    std::vector<uint64_t> file_off((size_t)ns * (B + 1));
    uint64_t text_bytes = 0;
    rc = fqtk_b200_demux_emit_device(m->device, esrc, n_sources, segments, n_segments, output_kinds, d_order.as<uint32_t>(),
                                     d_off.as<uint64_t>(), B, n, d_text.as<uint8_t>(), text_cap, file_off.data(), &text_bytes, st);
    if (rc != FQTK_B200_OK) return rc;
    lap("records");
    // every (stream, sample) run -> its own BGZF members
    std::vector<uint64_t> seg_off(n_seg + 1);
    for (uint32_t t = 0; t < ns; t++)
        for (uint32_t b = 0; b < B; b++) seg_off[(size_t)t * B + b] = file_off[(size_t)t * (B + 1) + b];
    seg_off[n_seg] = text_bytes;
    const uint64_t comp_cap = fqtk_b200_bgzf_bound(text_bytes) + 31ull * n_seg;
    CU(d_out.alloc(comp_cap, st));
    rc = fqtk_b200_bgzf_compress_segments_device(z, d_text.as<uint8_t>(), seg_off.data(), n_seg, level, d_out.as<uint8_t>(), comp_cap,
                                                 out_offsets, st);
    if (rc != FQTK_B200_OK) return rc;
    lap("BGZF");
    if (out_offsets[n_seg] > out_capacity)
        return fail(FQTK_B200_ERR_ARG, "output buffer too small: " + std::to_string(out_offsets[n_seg]) + " bytes needed");
    if (!out && out_offsets[n_seg]) return fail(FQTK_B200_ERR_ARG, "NULL output buffer");
    CU(cudaMemcpyAsync(out, d_out.p, out_offsets[n_seg], cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    lap("D2H");
    *n_reads = n;
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_assign(fqtk_b200_matcher* m, const uint8_t* read_bases, size_t len, uint32_t* result) {
    if (!m || !result || (len && !read_bases)) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    if (len > 0xFFFFFFFFull) return fail(FQTK_B200_ERR_ARG, "read too long");
    const uint32_t len32 = (uint32_t)len;
    if (len == 0) {  // nothing to ship to the device: :167-169 makes it None (L >= 1); count it as the caller does
        *result = FQTK_B200_NONE;
        CU(cudaSetDevice(m->device));
        CU(cudaDeviceSynchronize());
        unsigned long long c = 0;
        CU(cudaMemcpy(&c, m->d_counts + m->S, 8, cudaMemcpyDeviceToHost));
        c += 1;
        CU(cudaMemcpy(m->d_counts + m->S, &c, 8, cudaMemcpyHostToDevice));
        return FQTK_B200_OK;
    }
    return fqtk_b200_matcher_assign_batch(m, read_bases, 1, len, &len32, result);
}

int fqtk_b200_matcher_route_device(fqtk_b200_matcher* m, const uint32_t* d_results, uint64_t n, uint32_t* d_order,
                                   uint64_t* d_offsets, void* stream) {
    if (!m || !d_offsets || (n && (!d_results || !d_order))) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    if (n >= (1ull << 32)) return fail(FQTK_B200_ERR_ARG, "n_reads must be < 2^32 per device call");
    if (!fq::route_supported(m->S, m->geo)) return fail(FQTK_B200_ERR_UNSUPPORTED, "too many samples for routing");
    CU(cudaSetDevice(m->device));
    const size_t need = fq::route_workspace_bytes(n, m->S, m->geo);
    if (need > m->route_ws_bytes) {
        if (m->d_route_ws) cudaFree(m->d_route_ws);
        m->d_route_ws = nullptr;
        CU(cudaMalloc(&m->d_route_ws, need));
        m->route_ws_bytes = need;
    }
    CU(fq::launch_route(d_results, n, m->S, d_order, reinterpret_cast<unsigned long long*>(d_offsets), m->d_route_ws,
                        m->geo, (cudaStream_t)stream));
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_route(fqtk_b200_matcher* m, const uint32_t* results, uint64_t n, uint32_t* order,
                            uint64_t* offsets) {
    if (!m || !offsets || (n && (!results || !order))) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(m->device));
    uint32_t *d_res = nullptr, *d_ord = nullptr;
    uint64_t* d_off = nullptr;
    cudaStream_t st = m->streams[0];
    cudaError_t e = cudaMalloc(&d_res, std::max<size_t>(16, n * 4));
    if (e == cudaSuccess) e = cudaMalloc(&d_ord, std::max<size_t>(16, n * 4));
    if (e == cudaSuccess) e = cudaMalloc(&d_off, (size_t)(m->S + 2) * 8);
    int rc = FQTK_B200_OK;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_res, results, n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = fqtk_b200_matcher_route_device(m, d_res, n, d_ord, d_off, st);
    if (e == cudaSuccess && rc == FQTK_B200_OK) e = cudaMemcpyAsync(order, d_ord, n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == FQTK_B200_OK)
        e = cudaMemcpyAsync(offsets, d_off, (size_t)(m->S + 2) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_res);
    cudaFree(d_ord);
    cudaFree(d_off);
    if (e != cudaSuccess) return cuda_fail(e, "route");
    return rc;
}

int fqtk_b200_matcher_counts(fqtk_b200_matcher* m, uint64_t* out) {
    if (!m || !out) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(m->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, m->d_counts, (size_t)(m->S + 1) * 8, cudaMemcpyDeviceToHost));
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_counts_device(fqtk_b200_matcher* m, uint64_t** d_counts) {
    if (!m || !d_counts) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    *d_counts = reinterpret_cast<uint64_t*>(m->d_counts);
    return FQTK_B200_OK;
}

int fqtk_b200_matcher_reset_counts(fqtk_b200_matcher* m) {
    if (!m) return fail(FQTK_B200_ERR_ARG, "NULL matcher");
    CU(cudaSetDevice(m->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(m->d_counts, 0, (size_t)(m->S + 1) * 8));
    return FQTK_B200_OK;
}

int fqtk_b200_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return FQTK_B200_OK;
}

/* asynchronous copies between host memory (pinned for full rate) and device memory on `stream`: for hosts that hold their
 * device buffers elsewhere (the Python mirrors keep them in torch tensors) */
int fqtk_b200_copy_to_device(void* d_dst, const void* src, uint64_t bytes, void* stream) {
    if (bytes && (!d_dst || !src)) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    CU(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return FQTK_B200_OK;
}
int fqtk_b200_copy_to_host(void* dst, const void* d_src, uint64_t bytes, void* stream) {
    if (bytes && (!dst || !d_src)) return fail(FQTK_B200_ERR_ARG, "NULL argument");
    CU(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return FQTK_B200_OK;
}

int fqtk_b200_host_free(void* ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return FQTK_B200_OK;
}

int fqtk_b200_synth_panel(uint64_t seed, uint32_t S, uint32_t L, uint32_t min_distance, uint32_t n_degenerate,
                          uint8_t* out) {
    if (!out || S == 0 || L == 0 || L > FQTK_B200_MAX_BARCODE_LEN) return fail(FQTK_B200_ERR_ARG, "bad argument");
    if (fq::synth_panel_host(seed, S, L, min_distance, n_degenerate, out) != 0)
        return fail(FQTK_B200_ERR_ARG, "no panel with that many samples at that distance");
    return FQTK_B200_OK;
}

int fqtk_b200_synth_reads_host(const uint8_t* panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t first,
                               uint64_t n, uint8_t* out) {
    if (!panel || (n && !out) || S == 0 || L == 0 || L > FQTK_B200_MAX_BARCODE_LEN)
        return fail(FQTK_B200_ERR_ARG, "bad argument");
    fq::synth_reads_host(panel, S, L, seed, first, n, out);
    return FQTK_B200_OK;
}

int fqtk_b200_synth_reads_device(const uint8_t* panel, uint32_t S, uint32_t L, uint64_t seed, uint64_t first,
                                 uint64_t n, uint8_t* d_ascii, uint32_t* d_packed, void* stream) {
    if (!panel || S == 0 || L == 0 || L > FQTK_B200_MAX_BARCODE_LEN) return fail(FQTK_B200_ERR_ARG, "bad argument");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, dev));
    uint8_t* d_panel = nullptr;
    CU(cudaMalloc(&d_panel, (size_t)S * L));
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(d_panel, panel, (size_t)S * L, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = fq::synth_reads_device(d_panel, S, L, seed, first, n, d_ascii, d_packed,
                                                     prop.multiProcessorCount, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_panel);
    if (e != cudaSuccess) return cuda_fail(e, "synth_reads_device");
    return FQTK_B200_OK;
}

}  // extern "C"
