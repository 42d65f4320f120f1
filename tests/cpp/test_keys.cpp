// CPU check of the key primitives shared by host and device code (fqtk_b200/csrc/common.cuh):
// acgt_key / acgt_only must (i) flag as valid exactly the reads whose every nibble is one of 1,2,4,8 and acgt_key must
// (ii) map distinct valid reads to distinct keys — the two properties k_probe3's exactness rests on; k_probe4's rests on
// (i) plus nibble_distance being the reference's mismatch count (bitenc.rs:441-452) and g4_hashes staying in range.
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <unordered_set>

#include "../../fqtk_b200/csrc/common.cuh"

static bool all_one_hot(const uint32_t* w, int L) {
    for (int i = 0; i < L; i++) {
        const uint32_t n = (w[i >> 3] >> (4 * (i & 7))) & 0xFu;
        if (n != 1 && n != 2 && n != 4 && n != 8) return false;
    }
    return true;
}

template <int W>
static int check(int L, uint64_t seed, int trials) {
    std::mt19937_64 rng(seed);
    const uint32_t pad = fq::last_word_pad_for_len((uint32_t)L);
    std::unordered_set<uint64_t> seen_key;
    std::unordered_set<std::string> seen_read;
    int bad = 0;
    for (int t = 0; t < trials; t++) {
        uint32_t w[W] = {};
        const bool want_valid = (rng() & 3u) != 0u;  // 3 of 4 trials: a pure A/C/G/T read
        for (int i = 0; i < L; i++) {
            uint32_t n = 1u << (rng() & 3u);
            if (!want_valid && (rng() % (uint64_t)L) == 0) n = (uint32_t)(rng() & 15u);  // any nibble, incl. 0 and 15
            w[i >> 3] |= n << (4 * (i & 7));
        }
        const bool valid = fq::acgt_only<W>(w, pad);
        if (valid != all_one_hot(w, L)) {
            std::printf("validity mismatch at L=%d\n", L);
            bad++;
        }
        // the two forms of the table-alphabet predicate agree (and say what they should)
        {
            bool in_alphabet = true;
            for (int i = 0; i < L; i++) {
                const uint32_t n = (w[i >> 3] >> (4 * (i & 7))) & 0xFu;
                in_alphabet = in_alphabet && (n == 1 || n == 2 || n == 4 || n == 8 || n == 15);
            }
            if (fq::acgtn_only<W>(w, pad) != in_alphabet || fq::read_in_table_alphabet<W>(w, pad) != in_alphabet) {
                std::printf("alphabet predicate mismatch at L=%d\n", L);
                bad++;
            }
        }
        // fingerprint-table hashes: bucket in range
        {
            const uint32_t n_buckets = 1000003u;
            uint32_t b, fph;
            fq::g4_hashes<W>(w, (uint32_t)rng(), n_buckets, b, fph);
            if (b >= n_buckets) bad++;
        }
        // nibble_distance against the definition: symbols i with obs_i & ~exp_i != 0
        {
            uint32_t ne[W] = {};
            uint32_t want = 0;
            for (int i = 0; i < L; i++) {
                const uint32_t forbid = (uint32_t)(rng() & 15u);
                ne[i >> 3] |= forbid << (4 * (i & 7));
                want += (((w[i >> 3] >> (4 * (i & 7))) & 0xFu) & forbid) != 0u;
            }
            if (fq::nibble_distance<W>(w, ne) != want) {
                std::printf("nibble_distance mismatch at L=%d\n", L);
                bad++;
            }
        }
        if constexpr (W <= 2) {
            bool v2;
            const uint32_t lo = fq::acgt_key<W>(w, pad, v2);
            if (v2 != valid) bad++;
            if (valid) {
                std::string read(reinterpret_cast<const char*>(w), sizeof w);
                const bool new_read = seen_read.insert(read).second, new_key = seen_key.insert(lo).second;
                if (new_read != new_key) {
                    std::printf("key collision at L=%d\n", L);
                    bad++;
                }
            }
        }
    }
    return bad;
}

int main() {
    int bad = 0;
    // exhaustive for short barcodes: every nibble pattern of 3 symbols, then random for the rest
    for (uint32_t x = 0; x < 4096; x++) {
        uint32_t w[1] = {x};
        bool valid;
        (void)fq::acgt_key<1>(w, fq::last_word_pad_for_len(3), valid);
        if (valid != all_one_hot(w, 3)) bad++;
    }
    for (int L : {1, 2, 5, 8}) bad += check<1>(L, 100 + L, 200000);
    for (int L : {9, 12, 16}) bad += check<2>(L, 200 + L, 400000);
    for (int L : {17, 20, 24}) bad += check<3>(L, 300 + L, 400000);
    for (int L : {25, 29, 32}) bad += check<4>(L, 400 + L, 400000);
    std::printf(bad ? "FAILED: %d\n" : "compressed keys: valid <=> one-hot, injective on valid reads (%d errors)\n", bad);
    return bad ? 1 : 0;
}
