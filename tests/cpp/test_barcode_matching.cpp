// Mirrors the reference's #[cfg(test)] module for BarcodeMatcher::assign (src/lib/barcode_matching.rs:189-448 of
// fulcrumgenomics/fqtk) against the C++ host mirror in include/fqtk_b200.hpp.  Each case is run with use_cache
// true and false, like the reference's rstest parametrisation (true = memo-table kernels, false = brute-force kernels).
// Build + run: tests/test_cpp_mirror.py (needs a GPU).
#include <cstdio>
#include <cstdlib>
#include <optional>
#include <string>
#include <vector>

#include "fqtk_b200.hpp"

using fqtk_b200::BarcodeMatch;
using fqtk_b200::BarcodeMatcher;
using fqtk_b200::Panic;
using fqtk_b200::Sample;

static int failures = 0;
#define CHECK(cond)                                                               \
    do {                                                                          \
        if (!(cond)) {                                                            \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                           \
        }                                                                         \
    } while (0)

// barcode_matching.rs:205-213
static std::vector<Sample> barcodes_to_samples(const std::vector<std::string>& barcodes) {
    std::vector<Sample> out;
    for (std::size_t i = 0; i < barcodes.size(); i++) out.push_back(Sample{"sample_" + std::to_string(i), barcodes[i], i});
    return out;
}

static void run(bool use_cache) {
    {  // test_barcode_matcher_instantiation_can_succeed, :226-232
        BarcodeMatcher m(barcodes_to_samples({"ACGT"}), 2, 1, use_cache);
    }
    {  // test_barcode_matcher_fails_if_no_samples_provided, :234-243
        bool panicked = false;
        try {
            BarcodeMatcher m(barcodes_to_samples({}), 2, 1, use_cache);
        } catch (const Panic& p) {
            panicked = std::string(p.what()) == "Must provide at least one sample";
        }
        CHECK(panicked);
    }
    {  // test_assign_exact_match, :326-341
        auto samples = barcodes_to_samples({"ACGT", "AAAG", "CACA"});
        BarcodeMatcher m(samples, 2, 2, use_cache);
        CHECK(m.assign(samples[0].barcode) == std::optional<BarcodeMatch>(BarcodeMatch{0, 0, 3}));
    }
    {  // test_assign_imprecise_match, :343-355
        BarcodeMatcher m(barcodes_to_samples({"AAAT", "AGAG", "CACA"}), 2, 2, use_cache);
        CHECK(m.assign("GAAT") == std::optional<BarcodeMatch>(BarcodeMatch{0, 1, 3}));
    }
    {  // test_assign_precise_match_with_no_call, :357-369
        BarcodeMatcher m(barcodes_to_samples({"AAAT", "AGAG", "CACA"}), 2, 2, use_cache);
        CHECK(m.assign("NAAT") == std::optional<BarcodeMatch>(BarcodeMatch{0, 1, 3}));
    }
    {  // test_assign_imprecise_match_with_no_call, :371-385
        BarcodeMatcher m(barcodes_to_samples({"AAATTT", "AGAGGG", "CACAGG"}), 2, 2, use_cache);
        CHECK(m.assign("NAGTTT") == std::optional<BarcodeMatch>(BarcodeMatch{0, 2, 5}));
    }
    {  // test_sample_no_call_doesnt_contribute_to_mismatch_number, :387-401
        BarcodeMatcher m(barcodes_to_samples({"NAGTTT", "AGAGGG", "CACAGG"}), 1, 2, use_cache);
        CHECK(m.assign("AAATTT") == std::optional<BarcodeMatch>(BarcodeMatch{0, 1, 4}));
    }
    {  // test_read_no_call_contributes_to_mismatch_number, :404-417
        BarcodeMatcher m(barcodes_to_samples({"AAATTT", "AGAGGG", "CACAGG"}), 1, 2, use_cache);
        CHECK(m.assign("NAGTTT") == std::nullopt);
    }
    {  // test_produce_no_match_if_too_many_mismatches, :419-427
        BarcodeMatcher m(barcodes_to_samples({"AAGCTAG", "CAGCTAG", "GAGCTAG", "TAGCTAG"}), 0, 100, use_cache);
        CHECK(m.assign("ATCGATC") == std::nullopt);
    }
    {  // test_produce_no_match_if_within_mismatch_delta, :429-437
        auto samples = barcodes_to_samples({"AAAAAAAA", "CCCCCCCC", "GGGGGGGG", "GGGGGGTT"});
        BarcodeMatcher m(samples, 100, 3, use_cache);
        CHECK(m.assign(samples[3].barcode) == std::nullopt);
    }
    {  // test_produce_no_match_if_too_many_mismatches_via_nocalls, :439-447
        BarcodeMatcher m(barcodes_to_samples({"AAAAAAAA", "CCCCCCCC", "GGGGGGGG", "GGGGGGTT"}), 0, 100, use_cache);
        CHECK(m.assign("GGGGGGTN") == std::nullopt);
    }
    {  // length rules of assign / count_mismatches, :95-106,167-172,251-257,314-320
        BarcodeMatcher m(barcodes_to_samples({"CTATGT"}), 1, 0, use_cache);
        CHECK(m.assign("GATTA") == std::nullopt);  // shorter than the barcodes: None
        bool panicked = false;
        try {
            m.assign("CTATGTA");  // longer: the reference panics inside count_mismatches
        } catch (const Panic& p) {
            panicked = std::string(p.what()).find("differs from expected barcode (CTATGT) length (6)") != std::string::npos;
        }
        CHECK(panicked);
        auto c = m.counts();
        CHECK(c.size() == 2 && c[0] == 0 && c[1] == 1);
    }
}

// the demux caller's loop (demux.rs:967-975) over a GROUP of matchers: three handles on device 0 (so that it also runs on a
// one-GPU box) must fill results[] and ONE count table exactly like a single matcher
static void run_group() {
    const std::vector<std::string> bcs = {"AAAAAAAA", "CCCCCCCC", "GGGGGGGG", "GGGGGGTT"};
    auto samples = barcodes_to_samples(bcs);
    std::string rows;
    const std::uint64_t n = 40000;
    for (std::uint64_t i = 0; i < n; i++) {
        std::string r = bcs[i % 4];
        if (i % 7 == 0) r[i % 8] = 'N';
        if (i % 11 == 0) r = "ACGTACGT";
        rows += r;
    }
    std::vector<std::uint32_t> one(n), many(n);
    BarcodeMatcher m(samples, 1, 2, true);
    m.assign_batch(reinterpret_cast<const std::uint8_t*>(rows.data()), n, 8, one.data());
    fqtk_b200::MatcherGroup g(samples, 1, 2, true, std::vector<int>{0, 0, 0});
    CHECK(g.size() == 3);
    g.assign_batch(reinterpret_cast<const std::uint8_t*>(rows.data()), n, 8, many.data());
    CHECK(one == many);
    CHECK(m.counts() == g.counts());
    g.assign_batch(reinterpret_cast<const std::uint8_t*>(rows.data()), n, 8, many.data());
    auto c1 = m.counts(), c2 = g.counts();
    bool doubled = true;
    for (std::size_t j = 0; j < c1.size(); j++) doubled = doubled && c2[j] == 2 * c1[j];
    CHECK(doubled);
}

int main() {
    run(true);
    run(false);
    run_group();
    if (failures) {
        std::fprintf(stderr, "%d check(s) failed\n", failures);
        return 1;
    }
    std::printf("cpp mirror: all reference assign cases pass (use_cache = true and false)\n");
    return 0;
}
