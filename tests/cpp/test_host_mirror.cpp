// CPU-only test of the C++ host mirror's formatting pieces (include/fqtk_b200.hpp): the reference's six
// write_header_internal tests (src/bin/commands/demux.rs:2084-2196) and the DemuxMetric::update rule (:481-496).
#include <cmath>
#include <cstdio>
#include <string>

#include "fqtk_b200.hpp"

using namespace fqtk_b200;

static int fails = 0;
#define CHECK(cond)                                              \
    do {                                                         \
        if (!(cond)) {                                           \
            std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            fails++;                                             \
        }                                                        \
    } while (0)

int main() {
    const std::vector<std::string> bcs = {"ACGT", "GGTT"}, umi = {"AACCGGTT"};
    // test_write_header_standard_no_umi
    CHECK(write_header(1, "inst:123:ABCDE:1:204:1022:2108 1:N:0:0", bcs) == "@inst:123:ABCDE:1:204:1022:2108 1:N:0:ACGT+GGTT");
    // test_write_header_standard_with_umi
    CHECK(write_header(2, "inst:123:ABCDE:1:204:1022:2108 1:Y:0:0", bcs, umi) ==
          "@inst:123:ABCDE:1:204:1022:2108:AACCGGTT 2:Y:0:ACGT+GGTT");
    // test_write_header_append_barcode_and_umi
    CHECK(write_header(2, "inst:123:ABCDE:1:204:1022:2108:AAAA 1:Y:0:TTTT", bcs, umi) ==
          "@inst:123:ABCDE:1:204:1022:2108:AAAA+AACCGGTT 2:Y:0:TTTT+ACGT+GGTT");
    // test_write_header_short_name_no_comment
    CHECK(write_header(1, "q1", bcs, umi) == "@q1:AACCGGTT 1:N:0:ACGT+GGTT");
    // test_write_header_name_too_many_parts (should_panic "8 segments")
    try {
        write_header(1, "q1:1:2:3:4:5:6:7:8:9:10", bcs, umi);
        CHECK(false);
    } catch (const Panic& e) {
        CHECK(std::string(e.what()).find("8 segments") != std::string::npos);
    }
    // test_write_header_comment_too_few_parts
    CHECK(write_header(1, "q1 0:0", bcs, umi) == "@q1:AACCGGTT 0:0:ACGT+GGTT");
    // branches beyond the reference's tests, from the code
    CHECK(write_header(2, "q1 1:N:0:GATC", {"AC"}) == "@q1 2:N:0:GATC+AC");
    try {
        write_header(1, "q1 1:N:0:1:2", {"AC"});
        CHECK(false);
    } catch (const Panic& e) {
        CHECK(std::string(e.what()).find("4 segments") != std::string::npos);
    }

    // DemuxMetric::update
    const std::vector<Sample> samples = {{"a", "AAAA", 0}, {"b", "CCCC", 1}, {"c", "GGGG", 2}};
    const auto rows = demux_metrics(samples, {30, 10, 20, 40});
    CHECK(rows.size() == 4 && rows[3].sample_id == "unmatched" && rows[3].barcode == ".");
    CHECK(std::fabs(rows[0].frac_templates - 0.3) < 1e-12 && std::fabs(rows[3].frac_templates - 0.4) < 1e-12);
    CHECK(std::fabs(rows[0].ratio_to_mean - 1.5) < 1e-12 && std::fabs(rows[3].ratio_to_mean - 2.0) < 1e-12);
    CHECK(std::fabs(rows[1].ratio_to_best - 1.0 / 3.0) < 1e-12 && rows[0].ratio_to_best == 1.0);
    const auto zero = demux_metrics({{"a", "AAAA", 0}}, {0, 0});
    CHECK(std::isnan(zero[0].frac_templates) && std::isnan(zero[0].ratio_to_best));

    std::printf(fails ? "%d checks failed\n" : "host mirror: header rewrite + metrics match the reference's tests (%d failures)\n", fails);
    return fails ? 1 : 0;
}
