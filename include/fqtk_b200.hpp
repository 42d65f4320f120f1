// fqtk_b200.hpp — header-only C++ mirror of the reference's matcher interface over the C ABI (fqtk_b200.h).
//
// The reference's host side is Rust and this image has no Rust toolchain, so the host-side mirror above the C ABI
// is written in C++ (the reference is compiled code) and, for the pytest parity suite, in Python
// (fqtk_b200/barcode_matching.py).  Same names, argument meaning and error behaviour as
//   fqtk_lib::barcode_matching::{BarcodeMatch, BarcodeMatcher}   src/lib/barcode_matching.rs:15-186
// so tests/cpp/test_barcode_matching.cpp reads like the reference's own #[cfg(test)] module.
// A Rust panic becomes a thrown fqtk_b200::Panic carrying the reference's panic text.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "fqtk_b200.h"

namespace fqtk_b200 {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// src/lib/samples.rs:17-26
struct Sample {
    std::string sample_id;
    std::string barcode;
    std::size_t ordinal = 0;
};

// src/lib/barcode_matching.rs:16-25
struct BarcodeMatch {
    std::size_t best_match;
    std::uint8_t best_mismatches;
    std::uint8_t next_best_mismatches;
    bool operator==(const BarcodeMatch& o) const {
        return best_match == o.best_match && best_mismatches == o.best_mismatches &&
               next_best_mismatches == o.next_best_mismatches;
    }
};

inline std::optional<BarcodeMatch> unpack(std::uint32_t w) {
    if (w == FQTK_B200_NONE) return std::nullopt;
    return BarcodeMatch{FQTK_B200_BEST_MATCH(w), (std::uint8_t)FQTK_B200_BEST_MISMATCHES(w),
                        (std::uint8_t)FQTK_B200_NEXT_BEST_MISMATCHES(w)};
}

// src/lib/barcode_matching.rs:29-186
class BarcodeMatcher {
  public:
    // BarcodeMatcher::new, :55-86
    BarcodeMatcher(const std::vector<Sample>& samples, std::uint8_t max_mismatches, std::uint8_t min_mismatch_delta,
                   bool use_cache, int device = 0) {
        if (samples.empty()) throw Panic("Must provide at least one sample");
        for (const auto& s : samples)
            if (s.barcode.empty()) throw Panic("Sample barcode cannot be empty string");
        const std::size_t L = samples[0].barcode.size();
        std::string panel;
        for (const auto& s : samples) {
            if (s.barcode.size() != L) throw Panic("All barcodes must have the same length");
            panel += s.barcode;
        }
        n_samples_ = samples.size();
        check(fqtk_b200_matcher_create(reinterpret_cast<const std::uint8_t*>(panel.data()), (std::uint32_t)samples.size(),
                                       (std::uint32_t)L, max_mismatches, min_mismatch_delta, use_cache ? 1 : 0, device,
                                       &h_));
    }
    ~BarcodeMatcher() { fqtk_b200_matcher_destroy(h_); }
    BarcodeMatcher(const BarcodeMatcher&) = delete;
    BarcodeMatcher& operator=(const BarcodeMatcher&) = delete;

    // BarcodeMatcher::assign, :165-186
    std::optional<BarcodeMatch> assign(const std::string& read_bases) {
        std::uint32_t w = 0;
        check(fqtk_b200_matcher_assign(h_, reinterpret_cast<const std::uint8_t*>(read_bases.data()), read_bases.size(),
                                       &w));
        return unpack(w);
    }

    // the batched form the GPU wants: n rows of `stride` bytes -> one result word per row (host buffers)
    void assign_batch(const std::uint8_t* rows, std::uint64_t n, std::uint64_t stride, std::uint32_t* results,
                      const std::uint32_t* lengths = nullptr) {
        check(fqtk_b200_matcher_assign_batch(h_, rows, n, stride, lengths, results));
    }

    // DemuxMetric.templates per sample, last = unmatched (demux.rs:458,971,974)
    std::vector<std::uint64_t> counts() {
        std::vector<std::uint64_t> c(n_samples_ + 1);
        check(fqtk_b200_matcher_counts(h_, c.data()));
        return c;
    }
    void reset_counts() { check(fqtk_b200_matcher_reset_counts(h_)); }
    fqtk_b200_matcher* handle() { return h_; }

  private:
    static void check(int rc) {
        if (rc == FQTK_B200_OK) return;
        const std::string msg = fqtk_b200_last_error();
        if (rc == FQTK_B200_ERR_EMPTY_PANEL || rc == FQTK_B200_ERR_EMPTY_BARCODE || rc == FQTK_B200_ERR_LENGTH)
            throw Panic(msg);
        throw Error(rc, msg);
    }
    fqtk_b200_matcher* h_ = nullptr;
    std::size_t n_samples_ = 0;
};

}  // namespace fqtk_b200
