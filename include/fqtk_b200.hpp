// fqtk_b200.hpp — header-only C++ mirror of the reference's matcher interface over the C ABI (fqtk_b200.h).
//
// The reference's host side is Rust and this image has no Rust toolchain, so the host-side mirror above the C ABI
// is written in C++ (the reference is compiled code) and, for the pytest parity suite, in Python
// (fqtk_b200/barcode_matching.py).  Same names, argument meaning and error behaviour as
//   fqtk_lib::barcode_matching::{BarcodeMatch, BarcodeMatcher}   src/lib/barcode_matching.rs:15-186
// so tests/cpp/test_barcode_matching.cpp reads like the reference's own #[cfg(test)] module.
// A Rust panic becomes a thrown fqtk_b200::Panic carrying the reference's panic text.
#pragma once
#include <algorithm>
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
    // assign_batch with encode() done by `threads` host threads while the batch is in flight (-1 = auto, 0 = off): most
    // chunks cross PCIe as BitEnc words instead of ASCII rows; same results (fqtk_b200_matcher_set_host_pack)
    void set_host_pack(int threads) { check(fqtk_b200_matcher_set_host_pack(h_, threads)); }
    // stable partition of read indices by assignment (the batch form of demux.rs:970-975): order[offsets[s] ..
    // offsets[s + 1]) = the reads of sample s in input order, s = n_samples = unmatched; offsets has n_samples + 2 entries
    void route(const std::uint32_t* results, std::uint64_t n, std::vector<std::uint32_t>& order,
               std::vector<std::uint64_t>& offsets) {
        order.resize(n);
        offsets.assign(n_samples_ + 2, 0);
        check(fqtk_b200_matcher_route(h_, results, n, order.data(), offsets.data()));
    }
    fqtk_b200_matcher* handle() { return h_; }
    // same batch call for reads the host holds in BitEnc form (fqtk_b200_pack_host); either output may be null
    void assign_batch_packed(const std::uint32_t* packed, std::uint64_t n, std::uint32_t* results,
                             std::uint16_t* sample_index = nullptr) {
        check(fqtk_b200_matcher_assign_batch_packed(h_, packed, n, results, sample_index));
    }

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

// ReadSet::write_header_internal, src/bin/commands/demux.rs:171-267 — the rewritten FASTQ header line (with the leading
// '@', without a newline).  The reference's `ensure!` failures and `unwrap` panics become a thrown Panic.
inline std::string write_header(std::size_t read_num, const std::string& header,
                                const std::vector<std::string>& sample_barcode_segments,
                                const std::vector<std::string>& molecular_barcode_segments = {}) {
    auto count = [](const std::string& t, char c) {
        std::size_t n = 0;
        for (char x : t) n += x == c;
        return n;
    };
    auto join = [](const std::vector<std::string>& v) {
        std::string o;
        for (std::size_t i = 0; i < v.size(); i++) o += (i ? "+" : "") + v[i];
        return o;
    };
    const std::size_t sp = header.find(' ');
    const bool has_comment = sp != std::string::npos;
    const std::string name = has_comment ? header.substr(0, sp) : header;
    const std::string comment = has_comment ? header.substr(sp + 1) : std::string();
    std::string out = "@";
    if (!molecular_barcode_segments.empty()) {  // :189-213
        const std::size_t sep = count(name, ':');
        if (sep > 7) throw Panic("Can't handle read name with more than 8 segments: " + header);
        out += name;
        out += sep == 7 ? "+" : ":";
        out += join(molecular_barcode_segments);
    } else {
        out += name;
    }
    out += ' ';
    if (!has_comment) {  // :219-223
        out += std::to_string(read_num) + ":N:0:";
    } else {
        if (comment.empty()) throw Panic("empty comment after the read name: " + header);  // chars.last().unwrap(), :230
        const std::size_t sep = count(comment, ':');
        if (sep < 3) {  // :228-233
            out += comment;
            if (comment.back() != ':') out += ':';
        } else {
            if (sep != 3) throw Panic("Comment in did not have 4 segments: " + header);
            const std::size_t first = comment.find(':');
            const bool digit = comment.back() >= '0' && comment.back() <= '9';  // Illumina's "0" index, :241-246
            const std::string remainder = comment.substr(first + 1, comment.size() - first - 1 - (digit ? 1 : 0));
            out += std::to_string(read_num) + ":" + remainder;
            if (remainder.empty() || remainder.back() != ':') out += '+';
        }
    }
    out += join(sample_barcode_segments);  // :257-263
    return out;
}

// DemuxMetric + DemuxMetric::update, demux.rs:452-497: rows for the S samples and, last, the unmatched pseudo-sample
// (barcode "."); the ratios exclude the unmatched row from mean and best.  counts = S + 1 integers, last = unmatched.
struct DemuxMetric {
    std::string sample_id, barcode;
    std::uint64_t templates = 0;
    double frac_templates = 0, ratio_to_mean = 0, ratio_to_best = 0;
};
inline std::vector<DemuxMetric> demux_metrics(const std::vector<Sample>& samples, const std::vector<std::uint64_t>& counts,
                                              const std::string& unmatched_prefix = "unmatched") {
    if (counts.size() != samples.size() + 1) throw Error(FQTK_B200_ERR_ARG, "need S + 1 counts");
    std::vector<DemuxMetric> rows;
    double sample_total = 0, best = 0;
    for (std::size_t j = 0; j < samples.size(); j++) {
        rows.push_back({samples[j].sample_id, samples[j].barcode, counts[j]});
        sample_total += (double)counts[j];
        best = std::max(best, (double)counts[j]);
    }
    rows.push_back({unmatched_prefix, ".", counts.back()});
    const double total = sample_total + (double)counts.back(), mean = sample_total / (double)samples.size();
    for (auto& r : rows) {  // IEEE division, as Rust's f64: x / 0 = inf, 0 / 0 = NaN
        r.frac_templates = (double)r.templates / total;
        r.ratio_to_mean = (double)r.templates / mean;
        r.ratio_to_best = (double)r.templates / best;
    }
    return rows;
}

// One matcher per GPU of the box behind one handle (fqtk_b200_group_*): contiguous shards of every batch, ONE count
// table — what demux.rs:921-926 / 970-975 own, over several devices.  `devices` empty = every visible device.
class MatcherGroup {
  public:
    MatcherGroup(const std::vector<Sample>& samples, std::uint8_t max_mismatches, std::uint8_t min_mismatch_delta,
                 bool use_cache, const std::vector<int>& devices = {}) {
        if (samples.empty()) throw Panic("Must provide at least one sample");
        const std::size_t L = samples[0].barcode.size();
        std::string panel;
        for (const auto& s : samples) {
            if (s.barcode.empty()) throw Panic("Sample barcode cannot be empty string");
            if (s.barcode.size() != L) throw Panic("All barcodes must have the same length");
            panel += s.barcode;
        }
        n_samples_ = samples.size();
        const int rc = fqtk_b200_group_create(reinterpret_cast<const std::uint8_t*>(panel.data()), (std::uint32_t)samples.size(),
                                              (std::uint32_t)L, max_mismatches, min_mismatch_delta, use_cache ? 1 : 0,
                                              devices.empty() ? nullptr : devices.data(), (std::uint32_t)devices.size(),
                                              nullptr, &g_);
        if (rc != FQTK_B200_OK) throw Error(rc, fqtk_b200_last_error());
    }
    ~MatcherGroup() { fqtk_b200_group_destroy(g_); }
    MatcherGroup(const MatcherGroup&) = delete;
    MatcherGroup& operator=(const MatcherGroup&) = delete;

    std::uint32_t size() const { return fqtk_b200_group_size(g_); }
    void assign_batch(const std::uint8_t* rows, std::uint64_t n, std::uint64_t stride, std::uint32_t* results,
                      const std::uint32_t* lengths = nullptr) {
        const int rc = fqtk_b200_group_assign_batch(g_, rows, n, stride, lengths, results);
        if (rc == FQTK_B200_ERR_LENGTH) throw Panic(fqtk_b200_last_error());
        if (rc != FQTK_B200_OK) throw Error(rc, fqtk_b200_last_error());
    }
    std::vector<std::uint64_t> counts() {
        std::vector<std::uint64_t> c(n_samples_ + 1);
        const int rc = fqtk_b200_group_counts(g_, c.data());
        if (rc != FQTK_B200_OK) throw Error(rc, fqtk_b200_last_error());
        return c;
    }
    void reset_counts() { fqtk_b200_group_reset_counts(g_); }

  private:
    fqtk_b200_group* g_ = nullptr;
    std::size_t n_samples_ = 0;
};

// The reference's output side (src/bin/commands/demux.rs:755-798): `PooledWriter`s over `BgzfCompressor`.  One
// BgzfPool per device owns the GPU compressor; every BgzfWriter buffers its file's bytes (`write`) and the pool turns
// what is buffered into BGZF members in one device pass per writer (`flush`: only whole 65 280-byte blocks leave, as the
// reference's writers do; `finish`: the rest + the EOF block).  `sink` receives the file image bytes in order.
class BgzfPool {
  public:
    explicit BgzfPool(int device = 0, int compression_level = 5, std::uint64_t chunk_bytes = 0) : level_(compression_level) {
        const int rc = fqtk_b200_bgzf_create(device, chunk_bytes, &z_);
        if (rc != FQTK_B200_OK) throw Error(rc, fqtk_b200_last_error());
    }
    ~BgzfPool() { fqtk_b200_bgzf_destroy(z_); }
    BgzfPool(const BgzfPool&) = delete;
    BgzfPool& operator=(const BgzfPool&) = delete;

    // members for `n` bytes (+ EOF block if asked), appended to `out`
    void compress(const std::uint8_t* data, std::uint64_t n, bool eof, std::string& out) {
        const std::uint64_t cap = fqtk_b200_bgzf_bound(n);
        const std::size_t at = out.size();
        out.resize(at + cap);
        std::uint64_t written = 0;
        const int rc = fqtk_b200_bgzf_compress(z_, data, n, level_, eof ? 1 : 0, reinterpret_cast<std::uint8_t*>(&out[at]), cap, &written);
        if (rc != FQTK_B200_OK) throw Error(rc, fqtk_b200_last_error());
        out.resize(at + written);
    }

  private:
    fqtk_b200_bgzf* z_ = nullptr;
    int level_;
};

class BgzfWriter {
  public:
    static constexpr std::size_t BLOCK = 65280;  // bgzf crate BGZF_BLOCK_SIZE
    explicit BgzfWriter(BgzfPool& pool) : pool_(pool) {}
    void write_all(const void* p, std::size_t n) { buf_.append(static_cast<const char*>(p), n); }
    void write_all(const std::string& s) { buf_ += s; }
    // every whole block buffered so far -> members appended to `image`
    void flush(std::string& image) {
        const std::size_t whole = buf_.size() / BLOCK * BLOCK;
        if (!whole) return;
        pool_.compress(reinterpret_cast<const std::uint8_t*>(buf_.data()), whole, false, image);
        buf_.erase(0, whole);
    }
    // close(): the partial last block, then the EOF block
    void finish(std::string& image) {
        pool_.compress(reinterpret_cast<const std::uint8_t*>(buf_.data()), buf_.size(), true, image);
        buf_.clear();
    }

  private:
    BgzfPool& pool_;
    std::string buf_;
};

}  // namespace fqtk_b200
