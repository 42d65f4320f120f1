// BgzfPool / BgzfWriter of include/fqtk_b200.hpp (the reference's pooled BGZF writers, demux.rs:755-798) against zlib's
// own inflater: records written through a buffering writer, flushed in whole blocks, closed with the EOF block; every
// member must inflate to its 65 280-byte piece with the CRC32 / ISIZE / BSIZE the header and trailer state.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "fqtk_b200.hpp"

static bool inflate_members(const std::string& img, std::string& out, size_t& members) {
    size_t pos = 0;
    members = 0;
    while (pos < img.size()) {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(img.data()) + pos;
        if (img.size() - pos < 26 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || p[3] != 4 || p[12] != 'B' || p[13] != 'C') return false;
        const size_t bsize = (size_t)(p[16] | (p[17] << 8)) + 1;
        if (pos + bsize > img.size()) return false;
        const unsigned crc = p[bsize - 8] | (p[bsize - 7] << 8) | (p[bsize - 6] << 16) | ((unsigned)p[bsize - 5] << 24);
        const unsigned isize = p[bsize - 4] | (p[bsize - 3] << 8) | (p[bsize - 2] << 16) | ((unsigned)p[bsize - 1] << 24);
        if (isize > 65280) return false;
        std::string piece(isize, '\0');
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) return false;
        zs.next_in = const_cast<unsigned char*>(p + 18);
        zs.avail_in = (uInt)(bsize - 26);
        zs.next_out = reinterpret_cast<unsigned char*>(&piece[0]);
        zs.avail_out = (uInt)isize;
        const int rc = inflate(&zs, Z_FINISH);
        const bool ok = rc == Z_STREAM_END && zs.avail_in == 0 && zs.total_out == isize;
        inflateEnd(&zs);
        if (!ok) return false;
        if ((unsigned)crc32(0L, reinterpret_cast<const unsigned char*>(piece.data()), (uInt)isize) != crc) return false;
        out += piece;
        members++;
        pos += bsize;
    }
    return true;
}

int main() {
    fqtk_b200::BgzfPool pool(0, 5, 65280 * 16);
    fqtk_b200::BgzfWriter w(pool);
    std::string text, image;
    unsigned x = 12345;
    for (int i = 0; i < 4000; i++) {  // ~1.2 MB of FASTQ records, flushed every 500 records like a batch loop would
        char head[96];
        std::snprintf(head, sizeof head, "@A00123:45:HXXXXXX:1:1101:%d:%d 1:N:0:ACGTACGT\n", 1000 + i * 7 % 30000, 2000 + i * 13 % 30000);
        std::string rec = head;
        for (int k = 0; k < 150; k++) { x = x * 1664525u + 1013904223u; rec += "ACGT"[x >> 30]; }
        rec += "\n+\n";
        for (int k = 0; k < 150; k++) { x = x * 1664525u + 1013904223u; rec += (x >> 28) ? 'F' : ','; }
        rec += "\n";
        w.write_all(rec);
        text += rec;
        if (i % 500 == 499) w.flush(image);
    }
    w.finish(image);
    std::string back;
    size_t members = 0;
    if (!inflate_members(image, back, members)) { std::puts("FAIL: a member does not inflate"); return 1; }
    if (back != text) { std::puts("FAIL: payload differs"); return 1; }
    if (members != (text.size() + 65279) / 65280 + 1) { std::printf("FAIL: %zu members\n", members); return 1; }
    if (image.size() < 28 || std::memcmp(image.data() + image.size() - 28, "\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0\x1b\0\x03\0\0\0\0\0\0\0\0\0", 28) != 0) {
        std::puts("FAIL: no EOF block");
        return 1;
    }
    std::printf("bgzf writer: %zu bytes -> %zu bytes in %zu members, every member inflates (zlib)\n", text.size(), image.size(), members);

    // ---- the one-call batch form (fqtk_b200_demux_chunks): what the Rust main loop would call once per batch ----
    {
        const char* bcs[3] = {"AAAAAAAA", "CCCCCCCC", "GGGGGGGG"};
        const char* obs[4] = {"AAAAAAAA", "CCCCCCCC", "GGGGGGGG", "TTTTTTTT"};  // the 4th matches nothing
        std::string panel = std::string(bcs[0]) + bcs[1] + bcs[2];
        fqtk_b200_matcher* m = nullptr;
        if (fqtk_b200_matcher_create(reinterpret_cast<const std::uint8_t*>(panel.data()), 3, 8, 1, 2, 1, 0, &m) != FQTK_B200_OK) {
            std::printf("FAIL: matcher_create: %s\n", fqtk_b200_last_error());
            return 1;
        }
        fqtk_b200_bgzf* z = nullptr;
        if (fqtk_b200_bgzf_create(0, 0, &z) != FQTK_B200_OK) { std::puts("FAIL: bgzf_create"); return 1; }
        const int n = 30000;
        std::string chunk, expect[4];
        unsigned y = 99;
        for (int i = 0; i < n; i++) {
            std::string tmpl;
            for (int k = 0; k < 40 + i % 7; k++) { y = y * 1664525u + 1013904223u; tmpl += "ACGT"[y >> 30]; }
            const std::string q(tmpl.size(), 'F');
            const int b = i % 4;
            chunk += "@r" + std::to_string(i) + "\n" + obs[b] + tmpl + "\n+\n" + std::string(8, 'I') + q + "\n";
            expect[b] += "@r" + std::to_string(i) + " 1:N:0:" + obs[b] + "\n" + tmpl + "\n+\n" + q + "\n";
        }
        chunk += "@cut_off_record\nACGT";  // an incomplete record at the end of the chunk: carried over, not written
        const fqtk_b200_fastq_chunk ch{reinterpret_cast<const std::uint8_t*>(chunk.data()), chunk.size()};
        const fqtk_b200_read_segment segs[2] = {{0, 'B', 0, 8}, {0, 'T', 8, FQTK_B200_SEGMENT_REST}};
        std::string out(fqtk_b200_bgzf_bound(chunk.size() * 2), '\0');
        std::uint64_t offs[4 + 1], counts[4], n_reads = 0, consumed = 0;
        std::uint32_t n_streams = 0;
        const int rc = fqtk_b200_demux_chunks(m, z, &ch, 1, segs, 2, "T", 5, ~0ull, reinterpret_cast<std::uint8_t*>(&out[0]), out.size(), offs,
                                              &n_streams, counts, &n_reads, &consumed);
        if (rc != FQTK_B200_OK) { std::printf("FAIL: demux_chunks: %s\n", fqtk_b200_last_error()); return 1; }
        if (n_streams != 1 || n_reads != (std::uint64_t)n || consumed != chunk.size() - std::strlen("@cut_off_record\nACGT")) {
            std::printf("FAIL: %u streams, %llu reads, %llu consumed\n", n_streams, (unsigned long long)n_reads, (unsigned long long)consumed);
            return 1;
        }
        for (int b = 0; b < 4; b++) {
            std::string got;
            size_t mem = 0;
            if (!inflate_members(out.substr(offs[b], offs[b + 1] - offs[b]), got, mem) || got != expect[b] || counts[b] != (std::uint64_t)n / 4) {
                std::printf("FAIL: run of bucket %d differs\n", b);
                return 1;
            }
        }
        fqtk_b200_bgzf_destroy(z);
        fqtk_b200_matcher_destroy(m);
        std::printf("demux_chunks: %d reads -> 4 runs (3 samples + unmatched), every run inflates to its records\n", n);
    }
    return 0;
}
