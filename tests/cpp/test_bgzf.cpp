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
    return 0;
}
