"""TEST INFRASTRUCTURE — CPU restatement of the reference's BGZF output path, used only by tests/, smoke() and
bench.py's CPU legs (never by the product).

The reference's writers are `pooled_writer::PooledWriter` over `bgzf::BgzfCompressor`
(src/bin/commands/demux.rs:755-798; pooled-writer 0.4.0 and bgzf 0.2 from Cargo.lock, not vendored): each writer buffers
BGZF_BLOCK_SIZE = 65 280 bytes, every full buffer (and the last partial one) becomes one gzip member
    1f 8b 08 04 | MTIME 0 | XFL | OS ff | XLEN 6 | 'B' 'C' 2 0 BSIZE-1 | raw deflate | CRC32 | ISIZE      (SAM spec 4.1)
and close() appends the 28-byte EOF member.  The deflate bytes come from libdeflate there and from zlib here: deflate
output is implementation-defined, so what parity means for this row is (a) the same members: one per 65 280-byte piece,
in order, ISIZE / CRC32 of the piece, BSIZE of the member, then the EOF member; (b) every member inflates to its piece.
`parse()` checks exactly that for any producer."""
from __future__ import annotations

import struct
import zlib

BGZF_BLOCK_SIZE = 65280
BGZF_EOF = bytes([0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0,
                  0, 0, 0, 0, 0, 0, 0, 0])


def member(piece: bytes, level: int = 5) -> bytes:
    """One BGZF member for `piece` (<= 65 280 bytes)."""
    assert len(piece) <= BGZF_BLOCK_SIZE
    if level == 0:
        n = len(piece)
        cdata = bytes([1]) + struct.pack("<HH", n, n ^ 0xFFFF) + piece
    else:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        cdata = co.compress(piece) + co.flush()
    xfl = 2 if level >= 9 else (4 if level == 1 else 0)
    total = 18 + len(cdata) + 8
    assert total <= 65536
    head = bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, xfl, 0xff, 6, 0, 0x42, 0x43, 2, 0]) + struct.pack("<H", total - 1)
    return head + cdata + struct.pack("<II", zlib.crc32(piece) & 0xFFFFFFFF, len(piece))


def compress(data: bytes, level: int = 5, eof: bool = True) -> bytes:
    """What one pooled writer leaves in its file for `data`."""
    out = [member(data[i:i + BGZF_BLOCK_SIZE], level) for i in range(0, len(data), BGZF_BLOCK_SIZE)]
    if eof:
        out.append(BGZF_EOF)
    return b"".join(out)


def parse(image: bytes):
    """Strict BGZF reader: returns (payload, [ISIZE of every member]); raises ValueError on any framing / CRC error."""
    pos, pieces, sizes = 0, [], []
    while pos < len(image):
        if len(image) - pos < 26:
            raise ValueError("truncated member")
        if image[pos:pos + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError(f"bad gzip magic / flags at {pos}")
        if image[pos + 4:pos + 8] != b"\0\0\0\0" or image[pos + 9] != 0xff:
            raise ValueError("MTIME / OS differ from the bgzf crate's header")
        xlen = struct.unpack_from("<H", image, pos + 10)[0]
        if xlen != 6 or image[pos + 12:pos + 16] != b"BC\x02\x00":
            raise ValueError("missing BC extra subfield")
        bsize = struct.unpack_from("<H", image, pos + 16)[0] + 1
        if pos + bsize > len(image):
            raise ValueError("BSIZE runs past the end")
        cdata = image[pos + 18:pos + bsize - 8]
        crc, isize = struct.unpack_from("<II", image, pos + bsize - 8)
        d = zlib.decompressobj(-15)
        piece = d.decompress(cdata) + d.flush()
        if not d.eof or d.unused_data:
            raise ValueError("deflate stream does not end with the member")
        if len(piece) != isize or (zlib.crc32(piece) & 0xFFFFFFFF) != crc:
            raise ValueError("ISIZE / CRC32 mismatch")
        if isize > BGZF_BLOCK_SIZE:
            raise ValueError("member larger than BGZF_BLOCK_SIZE")
        pieces.append(piece)
        sizes.append(isize)
        pos += bsize
    return b"".join(pieces), sizes
