"""FASTQ header rewrite of the demuxed records — host-side mirror of `ReadSet::write_header_internal`
(src/bin/commands/demux.rs:171-267; SURVEY 8f "next" #4).  Pure byte formatting, done once per written record on the
host next to the writers; the GPU path only supplies which sample the record goes to.

    @<name>[:<UMI>|+<UMI>] <read_num>:<filter>:<control>:[<existing barcode>+]<sample barcode segments joined by '+'>
"""
from __future__ import annotations

from typing import Sequence


class HeaderError(ValueError):
    """The reference's `ensure!` failures (demux.rs:190-194, 229-233)."""


def write_header(read_num: int, header: bytes, sample_barcode_segments: Sequence[bytes],
                 molecular_barcode_segments: Sequence[bytes] = ()) -> bytes:
    """Returns the rewritten header line (with the leading '@', without a newline)."""
    # name and optional comment split at the first space (demux.rs:179-183)
    sp = header.find(b" ")
    name, comment = (header, None) if sp < 0 else (header[:sp], header[sp + 1:])
    out = bytearray(b"@")
    umis = list(molecular_barcode_segments)
    if umis:  # demux.rs:189-213
        sep_count = name.count(b":")
        if sep_count > 7:
            raise HeaderError("Can't handle read name with more than 8 segments: " + header.decode(errors="replace"))
        out += name
        out += b"+" if sep_count == 7 else b":"  # a UMI is already there: append to it
        out += b"+".join(umis)
    else:
        out += name
    out += b" "
    if comment is None:  # demux.rs:219-223: passing filter, non-control
        out += b"%d:N:0:" % read_num
    else:
        if not comment:  # `chars.last().unwrap()` (demux.rs:230) panics on a header that ends in its first space
            raise HeaderError("empty comment after the read name: " + header.decode(errors="replace"))
        sep_count = comment.count(b":")
        if sep_count < 3:  # demux.rs:228-233
            out += comment
            if comment[-1:] != b":":
                out += b":"
        else:
            if sep_count != 3:
                raise HeaderError("Comment in did not have 4 segments: " + header.decode(errors="replace"))
            first_colon = comment.index(b":")
            # Illumina may put a "0" in the index position of unmatched FASTQs (demux.rs:241-246)
            remainder = comment[first_colon + 1:-1] if comment[-1:].isdigit() else comment[first_colon + 1:]
            out += b"%d:" % read_num
            out += remainder
            if remainder[-1:] != b":":
                out += b"+"
    out += b"+".join(sample_barcode_segments)  # demux.rs:257-263
    return bytes(out)
