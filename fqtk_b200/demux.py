"""Batched host pipeline around the GPU matcher — SURVEY 8f "next" #2: the reference's per-read loop
(src/bin/commands/demux.rs:945-977) restructured per BATCH: K read-sets in, one matcher call, one routing call, then
per-sample bulk appends in input order.  Host-side mirror (Python here, Rust in the product):

    read structures (read-structure 0.2.0 semantics: <len><kind>..., a trailing '+' = the rest, >= 1 base)
      -> ReadSetIterator::next          demux.rs:288-342   segment extraction, too-few-bases rule
      -> ReadSet::sample_barcode_sequence   :121-123       B segments of all inputs, in order  -> barcode rows
      -> BarcodeMatcher::assign (GPU, batched)  :968       fqtk_b200_matcher_assign_batch
      -> routing (GPU)                   :970-975          fqtk_b200_matcher_route: stable partition by sample
      -> SampleWriters::write            :396-415          per output type T, B, M, C: one record per segment,
         ReadSet::write_header           :161-267          header rewritten with read number, UMIs, sample barcode
      -> DemuxMetric                     :452-497          from the matcher's count table
      -> pooled BGZF writers             :755-798          write_bgzf_files: every output file's records as one text
                                                           buffer -> fqtk_b200_bgzf_compress (GPU) -> the .fq.gz image

No FASTQ / gzip IO lives here on purpose: records come in and go out as (head, seq, qual) byte triples."""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from .headers import write_header
from .metrics import DemuxMetric, demux_metrics

FILE_TYPE_CODE = {"T": "R", "B": "I", "M": "U", "C": "C"}  # demux.rs:674-679
OUTPUT_ORDER = ("T", "B", "M", "C")                         # the order SampleWriters::write walks its writers
Record = tuple  # (head: bytes without '@', seq: bytes, qual: bytes)


class ReadStructureError(ValueError):
    pass


class TooFewBases(ValueError):
    """demux.rs:309-315: the reference panics unless --skip-reasons too-few-bases was given."""


def parse_read_structure(text: str) -> list[tuple[str, int | None]]:
    """`8B+T`, `10M8B7C100T`, `100S3B` ... -> [(kind, length | None)]; only the last segment may be `+`."""
    segs = re.findall(r"(\d+|\+)([TBMSC])", text.upper())
    if not segs or "".join(a + b for a, b in segs) != text.upper():
        raise ReadStructureError(f"cannot parse read structure {text!r}")
    out: list[tuple[str, int | None]] = []
    for i, (length, kind) in enumerate(segs):
        if length == "+":
            if i != len(segs) - 1:
                raise ReadStructureError(f"only the last segment may be of variable length: {text!r}")
            out.append((kind, None))
        else:
            if int(length) == 0:
                raise ReadStructureError(f"zero-length segment in {text!r}")
            out.append((kind, int(length)))
    return out


def min_length(structure: Sequence[tuple[str, int | None]]) -> int:
    return sum(n if n is not None else 1 for _, n in structure)  # demux.rs:298


def structure_text(structure: Sequence[tuple[str, int | None]]) -> str:
    """The read structure as the reference prints it (`{}` of read_structure::ReadStructure): `8B`, `+T`, `10M8B7C100T`."""
    return "".join(("+" if n is None else str(n)) + kind for kind, n in structure)


def too_few_bases_text(name: str, have: int, structure) -> str:
    """The reference's panic text (demux.rs:307-313)."""
    return (f"Read {name} had too few bases to demux {have} vs. {min_length(structure)} needed in read structure "
            f"{structure_text(structure)}.")


def extract_segments(structure, seq: bytes, qual: bytes) -> list[tuple[str, bytes, bytes]]:
    """ReadSegment::extract_bases_and_quals for every segment (fixed ones by offset, `+` = the rest)."""
    out, pos = [], 0
    for kind, n in structure:
        end = len(seq) if n is None else pos + n
        out.append((kind, seq[pos:end], qual[pos:end]))
        pos = end
    return out


@dataclass
class DemuxResult:
    files: dict = field(default_factory=dict)     # "<prefix>.<code><n>.fq.gz" -> [Record, ...] in input order
    counts: np.ndarray | None = None              # THIS batch: S + 1, last = unmatched (skipped read-sets are counted nowhere)
    metrics: list[DemuxMetric] = field(default_factory=list)
    skipped: int = 0


def demux_batch(matcher, sample_ids: Sequence[str], barcodes: Sequence[str], read_structures: Sequence[str],
                inputs: Sequence[Sequence[Record]], output_types: Sequence[str] = ("T",),
                unmatched_prefix: str = "unmatched", skip_too_few_bases: bool = False) -> DemuxResult:
    """One batch through the pipeline.  `inputs[k][i]` = record i of input FASTQ k (all inputs in lock step)."""
    structures = [parse_read_structure(s) for s in read_structures]
    if len(structures) != len(inputs):
        raise ValueError("The same number of read structures should be given as FASTQs")  # demux.rs:709-717
    n = len(inputs[0])
    if any(len(x) != n for x in inputs):
        raise ValueError("FASTQ sources out of sync")  # demux.rs:960-964
    # segment extraction (ReadSetIterator::next) and the barcode rows
    read_sets = []  # (header of the first input's record, [segments of all inputs in order])
    for i in range(n):
        segments, skip = [], False
        for st, recs in zip(structures, inputs):
            head, seq, qual = recs[i]
            if len(seq) < min_length(st):
                if not skip_too_few_bases:
                    raise TooFewBases(too_few_bases_text(head.decode(errors="replace"), len(seq), st))
                skip = True
                break
            segments.extend(extract_segments(st, seq, qual))
        if skip:
            continue  # demux.rs:954-957: skipped read-sets are neither matched nor counted
        read_sets.append((inputs[0][i][0], segments))
    result = DemuxResult(skipped=n - len(read_sets))
    S = len(sample_ids)
    if read_sets:
        rows = [b"".join(s for k, s, _ in segs if k == "B") for _, segs in read_sets]
        stride = max(max(len(r) for r in rows), 1)
        mat = np.zeros((len(rows), stride), dtype=np.uint8)
        lens = np.array([len(r) for r in rows], dtype=np.uint32)
        for j, r in enumerate(rows):
            mat[j, :len(r)] = np.frombuffer(r, dtype=np.uint8)
        words = matcher.assign_batch(mat, lengths=lens)   # GPU: one call per batch
        order, offsets = matcher.route(words)             # GPU: stable partition of read indices by sample
    else:
        order, offsets = np.zeros(0, np.uint32), np.zeros(S + 2, np.uint64)
    # per-sample bulk appends, input order inside every run (SampleWriters::write)
    kinds = [t for t in OUTPUT_ORDER if t in {x.upper() for x in output_types}]
    for b in range(S + 1):
        prefix = sample_ids[b] if b < S else unmatched_prefix
        for j in order[int(offsets[b]):int(offsets[b + 1])]:
            header, segs = read_sets[int(j)]
            sample_bcs = [s for k, s, _ in segs if k == "B"]
            umis = [s for k, s, _ in segs if k == "M"]
            for kind in kinds:
                for idx, (_, s, q) in enumerate([x for x in segs if x[0] == kind]):
                    head = write_header(idx + 1, header, sample_bcs, umis)[1:]
                    result.files.setdefault(f"{prefix}.{FILE_TYPE_CODE[kind]}{idx + 1}.fq.gz", []).append((head, s, q))
    # this batch's table, from the routing offsets (the matcher's own counters are running totals over every call
    # since the last reset_counts, single assign() calls included)
    result.counts = np.diff(np.asarray(offsets, dtype=np.uint64)).astype(np.uint64)
    result.metrics = demux_metrics(list(sample_ids), list(barcodes), [int(c) for c in result.counts], unmatched_prefix)
    return result


def fastq_text(records) -> bytes:
    """SampleWriters::write (demux.rs:396-415): `@head \\n seq \\n+\\n quals \\n` per record."""
    return b"".join(b"@" + h + b"\n" + q_s + b"\n+\n" + q + b"\n" for h, q_s, q in records)


def write_bgzf_files(result: DemuxResult, compressor, level: int = 5, eof: bool = True) -> dict:
    """The batch's output files as BGZF images (what the reference's pooled writers leave on disk, demux.rs:755-798,
    compression level demux.rs:641-643): file name -> bytes.  `compressor` is a fqtk_b200.bgzf.BgzfCompressor; pass
    eof=False for every batch of a file but the last (a writer appends members and closes with the EOF block)."""
    return {name: compressor.compress(fastq_text(recs), level, eof) for name, recs in result.files.items()}
