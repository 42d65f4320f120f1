"""The whole per-batch data path on the device — FASTQ text in, per-sample BGZF file images out:

    fqtk_b200_fastq_scan_device           records of every input chunk (ReadSetIterator::next, demux.rs:288-342)
    fqtk_b200_matcher_assign_fastq_device B segments gathered + encoded + matched (BarcodeMatcher::assign, :968)
    fqtk_b200_matcher_route_device        stable per-sample partition of the read indices (:970-975)
    fqtk_b200_demux_emit_device           every output record written, header rewritten (SampleWriters::write :396-415,
                                          write_header_internal :171-267), one contiguous run per (stream, sample)
    fqtk_b200_bgzf_compress_segments_device   every run deflated into its own BGZF members (pooled writers, :755-798)

Host-side mirror of the reference's main loop for one batch; torch only holds the device buffers.  The host sees the
chunks once on the way in and the compressed images once on the way out; nothing walks records on the host.
`fqtk_b200.fastq.demux_fastq_batch` is the same pipeline with the record formatting on the host (and is what the tests
compare this one with)."""
from __future__ import annotations

import ctypes as C
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Iterable, Iterator, Sequence

import numpy as np

from . import _lib
from .bgzf import BGZF_EOF
from .demux import FILE_TYPE_CODE, TooFewBases, min_length, parse_read_structure, too_few_bases_text
from .metrics import DemuxMetric, demux_metrics


@dataclass
class GpuDemuxResult:
    files: dict = field(default_factory=dict)   # "<prefix>.<code><n>.fq.gz" -> BGZF image (bytes) of this batch's records
    counts: np.ndarray | None = None            # S + 1, last = unmatched
    metrics: list[DemuxMetric] = field(default_factory=list)
    skipped: int = 0
    text_bytes: int = 0                         # uncompressed FASTQ bytes written on the device


class _PinnedTls(threading.local):  # one grow-only buffer per host thread (DemuxLanes runs one thread per lane)
    def __init__(self):
        self.d = {"ptr": None, "cap": 0}


_PINNED_TLS = _PinnedTls()


def _pinned_out(nbytes: int) -> np.ndarray:
    """A grow-only pinned host buffer for the compressed images (fqtk_b200_host_alloc): the D2H copy runs at full rate."""
    lib = _lib.lib()
    _PINNED = _PINNED_TLS.d
    if nbytes > _PINNED["cap"]:
        if _PINNED["ptr"]:
            lib.fqtk_b200_host_free(_PINNED["ptr"])
        p = C.c_void_p()
        cap = max(nbytes * 5 // 4, 1 << 20)
        _lib.check(lib.fqtk_b200_host_alloc(C.byref(p), cap))
        _PINNED["ptr"], _PINNED["cap"] = p, cap
    if nbytes == 0:
        return np.zeros(0, dtype=np.uint8)
    return np.ctypeslib.as_array(C.cast(_PINNED["ptr"], C.POINTER(C.c_uint8)), shape=(nbytes,))


def _cuda_memcpy(dst: int, src: int, nbytes: int, stream: int, to_device: bool) -> None:
    lib = _lib.lib()
    _lib.check((lib.fqtk_b200_copy_to_device if to_device else lib.fqtk_b200_copy_to_host)(dst, src, nbytes, stream or None))


def read_segments(structures) -> list[tuple[int, str, int, int]]:
    """(source, kind, offset, length | SEGMENT_REST) of every segment of the read structures, inputs in order."""
    out = []
    for s, st in enumerate(structures):
        pos = 0
        for kind, n in st:
            out.append((s, kind, pos, _lib.SEGMENT_REST if n is None else n))
            pos += 0 if n is None else n
    return out


def demux_fastq_batch_gpu(matcher, compressor, sample_ids: Sequence[str], barcodes: Sequence[str],
                          read_structures: Sequence[str], fastq_texts: Sequence[bytes], output_types: Sequence[str] = ("T",),
                          unmatched_prefix: str = "unmatched", level: int = 5, eof: bool = True,
                          skip_too_few_bases: bool = False, device: int = 0) -> GpuDemuxResult:
    import torch

    lib = _lib.lib()
    structures = [parse_read_structure(s) for s in read_structures]
    if len(structures) != len(fastq_texts):
        raise ValueError("The same number of read structures should be given as FASTQs")  # demux.rs:709-717
    dev = torch.device("cuda", device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    S = len(sample_ids)
    # 1. chunks to the device, records found there
    d_chunks, tables = [], []
    for text in fastq_texts:
        arr = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
        d = torch.empty(max(int(arr.size), 1), dtype=torch.uint8, device=dev)
        if arr.size:  # straight from the caller's buffer (pinned memory makes this a full-rate asynchronous copy)
            _cuda_memcpy(d.data_ptr(), arr.ctypes.data, int(arr.size), stream, to_device=True)
        cap = max(1, int(arr.size) // 7 + 1)
        d_head = torch.empty(cap, dtype=torch.int64, device=dev)
        d_seq = torch.empty(cap, dtype=torch.int64, device=dev)
        d_len = torch.empty(cap, dtype=torch.int32, device=dev)
        n_rec, used = C.c_uint64(), C.c_uint64()
        _lib.check(lib.fqtk_b200_fastq_scan_device(device, d.data_ptr(), int(arr.size), cap, d_head.data_ptr(), d_seq.data_ptr(),
                                                   d_len.data_ptr(), C.byref(n_rec), C.byref(used), stream))
        k = int(n_rec.value)
        d_chunks.append((d, int(arr.size)))
        tables.append([d_head[:k], d_seq[:k], d_len[:k]])
    n = tables[0][0].shape[0]
    if any(t[0].shape[0] != n for t in tables):
        raise ValueError("FASTQ sources out of sync")  # demux.rs:960-964
    # 2. the too-few-bases rule (demux.rs:298-315), on the device tables
    keep = torch.ones(n, dtype=torch.bool, device=dev)
    for st, t in zip(structures, tables):
        keep &= t[2] >= min_length(st)
    skipped = int(n - int(keep.sum().item()))
    if skipped:
        if not skip_too_few_bases:
            i = int(torch.nonzero(~keep)[0].item())
            st, t = next((st, t) for st, t in zip(structures, tables) if int(t[2][i].item()) < min_length(st))
            h0, s0 = int(tables[0][0][i].item()), int(tables[0][1][i].item())
            name = bytes(fastq_texts[0][h0 + 1:s0 - 1]).rstrip(b"\r").decode(errors="replace")
            raise TooFewBases(too_few_bases_text(name, int(t[2][i].item()), st))
        tables = [[x[keep].contiguous() for x in t] for t in tables]
        n = int(tables[0][0].shape[0])
    res = GpuDemuxResult(skipped=skipped)
    segs_all = read_segments(structures)
    d_off = torch.zeros(S + 2, dtype=torch.int64, device=dev)
    if n:
        # 3. match: B segments gathered by offset out of the chunks
        srcs = (_lib.FastqSource * len(tables))(*[_lib.FastqSource(d.data_ptr(), nb, t[1].data_ptr(), t[2].data_ptr())
                                                 for (d, nb), t in zip(d_chunks, tables)])
        bsegs = [(s, off, ln) for s, kind, off, ln in segs_all if kind == "B"]
        fsegs = (_lib.FastqSegment * len(bsegs))(*[_lib.FastqSegment(*x) for x in bsegs])
        d_res = torch.empty(n, dtype=torch.int32, device=dev)
        rc = lib.fqtk_b200_matcher_assign_fastq_device(matcher._h, srcs, len(tables), fsegs, len(bsegs), n, d_res.data_ptr(), stream)
        if rc != _lib.OK:
            from .barcode_matching import _raise

            _raise(rc, getattr(matcher, "_sample0_id", None))
        # 4. route
        d_order = torch.empty(n, dtype=torch.int32, device=dev)
        matcher.route_device(d_res.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr(), stream)
        # 5. records
        esrc = (_lib.EmitSource * len(tables))(*[_lib.EmitSource(d.data_ptr(), nb, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr())
                                                 for (d, nb), t in zip(d_chunks, tables)])
        rsegs = (_lib.ReadSegment * len(segs_all))(*[_lib.ReadSegment(s, ord(kind), off, ln) for s, kind, off, ln in segs_all])
        kinds = "".join(t.upper() for t in output_types).encode()
        ns = C.c_uint32()
        skinds = C.create_string_buffer(16)
        snums = (C.c_uint32 * 16)()
        _lib.check(lib.fqtk_b200_emit_streams(rsegs, len(segs_all), kinds, C.byref(ns), skinds, snums))
        n_streams = int(ns.value)
        in_bytes = sum(nb for _, nb in d_chunks)
        bc_bytes = sum(ln if ln != _lib.SEGMENT_REST else 0 for _, kind, _, ln in segs_all if kind in "BM") + 16
        cap = max(n_streams, 1) * (in_bytes + n * (bc_bytes + 16)) + 64
        d_text = torch.empty(cap, dtype=torch.uint8, device=dev)
        file_off = (C.c_uint64 * (max(n_streams, 1) * (S + 2)))()
        text_bytes = C.c_uint64()
        _lib.check(lib.fqtk_b200_demux_emit_device(device, esrc, len(tables), rsegs, len(segs_all), kinds, d_order.data_ptr(),
                                                   d_off.data_ptr(), S + 1, n, d_text.data_ptr(), cap, file_off, C.byref(text_bytes),
                                                   stream))
        res.text_bytes = int(text_bytes.value)
        # 6. every (stream, sample) run -> its own BGZF members
        seg_off = np.zeros(n_streams * (S + 1) + 1, dtype=np.uint64)
        for t in range(n_streams):
            for b in range(S + 1):
                seg_off[t * (S + 1) + b] = file_off[t * (S + 2) + b]
        seg_off[-1] = res.text_bytes
        n_seg = n_streams * (S + 1)
        out_cap = compressor.bound(res.text_bytes) + 31 * n_seg
        d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
        out_off = np.zeros(n_seg + 1, dtype=np.uint64)
        _lib.check(lib.fqtk_b200_bgzf_compress_segments_device(compressor._h, d_text.data_ptr(), seg_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                               n_seg, level, d_out.data_ptr(), out_cap,
                                                               out_off.ctypes.data_as(C.POINTER(C.c_uint64)), stream))
        image = _pinned_out(int(out_off[-1]))
        if image.size:
            _cuda_memcpy(image.ctypes.data, d_out.data_ptr(), int(image.size), stream, to_device=False)
            torch.cuda.current_stream(dev).synchronize()
        for t in range(n_streams):
            code = FILE_TYPE_CODE[skinds.raw[t:t + 1].decode()]
            for b in range(S + 1):
                k = t * (S + 1) + b
                if seg_off[k + 1] == seg_off[k]:
                    continue  # no record of this sample in the batch: the host pipeline has no entry either
                prefix = sample_ids[b] if b < S else unmatched_prefix
                body = image[int(out_off[k]):int(out_off[k + 1])].tobytes()
                res.files[f"{prefix}.{code}{int(snums[t])}.fq.gz"] = body + (BGZF_EOF if eof else b"")
    off = d_off.cpu().numpy().astype(np.uint64)
    res.counts = np.diff(off).astype(np.uint64)
    res.metrics = demux_metrics(list(sample_ids), list(barcodes), [int(c) for c in res.counts], unmatched_prefix)
    return res


def demux_chunks(matcher, compressor, sample_ids: Sequence[str], barcodes: Sequence[str], read_structures: Sequence[str],
                 fastq_texts: Sequence, output_types: Sequence[str] = ("T",), unmatched_prefix: str = "unmatched",
                 level: int = 5, eof: bool = True, max_reads: int | None = None, raw: bool = False):
    """The same batch through ONE C-ABI call (fqtk_b200_demux_chunks): chunks in host memory in, per-file BGZF bytes out.
    Returns (GpuDemuxResult, bytes consumed per input).  Reads that are too short for their read structure fail the call
    with the reference's text (there is no skip option in this form).  raw=True skips the per-file byte strings."""
    lib = _lib.lib()
    structures = [parse_read_structure(s) for s in read_structures]
    if len(structures) != len(fastq_texts):
        raise ValueError("The same number of read structures should be given as FASTQs")
    arrs = [np.frombuffer(t, dtype=np.uint8) if not isinstance(t, np.ndarray) else t for t in fastq_texts]
    S = len(sample_ids)
    segs_all = read_segments(structures)
    rsegs = (_lib.ReadSegment * len(segs_all))(*[_lib.ReadSegment(s, ord(kind), off, ln) for s, kind, off, ln in segs_all])
    kinds = "".join(t.upper() for t in output_types).encode()
    ns = C.c_uint32()
    skinds = C.create_string_buffer(16)
    snums = (C.c_uint32 * 16)()
    _lib.check(lib.fqtk_b200_emit_streams(rsegs, len(segs_all), kinds, C.byref(ns), skinds, snums))
    n_streams = int(ns.value)
    chunks = (_lib.FastqChunk * len(arrs))(*[_lib.FastqChunk(a.ctypes.data if a.size else None, a.size) for a in arrs])
    in_bytes = sum(int(a.size) for a in arrs)
    cap = compressor.bound(max(n_streams, 1) * (in_bytes * 2 + 64)) + 31 * max(n_streams, 1) * (S + 1)
    out = _pinned_out(cap)
    n_seg = n_streams * (S + 1)
    out_off = (C.c_uint64 * (n_seg + 1))()
    counts = (C.c_uint64 * (S + 1))()
    n_reads = C.c_uint64()
    consumed = (C.c_uint64 * len(arrs))()
    cap_reads = (1 << 32) - 1 if max_reads is None else int(max_reads)
    rc = lib.fqtk_b200_demux_chunks(matcher._h, compressor._h, chunks, len(arrs), rsegs, len(segs_all), kinds, level, cap_reads,
                                    out.ctypes.data, out.size, out_off, C.byref(ns), counts, C.byref(n_reads), consumed)
    if rc != _lib.OK:
        msg = _lib.last_error()
        if "had too few bases to demux" in msg:
            raise TooFewBases(msg)
        from .barcode_matching import _raise

        _raise(rc, getattr(matcher, "_sample0_id", None))
    if raw:  # what the C call returns, untouched: (pinned buffer view, run offsets, counts), consumed
        return (out, np.array(list(out_off), dtype=np.uint64), np.array(list(counts), dtype=np.uint64)), [int(c) for c in consumed]
    res = GpuDemuxResult()
    for t in range(n_streams):
        code = FILE_TYPE_CODE[skinds.raw[t:t + 1].decode()]
        for b in range(S + 1):
            k = t * (S + 1) + b
            if out_off[k + 1] == out_off[k]:
                continue
            prefix = sample_ids[b] if b < S else unmatched_prefix
            res.files[f"{prefix}.{code}{int(snums[t])}.fq.gz"] = out[int(out_off[k]):int(out_off[k + 1])].tobytes() + (BGZF_EOF if eof else b"")
    res.counts = np.array(list(counts), dtype=np.uint64)
    res.metrics = demux_metrics(list(sample_ids), list(barcodes), [int(c) for c in res.counts], unmatched_prefix)
    return res, [int(c) for c in consumed]


class DemuxLanes:
    """K (matcher, compressor) pairs on ONE device, one host thread each; batch i goes to lane i % K.

    The stages of one fqtk_b200_demux_chunks call run back to back (copy in, scan, match, route, records, BGZF, copy out),
    so one lane leaves the copy engines idle while its kernels run and the SMs idle while it copies.  Two lanes overlap
    the copy of one batch with the kernels of the other — the device-side form of the reference's reader thread, main loop
    and writer pool running concurrently (demux.rs:921-934, :755-798).  Results come back in batch order, which is what keeps
    every output file's records in input order (demux.rs:1505-1523); counts() is the sum over the lanes' matchers.
    A Rust host does the same with two matcher / compressor handles and two threads."""

    def __init__(self, make_lane, lanes: int = 2):
        # make_lane() -> (BarcodeMatcher, BgzfCompressor): called once per lane, on the caller's thread
        self.lanes = [make_lane() for _ in range(lanes)]
        self._pools = [ThreadPoolExecutor(max_workers=1) for _ in range(lanes)]

    def map(self, batches: Iterable[Sequence], sample_ids, barcodes, read_structures, output_types=("T",), raw: bool = False,
            depth: int | None = None, **kw) -> Iterator:
        """Yields demux_chunks(...) of every batch (a sequence of FASTQ texts, one per input), in batch order; at most
        `depth` (default: the number of lanes) batches are in flight."""
        depth = depth or len(self.lanes)
        pending = []
        for i, texts in enumerate(batches):
            m, z = self.lanes[i % len(self.lanes)]
            pending.append(self._pools[i % len(self.lanes)].submit(demux_chunks, m, z, sample_ids, barcodes, read_structures, texts,
                                                                   output_types, raw=raw, **kw))
            if len(pending) >= depth:
                yield pending.pop(0).result()
        while pending:
            yield pending.pop(0).result()

    def counts(self) -> np.ndarray:
        return np.sum([m.counts() for m, _ in self.lanes], axis=0)

    def close(self):
        for p in self._pools:
            p.shutdown(wait=True)
        for m, z in self.lanes:
            z.close()
            m.close()
        self.lanes = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
