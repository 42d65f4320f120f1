"""Host-side mirror of the reference's matcher interface (src/lib/barcode_matching.rs:15-186) over the C ABI.

Same names, argument meaning and error behaviour as the reference so the parity tests read like its own tests:

    BarcodeMatcher(samples, max_mismatches, min_mismatch_delta, use_cache)   # ::new, :55-86
    matcher.assign(read_bases) -> Optional[BarcodeMatch]                     # ::assign, :165-186
    BarcodeMatch(best_match, best_mismatches, next_best_mismatches)          # :16-25

plus the batched forms the GPU needs (`assign_batch` on host arrays, `assign_packed_device` /
`assign_ascii_device` on raw device pointers) and the caller's per-sample counts (demux.rs:970-974).
Every call goes to the CUDA library; there is no Python or CPU matching path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .samples import barcodes_of


class MatcherPanic(Exception):
    """Stands in for a Rust panic on the reference path; `.args[0]` is the reference's panic text."""


@dataclass(frozen=True)
class BarcodeMatch:
    """barcode_matching.rs:16-25"""
    best_match: int
    best_mismatches: int
    next_best_mismatches: int


def unpack(word: int) -> Optional[BarcodeMatch]:
    word = int(word)
    if word == _lib.NONE:
        return None
    return BarcodeMatch(word >> 16, (word >> 8) & 0xFF, word & 0xFF)


def _raise(rc: int, sample0_id: Optional[str] = None) -> None:
    msg = _lib.last_error()
    if rc == _lib.ERR_LENGTH and sample0_id is not None and msg.endswith(" for sample 0"):
        # the C ABI does not know sample ids; the reference's text names the first sample (barcode_matching.rs:99-105)
        msg = msg[: -len("0")] + sample0_id
    if rc in (_lib.ERR_EMPTY_PANEL, _lib.ERR_EMPTY_BARCODE, _lib.ERR_LENGTH):
        raise MatcherPanic(msg)
    raise _lib.Fqtk_b200Error(rc, msg)


class BarcodeMatcher:
    """BarcodeMatcher (barcode_matching.rs:29-186).  `samples`: Sample objects, str or bytes barcodes, in sheet order."""

    def __init__(self, samples: Sequence, max_mismatches: int, min_mismatch_delta: int, use_cache: bool = True,
                 device: int = 0, **options):
        """`options`: fields of fqtk_b200_options (kernel, table_budget, chunk_bytes, l2_table_load_pct) for this handle
        only; without any the thread's defaults apply (fqtk_b200_matcher_create)."""
        if not (0 <= max_mismatches <= 255 and 0 <= min_mismatch_delta <= 255):
            raise OverflowError("max_mismatches / min_mismatch_delta must fit in u8 (demux.rs:923-924)")
        bcs = barcodes_of(samples)
        if len(bcs) == 0:
            raise MatcherPanic("Must provide at least one sample")
        if any(len(b) == 0 for b in bcs):
            raise MatcherPanic("Sample barcode cannot be empty string")
        L = len(bcs[0])
        if any(len(b) != L for b in bcs):
            # the reference only notices at match time (count_mismatches panics, :95-106); a dense panel cannot hold it
            raise MatcherPanic("All barcodes must have the same length")
        self.n_samples = len(bcs)
        self._sample0_id = getattr(samples[0], "sample_id", None)
        self.barcode_len = L
        self.max_mismatches = max_mismatches
        self.min_mismatch_delta = min_mismatch_delta
        self.use_cache = bool(use_cache)
        panel = np.frombuffer(b"".join(bcs), dtype=np.uint8)
        self._h = C.c_void_p()
        if options:
            opts = _lib.Options()
            _lib.lib().fqtk_b200_options_init(C.byref(opts))
            for name, value in options.items():
                if name not in ("kernel", "table_budget", "chunk_bytes", "l2_table_load_pct"):
                    raise TypeError(f"unknown matcher option {name!r}")
                setattr(opts, name, int(value))
            rc = _lib.lib().fqtk_b200_matcher_create_ex(panel.ctypes.data, self.n_samples, L, max_mismatches,
                                                        min_mismatch_delta, int(self.use_cache), device,
                                                        C.byref(opts), C.byref(self._h))
        else:
            rc = _lib.lib().fqtk_b200_matcher_create(panel.ctypes.data, self.n_samples, L, max_mismatches,
                                                     min_mismatch_delta, int(self.use_cache), device, C.byref(self._h))
        if rc != _lib.OK:
            self._h = None
            _raise(rc)

    # -- lifecycle ----------------------------------------------------------------------------------
    def close(self) -> None:
        h = getattr(self, "_h", None)
        if h:
            _lib.lib().fqtk_b200_matcher_destroy(h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- introspection ------------------------------------------------------------------------------
    def info(self) -> _lib.MatcherInfo:
        info = _lib.MatcherInfo()
        _lib.check(_lib.lib().fqtk_b200_matcher_get_info(self._h, C.byref(info)))
        return info

    @property
    def mode(self) -> str:
        return {_lib.MODE_BRUTE: "brute", _lib.MODE_TABLE: "table"}[self.info().mode]

    def set_mode(self, mode: str) -> None:
        _lib.check(_lib.lib().fqtk_b200_matcher_set_mode(self._h, {"brute": _lib.MODE_BRUTE, "table": _lib.MODE_TABLE}[mode]))

    def set_host_pack(self, threads: int) -> None:
        """assign_batch with encode() done by `threads` host threads while the batch is in flight (-1 = auto, 0 = off): the
        packed words, half the bytes of the ASCII rows, cross PCIe (fqtk_b200_matcher_set_host_pack)."""
        _lib.check(_lib.lib().fqtk_b200_matcher_set_host_pack(self._h, int(threads)))

    @property
    def words_per_read(self) -> int:
        return (self.barcode_len + 7) // 8

    # -- the reference's call ------------------------------------------------------------------------
    def assign_word(self, read_bases: bytes) -> int:
        out = C.c_uint32()
        rb = bytes(read_bases)
        rc = _lib.lib().fqtk_b200_matcher_assign(self._h, rb, len(rb), C.byref(out))
        if rc != _lib.OK:
            _raise(rc, self._sample0_id)
        return int(out.value)

    def assign(self, read_bases: bytes) -> Optional[BarcodeMatch]:
        """BarcodeMatcher::assign, barcode_matching.rs:165-186."""
        return unpack(self.assign_word(read_bases))

    # -- batched forms ----------------------------------------------------------------------------------
    def assign_batch(self, reads: np.ndarray, lengths: Optional[np.ndarray] = None,
                     out: Optional[np.ndarray] = None) -> np.ndarray:
        """reads: (N, stride) uint8 host array, row i = one read's concatenated sample-barcode bases
        (demux.rs:121-123).  Returns uint32[N] result words."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        assert reads.ndim == 2
        n, stride = reads.shape
        if out is None:
            out = np.empty(n, dtype=np.uint32)
        assert out.dtype == np.uint32 and out.shape == (n,) and out.flags.c_contiguous
        lp = None
        if lengths is not None:
            lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
            assert lengths.shape == (n,)
            lp = lengths.ctypes.data
        rc = _lib.lib().fqtk_b200_matcher_assign_batch(self._h, reads.ctypes.data, n, stride, lp, out.ctypes.data)
        if rc != _lib.OK:
            _raise(rc, self._sample0_id)
        return out

    def assign_batch_ptr(self, rows_ptr: int, n: int, stride: int, results_ptr: int, lengths_ptr: int = 0) -> None:
        """Same call on raw host pointers (e.g. pinned buffers from fqtk_b200_host_alloc / torch pin_memory)."""
        rc = _lib.lib().fqtk_b200_matcher_assign_batch(self._h, rows_ptr, n, stride, lengths_ptr or None, results_ptr)
        if rc != _lib.OK:
            _raise(rc)

    def assign_batch_packed(self, packed: np.ndarray, want_words: bool = True, want_index: bool = False):
        """The same batch call for reads the host already holds in BitEnc form (`pack_host`): (N, W) uint32 in, result
        words (uint32[N]) and / or bare sample indices (uint16[N], 0xFFFF = None) out."""
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        assert packed.ndim == 2 and packed.shape[1] == self.words_per_read
        n = packed.shape[0]
        words = np.empty(n, dtype=np.uint32) if want_words else None
        index = np.empty(n, dtype=np.uint16) if want_index else None
        rc = _lib.lib().fqtk_b200_matcher_assign_batch_packed(
            self._h, packed.ctypes.data, n, words.ctypes.data if want_words else None,
            index.ctypes.data if want_index else None)
        if rc != _lib.OK:
            _raise(rc)
        return words, index

    def assign_segments(self, segments, n: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Barcodes that arrive in pieces (demux.rs:121-123): `segments` = [(rows, offset, length), ...] where `rows` is
        a C-contiguous (n, stride) uint8 host array; the pieces are gathered + encoded on the device, in order."""
        segs = (_lib.Segment * len(segments))()
        keep = []
        for k, (rows, offset, length) in enumerate(segments):
            rows = np.ascontiguousarray(rows, dtype=np.uint8)
            assert rows.ndim == 2 and rows.shape[0] >= n
            keep.append(rows)
            segs[k] = _lib.Segment(rows.ctypes.data, rows.shape[1], offset, length)
        if out is None:
            out = np.empty(n, dtype=np.uint32)
        rc = _lib.lib().fqtk_b200_matcher_assign_segments(self._h, segs, len(segments), n, out.ctypes.data)
        if rc != _lib.OK:
            _raise(rc)
        return out

    def assign_segments_device(self, segments, n: int, d_results: int, stream: int = 0) -> None:
        """Device form: `segments` = [(device_ptr, row_stride, offset, length), ...]."""
        segs = (_lib.Segment * len(segments))()
        for k, (ptr, stride, offset, length) in enumerate(segments):
            segs[k] = _lib.Segment(ptr, stride, offset, length)
        rc = _lib.lib().fqtk_b200_matcher_assign_segments_device(self._h, segs, len(segments), n, d_results,
                                                                 stream or None)
        if rc != _lib.OK:
            _raise(rc)

    def assign_packed_device(self, d_packed: int, n: int, d_results: int, stream: int = 0) -> None:
        """HBM-resident batch: raw device pointers (packed BitEnc words in, result words out), async on `stream`."""
        rc = _lib.lib().fqtk_b200_matcher_assign_packed_device(self._h, d_packed, n, d_results, stream or None)
        if rc != _lib.OK:
            _raise(rc)

    def assign_ascii_device(self, d_ascii: int, n: int, stride: int, d_results: int, d_lengths: int = 0,
                            stream: int = 0) -> None:
        rc = _lib.lib().fqtk_b200_matcher_assign_ascii_device(self._h, d_ascii, n, stride, d_lengths or None,
                                                              d_results, stream or None)
        if rc != _lib.OK:
            _raise(rc)

    # -- per-sample routing of a batch (demux.rs:970-975: every sample's output keeps input order) --------------
    def route(self, results: np.ndarray):
        """Stable partition of read indices by assignment.  Returns (order uint32[n], offsets uint64[S+2]):
        order[offsets[j]:offsets[j+1]] = ascending indices of the reads of sample j (j = S: unmatched)."""
        results = np.ascontiguousarray(results, dtype=np.uint32)
        n = results.shape[0]
        order = np.empty(n, dtype=np.uint32)
        offsets = np.zeros(self.n_samples + 2, dtype=np.uint64)
        rc = _lib.lib().fqtk_b200_matcher_route(self._h, results.ctypes.data, n, order.ctypes.data, offsets.ctypes.data)
        if rc != _lib.OK:
            _raise(rc)
        return order, offsets

    def route_device(self, d_results: int, n: int, d_order: int, d_offsets: int, stream: int = 0) -> None:
        rc = _lib.lib().fqtk_b200_matcher_route_device(self._h, d_results, n, d_order, d_offsets, stream or None)
        if rc != _lib.OK:
            _raise(rc)

    # -- the caller's counters (demux.rs:970-974) ------------------------------------------------------
    def counts(self) -> np.ndarray:
        """uint64[S + 1]; last element = unmatched."""
        out = np.zeros(self.n_samples + 1, dtype=np.uint64)
        _lib.check(_lib.lib().fqtk_b200_matcher_counts(self._h, out.ctypes.data))
        return out

    def counts_device_ptr(self) -> int:
        p = C.c_void_p()
        _lib.check(_lib.lib().fqtk_b200_matcher_counts_device(self._h, C.byref(p)))
        return int(p.value)

    def reset_counts(self) -> None:
        _lib.check(_lib.lib().fqtk_b200_matcher_reset_counts(self._h))


class MatcherGroup:
    """One matcher per GPU of the box behind one handle (fqtk_b200_group_*): contiguous shards of every batch, ONE count
    table (demux.rs:921-926, 970-975).  `devices` None = every visible device."""

    def __init__(self, samples: Sequence, max_mismatches: int, min_mismatch_delta: int, use_cache: bool = True,
                 devices: Optional[Sequence[int]] = None, **options):
        bcs = barcodes_of(samples)
        if len(bcs) == 0:
            raise MatcherPanic("Must provide at least one sample")
        L = len(bcs[0])
        if any(len(b) != L for b in bcs):
            raise MatcherPanic("All barcodes must have the same length")
        self.n_samples, self.barcode_len = len(bcs), L
        self._sample0_id = getattr(samples[0], "sample_id", None)
        panel = np.frombuffer(b"".join(bcs), dtype=np.uint8)
        opts = _lib.Options()
        _lib.lib().fqtk_b200_options_init(C.byref(opts))
        for name, value in options.items():
            if name not in ("kernel", "table_budget", "chunk_bytes", "l2_table_load_pct"):
                raise TypeError(f"unknown matcher option {name!r}")
            setattr(opts, name, int(value))
        devs = (C.c_int * len(devices))(*devices) if devices else None
        self._h = C.c_void_p()
        rc = _lib.lib().fqtk_b200_group_create(panel.ctypes.data, self.n_samples, L, max_mismatches, min_mismatch_delta,
                                               int(bool(use_cache)), devs, len(devices) if devices else 0,
                                               C.byref(opts), C.byref(self._h))
        if rc != _lib.OK:
            self._h = None
            _raise(rc)

    def close(self) -> None:
        h = getattr(self, "_h", None)
        if h:
            _lib.lib().fqtk_b200_group_destroy(h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def size(self) -> int:
        return int(_lib.lib().fqtk_b200_group_size(self._h))

    @property
    def devices(self) -> list:
        return [int(_lib.lib().fqtk_b200_group_device(self._h, k)) for k in range(self.size)]

    def shard(self, n_reads: int, k: int):
        first, count = C.c_uint64(), C.c_uint64()
        _lib.lib().fqtk_b200_group_shard(self._h, n_reads, k, C.byref(first), C.byref(count))
        return int(first.value), int(count.value)

    def assign_batch(self, reads: np.ndarray, lengths: Optional[np.ndarray] = None) -> np.ndarray:
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n, stride = reads.shape
        out = np.empty(n, dtype=np.uint32)
        lp = None
        if lengths is not None:
            lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
            lp = lengths.ctypes.data
        rc = _lib.lib().fqtk_b200_group_assign_batch(self._h, reads.ctypes.data, n, stride, lp, out.ctypes.data)
        if rc != _lib.OK:
            _raise(rc, self._sample0_id)
        return out

    def assign_batch_ptr(self, rows_ptr: int, n: int, stride: int, results_ptr: int) -> None:
        rc = _lib.lib().fqtk_b200_group_assign_batch(self._h, rows_ptr, n, stride, None, results_ptr)
        if rc != _lib.OK:
            _raise(rc, self._sample0_id)

    def assign_batch_packed(self, packed: np.ndarray, want_words: bool = True, want_index: bool = False):
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        n = packed.shape[0]
        words = np.empty(n, dtype=np.uint32) if want_words else None
        index = np.empty(n, dtype=np.uint16) if want_index else None
        rc = _lib.lib().fqtk_b200_group_assign_batch_packed(
            self._h, packed.ctypes.data, n, words.ctypes.data if want_words else None,
            index.ctypes.data if want_index else None)
        if rc != _lib.OK:
            _raise(rc)
        return words, index

    def assign_batch_packed_ptr(self, packed_ptr: int, n: int, results_ptr: int = 0, index_ptr: int = 0) -> None:
        rc = _lib.lib().fqtk_b200_group_assign_batch_packed(self._h, packed_ptr, n, results_ptr or None, index_ptr or None)
        if rc != _lib.OK:
            _raise(rc)

    def assign_packed_device(self, d_packed: Sequence[int], n_reads: Sequence[int], d_results: Sequence[int],
                             streams: Optional[Sequence[int]] = None) -> None:
        G = self.size
        assert len(d_packed) == len(n_reads) == len(d_results) == G
        pk = (C.c_void_p * G)(*d_packed)
        rs = (C.c_void_p * G)(*d_results)
        ns = (C.c_uint64 * G)(*n_reads)
        st = (C.c_void_p * G)(*streams) if streams else None
        rc = _lib.lib().fqtk_b200_group_assign_packed_device(self._h, pk, ns, rs, st)
        if rc != _lib.OK:
            _raise(rc)

    def counts(self) -> np.ndarray:
        out = np.zeros(self.n_samples + 1, dtype=np.uint64)
        _lib.check(_lib.lib().fqtk_b200_group_counts(self._h, out.ctypes.data))
        return out

    def reset_counts(self) -> None:
        _lib.check(_lib.lib().fqtk_b200_group_reset_counts(self._h))


def pack_host(reads: np.ndarray, threads: int = 0) -> np.ndarray:
    """encode() of every row on the host (fqtk_b200_pack_host): (N, stride >= L) uint8 -> (N, W) uint32 BitEnc words."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    n, L = reads.shape
    out = np.empty((n, (L + 7) // 8), dtype=np.uint32)
    _lib.check(_lib.lib().fqtk_b200_pack_host(reads.ctypes.data, n, L, L, out.ctypes.data, threads))
    return out


def encode(bases: bytes) -> list[int]:
    """encode(), src/lib/mod.rs:49-61 -> u32 blocks of the width-4 BitEnc layout."""
    rb = bytes(bases)
    out = (C.c_uint32 * max(1, (len(rb) + 7) // 8))()
    _lib.check(_lib.lib().fqtk_b200_encode_host(rb, len(rb), out))
    return [int(out[i]) for i in range((len(rb) + 7) // 8)]


def kernel_launches() -> int:
    return int(_lib.lib().fqtk_b200_kernel_launches())
