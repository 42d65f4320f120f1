"""CPU oracle for the fqtk demux matcher path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``fqtk_b200``) never does.

ctypes bindings over ``oracle/liboracle.so`` (built from ``oracle/fqtk_oracle.c`` by ``oracle/Makefile``);
see ``fqtk_oracle.h`` for the reference citations of every function.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

NONE = 0xFFFFFFFF
OK = 0
ERR_EMPTY_PANEL = -1
ERR_EMPTY_BARCODE = -2
ERR_LENGTH = -3
ERR_ARG = -4


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc).  Returns the path."""
    srcs = [os.path.join(_HERE, f) for f in ("fqtk_oracle.c", "fqtk_synth.c", "fqtk_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "clean", "all"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.fqo_iupac_mask.restype = C.c_uint8
    L.fqo_iupac_mask.argtypes = [C.c_uint8]
    L.fqo_byte_is_nocall.restype = C.c_int
    L.fqo_byte_is_nocall.argtypes = [C.c_uint8]
    L.fqo_is_valid_iupac.restype = C.c_int
    L.fqo_is_valid_iupac.argtypes = [C.c_uint8]
    L.fqo_encode.restype = C.c_int
    L.fqo_encode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    L.fqo_decode.restype = C.c_int
    L.fqo_decode.argtypes = [C.c_void_p, C.c_char_p]
    L.fqo_hamming.restype = C.c_uint32
    L.fqo_hamming.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.fqo_count_mismatches.restype = C.c_int
    L.fqo_count_mismatches.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_uint8, C.c_char_p]
    L.fqo_matcher_new.restype = C.c_int
    L.fqo_matcher_new.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, C.c_int, C.POINTER(C.c_void_p)]
    L.fqo_matcher_free.restype = None
    L.fqo_matcher_free.argtypes = [C.c_void_p]
    L.fqo_matcher_max_ns.restype = C.c_uint32
    L.fqo_matcher_max_ns.argtypes = [C.c_void_p]
    L.fqo_matcher_cache_len.restype = C.c_uint64
    L.fqo_matcher_cache_len.argtypes = [C.c_void_p]
    for name in ("fqo_assign", "fqo_assign_internal", "fqo_assign_closed"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, u32p]
    L.fqo_synth_reads.restype = None
    L.fqo_synth_reads.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    L.fqo_synth_panel.restype = C.c_int
    L.fqo_synth_panel.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.fqo_assign_batch.restype = C.c_int
    L.fqo_assign_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
    L.fqo_assign_batch_mt.restype = C.c_int
    L.fqo_assign_batch_mt.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, C.c_int,
                                      C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
    L.fqo_max_threads.restype = C.c_int
    L.fqo_max_threads.argtypes = []
    _lib = L
    return L


class _BitEnc(C.Structure):
    _fields_ = [("blk", C.c_uint32 * 32), ("len", C.c_uint32)]


class OraclePanic(Exception):
    """Stands in for a Rust panic on the reference path."""


@dataclass(frozen=True)
class BarcodeMatch:
    """barcode_matching.rs:16-25"""
    best_match: int
    best_mismatches: int
    next_best_mismatches: int


def unpack(word: int) -> Optional[BarcodeMatch]:
    word = int(word)
    if word == NONE:
        return None
    return BarcodeMatch(word >> 16, (word >> 8) & 0xFF, word & 0xFF)


def encode(bases: bytes) -> tuple[list[int], int]:
    """mod.rs:49-61 -> (u32 blocks, nr_symbols)"""
    e = _BitEnc()
    rc = lib().fqo_encode(bases, len(bases), C.byref(e))
    if rc != OK:
        raise ValueError("encode: too long")
    return [int(e.blk[i]) for i in range((e.len + 7) // 8)], int(e.len)


def decode_blocks(blocks: Sequence[int], n: int) -> str:
    """mod.rs:68-82"""
    e = _BitEnc()
    for i, b in enumerate(blocks):
        e.blk[i] = b
    e.len = n
    out = C.create_string_buffer(n + 1)
    if lib().fqo_decode(C.byref(e), out) != 0:
        raise OraclePanic("Invalid bit mask for base")
    return out.value.decode()


def hamming_nibbles(a: Sequence[int], b: Sequence[int], max_mismatches: int) -> int:
    """bitenc.rs:432-459 on two sequences of 4-bit values (push order)."""
    ea, eb = _BitEnc(), _BitEnc()
    for e, vals in ((ea, a), (eb, b)):
        e.len = len(vals)
        for i, v in enumerate(vals):
            e.blk[i // 8] |= (v & 0xF) << (4 * (i % 8))
    r = lib().fqo_hamming(C.byref(ea), C.byref(eb), max_mismatches)
    if r == 0xFFFFFFFF:
        raise OraclePanic("Both bitenc sequences must have the same length")
    return int(r)


def count_mismatches(observed: bytes, expected: bytes, sample_id: str = "sample_0", max_mismatches: int = 255) -> int:
    """barcode_matching.rs:89-110"""
    msg = C.create_string_buffer(512)
    rc = lib().fqo_count_mismatches(observed, len(observed), expected, len(expected), sample_id.encode(),
                                    max_mismatches, msg)
    if rc == ERR_LENGTH:
        raise OraclePanic(msg.value.decode())
    if rc < 0:
        raise ValueError(rc)
    return rc


class OracleMatcher:
    """BarcodeMatcher (barcode_matching.rs:29-186), literal restatement with the closed form beside it."""

    def __init__(self, barcodes: Sequence[bytes | str], max_mismatches: int, min_mismatch_delta: int,
                 use_cache: bool = True):
        bcs = [b.encode() if isinstance(b, str) else bytes(b) for b in barcodes]
        if len(bcs) == 0:
            raise OraclePanic("Must provide at least one sample")
        if any(len(b) == 0 for b in bcs):
            raise OraclePanic("Sample barcode cannot be empty string")
        self.S = len(bcs)
        self.L = len(bcs[0])
        if any(len(b) != self.L for b in bcs):
            raise ValueError("oracle takes a dense S x L panel")
        self.panel = np.frombuffer(b"".join(bcs), dtype=np.uint8).copy()
        self._h = C.c_void_p()
        rc = lib().fqo_matcher_new(self.panel.ctypes.data, self.S, self.L, max_mismatches, min_mismatch_delta,
                                   int(use_cache), C.byref(self._h))
        if rc != OK:
            raise OraclePanic(f"fqo_matcher_new rc={rc}")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().fqo_matcher_free(h)
            self._h = None

    @property
    def max_ns_in_barcodes(self) -> int:
        return int(lib().fqo_matcher_max_ns(self._h))

    @property
    def cache_len(self) -> int:
        return int(lib().fqo_matcher_cache_len(self._h))

    def _call(self, fn, read: bytes) -> int:
        out = C.c_uint32()
        rc = fn(self._h, read, len(read), C.byref(out))
        if rc == ERR_LENGTH:
            raise OraclePanic("Read barcode length differs from expected barcode length")
        if rc != OK:
            raise ValueError(rc)
        return int(out.value)

    def assign_word(self, read: bytes) -> int:
        return self._call(lib().fqo_assign, read)

    def assign(self, read: bytes) -> Optional[BarcodeMatch]:
        return unpack(self.assign_word(read))

    def assign_internal(self, read: bytes) -> Optional[BarcodeMatch]:
        return unpack(self._call(lib().fqo_assign_internal, read))

    def assign_closed(self, read: bytes) -> Optional[BarcodeMatch]:
        return unpack(self._call(lib().fqo_assign_closed, read))

    def assign_batch(self, reads: np.ndarray, mode: int = 0, want_results: bool = True):
        """reads: (N, L) uint8.  Returns (results uint32[N] | None, counts uint64[S+1])."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        assert reads.ndim == 2 and reads.shape[1] == self.L
        n = reads.shape[0]
        res = np.empty(n, dtype=np.uint32) if want_results else None
        counts = np.zeros(self.S + 1, dtype=np.uint64)
        rc = lib().fqo_assign_batch(self._h, reads.ctypes.data, n, res.ctypes.data if want_results else None,
                                    counts.ctypes.data, mode)
        if rc != OK:
            raise OraclePanic(f"assign_batch rc={rc}")
        return res, counts


def assign_batch_mt(barcodes_panel: np.ndarray, max_mismatches: int, min_mismatch_delta: int, reads: np.ndarray,
                    threads: int = 0, use_cache: bool = True, want_results: bool = True):
    """All-host-cores upper bound (NOT what the reference does).  Returns (results, counts, threads_used)."""
    panel = np.ascontiguousarray(barcodes_panel, dtype=np.uint8)
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    S, L = panel.shape
    n = reads.shape[0]
    res = np.empty(n, dtype=np.uint32) if want_results else None
    counts = np.zeros(S + 1, dtype=np.uint64)
    used = lib().fqo_assign_batch_mt(panel.ctypes.data, S, L, max_mismatches, min_mismatch_delta, int(use_cache),
                                     reads.ctypes.data, n, res.ctypes.data if want_results else None,
                                     counts.ctypes.data, threads)
    if used < 0:
        raise OraclePanic(f"assign_batch_mt rc={used}")
    return res, counts, used


def synth_panel(seed: int, n_samples: int, barcode_len: int, min_distance: int = 3, n_degenerate: int = 0) -> np.ndarray:
    """(S, L) uint8 ASCII panel of the synthetic workload (fqtk_synth.c) — the same bytes fqtk_b200.synth.make_panel gives."""
    out = np.empty((n_samples, barcode_len), dtype=np.uint8)
    if lib().fqo_synth_panel(seed, n_samples, barcode_len, min_distance, n_degenerate, out.ctypes.data) != 0:
        raise ValueError("no panel with that many samples at that distance")
    return out


def synth_reads(panel: np.ndarray, seed: int, first: int, n: int) -> np.ndarray:
    """(n, L) uint8 ASCII reads [first, first + n) of the synthetic stream — the same bytes fqtk_b200.synth.reads_host gives."""
    panel = np.ascontiguousarray(panel, dtype=np.uint8)
    S, L = panel.shape
    out = np.empty((n, L), dtype=np.uint8)
    lib().fqo_synth_reads(panel.ctypes.data, S, L, seed, first, n, out.ctypes.data)
    return out


def max_threads() -> int:
    return int(lib().fqo_max_threads())
