"""ctypes binding of the C ABI in include/fqtk_b200.h.  Loading fails loudly when the CUDA library has not been
built (``python -m fqtk_b200.build`` / ``__graft_entry__.build()``): there is no CPU fallback to hide behind."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FQTK_B200_LIB") or os.path.join(_HERE, "libfqtk_b200.so")  # env: A/B builds only

NONE = 0xFFFFFFFF
OK = 0
ERR_EMPTY_PANEL = -1
ERR_EMPTY_BARCODE = -2
ERR_LENGTH = -3
ERR_ARG = -4
ERR_CUDA = -5
ERR_UNSUPPORTED = -6
MODE_BRUTE = 1
MODE_TABLE = 2

# every symbol include/fqtk_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "fqtk_b200_matcher_create", "fqtk_b200_matcher_create_ex", "fqtk_b200_options_init", "fqtk_b200_matcher_destroy", "fqtk_b200_matcher_get_info",
    "fqtk_b200_set_table_budget", "fqtk_b200_set_cuckoo_arity", "fqtk_b200_matcher_assign", "fqtk_b200_matcher_assign_batch",
    "fqtk_b200_matcher_assign_segments", "fqtk_b200_matcher_assign_segments_device",
    "fqtk_b200_matcher_route_device", "fqtk_b200_matcher_route",
    "fqtk_b200_matcher_assign_packed_device", "fqtk_b200_matcher_assign_ascii_device", "fqtk_b200_pack_device",
    "fqtk_b200_encode_host", "fqtk_b200_matcher_counts", "fqtk_b200_matcher_counts_device",
    "fqtk_b200_matcher_reset_counts", "fqtk_b200_matcher_set_mode", "fqtk_b200_matcher_set_host_pack",
    "fqtk_b200_kernel_launches",
    "fqtk_b200_last_error", "fqtk_b200_device_count", "fqtk_b200_host_alloc", "fqtk_b200_host_free",
    "fqtk_b200_synth_panel", "fqtk_b200_synth_reads_host", "fqtk_b200_synth_reads_device",
    "fqtk_b200_matcher_assign_batch_packed", "fqtk_b200_pack_host", "fqtk_b200_copy_ceiling",
    "fqtk_b200_group_create", "fqtk_b200_group_destroy", "fqtk_b200_group_size", "fqtk_b200_group_device",
    "fqtk_b200_group_matcher", "fqtk_b200_group_shard", "fqtk_b200_group_assign_batch",
    "fqtk_b200_group_assign_batch_packed", "fqtk_b200_group_assign_packed_device", "fqtk_b200_group_counts",
    "fqtk_b200_group_reset_counts", "fqtk_b200_fastq_scan", "fqtk_b200_matcher_assign_fastq",
    "fqtk_b200_matcher_assign_fastq_device",
    "fqtk_b200_fastq_scan_device", "fqtk_b200_matcher_assign_fastq_chunks",
    "fqtk_b200_bgzf_create", "fqtk_b200_bgzf_destroy", "fqtk_b200_bgzf_chunk_bytes", "fqtk_b200_bgzf_bound",
    "fqtk_b200_bgzf_compress", "fqtk_b200_bgzf_compress_device", "fqtk_b200_bgzf_compress_segments_device",
    "fqtk_b200_emit_streams", "fqtk_b200_demux_emit_device", "fqtk_b200_copy_to_device", "fqtk_b200_copy_to_host",
    "fqtk_b200_demux_chunks",
]


class MatcherInfo(C.Structure):
    _fields_ = [
        ("n_samples", C.c_uint32), ("barcode_len", C.c_uint32), ("words_per_read", C.c_uint32),
        ("max_ns_in_barcodes", C.c_uint32), ("mode", C.c_uint32), ("device", C.c_uint32),
        ("table_entries", C.c_uint64), ("table_slots", C.c_uint64), ("table_bytes", C.c_uint64),
        ("table_candidates", C.c_uint64), ("tier_entries", C.c_uint64), ("tier_slots", C.c_uint64),
        ("cuckoo_entries", C.c_uint64), ("cuckoo_probes", C.c_uint32), ("cuckoo_slots", C.c_uint32),
        ("l2_table_entries", C.c_uint64), ("l2_table_bytes", C.c_uint64),
        ("l2_table_slow_keys", C.c_uint64),
    ]


class Options(C.Structure):
    """fqtk_b200_options"""
    _fields_ = [("struct_size", C.c_uint32), ("kernel", C.c_int32), ("table_budget", C.c_uint64),
                ("chunk_bytes", C.c_uint64), ("l2_table_load_pct", C.c_uint32), ("reserved", C.c_uint32)]


class Segment(C.Structure):
    """fqtk_b200_segment"""
    _fields_ = [("base", C.c_void_p), ("row_stride", C.c_uint64), ("offset", C.c_uint32), ("length", C.c_uint32)]


class FastqSource(C.Structure):
    """fqtk_b200_fastq_source"""
    _fields_ = [("chunk", C.c_void_p), ("chunk_bytes", C.c_uint64), ("seq_offsets", C.c_void_p), ("seq_lengths", C.c_void_p)]


class EmitSource(C.Structure):
    """fqtk_b200_emit_source"""
    _fields_ = [("d_chunk", C.c_void_p), ("chunk_bytes", C.c_uint64), ("d_head_offsets", C.c_void_p),
                ("d_seq_offsets", C.c_void_p), ("d_seq_lengths", C.c_void_p)]


class ReadSegment(C.Structure):
    """fqtk_b200_read_segment"""
    _fields_ = [("source", C.c_uint32), ("kind", C.c_uint32), ("offset", C.c_uint32), ("length", C.c_uint32)]


class FastqChunk(C.Structure):
    """fqtk_b200_fastq_chunk"""
    _fields_ = [("data", C.c_void_p), ("bytes", C.c_uint64)]


class FastqSegment(C.Structure):
    """fqtk_b200_fastq_segment"""
    _fields_ = [("source", C.c_uint32), ("offset", C.c_uint32), ("length", C.c_uint32)]


SEGMENT_REST = 0xFFFFFFFF


class Fqtk_b200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fqtk_b200 error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the sm_100a library first (python -m fqtk_b200.build). "
            "fqtk_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    sig = {
        "fqtk_b200_matcher_create": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, C.c_int, C.c_int, C.POINTER(vp)]),
        "fqtk_b200_matcher_create_ex": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(vp)]),
        "fqtk_b200_options_init": (None, [C.POINTER(Options)]),
        "fqtk_b200_matcher_destroy": (None, [vp]),
        "fqtk_b200_matcher_get_info": (C.c_int, [vp, C.POINTER(MatcherInfo)]),
        "fqtk_b200_set_table_budget": (None, [C.c_uint64]),
        "fqtk_b200_set_cuckoo_arity": (None, [C.c_int]),
        "fqtk_b200_matcher_assign": (C.c_int, [vp, C.c_char_p, C.c_size_t, u32p]),
        "fqtk_b200_matcher_assign_batch": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp]),
        "fqtk_b200_matcher_assign_segments": (C.c_int, [vp, C.POINTER(Segment), C.c_uint32, C.c_uint64, vp]),
        "fqtk_b200_matcher_assign_segments_device": (C.c_int, [vp, C.POINTER(Segment), C.c_uint32, C.c_uint64, vp, vp]),
        "fqtk_b200_matcher_route_device": (C.c_int, [vp, vp, C.c_uint64, vp, vp, vp]),
        "fqtk_b200_matcher_route": (C.c_int, [vp, vp, C.c_uint64, vp, vp]),
        "fqtk_b200_matcher_assign_packed_device": (C.c_int, [vp, vp, C.c_uint64, vp, vp]),
        "fqtk_b200_matcher_assign_ascii_device": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp]),
        "fqtk_b200_pack_device": (C.c_int, [vp, C.c_uint64, C.c_uint32, C.c_uint64, vp, vp]),
        "fqtk_b200_encode_host": (C.c_int, [C.c_char_p, C.c_size_t, u32p]),
        "fqtk_b200_matcher_counts": (C.c_int, [vp, vp]),
        "fqtk_b200_matcher_counts_device": (C.c_int, [vp, C.POINTER(vp)]),
        "fqtk_b200_matcher_reset_counts": (C.c_int, [vp]),
        "fqtk_b200_matcher_set_mode": (C.c_int, [vp, C.c_int]),
        "fqtk_b200_matcher_set_host_pack": (C.c_int, [vp, C.c_int]),
        "fqtk_b200_kernel_launches": (C.c_uint64, []),
        "fqtk_b200_last_error": (C.c_char_p, []),
        "fqtk_b200_device_count": (C.c_int, []),
        "fqtk_b200_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
        "fqtk_b200_host_free": (C.c_int, [vp]),
        "fqtk_b200_matcher_assign_batch_packed": (C.c_int, [vp, vp, C.c_uint64, vp, vp]),
        "fqtk_b200_pack_host": (C.c_int, [vp, C.c_uint64, C.c_uint32, C.c_uint64, vp, C.c_int]),
        "fqtk_b200_copy_ceiling": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
        "fqtk_b200_group_create": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, C.c_int, C.POINTER(C.c_int),
                                             C.c_uint32, C.POINTER(Options), C.POINTER(vp)]),
        "fqtk_b200_group_destroy": (None, [vp]),
        "fqtk_b200_group_size": (C.c_uint32, [vp]),
        "fqtk_b200_group_device": (C.c_int, [vp, C.c_uint32]),
        "fqtk_b200_group_matcher": (vp, [vp, C.c_uint32]),
        "fqtk_b200_group_shard": (None, [vp, C.c_uint64, C.c_uint32, u64p, u64p]),
        "fqtk_b200_group_assign_batch": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp]),
        "fqtk_b200_group_assign_batch_packed": (C.c_int, [vp, vp, C.c_uint64, vp, vp]),
        "fqtk_b200_group_assign_packed_device": (C.c_int, [vp, C.POINTER(vp), u64p, C.POINTER(vp), C.POINTER(vp)]),
        "fqtk_b200_group_counts": (C.c_int, [vp, vp]),
        "fqtk_b200_group_reset_counts": (C.c_int, [vp]),
        "fqtk_b200_fastq_scan": (C.c_int, [vp, C.c_uint64, C.c_uint64, vp, vp, vp, u64p, u64p]),
        "fqtk_b200_matcher_assign_fastq": (C.c_int, [vp, C.POINTER(FastqSource), C.c_uint32, C.POINTER(FastqSegment), C.c_uint32,
                                                     C.c_uint64, vp]),
        "fqtk_b200_matcher_assign_fastq_device": (C.c_int, [vp, C.POINTER(FastqSource), C.c_uint32, C.POINTER(FastqSegment),
                                                            C.c_uint32, C.c_uint64, vp, vp]),
        "fqtk_b200_fastq_scan_device": (C.c_int, [C.c_int, vp, C.c_uint64, C.c_uint64, vp, vp, vp, u64p, u64p, vp]),
        "fqtk_b200_matcher_assign_fastq_chunks": (C.c_int, [vp, C.POINTER(FastqChunk), C.c_uint32, C.POINTER(FastqSegment), C.c_uint32,
                                                            C.c_uint64, vp, u64p, u64p]),
        "fqtk_b200_bgzf_create": (C.c_int, [C.c_int, C.c_uint64, C.POINTER(vp)]),
        "fqtk_b200_bgzf_destroy": (None, [vp]),
        "fqtk_b200_bgzf_chunk_bytes": (C.c_uint64, [vp]),
        "fqtk_b200_bgzf_bound": (C.c_uint64, [C.c_uint64]),
        "fqtk_b200_bgzf_compress": (C.c_int, [vp, vp, C.c_uint64, C.c_int, C.c_int, vp, C.c_uint64, u64p]),
        "fqtk_b200_bgzf_compress_device": (C.c_int, [vp, vp, C.c_uint64, C.c_int, vp, C.c_uint64, vp, vp]),
        "fqtk_b200_bgzf_compress_segments_device": (C.c_int, [vp, vp, u64p, C.c_uint32, C.c_int, vp, C.c_uint64, u64p, vp]),
        "fqtk_b200_emit_streams": (C.c_int, [C.POINTER(ReadSegment), C.c_uint32, C.c_char_p, u32p, C.c_char_p, u32p]),
        "fqtk_b200_demux_emit_device": (C.c_int, [C.c_int, C.POINTER(EmitSource), C.c_uint32, C.POINTER(ReadSegment), C.c_uint32,
                                                  C.c_char_p, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_uint64, u64p, u64p, vp]),
        "fqtk_b200_demux_chunks": (C.c_int, [vp, vp, C.POINTER(FastqChunk), C.c_uint32, C.POINTER(ReadSegment), C.c_uint32, C.c_char_p,
                                             C.c_int, C.c_uint64, vp, C.c_uint64, u64p, u32p, u64p, u64p, u64p]),
        "fqtk_b200_copy_to_device": (C.c_int, [vp, vp, C.c_uint64, vp]),
        "fqtk_b200_copy_to_host": (C.c_int, [vp, vp, C.c_uint64, vp]),
        "fqtk_b200_synth_panel": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
        "fqtk_b200_synth_reads_host": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, vp]),
        "fqtk_b200_synth_reads_device": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def last_error() -> str:
    return (lib().fqtk_b200_last_error() or b"").decode(errors="replace")


def check(rc: int) -> None:
    if rc != OK:
        raise Fqtk_b200Error(rc, last_error())
