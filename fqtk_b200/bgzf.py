"""Host-side mirror of the reference's pooled BGZF writers (src/bin/commands/demux.rs:755-798: `PoolBuilder::<_,
BgzfCompressor>`, `compression_level` demux.rs:641-643) over the C ABI's GPU compressor.  One `BgzfCompressor` per
device; `compress(data)` = what one writer would have put in its file for `data` (65 280-byte blocks, EOF block)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

BGZF_BLOCK_SIZE = 65280  # bgzf crate
BGZF_EOF = bytes([0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0,
                  0, 0, 0, 0, 0, 0, 0, 0])


class BgzfCompressor:
    def __init__(self, device: int = 0, chunk_bytes: int = 0):
        self._h = C.c_void_p()
        _lib.check(_lib.lib().fqtk_b200_bgzf_create(device, chunk_bytes, C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            _lib.lib().fqtk_b200_bgzf_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def chunk_bytes(self) -> int:
        return int(_lib.lib().fqtk_b200_bgzf_chunk_bytes(self._h))

    @staticmethod
    def bound(n: int) -> int:
        return int(_lib.lib().fqtk_b200_bgzf_bound(n))

    def compress(self, data, level: int = 5, eof: bool = True) -> bytes:
        """bytes-like in, the BGZF file image out."""
        src = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(self.bound(src.size), dtype=np.uint8)
        n = self.compress_into(src, out, level, eof)
        return out[:n].tobytes()

    def compress_into(self, src: np.ndarray, out: np.ndarray, level: int = 5, eof: bool = True) -> int:
        """numpy uint8 buffers (pinned ones make the copies asynchronous); returns the bytes written."""
        n_out = C.c_uint64(0)
        _lib.check(_lib.lib().fqtk_b200_bgzf_compress(self._h, src.ctypes.data, src.size, level, 1 if eof else 0,
                                                      out.ctypes.data, out.size, C.byref(n_out)))
        return int(n_out.value)

    def compress_device(self, d_in: int, n: int, d_out: int, out_capacity: int, d_out_bytes: int, level: int = 5,
                        stream: int = 0) -> None:
        _lib.check(_lib.lib().fqtk_b200_bgzf_compress_device(self._h, d_in, n, level, d_out, out_capacity, d_out_bytes,
                                                             stream or None))
