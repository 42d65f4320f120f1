"""Deterministic synthetic workloads for BASELINE.json's configs: counter-based, identical on host
and device.  Workload generation only — nothing here matches reads."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass(frozen=True)
class Config:
    name: str
    read_structures: str
    n_samples: int
    barcode_len: int
    max_mismatches: int
    min_mismatch_delta: int
    n_reads: int
    min_distance: int = 3
    n_degenerate: int = 0
    cfg: int = 0

    @property
    def seed_panel(self) -> int:
        return 0xF07C0000 + self.cfg

    @property
    def seed_reads(self) -> int:
        return 0xBA5E0000 + self.cfg

    @property
    def words_per_read(self) -> int:
        return (self.barcode_len + 7) // 8

    @property
    def algorithmic_bytes_per_read(self) -> int:
        """B(L) = 4*ceil(L/8) + 4: packed 4-bit barcode words in, one result word out."""
        return 4 * self.words_per_read + 4


# BASELINE.json configs[0..4]
CONFIGS = {
    1: Config("cfg1 single-end 8B+T, 4 samples", "8B+T", 4, 8, 1, 2, 10_000, cfg=1),
    2: Config("cfg2 PE150 + I1 8bp, 96 samples", "+T +T 8B", 96, 8, 1, 2, 100_000_000, cfg=2),
    3: Config("cfg3 dual-index 8+8bp PE150, 384 samples", "8B 8B +T +T", 384, 16, 1, 2, 500_000_000, cfg=3),
    4: Config("cfg4 inline 16B134T, 1536 samples", "16B134T", 1536, 16, 2, 2, 1_000_000_000, cfg=4),
    5: Config("cfg5 dual 10+10bp IUPAC sheet, 6144 samples", "10B 10B +T +T", 6144, 20, 1, 2, 1_000_000_000,
              n_degenerate=2, cfg=5),
}


def panel(cfg: Config) -> np.ndarray:
    """(S, L) uint8 ASCII panel for a config."""
    out = np.empty((cfg.n_samples, cfg.barcode_len), dtype=np.uint8)
    _lib.check(_lib.lib().fqtk_b200_synth_panel(cfg.seed_panel, cfg.n_samples, cfg.barcode_len, cfg.min_distance,
                                                cfg.n_degenerate, out.ctypes.data))
    return out


def make_panel(seed: int, n_samples: int, barcode_len: int, min_distance: int = 3, n_degenerate: int = 0) -> np.ndarray:
    out = np.empty((n_samples, barcode_len), dtype=np.uint8)
    _lib.check(_lib.lib().fqtk_b200_synth_panel(seed, n_samples, barcode_len, min_distance, n_degenerate,
                                                out.ctypes.data))
    return out


def reads_host(panel_ascii: np.ndarray, seed: int, first: int, n: int) -> np.ndarray:
    """(n, L) uint8 ASCII reads [first, first + n) of the stream, generated on the host."""
    panel_ascii = np.ascontiguousarray(panel_ascii, dtype=np.uint8)
    S, L = panel_ascii.shape
    out = np.empty((n, L), dtype=np.uint8)
    _lib.check(_lib.lib().fqtk_b200_synth_reads_host(panel_ascii.ctypes.data, S, L, seed, first, n, out.ctypes.data))
    return out


def reads_device(panel_ascii: np.ndarray, seed: int, first: int, n: int, d_ascii: int = 0, d_packed: int = 0,
                 stream: int = 0) -> None:
    """Same stream generated straight into device memory (raw pointers; either may be 0)."""
    panel_ascii = np.ascontiguousarray(panel_ascii, dtype=np.uint8)
    S, L = panel_ascii.shape
    _lib.check(_lib.lib().fqtk_b200_synth_reads_device(panel_ascii.ctypes.data, S, L, seed, first, n,
                                                       d_ascii or None, d_packed or None, stream or None))


def pack_host(reads_ascii: np.ndarray) -> np.ndarray:
    """encode() of every row (mod.rs:49-61) with numpy, for tests: (n, L) ASCII -> (n, W) uint32."""
    reads_ascii = np.ascontiguousarray(reads_ascii, dtype=np.uint8)
    n, L = reads_ascii.shape
    lut = np.zeros(256, dtype=np.uint32)
    for ch, m in {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10, "K": 12, "V": 7,
                  "H": 11, "D": 13, "B": 14, "N": 15}.items():
        lut[ord(ch)] = m
        lut[ord(ch.lower())] = m
    lut[ord(".")] = 15
    W = (L + 7) // 8
    masks = lut[reads_ascii]
    out = np.zeros((n, W), dtype=np.uint32)
    for i in range(L):
        out[:, i // 8] |= masks[:, i] << np.uint32(4 * (i % 8))
    return out
