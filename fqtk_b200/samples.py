"""Minimal host-side mirror of the reference's sample sheet (src/lib/samples.rs:17-148): just enough to build the
panels the matcher needs, with the same four validations.  The real product keeps this in Rust (out of scope)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Sequence

from . import _lib

HEADER = "sample_id\tbarcode"  # Sample::deserialize_header_line, samples.rs:59-70


class SampleSheetError(ValueError):
    """Stands in for the panics / FgError of samples.rs."""


def is_valid_iupac(byte: int) -> bool:
    """src/lib/mod.rs:90-92 — uppercase IUPAC, 'U', or a no-call ('N', 'n', '.')."""
    return chr(byte) in "ACGTUMRWSYKVHDBNn."


@dataclass(frozen=True)
class Sample:
    """samples.rs:17-26"""
    sample_id: str
    barcode: str
    ordinal: int = 0

    @staticmethod
    def new(ordinal: int, name: str, barcode: str) -> "Sample":
        """samples.rs:49-57"""
        if not name:
            raise SampleSheetError("Sample name cannot be empty")
        if not barcode:
            raise SampleSheetError("Sample barcode cannot be empty")
        if not all(is_valid_iupac(b) for b in barcode.encode()):
            raise SampleSheetError(
                "All sample barcode bases must be one of A, C, G, T, U, R, Y, S, W, K, M, D, V, H, B, N")
        return Sample(name, barcode, ordinal)


@dataclass(frozen=True)
class SampleGroup:
    """samples.rs:80-148"""
    samples: tuple

    @staticmethod
    def from_samples(samples: Sequence[Sample]) -> "SampleGroup":
        """samples.rs:101-133"""
        if len(samples) == 0:
            raise SampleSheetError("Must provide one or more sample")
        if len({s.sample_id for s in samples}) != len(samples):
            raise SampleSheetError("Each sample name must be unique, duplicate identified")
        if len({s.barcode for s in samples}) != len(samples):
            raise SampleSheetError("Each sample barcode must be unique, duplicate identified")
        first = len(samples[0].barcode)
        if any(len(s.barcode) != first for s in samples):
            raise SampleSheetError("All barcodes must have the same length")
        return SampleGroup(tuple(Sample.new(i, s.sample_id, s.barcode) for i, s in enumerate(samples)))

    @staticmethod
    def from_file(path: str) -> "SampleGroup":
        """samples.rs:144-147 — tab-delimited, header `sample_id<TAB>barcode`, trailing empty lines ignored."""
        with open(path, "r", encoding="utf-8") as fh:
            lines = fh.read().split("\n")
        while lines and lines[-1] == "":
            lines.pop()
        if not lines or lines[0] != HEADER:
            found = lines[0] if lines else ""
            raise SampleSheetError(f"DelimFileHeaderError: expected {HEADER!r}, found {found!r}")
        rows = []
        for ln in lines[1:]:
            parts = ln.split("\t")
            if len(parts) != 2:
                raise SampleSheetError(f"malformed sample sheet line: {ln!r}")
            rows.append(Sample(parts[0], parts[1], 0))
        return SampleGroup.from_samples(rows)

    def barcodes(self) -> list[str]:
        return [s.barcode for s in self.samples]


def barcodes_of(samples: Iterable) -> list[bytes]:
    out = []
    for s in samples:
        if isinstance(s, Sample):
            out.append(s.barcode.encode())
        elif isinstance(s, str):
            out.append(s.encode())
        else:
            out.append(bytes(s))
    return out


__all__ = ["Sample", "SampleGroup", "SampleSheetError", "is_valid_iupac", "barcodes_of", "HEADER", "_lib"]
