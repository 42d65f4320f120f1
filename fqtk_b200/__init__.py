"""fqtk_b200 — B200-native (sm_100a) core for fqtk's `demux` barcode-matcher hot path.

Host-side mirror of the reference's `fqtk_lib::barcode_matching` interface over the C ABI in
``include/fqtk_b200.h``; the compute lives in ``fqtk_b200/csrc`` (hand-written CUDA) and nowhere else.
"""
from .barcode_matching import BarcodeMatch, BarcodeMatcher, MatcherGroup, MatcherPanic  # noqa: F401
from .samples import Sample, SampleGroup  # noqa: F401

__all__ = ["BarcodeMatch", "BarcodeMatcher", "MatcherGroup", "MatcherPanic", "Sample", "SampleGroup"]
