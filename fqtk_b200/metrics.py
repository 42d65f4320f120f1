"""`demux-metrics.txt` rows from the matcher's count table — host-side mirror of DemuxMetric
(src/bin/commands/demux.rs:452-497): integers come from the GPU (`BarcodeMatcher.counts()`), the f64 ratios are
derived here exactly as `DemuxMetric::update` does (ratios exclude the unmatched pseudo-sample from mean and best)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

HEADER = ("sample_id", "barcode", "templates", "frac_templates", "ratio_to_mean", "ratio_to_best")


@dataclass
class DemuxMetric:
    """demux.rs:452-465"""
    sample_id: str
    barcode: str
    templates: int = 0
    frac_templates: float = 0.0
    ratio_to_mean: float = 0.0
    ratio_to_best: float = 0.0


def _div(a: float, b: float) -> float:
    """f64 division with Rust's semantics (x/0 = inf, 0/0 = NaN) instead of Python's ZeroDivisionError."""
    if b == 0.0:
        return float("nan") if a == 0.0 else float("inf")
    return a / b


def demux_metrics(sample_ids: Sequence[str], barcodes: Sequence[str], counts: Sequence[int],
                  unmatched_prefix: str = "unmatched") -> list[DemuxMetric]:
    """counts: S + 1 integers, last = unmatched (demux.rs:970-974).  Returns S sample rows + the unmatched row last
    (barcode "."), with the derived fields of `DemuxMetric::update` (demux.rs:481-496)."""
    S = len(sample_ids)
    if len(barcodes) != S or len(counts) != S + 1:
        raise ValueError("need S sample ids, S barcodes and S + 1 counts")
    rows = [DemuxMetric(sid, bc, int(c)) for sid, bc, c in zip(sample_ids, barcodes, counts[:S])]
    unmatched = DemuxMetric(unmatched_prefix, ".", int(counts[S]))
    sample_total = float(sum(r.templates for r in rows))
    total = sample_total + float(unmatched.templates)
    mean = _div(sample_total, float(S))
    best = float(max((r.templates for r in rows), default=0))
    for r in rows + [unmatched]:
        r.frac_templates = _div(float(r.templates), total)
        r.ratio_to_mean = _div(float(r.templates), mean)
        r.ratio_to_best = _div(float(r.templates), best)
    return rows + [unmatched]


def write_tsv(path: str, metrics: Sequence[DemuxMetric]) -> None:
    """Tab-separated with the reference's column order (serde field order of DemuxMetric)."""
    with open(path, "w", encoding="utf-8") as fh:
        fh.write("\t".join(HEADER) + "\n")
        for m in metrics:
            fh.write(f"{m.sample_id}\t{m.barcode}\t{m.templates}\t{m.frac_templates!r}\t{m.ratio_to_mean!r}\t"
                     f"{m.ratio_to_best!r}\n")
