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


def format_f64(x: float) -> str:
    """An f64 as the reference's csv writer prints it (serde -> csv -> ryu: shortest round-trip digits; exponent form
    `1e-5` / `1.5e16` below 1e-5 and from 1e16 up, without '+' or leading zeros; `NaN`, `inf`, `-inf`)."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    if x == 0.0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    r = repr(float(x))  # shortest round-trip digits, like ryu
    sign = "-" if r.startswith("-") else ""
    r = r.lstrip("-")
    if "e" in r:
        mant, exp = r.split("e")
        e10 = int(exp)
    else:
        mant, e10 = r, 0
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    if ip.strip("0"):
        e10 += len(ip.lstrip("0")) - 1
    else:
        e10 -= len(fp) - len(fp.lstrip("0")) + 1
    digits = digits.rstrip("0") or "0"
    if -5 <= e10 < 16:  # ryu's plain-decimal range
        if e10 >= 0:
            whole, frac = digits[: e10 + 1].ljust(e10 + 1, "0"), digits[e10 + 1:]
            return f"{sign}{whole}.{frac or '0'}"
        return f"{sign}0.{'0' * (-e10 - 1)}{digits}"
    mant = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
    return f"{sign}{mant}e{e10}"


def write_tsv(path: str, metrics: Sequence[DemuxMetric]) -> None:
    """Tab-separated with the reference's column order (serde field order of DemuxMetric)."""
    with open(path, "w", encoding="utf-8") as fh:
        fh.write("\t".join(HEADER) + "\n")
        for m in metrics:
            fh.write(f"{m.sample_id}\t{m.barcode}\t{m.templates}\t{format_f64(m.frac_templates)}\t"
                     f"{format_f64(m.ratio_to_mean)}\t{format_f64(m.ratio_to_best)}\n")
