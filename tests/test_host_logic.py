"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the header
declares, the host encoder agrees with the oracle, the synthetic generator is deterministic and replayable, the
sample-sheet mirror validates like the reference, and matching refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from fqtk_b200 import _lib, synth
from fqtk_b200.barcode_matching import BarcodeMatcher, encode
from fqtk_b200.samples import Sample, SampleGroup, SampleSheetError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fqtk_b200.h")).read()
    declared = set(re.findall(r"FQTK_B200_API\s+[\w\s\*]+?\b(fqtk_b200_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SYMBOLS)
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_product_does_not_link_or_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fqtk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "fqtk_oracle" not in text and "liboracle" not in text, f
    header = open(os.path.join(ROOT, "include", "fqtk_b200.h")).read()
    assert "oracle" not in header.lower()


def test_encode_host_matches_oracle():
    rng = np.random.default_rng(3)
    alphabet = np.frombuffer(b"ACGTUMRWSYKVHDBNacgtumrwsykvhdbn.-*X019 \x00\xff", dtype=np.uint8)
    for _ in range(500):
        n = int(rng.integers(0, 70))
        s = bytes(rng.choice(alphabet, n))
        want, _ = oracle.encode(s)
        assert encode(s) == want


def test_pack_host_matches_encode():
    rng = np.random.default_rng(4)
    for L in (1, 7, 8, 9, 16, 20, 33):
        reads = rng.choice(np.frombuffer(b"ACGTNacgtn.RYKM-", dtype=np.uint8), size=(50, L))
        packed = synth.pack_host(reads)
        for r, p in zip(reads, packed):
            assert encode(bytes(r)) == [int(x) for x in p]


@pytest.mark.parametrize("cfg_id", [1, 2, 3, 4])
def test_synth_panel_properties(cfg_id):
    cfg = synth.CONFIGS[cfg_id]
    p = synth.panel(cfg)
    assert p.shape == (cfg.n_samples, cfg.barcode_len)
    assert set(np.unique(p).tolist()) <= set(b"ACGT")
    assert np.array_equal(p, synth.panel(cfg)), "deterministic"
    # pairwise Hamming distance >= 3 (sampled for the big panel)
    rows = p if cfg.n_samples <= 400 else p[:400]
    d = (rows[:, None, :] != rows[None, :, :]).sum(-1)
    d[np.arange(len(rows)), np.arange(len(rows))] = 99
    assert d.min() >= cfg.min_distance


def test_synth_panel_cfg5_is_a_valid_degenerate_sheet():
    cfg = synth.CONFIGS[5]
    p = synth.panel(cfg)
    strings = [bytes(r).decode() for r in p]
    group = SampleGroup.from_samples([Sample(f"s{i}", b, i) for i, b in enumerate(strings)])  # samples.rs validations
    assert len(group.samples) == cfg.n_samples
    degenerate = np.isin(p, np.frombuffer(b"RYSWKMBDHVN", dtype=np.uint8)).sum(1)
    assert (degenerate == cfg.n_degenerate).all()


def test_synth_reads_are_counter_based():
    cfg = synth.CONFIGS[3]
    p = synth.panel(cfg)
    whole = synth.reads_host(p, cfg.seed_reads, 0, 5000)
    part = synth.reads_host(p, cfg.seed_reads, 1234, 777)
    assert np.array_equal(whole[1234:1234 + 777], part)
    assert set(np.unique(whole).tolist()) <= set(b"ACGTN")
    other = synth.reads_host(p, cfg.seed_reads + 1, 0, 5000)
    assert not np.array_equal(whole, other)


def test_synth_mix_matches_spec():
    """~90 % true / 8 % near-miss / 2 % random: judged through the oracle's view of the reads."""
    cfg = synth.CONFIGS[3]
    p = synth.panel(cfg)
    reads = synth.reads_host(p, cfg.seed_reads, 0, 40000)
    m = oracle.OracleMatcher([bytes(r) for r in p], cfg.max_mismatches, cfg.min_mismatch_delta)
    res, counts = m.assign_batch(reads)
    matched = float((res != oracle.NONE).mean())
    assert 0.85 < matched < 0.93
    exact = float(((res & 0xFF00) == 0)[res != oracle.NONE].mean())
    assert exact > 0.85
    assert int(counts.sum()) == reads.shape[0]
    per_sample = counts[:-1].astype(float)
    assert per_sample.min() > 0.4 * per_sample.mean()


def test_sample_sheet_validations(tmp_path):
    assert Sample.new(0, "s_1_example_name", "GATTANN").barcode == "GATTANN"  # samples.rs:205-211
    with pytest.raises(SampleSheetError, match="Sample name cannot be empty"):
        Sample.new(0, "", "ACGT")
    with pytest.raises(SampleSheetError, match="Sample barcode cannot be empty"):
        Sample.new(0, "s", "")
    with pytest.raises(SampleSheetError, match="All sample barcode bases"):
        Sample.new(0, "s", "ACGTX")
    with pytest.raises(SampleSheetError, match="All sample barcode bases"):
        Sample.new(0, "s", "acgt")  # lowercase is invalid in the sheet (mod.rs:121-124)
    with pytest.raises(SampleSheetError, match="Must provide one or more sample"):
        SampleGroup.from_samples([])
    with pytest.raises(SampleSheetError, match="Each sample name must be unique"):
        SampleGroup.from_samples([Sample("a", "ACGT"), Sample("a", "TTTT")])
    with pytest.raises(SampleSheetError, match="Each sample barcode must be unique"):
        SampleGroup.from_samples([Sample("a", "ACGT"), Sample("b", "ACGT")])
    with pytest.raises(SampleSheetError, match="All barcodes must have the same length"):
        SampleGroup.from_samples([Sample("a", "ACGT"), Sample("b", "ACGTA")])
    f = tmp_path / "sheet.tsv"
    f.write_text("sample_id\tbarcode\nsample1\tGATTACA\nsample2\tCATGCTA\n\n\n")  # samples.rs:160-199
    g = SampleGroup.from_file(str(f))
    assert [s.sample_id for s in g.samples] == ["sample1", "sample2"]
    assert g.barcodes() == ["GATTACA", "CATGCTA"]
    assert [s.ordinal for s in g.samples] == [0, 1]
    f.write_text("sample_id,barcode\nsample1,GATTACA\n")  # samples.rs:213-233
    with pytest.raises(SampleSheetError, match="DelimFileHeaderError"):
        SampleGroup.from_file(str(f))
    f.write_text("sample1\tGATTACA\nsample2\tCATGCTA\n")  # no header, samples.rs:238-255
    with pytest.raises(SampleSheetError, match="DelimFileHeaderError"):
        SampleGroup.from_file(str(f))


def test_matcher_argument_errors_without_touching_a_gpu():
    from fqtk_b200.barcode_matching import MatcherPanic

    with pytest.raises(MatcherPanic, match="Must provide at least one sample"):
        BarcodeMatcher([], 2, 1)
    with pytest.raises(MatcherPanic, match="Sample barcode cannot be empty string"):
        BarcodeMatcher([""], 2, 1)
    with pytest.raises(OverflowError):
        BarcodeMatcher(["ACGT"], 256, 1)  # demux.rs:923-924: u8::try_from


def test_no_cpu_fallback_when_no_gpu_is_visible():
    if _lib.lib().fqtk_b200_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(_lib.Fqtk_b200Error) as ei:
        BarcodeMatcher(["ACGT", "TTTT"], 1, 1)
    assert ei.value.code == _lib.ERR_CUDA
    assert "no CPU fallback" in ei.value.message


def test_demux_metrics_follow_the_reference_update_rule(tmp_path):
    """DemuxMetric::update (demux.rs:481-496): mean and best exclude the unmatched pseudo-sample; unmatched row last,
    barcode '.' (demux.rs:918).  The reference's own check (demux.rs:2058-2064): templates sum = 2, Sample0000 = 2."""
    from fqtk_b200.metrics import demux_metrics, write_tsv

    rows = demux_metrics(["Sample0000"], ["GATTGGG"], [2, 0])
    assert sum(r.templates for r in rows) == 2 and rows[0].templates == 2
    assert rows[-1].sample_id == "unmatched" and rows[-1].barcode == "."
    rows = demux_metrics(["a", "b", "c"], ["AAAA", "CCCC", "GGGG"], [30, 10, 20, 40])
    total, mean, best = 100.0, 20.0, 30.0
    for r, t in zip(rows, [30, 10, 20, 40]):
        assert r.templates == t
        assert r.frac_templates == t / total and r.ratio_to_mean == t / mean and r.ratio_to_best == t / best
    f = tmp_path / "demux-metrics.txt"
    write_tsv(str(f), rows)
    lines = f.read_text().splitlines()
    assert lines[0] == "sample_id\tbarcode\ttemplates\tfrac_templates\tratio_to_mean\tratio_to_best"
    assert lines[-1].split("\t")[:3] == ["unmatched", ".", "40"]
    # an empty run divides by zero the way f64 does in the reference (NaN), it does not raise
    rows = demux_metrics(["a"], ["AAAA"], [0, 0])
    assert rows[0].frac_templates != rows[0].frac_templates


def test_metrics_floats_print_like_the_reference_csv_writer(tmp_path):
    """serde -> csv -> ryu: shortest round-trip digits, exponent form without '+' / leading zeros below 1e-5 and from
    1e16 up, NaN / inf spelled the Rust way (Python's repr gives 1e-05 / 1e+16 / nan)."""
    from fqtk_b200.metrics import demux_metrics, format_f64, write_tsv

    want = {0.5: "0.5", 1.0: "1.0", 0.0: "0.0", 2 / 3: "0.6666666666666666", 1e-5: "0.00001", 1.5e-5: "0.000015",
            9.999e-6: "9.999e-6", 1e-7: "1e-7", 1.234e-10: "1.234e-10", 1e15: "1000000000000000.0", 1e16: "1e16",
            1.5e16: "1.5e16", 123456.789: "123456.789", float("inf"): "inf", float("-inf"): "-inf", -2.5e-9: "-2.5e-9"}
    for x, text in want.items():
        assert format_f64(x) == text, (x, format_f64(x), text)
        if x == x and abs(x) != float("inf"):
            assert float(format_f64(x)) == x  # round-trips
    assert format_f64(float("nan")) == "NaN"
    # a rare sample in a large run lands in the exponent range
    rows = demux_metrics(["rare", "big"], ["AAAA", "CCCC"], [3, 2_000_000_000, 0])
    f = tmp_path / "m.txt"
    write_tsv(str(f), rows)
    rare = f.read_text().splitlines()[1].split("\t")
    assert rare[3] == "1.49999999775e-9" and "e-0" not in rare[3] and "e+" not in "".join(rare)
    empty = demux_metrics(["a"], ["AAAA"], [0, 0])
    write_tsv(str(f), empty)
    assert f.read_text().splitlines()[1].split("\t")[3:] == ["NaN", "NaN", "NaN"]


def test_write_header_reference_kats(kats):
    """ReadSet::write_header_internal (demux.rs:171-267) against the reference's own six tests (:2084-2196)."""
    from fqtk_b200.headers import HeaderError, write_header

    for case in kats["write_header"]:
        args = (case["read_num"], case["header"].encode(), [b.encode() for b in case["sample_barcodes"]],
                [u.encode() for u in case["umis"]])
        if "expect_error" in case:
            with pytest.raises(HeaderError, match=case["expect_error"]):
                write_header(*args)
        else:
            assert write_header(*args).decode() == case["expect"], case["source"]
    # branches the reference tests do not reach, derived from the code: two UMI segments, a five-part comment,
    # a three-part comment that already ends in ':', a barcode kept when the comment does not end in a digit
    assert write_header(1, b"q1", [b"AC"], [b"GG", b"TT"]) == b"@q1:GG+TT 1:N:0:AC"
    with pytest.raises(HeaderError, match="4 segments"):
        write_header(1, b"q1 1:N:0:1:2", [b"AC"])
    assert write_header(1, b"q1 a:b:", [b"AC"]) == b"@q1 a:b:AC"
    assert write_header(2, b"q1 1:N:0:GATC", [b"AC"]) == b"@q1 2:N:0:GATC+AC"
    assert write_header(2, b"q1 1:N:0:", [b"AC"]) == b"@q1 2:N:0:AC"


def test_read_structures_and_segment_extraction():
    """Host-only parts of the batched pipeline (fqtk_b200/demux.py): read-structure parsing, minimum length
    (demux.rs:298), segment extraction with a trailing `+` segment."""
    from fqtk_b200.demux import ReadStructureError, extract_segments, min_length, parse_read_structure

    assert parse_read_structure("8B+T") == [("B", 8), ("T", None)]
    assert parse_read_structure("10m8b7c100t") == [("M", 10), ("B", 8), ("C", 7), ("T", 100)]
    assert min_length(parse_read_structure("8B+T")) == 9
    assert min_length(parse_read_structure("6B1S1M1T")) == 9
    for bad in ("", "8", "B8", "+B8T", "0B+T", "8X+T"):
        with pytest.raises(ReadStructureError):
            parse_read_structure(bad)
    segs = extract_segments(parse_read_structure("4B4M8S"), b"AAAACCCCGGGGTTTT", b"0123456789abcdef")
    assert segs == [("B", b"AAAA", b"0123"), ("M", b"CCCC", b"4567"), ("S", b"GGGGTTTT", b"89abcdef")]
    segs = extract_segments(parse_read_structure("2B+T"), b"ACGTA", b"!!###")
    assert segs == [("B", b"AC", b"!!"), ("T", b"GTA", b"###")]


class _OracleBackedMatcher:
    """Stand-in for the GPU matcher so that the HOST logic of fqtk_b200/demux.py runs in the CPU suite: assignments from
    the oracle, routing = numpy's stable argsort (what fqtk_b200_matcher_route computes on the device)."""

    def __init__(self, barcodes, max_mismatches, min_mismatch_delta):
        self.om = oracle.OracleMatcher([b.encode() for b in barcodes], max_mismatches, min_mismatch_delta)
        self.S = len(barcodes)
        self._counts = np.zeros(self.S + 1, dtype=np.uint64)

    def assign_batch(self, reads, lengths=None):
        out = np.empty(len(reads), dtype=np.uint32)
        for i in range(len(reads)):
            row = bytes(reads[i, :int(lengths[i])] if lengths is not None else reads[i])
            out[i] = self.om.assign_word(row)
            self._counts[self.S if out[i] == _lib.NONE else int(out[i]) >> 16] += 1
        return out

    def route(self, words):
        bucket = np.where(words == _lib.NONE, self.S, words >> 16).astype(np.int64)
        order = np.argsort(bucket, kind="stable").astype(np.uint32)
        offsets = np.zeros(self.S + 2, dtype=np.uint64)
        offsets[1:] = np.cumsum(np.bincount(bucket, minlength=self.S + 1))
        return order, offsets

    def counts(self):
        return self._counts.copy()


def test_batched_pipeline_host_logic_with_the_oracle_as_matcher(kats):
    """The reference's end-to-end demux vectors through fqtk_b200/demux.py with the oracle standing in for the GPU
    matcher (the GPU suite runs the same vectors through the C ABI)."""
    from fqtk_b200.demux import demux_batch

    for case in kats["demux_e2e"]:
        S = len(case["barcodes"])
        ids = [f"Sample{j:04d}" for j in range(S)]
        inputs = [[(f"ex_{i}".encode(), b.encode(), b";" * len(b)) for i, b in enumerate(col)] for col in case["inputs"]]
        m = _OracleBackedMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"])
        out = demux_batch(m, ids, case["barcodes"], case["read_structures"], inputs, case["output_types"])
        got = {name: [[h.decode(), s.decode()] for h, s, _ in recs] for name, recs in out.files.items()}
        assert got == case["expect"], case["source"]
        if "expect_counts" in case:
            assert out.counts.tolist() == case["expect_counts"]
            assert [r.templates for r in out.metrics] == case["expect_counts"]


def test_too_few_bases_panic_text_is_the_reference_text(kats):
    """demux.rs:1983-2020: a read shorter than its read structure needs panics with the read's name, both lengths AND the
    read structure as the reference prints it (demux.rs:307-313)."""
    from fqtk_b200.demux import TooFewBases, demux_batch, parse_read_structure, structure_text

    for text in ("+T", "7B", "10M8B7C100T", "4B4M8S", "17B20T20S20T20S20T", "6M+T"):
        assert structure_text(parse_read_structure(text)) == text
    for case in kats["demux_panics"]:
        ids = [f"Sample{j:04d}" for j in range(len(case["barcodes"]))]
        inputs = [[(f"ex_{i}".encode(), b.encode(), b";" * len(b)) for i, b in enumerate(col)] for col in case["inputs"]]
        m = _OracleBackedMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"])
        with pytest.raises(TooFewBases) as e:
            demux_batch(m, ids, case["barcodes"], case["read_structures"], inputs, case["output_types"])
        assert str(e.value) == case["expect_panic"], case["source"]


def test_pack_host_is_encode_for_every_row():
    """fqtk_b200_pack_host (two symbols per table lookup, several host threads) == encode() (mod.rs:49-61) of every row:
    checked against the numpy restatement and, per row, against the oracle's encode; all 256 byte values, odd and even L,
    row strides wider than L."""
    import oracle
    from fqtk_b200 import _lib, synth
    from fqtk_b200.barcode_matching import pack_host

    rng = np.random.default_rng(2024)
    for L in (1, 2, 7, 8, 9, 16, 17, 20, 31, 32, 33, 40, 254):
        n = 200_000 if L == 16 else 3000
        reads = rng.integers(0, 256, size=(n, L), dtype=np.uint8)
        reads[::3] = np.frombuffer(b"ACGTNacgtn.RYKMSWBDHVUu-*", dtype=np.uint8)[rng.integers(0, 25, size=reads[::3].shape)]
        got = pack_host(reads, threads=0 if L == 16 else 1)
        assert np.array_equal(got, synth.pack_host(reads)), L
        W = (L + 7) // 8
        for i in (0, 1, n - 1):
            blocks, n_symbols = oracle.encode(bytes(reads[i]))
            assert n_symbols == L and blocks == got[i].tolist(), (L, i)
        # a wider row stride: only the first L bytes of a row count
        wide = np.full((n, L + 5), ord("T"), dtype=np.uint8)
        wide[:, :L] = reads
        out = np.empty((n, W), dtype=np.uint32)
        _lib.check(_lib.lib().fqtk_b200_pack_host(wide.ctypes.data, n, L, L + 5, out.ctypes.data, 2))
        assert np.array_equal(out, got), L


def test_pack_host_stream_path_is_encode_for_every_byte_value():
    """Rows back to back with L a multiple of 8 take the AVX2 stream form (host_pack.cpp): A/C/G/T/N in either case by byte
    shuffles, any 32-byte step with another byte in it through the table.  Mostly-clean streams with sparse IUPAC codes,
    dots, lower case and arbitrary bytes (every value 0 .. 255 appears), sizes that leave scalar tails, 1 / 3 / all threads."""
    from fqtk_b200 import synth
    from fqtk_b200.barcode_matching import pack_host

    rng = np.random.default_rng(77)
    clean = np.frombuffer(b"ACGTNacgtn", dtype=np.uint8)
    for L in (8, 16, 24, 32, 40):
        for n in (1, 3, 5, 100_003):
            reads = clean[rng.integers(0, 5 if n > 5 else 10, size=(n, L))].copy()
            if n > 5:
                odd = rng.random(size=reads.shape) < 0.002
                reads[odd] = rng.integers(0, 256, size=int(odd.sum()), dtype=np.uint8)
                reads[7::997, L - 1] = np.arange(len(reads[7::997]), dtype=np.uint8)  # every byte value, last column
                reads[11::31, 0] = clean[5 + rng.integers(0, 5, size=len(reads[11::31]))]  # lower case
                reads[13::53, 3] = np.frombuffer(b".RYKMSWBDHVUu", dtype=np.uint8)[rng.integers(0, 13, size=len(reads[13::53]))]
            want = synth.pack_host(reads)
            for threads in (1, 3, 0):
                assert np.array_equal(pack_host(reads, threads=threads), want), (L, n, threads)


def test_fastq_scanner():
    """fqtk_b200_fastq_scan: offsets / lengths of every complete 4-line record, carry-over of a record cut by the chunk
    boundary, CRLF line ends, and the three malformed-record errors.  Host only."""
    from fqtk_b200 import _lib, fastq

    recs = [(f"r{i} 1:N:0:{i}", "ACGTN"[: 1 + i % 5] * (1 + i % 7), i) for i in range(1000)]
    text = "".join(f"@{h}\n{s}\n+\n{'I' * len(s)}\n" for h, s, _ in recs).encode()
    ix = fastq.scan(text)
    assert len(ix) == 1000 and ix.consumed == len(text)
    for i in (0, 1, 499, 999):
        assert ix.header(i).decode() == recs[i][0]
        assert ix.bases(i).decode() == recs[i][1] and ix.quals(i) == b"I" * len(recs[i][1])
        assert int(ix.seq_lengths[i]) == len(recs[i][1])
        assert text[int(ix.head_offsets[i])] == ord("@")
    # a chunk boundary in the middle of a record: only whole records are reported, `consumed` says where to resume
    for cut in (len(text) - 1, len(text) - 3, len(text) // 2, 5):
        part = fastq.scan(text[:cut])
        assert text[:part.consumed].count(b"\n") == 4 * len(part)
        rest = fastq.scan(text[part.consumed:])
        assert len(part) + len(rest) == 1000 and rest.consumed == len(text) - part.consumed
    # max_records stops early
    few = fastq.scan(text, max_records=10)
    assert len(few) == 10 and few.consumed == int(ix.head_offsets[10])
    # CRLF
    crlf = fastq.scan(b"@a x\r\nACGT\r\n+\r\n!!!!\r\n")
    assert len(crlf) == 1 and crlf.header(0) == b"a x" and crlf.bases(0) == b"ACGT" and crlf.quals(0) == b"!!!!"
    assert len(fastq.scan(b"")) == 0
    for bad, what in ((b"r\nAC\n+\n!!\n", "'@'"), (b"@r\nAC\n-\n!!\n", "'+'"), (b"@r\nAC\n+\n!\n", "lengths differ")):
        with pytest.raises(_lib.Fqtk_b200Error, match=re.escape(what)):
            fastq.scan(bad)


def test_barcode_segments_of_read_structures():
    from fqtk_b200 import _lib
    from fqtk_b200.demux import parse_read_structure
    from fqtk_b200.fastq import barcode_segments

    st = [parse_read_structure(x) for x in ("8B92T", "10M8B7C100T", "+T", "3T+B")]
    assert barcode_segments(st) == [(0, 0, 8), (1, 10, 8), (3, 3, _lib.SEGMENT_REST)]
