"""SURVEY 8f "next" #4 tail: BGZF output compression (src/bin/commands/demux.rs:755-798, pooled BGZF writers).
CPU part: the oracle's restatement of the framing against Python's own gzip reader.  GPU part: the device compressor
through the C ABI — every member must inflate to its 65 280-byte piece (strict parser), same member boundaries as the
oracle, EOF block, stored fallback, multi-chunk pipeline, determinism."""
import ctypes as C
import gzip
import struct
import zlib

import numpy as np
import pytest

import oracle.bgzf as ob
from fqtk_b200 import _lib


def fastq_text(n_records: int, seed: int = 7, read_len: int = 150) -> bytes:
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    quals = np.frombuffer(b"F:,#", dtype=np.uint8)
    seqs = bases[rng.integers(0, 4, size=(n_records, read_len))]
    seqs[rng.random((n_records, read_len)) < 0.002] = ord("N")
    q = quals[np.minimum(3, rng.geometric(0.75, size=(n_records, read_len)) - 1)]
    out = []
    for i in range(n_records):
        out.append(b"@A00123:45:HXXXXXX:1:%d:%d:%d 1:N:0:ACGTACGT+TTGCAATC\n" % (1101 + i // 9000, 1000 + (i * 37) % 30000, 1000 + (i * 91) % 35000))
        out.append(seqs[i].tobytes() + b"\n+\n" + q[i].tobytes() + b"\n")
    return b"".join(out)


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_members_are_what_gzip_reads():
    data = fastq_text(700)
    for level in (0, 1, 5, 9):
        img = ob.compress(data, level)
        assert gzip.decompress(img) == data
        payload, sizes = ob.parse(img)
        assert payload == data
        assert sizes[:-1] == [ob.BGZF_BLOCK_SIZE] * (len(data) // ob.BGZF_BLOCK_SIZE) + [len(data) % ob.BGZF_BLOCK_SIZE]
        assert sizes[-1] == 0 and img.endswith(ob.BGZF_EOF)


def test_oracle_eof_block_is_the_spec_constant_and_parse_is_strict():
    assert len(ob.BGZF_EOF) == 28 and gzip.decompress(ob.BGZF_EOF) == b""
    assert ob.parse(ob.BGZF_EOF) == (b"", [0])
    assert ob.compress(b"") == ob.BGZF_EOF  # a writer that saw no record leaves only the EOF block
    img = bytearray(ob.compress(b"hello world" * 100))
    img[-40] ^= 1  # inside the first member's trailer / data
    with pytest.raises((ValueError, zlib.error)):
        ob.parse(bytes(img))
    with pytest.raises(ValueError):
        ob.parse(ob.compress(b"abc")[:-3])


def test_bound_and_symbols_without_a_gpu():
    L = _lib.lib()
    for n in (0, 1, 65280, 65281, 10 ** 9):
        blocks = (n + 65279) // 65280
        assert L.fqtk_b200_bgzf_bound(n) == n + blocks * 31 + 28
    if L.fqtk_b200_device_count() == 0:
        h = C.c_void_p()
        assert L.fqtk_b200_bgzf_create(0, 0, C.byref(h)) == _lib.ERR_CUDA  # no CPU compressor behind the ABI


# ------------------------------------------------------------------------------------------------ GPU
def _cases():
    rng = np.random.default_rng(11)
    fq = fastq_text(1500)
    yield "one_byte", b"A"
    yield "tiny", b"@r\nACGT\n+\nFFFF\n"
    yield "fastq_3_blocks", fq[:3 * 65280 + 17]
    yield "exactly_one_block", fq[:65280]
    yield "one_block_plus_one", fq[:65281]
    yield "zeros", bytes(200_000)
    yield "random_incompressible", rng.integers(0, 256, size=150_000, dtype=np.uint8).tobytes()
    yield "all_byte_values", bytes(range(256)) * 600
    yield "long_runs", b"".join(bytes([65 + i % 5]) * (1 + (i * 7919) % 700) for i in range(600))
    yield "period_3", b"ACG" * 50_000
    yield "period_300", (fq[:300]) * 700
    yield "two_symbols", bytes(rng.integers(0, 2, size=100_000, dtype=np.uint8) + 65)
    yield "short_tail", fq[:65280 + 3]


@pytest.mark.gpu
@pytest.mark.parametrize("name,data", list(_cases()), ids=[c[0] for c in _cases()])
def test_gpu_members_inflate_to_their_pieces(name, data):
    from fqtk_b200.bgzf import BgzfCompressor
    with BgzfCompressor(0, chunk_bytes=4 * 65280) as z:
        assert z.chunk_bytes == 4 * 65280
        for level in (5, 0):
            img = z.compress(data, level)
            payload, sizes = ob.parse(img)  # strict framing, CRC32, ISIZE, stream end
            assert payload == data
            _, want_sizes = ob.parse(ob.compress(data, level))
            assert sizes == want_sizes  # same members as the reference's writer: one per 65 280 bytes, then EOF
            assert img.endswith(ob.BGZF_EOF) and gzip.decompress(img) == data
            assert len(img) <= z.bound(len(data))
            if level == 0:
                assert len(img) == len(data) + (len(sizes) - 1) * 31 + 28  # stored blocks
        assert z.compress(data, 5) == z.compress(data, 5)  # deterministic bytes
        no_eof = z.compress(data, 5, eof=False)
        assert no_eof + ob.BGZF_EOF == z.compress(data, 5)


@pytest.mark.gpu
def test_gpu_empty_input_and_arguments():
    from fqtk_b200.bgzf import BgzfCompressor
    with BgzfCompressor(0) as z:
        assert z.compress(b"") == ob.BGZF_EOF
        assert z.compress(b"", eof=False) == b""
        with pytest.raises(_lib.Fqtk_b200Error):
            z.compress(b"abc", level=13)
        src = np.frombuffer(fastq_text(300), dtype=np.uint8)
        with pytest.raises(_lib.Fqtk_b200Error):
            z.compress_into(src, np.empty(100, dtype=np.uint8))  # output too small: an error, not a truncation


@pytest.mark.gpu
def test_gpu_fastq_ratio_and_xfl_levels():
    """Compression quality on FASTQ text: within 25 % of zlib level 5 (the reference's default level, libdeflate there)."""
    from fqtk_b200.bgzf import BgzfCompressor
    data = fastq_text(20_000, seed=3)
    ref5 = len(ob.compress(data, 5))
    with BgzfCompressor(0) as z:
        img = z.compress(data, 5)
        assert ob.parse(img)[0] == data
        assert len(img) <= 1.25 * ref5, (len(img), ref5, len(data))
        assert z.compress(data, 1)[8] == 4 and z.compress(data, 9)[8] == 2 and img[8] == 0  # XFL as the bgzf crate sets it


@pytest.mark.gpu
def test_gpu_multi_chunk_pipeline_at_size_and_device_entry():
    """64 MiB chunks, 300 MB of FASTQ text through the host entry (several chunks, two streams), and one chunk through
    the device entry; the image is checked member by member."""
    import torch
    from fqtk_b200.bgzf import BgzfCompressor
    unit = fastq_text(30_000, seed=5)
    data = unit * (300_000_000 // len(unit))
    with BgzfCompressor(0) as z:
        img = z.compress(data, 5)
        payload, sizes = ob.parse(img)
        assert payload == data and sizes[:-1].count(65280) == len(data) // 65280
        n = min(len(data), z.chunk_bytes)
        d_in = torch.frombuffer(bytearray(data[:n]), dtype=torch.uint8).cuda()
        d_out = torch.empty(z.bound(n), dtype=torch.uint8, device="cuda")
        d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
        z.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), d_n.data_ptr(), 5,
                          torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = d_out[:int(d_n.item())].cpu().numpy().tobytes()
        assert ob.parse(got)[0] == data[:n]
        assert got == img[:len(got)]  # the same members the host entry produced for the first chunk


@pytest.mark.gpu
def test_gpu_demux_outputs_as_bgzf_files(tmp_path):
    """The reference's end-to-end demux vectors with the output side attached: batched pipeline -> per-sample records ->
    GPU BGZF images; the files read back with Python's gzip reader hold exactly the expected records (the reference's tests
    read their .fq.gz outputs back the same way, demux.rs:1293-1333 ...)."""
    import json
    import os

    from fqtk_b200 import BarcodeMatcher
    from fqtk_b200.bgzf import BgzfCompressor
    from fqtk_b200.demux import demux_batch, write_bgzf_files

    kats = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
    with BgzfCompressor(0, chunk_bytes=65280 * 8) as z:
        for case in kats["demux_e2e"]:
            S = len(case["barcodes"])
            ids = [f"Sample{j:04d}" for j in range(S)]
            inputs = [[(f"ex_{i}".encode(), b.encode(), b";" * len(b)) for i, b in enumerate(col)] for col in case["inputs"]]
            with BarcodeMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"]) as m:
                out = demux_batch(m, ids, case["barcodes"], case["read_structures"], inputs, case["output_types"])
            files = write_bgzf_files(out, z)
            assert set(files) == set(case["expect"])
            for name, image in files.items():
                path = tmp_path / name
                path.write_bytes(image)
                with gzip.open(path, "rb") as fh:
                    lines = fh.read().split(b"\n")
                assert lines[-1] == b""
                recs = [[lines[k][1:].decode(), lines[k + 1].decode()] for k in range(0, len(lines) - 1, 4)]
                assert recs == case["expect"][name], (case["source"], name)
                assert all(lines[k + 2] == b"+" and lines[k + 3] == b";" * len(lines[k + 1]) for k in range(0, len(lines) - 1, 4))
                assert image.endswith(ob.BGZF_EOF)


@pytest.mark.gpu
def test_gpu_fuzz_structured_inputs():
    """400 generated inputs — every size from 1 to 40 bytes, sizes around one and two blocks, alphabets of 1 ... 256 symbols,
    copies at distances 1 ... 9 000 (inside and across the 4 080-byte parts a warp parses), long runs, matches of exactly
    3 / 4 / 258 / 259 bytes, Fibonacci-skewed symbol counts (code lengths above the 15-bit limit before the repair) — all
    through the strict member parser."""
    from fqtk_b200.bgzf import BgzfCompressor
    rng = np.random.default_rng(2024)
    cases = [bytes(rng.integers(65, 69, size=n, dtype=np.uint8)) for n in range(1, 41)]
    for n in (65279, 65280, 65281, 2 * 65280 - 1, 2 * 65280, 2 * 65280 + 1, 4079, 4080, 4081, 8160, 8161):
        cases.append(bytes(rng.integers(0, 256, size=n, dtype=np.uint8) & 0x0F))
    for _ in range(200):
        n = int(rng.integers(1, 140_000))
        alpha = int(rng.choice([1, 2, 3, 4, 5, 16, 64, 256]))
        buf = bytearray(rng.integers(0, alpha, size=n, dtype=np.uint8).tobytes())
        for _ in range(int(rng.integers(0, 60))):  # plant copies
            ln = int(rng.choice([3, 4, 5, 17, 100, 257, 258, 259, 300, 1000, 5000]))
            dist = int(rng.choice([1, 2, 3, 7, 31, 32, 33, 100, 1000, 4079, 4080, 4081, 9000]))
            if n <= dist + ln:
                continue
            dst = int(rng.integers(dist, n - ln))
            for k in range(ln):  # byte by byte: overlapping copies behave like deflate's
                buf[dst + k] = buf[dst + k - dist]
        cases.append(bytes(buf))
    fib = [1, 1]
    while sum(fib) < 60_000:
        fib.append(fib[-1] + fib[-2])
    cases.append(b"".join(bytes([i]) * c for i, c in enumerate(fib)))  # skewed: forces the length-limit repair
    skew = np.concatenate([np.full(c, i, dtype=np.uint8) for i, c in enumerate(fib)])
    rng.shuffle(skew)
    cases.append(skew.tobytes())
    for _ in range(140):
        n = int(rng.integers(1, 70_000))
        run = rng.geometric(0.05, size=n // 4 + 1)
        vals = rng.integers(33, 75, size=run.size, dtype=np.uint8)
        cases.append(np.repeat(vals, run)[:n].tobytes())
    with BgzfCompressor(0, chunk_bytes=3 * 65280) as z:
        for i, data in enumerate(cases):
            img = z.compress(data, 5)
            payload, sizes = ob.parse(img)
            assert payload == data, (i, len(data))
            assert sizes[:-1] == [65280] * (len(data) // 65280) + ([len(data) % 65280] if len(data) % 65280 else [])
            assert len(img) <= len(data) + 31 * (len(sizes) - 1) + 28  # never larger than stored blocks


@pytest.mark.gpu
def test_gpu_whole_data_path_equals_the_host_formatted_pipeline():
    """FASTQ text in -> per-sample BGZF images out with everything on the device (scan, match, route, record writing with
    header rewrite, BGZF per sample) against the host-formatted pipeline (fqtk_b200.fastq.demux_fastq_batch, itself pinned by
    the reference's end-to-end tests): same files, same bytes after inflation — the reference's five end-to-end vectors
    (multiple output types, UMIs, four inputs, several templates per read, +T), then 60 000 dual-index paired-end reads with
    CRLF lines, comments of every kind the header rules distinguish, a trailing +B, and reads that are too short."""
    import json
    import os

    from fqtk_b200 import BarcodeMatcher, synth
    from fqtk_b200.bgzf import BgzfCompressor
    from fqtk_b200.demux import TooFewBases, fastq_text as records_text
    from fqtk_b200.fastq import demux_fastq_batch
    from fqtk_b200.gpu_demux import demux_chunks, demux_fastq_batch_gpu

    def check(m, z, ids, bcs, structures, texts, types, **kw):
        want = demux_fastq_batch(m, ids, bcs, structures, texts, types, **kw)
        m.reset_counts()
        got = demux_fastq_batch_gpu(m, z, ids, bcs, structures, texts, types, **kw)
        assert set(got.files) == set(want.files)
        for name, recs in want.files.items():
            payload, sizes = ob.parse(got.files[name])
            assert payload == records_text(recs), name
            assert sizes[-1] == 0 and all(x == 65280 for x in sizes[:-2])
        assert np.array_equal(got.counts, want.counts) and got.skipped == want.skipped
        assert [(x.sample_id, x.templates) for x in got.metrics] == [(x.sample_id, x.templates) for x in want.metrics]
        if not kw.get("skip_too_few_bases"):  # the one-call C form (fqtk_b200_demux_chunks) must leave the very same files
            m.reset_counts()
            one, used = demux_chunks(m, z, ids, bcs, structures, texts, types)
            assert used == [len(t) for t in texts]
            assert one.files == got.files and np.array_equal(one.counts, want.counts)
        return got

    kats = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
    with BgzfCompressor(0) as z:
        for case in kats["demux_e2e"]:
            S = len(case["barcodes"])
            ids = [f"Sample{j:04d}" for j in range(S)]
            texts = ["".join(f"@ex_{i}\n{b}\n+\n{';' * len(b)}\n" for i, b in enumerate(col)).encode() for col in case["inputs"]]
            with BarcodeMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"]) as m:
                got = check(m, z, ids, case["barcodes"], case["read_structures"], texts, case["output_types"])
                assert set(got.files) == set(case["expect"])
        # at size
        cfg = synth.CONFIGS[3]
        panel = synth.panel(cfg)
        bcs = [bytes(r).decode() for r in panel]
        ids = [f"S{j}" for j in range(len(bcs))]
        n = 60_000
        reads = synth.reads_host(panel, cfg.seed_reads, 7, n)
        rng = np.random.default_rng(3)
        comments = [b"", b" 1:N:0:0", b" 2:Y:18:ACGTACGT", b" 1:N:0:", b" 0:0", b" x"]
        r1, r2, i1, i2 = [], [], [], []
        for i in range(n):
            name = b"@A00:1:HX:1:1101:%d:%d" % (1000 + i % 977, 2000 + i // 7)
            cm = comments[i % len(comments)]
            eol = b"\r\n" if i % 11 == 0 else b"\n"
            t1 = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(20, 60))))
            t2 = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(20, 60))))
            q = lambda s: bytes(rng.integers(35, 74, size=len(s), dtype=np.uint8))
            umi = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=6))
            r1.append(name + cm + eol + umi + t1 + eol + b"+" + eol + q(umi + t1) + eol)
            r2.append(name + cm.replace(b" 1:", b" 2:") + eol + t2 + eol + b"+" + eol + q(t2) + eol)
            b1, b2 = bytes(reads[i, 0:8]), bytes(reads[i, 8:16])
            i1.append(name + cm + eol + b1 + eol + b"+" + eol + b"F" * 8 + eol)
            i2.append(name + cm + eol + b2 + eol + b"+" + eol + b"F" * 8 + eol)
        texts = [b"".join(x) for x in (r1, r2, i1, i2)]
        with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta) as m:
            got = check(m, z, ids, bcs, ["6M+T", "+T", "8B", "+B"], texts, ["T", "B", "M"])
            assert len(got.files) > 3 * 300 and got.text_bytes > 10_000_000
            # too few bases: an error by default, skipped on request
            texts[2] = texts[2].replace(b"\n" + bytes(reads[5, 0:8]) + b"\n+\nFFFFFFFF\n", b"\n" + bytes(reads[5, 0:3]) + b"\n+\nFFF\n", 1)
            with pytest.raises(TooFewBases) as e1:
                demux_fastq_batch_gpu(m, z, ids, bcs, ["6M+T", "+T", "8B", "+B"], texts, ["T"])
            with pytest.raises(TooFewBases) as e2:
                demux_fastq_batch(m, ids, bcs, ["6M+T", "+T", "8B", "+B"], texts, ["T"])
            assert str(e1.value) == str(e2.value)
            with pytest.raises(TooFewBases) as e3:
                demux_chunks(m, z, ids, bcs, ["6M+T", "+T", "8B", "+B"], texts, ["T"])
            assert str(e3.value).endswith(str(e2.value))
            m.reset_counts()
            got = check(m, z, ids, bcs, ["6M+T", "+T", "8B", "+B"], texts, ["T"], skip_too_few_bases=True)
            assert got.skipped == 1
            # chunks that end in the middle of a record: the complete read sets are written, the rest is carried over
            good = [b"".join(x) for x in (r1, r2, i1, i2)]
            cut = [good[0][: len(good[0]) // 2], good[1][: len(good[1]) // 2 + 777], good[2][: len(good[2]) // 3], good[3]]
            m.reset_counts()
            part, used = demux_chunks(m, z, ids, bcs, ["6M+T", "+T", "8B", "+B"], cut, ["T"])
            k = int(part.counts.sum())
            assert 0 < k < n and used == [sum(len(x) for x in col[:k]) for col in (r1, r2, i1, i2)]
            m.reset_counts()
            rest, used2 = demux_chunks(m, z, ids, bcs, ["6M+T", "+T", "8B", "+B"], [g[u:] for g, u in zip(good, used)], ["T"])
            assert int(rest.counts.sum()) == n - k
            whole = demux_fastq_batch(m, ids, bcs, ["6M+T", "+T", "8B", "+B"], good, ["T"])
            for name, recs in whole.files.items():  # a file = the members of its batches, in order, then the EOF block
                image = part.files.get(name, ob.BGZF_EOF)[:-28] + rest.files.get(name, ob.BGZF_EOF)
                assert ob.parse(image)[0] == records_text(recs), name


@pytest.mark.gpu
def test_gpu_two_lanes_leave_the_files_of_one_lane():
    """DemuxLanes: batches alternate between two (matcher, compressor) pairs on two host threads so that one batch's copies
    overlap the other's kernels.  Same per-batch results, in batch order, as one lane working through the batches; every
    file = its batches' members in order (records of a sample stay in input order, demux.rs:1505-1523); counts add up."""
    from fqtk_b200 import BarcodeMatcher, synth
    from fqtk_b200.bgzf import BgzfCompressor
    from fqtk_b200.gpu_demux import DemuxLanes, demux_chunks

    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    bcs = [bytes(r).decode() for r in panel]
    ids = [f"S{j}" for j in range(len(bcs))]
    rng = np.random.default_rng(5)
    batches = []
    for b in range(7):
        n = int(rng.integers(1500, 6000))
        reads = synth.reads_host(panel, cfg.seed_reads, 100_000 * b, n)
        t = [b"".join(b"@r%d_%d 1:N:0:0\n%s\n+\n%s\n" % (b, i, bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=40)), b"F" * 40)
                      for i in range(n)),
             b"".join(b"@r%d_%d 1:N:0:0\n%s\n+\n%s\n" % (b, i, bytes(reads[i]), b"F" * 16) for i in range(n))]
        batches.append(t)
    structures = ["+T", "8B8B"]
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta) as m, BgzfCompressor(0) as z:
        want = [demux_chunks(m, z, ids, bcs, structures, t, ["T", "B"]) for t in batches]
        total = m.counts()
    with DemuxLanes(lambda: (BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta), BgzfCompressor(0)), lanes=2) as lanes:
        got = list(lanes.map(batches, ids, bcs, structures, ["T", "B"]))
        assert np.array_equal(lanes.counts(), total)
    assert len(got) == len(want)
    for (g, gu), (w, wu) in zip(got, want):
        assert gu == wu and g.files == w.files and np.array_equal(g.counts, w.counts)
    assert int(total.sum()) == sum(int(w.counts.sum()) for w, _ in want)


@pytest.mark.gpu
def test_gpu_header_rewrite_reference_vectors():
    """The reference's six write_header tests (demux.rs:2084-2196) through the device kernel: a one-record FASTQ whose read
    structure yields the test's sample-barcode and UMI segments, two template streams so that read number 2 exists."""
    import json
    import os

    from fqtk_b200 import BarcodeMatcher
    from fqtk_b200.bgzf import BgzfCompressor
    from fqtk_b200.gpu_demux import demux_fastq_batch_gpu

    kats = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
    with BgzfCompressor(0) as z:
        for case in kats["write_header"]:
            bcs, umis = case["sample_barcodes"], case["umis"]
            seq = "".join(bcs) + "".join(umis) + "AC"
            structure = "".join(f"{len(b)}B" for b in bcs) + "".join(f"{len(u)}M" for u in umis) + "1T1T"
            text = f"@{case['header']}\n{seq}\n+\n{'I' * len(seq)}\n".encode()
            with BarcodeMatcher(["".join(bcs)], 0, 0) as m:
                if "expect_error" in case:
                    with pytest.raises(_lib.Fqtk_b200Error) as e:
                        demux_fastq_batch_gpu(m, z, ["s"], ["".join(bcs)], [structure], [text], ["T"])
                    assert case["expect_error"] in e.value.message, case["source"]
                    continue
                got = demux_fastq_batch_gpu(m, z, ["s"], ["".join(bcs)], [structure], [text], ["T"])
            lines = gzip.decompress(got.files[f"s.R{case['read_num']}.fq.gz"]).split(b"\n")
            assert lines[0].decode() == case["expect"], case["source"]
            assert lines[1] in (b"A", b"C") and lines[2] == b"+" and lines[3] == b"I"
