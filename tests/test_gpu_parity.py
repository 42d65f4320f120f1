"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact (integer / index work): result word per read and counts[S+1].  Needs a B200 (run via gpurun)."""
import numpy as np
import pytest

import oracle
from fqtk_b200 import BarcodeMatch, BarcodeMatcher, MatcherPanic, _lib, synth

pytestmark = pytest.mark.gpu

MODES = [True, False]  # use_cache: True -> memo-table kernels, False -> brute-force kernels (rstest cases in the reference)


def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    return torch


def to_match(t):
    return None if t is None else BarcodeMatch(*t)


# ---------------------------------------------------------------------------------------------------------
# the reference's own known-answer tests, through the mirrored interface
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("use_cache", MODES)
def test_reference_assign_kats(kats, use_cache):
    for case in kats["assign"]:
        with BarcodeMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"], use_cache) as m:
            assert m.assign(case["read"].encode()) == to_match(case["expect"]), case


@pytest.mark.parametrize("use_cache", MODES)
def test_reference_count_mismatches_kats_via_single_sample_panel(kats, use_cache):
    # count_mismatches (barcode_matching.rs:89-110) is observable as best_mismatches against a 1-sample panel
    for case in kats["count_mismatches"]:
        if case["expected"] == "":
            continue  # an empty barcode cannot be constructed (barcode_matching.rs:62-65)
        with BarcodeMatcher([case["expected"]], 255, 0, use_cache) as m:
            got = m.assign(case["observed"].encode())
            assert got == BarcodeMatch(0, case["mismatches"], 255), case


@pytest.mark.parametrize("use_cache", MODES)
def test_reference_demux_caller_vectors(kats, use_cache):
    for case in kats["demux_caller"]:
        with BarcodeMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"], use_cache) as m:
            reads = np.frombuffer("".join(case["reads"]).encode(), dtype=np.uint8).reshape(len(case["reads"]), -1)
            res = m.assign_batch(reads)
            got = [None if r == _lib.NONE else int(r) >> 16 for r in res]
            assert got == case["expect_sample"], case["source"]
            assert m.counts().tolist() == case["expect_counts"], case["source"]


def test_matcher_instantiation(kats):
    ok = kats["matcher_new"]["ok"]
    BarcodeMatcher(ok["barcodes"], ok["max_mismatches"], ok["min_mismatch_delta"]).close()
    with pytest.raises(MatcherPanic, match="Must provide at least one sample"):
        BarcodeMatcher([], 2, 1)


@pytest.mark.parametrize("use_cache", MODES)
def test_length_rules(use_cache):
    with BarcodeMatcher(["ACGTAC", "TTTTTT"], 1, 1, use_cache) as m:
        assert m.assign(b"ACGTA") is None            # barcode_matching.rs:167-169
        assert m.assign(b"") is None
        with pytest.raises(MatcherPanic, match="differs from expected barcode"):
            m.assign(b"ACGTACG")                      # :95-106
        assert m.assign(b"NNNNNNN") is None           # no-call pre-filter (:170-172) fires before the panic
        assert m.assign(b"ACGTAC") == BarcodeMatch(0, 0, 5)
        assert m.counts().tolist() == [1, 0, 3]       # the failed call counted nothing
        # batched with per-row lengths
        rows = np.zeros((4, 8), dtype=np.uint8)
        lens = np.array([6, 5, 6, 7], dtype=np.uint32)
        for i, s in enumerate([b"ACGTAC", b"TTTTT", b"TTTTTA", b"NNNNNNN"]):
            rows[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
        m.reset_counts()
        res = m.assign_batch(rows, lengths=lens)
        assert [None if r == _lib.NONE else int(r) >> 16 for r in res] == [0, None, 1, None]
        assert m.counts().tolist() == [1, 1, 2]
        lens[3] = 8
        rows[3] = np.frombuffer(b"ACGTACGT", dtype=np.uint8)
        with pytest.raises(MatcherPanic):
            m.assign_batch(rows, lengths=lens)
        assert m.counts().tolist() == [1, 1, 2], "a failing batch counts nothing"


# ---------------------------------------------------------------------------------------------------------
# randomized parity against the oracle: both kernel families, host ASCII path and device packed / ASCII paths
# ---------------------------------------------------------------------------------------------------------
ALPHABETS = {
    "acgt": b"ACGT",
    "acgtn": b"ACGTN",
    "iupac": b"ACGTUMRWSYKVHDBNn.",
    "dirty": b"ACGTNn.acgtRYKMXx-*0 ",
}


def random_panel(rng, S, L, alphabet):
    seen, bcs = set(), []
    guard = 0
    while len(bcs) < S and guard < 100 * S:
        guard += 1
        if bcs and rng.random() < 0.4:
            b = bytearray(bcs[int(rng.integers(0, len(bcs)))])
            b[int(rng.integers(0, L))] = alphabet[int(rng.integers(0, len(alphabet)))]
            b = bytes(b)
        else:
            b = bytes(alphabet[i] for i in rng.integers(0, len(alphabet), L))
        if b not in seen:
            seen.add(b)
            bcs.append(b)
    return bcs


def random_reads(rng, bcs, L, n, alphabet):
    panel = np.frombuffer(b"".join(bcs), dtype=np.uint8).reshape(len(bcs), L)
    reads = panel[rng.integers(0, len(bcs), n)].copy()
    nsub = rng.integers(0, 4, n)
    for k in range(3):
        rows = np.nonzero(nsub > k)[0]
        reads[rows, rng.integers(0, L, rows.size)] = np.frombuffer(alphabet, dtype=np.uint8)[
            rng.integers(0, len(alphabet), rows.size)]
    rnd = rng.random(n) < 0.2
    reads[rnd] = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), (int(rnd.sum()), L))]
    return reads


def check_against_oracle(bcs, mm, delta, reads, use_cache, expect_mode=None):
    torch = torch_cuda()
    om = oracle.OracleMatcher(bcs, mm, delta, use_cache=True)
    want, want_counts = om.assign_batch(reads, mode=0)
    n, L = reads.shape
    with BarcodeMatcher(bcs, mm, delta, use_cache) as m:
        if expect_mode:
            assert m.mode == expect_mode
        # 1) reference-facing host-buffer call
        got = m.assign_batch(reads)
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (m.mode, bytes(reads[bad[0]]), hex(int(got[bad[0]])), hex(int(want[bad[0]])))
        assert np.array_equal(m.counts(), want_counts)
        # 2) HBM-resident packed call
        m.reset_counts()
        packed = synth.pack_host(reads)
        d_packed = torch.from_numpy(packed.view(np.int32)).cuda()
        d_res = torch.empty(n, dtype=torch.int32, device="cuda")
        m.assign_packed_device(d_packed.data_ptr(), n, d_res.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got2 = d_res.cpu().numpy().view(np.uint32)
        assert np.array_equal(got2, want)
        assert np.array_equal(m.counts(), want_counts)
        # 3) device ASCII call with a padded row stride
        m.reset_counts()
        stride = L + 3
        padded = np.full((n, stride), ord("#"), dtype=np.uint8)
        padded[:, :L] = reads
        d_ascii = torch.from_numpy(padded).cuda()
        d_res.zero_()
        m.assign_ascii_device(d_ascii.data_ptr(), n, stride, d_res.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_res.cpu().numpy().view(np.uint32), want)
        assert np.array_equal(m.counts(), want_counts)


@pytest.mark.parametrize("use_cache", MODES)
@pytest.mark.parametrize("seed", range(4))
def test_fuzz_small_panels(seed, use_cache):
    rng = np.random.default_rng(100 + seed)
    for _ in range(12):
        L = int(rng.choice([1, 2, 4, 7, 8, 9, 15, 16, 17, 20, 24, 31, 32]))
        S = int(rng.choice([1, 2, 3, 8, 33, 100]))
        pa = ALPHABETS[["acgt", "acgtn", "iupac"][int(rng.integers(0, 3))]]
        ra = ALPHABETS[["acgt", "acgtn", "dirty"][int(rng.integers(0, 3))]]
        bcs = random_panel(rng, S, L, pa)
        mm = int(rng.choice([0, 1, 2, 3]))
        delta = int(rng.choice([0, 1, 2, 3, 100]))
        reads = random_reads(rng, bcs, L, 3000 + int(rng.integers(0, 7)), ra)
        check_against_oracle(bcs, mm, delta, reads, use_cache)


@pytest.mark.parametrize("use_cache", MODES)
def test_fuzz_extreme_parameters(use_cache):
    rng = np.random.default_rng(77)
    for mm, delta in [(100, 0), (100, 3), (255, 255), (0, 0), (0, 255), (8, 1)]:
        bcs = random_panel(rng, 20, 8, ALPHABETS["acgtn"])
        reads = random_reads(rng, bcs, 8, 2000, ALPHABETS["dirty"])
        check_against_oracle(bcs, mm, delta, reads, use_cache)


def test_long_barcodes_take_the_generic_kernel():
    rng = np.random.default_rng(5)
    for L in (33, 40, 64, 100, 254):
        bcs = random_panel(rng, 17, L, ALPHABETS["iupac"])
        reads = random_reads(rng, bcs, L, 1500, ALPHABETS["dirty"])
        check_against_oracle(bcs, 2, 1, reads, use_cache=True, expect_mode="brute")


def test_long_barcodes_through_a_multi_chunk_pipeline_and_with_row_lengths():
    """the L > 32 host-buffer route packs into scratch before the long-barcode kernel; consecutive chunks run
    on different streams, so the scratch must be per pipeline slot (a shared one is overwritten while it is read).
    chunk_bytes = 64 KiB makes 30 000 reads of 40 bases span ~19 chunks.  Also: per-row lengths at L > 32 (the
    reference-facing assign() always passes one) — shorter rows are None, longer rows panic unless pre-filtered."""
    rng = np.random.default_rng(55)
    for L, S in ((40, 23), (70, 9)):
        bcs = random_panel(rng, S, L, ALPHABETS["iupac"])
        n = 30_000
        reads = random_reads(rng, bcs, L, n, ALPHABETS["dirty"])
        om = oracle.OracleMatcher(bcs, 3, 1, use_cache=True)
        want, want_counts = om.assign_batch(reads, mode=0)
        assert (want != _lib.NONE).sum() > n // 10
        with BarcodeMatcher(bcs, 3, 1, use_cache=True, chunk_bytes=64 << 10) as m:
            assert m.mode == "brute"
            for _ in range(3):  # races are timing dependent: several passes
                m.reset_counts()
                got = m.assign_batch(reads)
                assert np.array_equal(got, want)
                assert np.array_equal(m.counts(), want_counts)
            # ragged rows: every third row loses its tail (-> None), lengths passed explicitly
            lens = np.full(n, L, dtype=np.uint32)
            lens[::3] = rng.integers(0, L, size=len(lens[::3]))
            want2 = want.copy()
            want2[::3] = _lib.NONE
            m.reset_counts()
            got2 = m.assign_batch(reads, lengths=lens)
            assert np.array_equal(got2, want2)
            assert int(m.counts()[-1]) == int((want2 == _lib.NONE).sum())
            # the single-read call of the reference interface
            for i in (0, 1, 2, 5, 11):
                assert m.assign(bytes(reads[i])) == oracle_match(om, bytes(reads[i]))
            assert m.assign(bytes(reads[0][: L - 1])) is None
            with pytest.raises(MatcherPanic, match=r"length \(%d\) differs from expected barcode" % (L + 1)):
                m.assign(bcs[0] + b"A")  # its no-calls are within max_mismatches + max_ns_in_barcodes: no pre-filter
            assert m.assign(b"N" * (L + 1)) is None  # the no-call pre-filter fires before the length panic


def oracle_match(om, read):
    w = int(om.assign_batch(np.frombuffer(read, dtype=np.uint8).reshape(1, -1), mode=0)[0][0])
    return None if w == _lib.NONE else BarcodeMatch(w >> 16, (w >> 8) & 0xFF, w & 0xFF)


def test_length_panic_text_is_the_reference_text():
    """barcode_matching.rs:99-105 and its test (:326-341): decoded read, both lengths, the first sample's barcode and id."""
    from fqtk_b200.samples import Sample

    samples = [Sample("sample_0", "CTATGT", 0), Sample("sample_1", "GGGGGG", 1)]
    with BarcodeMatcher(samples, 1, 1) as m:
        with pytest.raises(MatcherPanic) as ei:
            m.assign(b"GATTACAn")
        assert ei.value.args[0] == ("Read barcode (GATTACAN) length (8) differs from expected barcode (CTATGT) length (6) "
                                    "for sample sample_0")
    with BarcodeMatcher(["CTATGT"], 1, 1) as m:  # bare barcodes: the sample index stands in for the id
        with pytest.raises(MatcherPanic, match=r"length \(7\) differs from expected barcode \(CTATGT\) length \(6\) for sample 0"):
            m.assign(b"GATTACA")


def test_matchers_created_concurrently_with_different_options():
    """Per-handle options (create_ex): two host threads creating matchers at the same time cannot influence each other."""
    import threading

    cfg = synth.CONFIGS[2]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    reads = synth.reads_host(panel, cfg.seed_reads, 0, 50_000)
    want, _ = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta).assign_batch(reads)
    out = {}

    def work(kernel):
        for _ in range(3):
            with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True, kernel=kernel) as m:
                info = m.info()
                out.setdefault(kernel, []).append((int(info.cuckoo_probes), int(info.l2_table_entries),
                                                   np.array_equal(m.assign_batch(reads), want)))

    ts = [threading.Thread(target=work, args=(k,)) for k in (0, 1, 2, 3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for k in (0, 1, 2, 3):
        assert len(out[k]) == 3
        for probes, l2e, ok in out[k]:
            assert ok
            assert probes == (k if k >= 2 else 0), (k, probes)
            assert (l2e > 0) == (k == 1), (k, l2e)


def test_packed_host_wire_format():
    """fqtk_b200_matcher_assign_batch_packed: BitEnc words in HOST memory in (what encode() returns; half the PCIe bytes of
    ASCII rows), result words and / or 2-byte sample indices out — the same answers as the ASCII call, chunked pipeline
    included (chunk_bytes = 256 KiB -> ~30 chunks)."""
    from fqtk_b200.barcode_matching import pack_host

    for cfg_id, n in ((3, 250_003), (5, 60_001), (2, 100_000)):
        cfg = synth.CONFIGS[cfg_id]
        panel = synth.panel(cfg)
        bcs = [bytes(r) for r in panel]
        reads = synth.reads_host(panel, cfg.seed_reads, 5, n)
        reads[::53, 1] = ord("r")
        want, want_counts = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta).assign_batch(reads)
        packed = pack_host(reads)
        for use_cache in MODES:
            with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, use_cache, chunk_bytes=256 << 10) as m:
                words, index = m.assign_batch_packed(packed, want_words=True, want_index=True)
                assert np.array_equal(words, want)
                assert np.array_equal(index, np.where(want == _lib.NONE, 0xFFFF, want >> 16).astype(np.uint16))
                assert np.array_equal(m.counts(), want_counts)
                m.reset_counts()
                words, index = m.assign_batch_packed(packed, want_words=False, want_index=True)
                assert words is None and np.array_equal(index.astype(np.uint32) == 0xFFFF, want == _lib.NONE)
                assert np.array_equal(m.counts(), want_counts)
    # barcodes longer than 32 bases go through the same call
    rng = np.random.default_rng(3)
    bcs = random_panel(rng, 11, 45, ALPHABETS["acgtn"])
    reads = random_reads(rng, bcs, 45, 20_000, ALPHABETS["dirty"])
    want, _ = oracle.OracleMatcher(bcs, 2, 1).assign_batch(reads, mode=0)
    with BarcodeMatcher(bcs, 2, 1, True, chunk_bytes=64 << 10) as m:
        words, _ = m.assign_batch_packed(pack_host(reads))
        assert np.array_equal(words, want)


@pytest.mark.parametrize("devices", [None, [0, 0, 0]])
def test_group_of_matchers_is_one_matcher(devices):
    """fqtk_b200_group_*: a batch split into contiguous shards over the devices of the box (every visible device; and
    three handles on device 0, which exercises the shard / reduce logic on a one-GPU box) gives exactly the result words of
    a single matcher, in input order, and ONE count table — host ASCII rows, host packed words and HBM-resident shards."""
    torch = torch_cuda()
    from fqtk_b200 import MatcherGroup
    from fqtk_b200.barcode_matching import pack_host

    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    n = 1_000_003
    reads = synth.reads_host(panel, cfg.seed_reads, 99, n)
    reads[::97, 3] = ord("R")
    want, want_counts = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta).assign_batch(reads)
    with MatcherGroup(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True, devices=devices, chunk_bytes=1 << 20) as g:
        G = g.size
        assert G == (len(devices) if devices else torch.cuda.device_count())
        bounds = [g.shard(n, k) for k in range(G)]
        assert bounds[0][0] == 0 and sum(c for _, c in bounds) == n
        assert all(bounds[k][0] + bounds[k][1] == bounds[k + 1][0] for k in range(G - 1))
        assert np.array_equal(g.assign_batch(reads), want)
        assert np.array_equal(g.counts(), want_counts)
        g.reset_counts()
        packed = pack_host(reads)
        words, index = g.assign_batch_packed(packed, want_words=True, want_index=True)
        assert np.array_equal(words, want)
        assert np.array_equal(index.astype(np.uint32) == 0xFFFF, want == _lib.NONE)
        assert np.array_equal(g.counts(), want_counts)
        g.reset_counts()
        # HBM-resident shards, one per device, asynchronous launches
        d_packed, d_res = [], []
        for k, (first, count) in enumerate(bounds):
            dev = torch.device("cuda", g.devices[k])
            d_packed.append(torch.from_numpy(packed[first:first + count].view(np.int32)).to(dev))
            d_res.append(torch.empty(count, dtype=torch.int32, device=dev))
        for _ in range(2):
            g.assign_packed_device([t.data_ptr() for t in d_packed], [c for _, c in bounds], [t.data_ptr() for t in d_res])
        got = np.concatenate([t.cpu().numpy().view(np.uint32) for t in d_res])
        assert np.array_equal(got, want)
        assert np.array_equal(g.counts(), 2 * want_counts)  # ONE table over all devices, accumulated over both passes
    with pytest.raises(_lib.Fqtk_b200Error):
        MatcherGroup(bcs, 1, 2, True, devices=[99])


def test_many_samples_use_global_histogram_and_big_panel():
    rng = np.random.default_rng(6)
    S, L = 9000, 12  # S + 1 > 8192 shared-memory bins
    panel = synth.make_panel(99, S, L, 3)
    bcs = [bytes(r) for r in panel]
    reads = random_reads(rng, bcs, L, 20000, ALPHABETS["acgtn"])
    check_against_oracle(bcs, 1, 2, reads, use_cache=True, expect_mode="table")
    check_against_oracle(bcs, 1, 2, reads, use_cache=False, expect_mode="brute")
    # panel planes larger than shared memory: S * 16 B > 227 KB
    S2 = 15000
    panel2 = synth.make_panel(98, S2, 14, 3)
    bcs2 = [bytes(r) for r in panel2]
    reads2 = random_reads(rng, bcs2, 14, 8000, ALPHABETS["acgtn"])
    check_against_oracle(bcs2, 1, 2, reads2, use_cache=False, expect_mode="brute")


@pytest.mark.parametrize("L,S", [(8, 1), (8, 2), (8, 31), (8, 33), (16, 64), (16, 545), (16, 3072), (16, 3073), (16, 6200),
                                 (20, 2047), (20, 2049), (24, 4100), (32, 1536), (32, 1537), (29, 3200)])
def test_brute_sliced_ties_across_groups_and_chunk_launches(L, S):
    """k_brute_sliced keeps its running minima bit-sliced per slot (barcodes 32g + j of all groups g), reduces them after the
    first shared-memory chunk and walks the rest of a large panel one launch per chunk with the interim state in the result
    word.  Panels made of near-duplicates of a few barcodes put equal best distances in different slots, groups and chunks
    (FIRST index wins, barcode_matching.rs:132) and equal second-best distances everywhere (:140); S around the group and
    chunk boundaries (3 072 barcodes per chunk at L <= 16, 2 048 at L <= 24, 1 536 at L <= 32) covers the partial last group
    of the first and of a later launch."""
    rng = np.random.default_rng(1000 * L + S)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    roots = acgt[rng.integers(0, 4, size=(7, L))]
    panel = roots[rng.integers(0, len(roots), size=S)].copy()
    for j in range(S):  # 0 - 2 substitutions; a few IUPAC codes and no-calls in the sheet
        for _ in range(int(rng.integers(0, 3))):
            panel[j, rng.integers(0, L)] = acgt[rng.integers(0, 4)]
        if rng.random() < 0.05:
            panel[j, rng.integers(0, L)] = np.frombuffer(b"RYKMSWN", dtype=np.uint8)[rng.integers(0, 7)]
    bcs = [bytes(r) for r in panel]
    n = 6000 + int(rng.integers(0, 40))
    reads = panel[rng.integers(0, S, size=n)].copy()
    reads[:n // 8] = roots[rng.integers(0, len(roots), size=n // 8)]
    sub = rng.random(size=reads.shape) < 0.06
    reads[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()))]
    reads[rng.random(size=reads.shape) < 0.01] = ord("N")
    reads[rng.random(size=reads.shape) < 0.003] = ord("S")
    for mm, delta in [(3, 0), (1, 2), (L, 1)]:
        check_against_oracle(bcs, mm, delta, reads, use_cache=False, expect_mode="brute")


@pytest.mark.parametrize("cfg_id,L", [(3, 16), (2, 8), (None, 24)])
def test_host_pack_route_of_assign_batch_equals_the_plain_route(cfg_id, L):
    """fqtk_b200_matcher_set_host_pack: encode() done by host threads while the batch is in flight (AVX2 stream form, pinned
    staging ring, packed words over PCIe).  Same result words and counts as the plain ASCII route and as the oracle, for
    1 / 3 / 5 packer threads, sizes with chunk tails, reads with no-calls, lower case, IUPAC codes and arbitrary bytes."""
    rng = np.random.default_rng(40 + L)
    if cfg_id is None:
        panel = synth.make_panel(5, 200, L, 3)
        mm, delta = 1, 2
    else:
        cfg = synth.CONFIGS[cfg_id]
        panel, mm, delta = synth.panel(cfg), cfg.max_mismatches, cfg.min_mismatch_delta
    bcs = [bytes(r) for r in panel]
    n = (3 << 20) + 4099
    reads = panel[rng.integers(0, len(bcs), size=n)].copy()
    sub = rng.random(size=reads.shape) < 0.01
    reads[sub] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(sub.sum()))]
    reads[rng.random(size=reads.shape) < 0.003] = ord("N")
    reads[5::97, 1] = np.frombuffer(b"acgtnRYK.", dtype=np.uint8)[rng.integers(0, 9, size=len(reads[5::97]))]
    reads[11::4001, L - 1] = rng.integers(0, 256, size=len(reads[11::4001]), dtype=np.uint8)
    with BarcodeMatcher(bcs, mm, delta, True) as m:
        want = m.assign_batch(reads)
        want_counts = m.counts()
        win = slice(1_000_000, 1_020_000)
        ow, _ = oracle.OracleMatcher(bcs, mm, delta, use_cache=True).assign_batch(reads[win], mode=0)
        assert np.array_equal(want[win], ow)
        for threads, nn in [(1, n), (3, n), (5, (1 << 20) + 3), (-1, n)]:
            m.reset_counts()
            m.set_host_pack(threads)
            got = m.assign_batch(reads[:nn])
            m.set_host_pack(0)
            assert np.array_equal(got, want[:nn]), (threads, nn)
            if nn == n:
                assert np.array_equal(m.counts(), want_counts)


def test_table_budget_falls_back_to_brute():
    L = _lib.lib()
    try:
        L.fqtk_b200_set_table_budget(10)
        with BarcodeMatcher(["ACGTACGT", "TTTTACGT"], 1, 1, use_cache=True) as m:
            assert m.mode == "brute"
            assert m.assign(b"ACGTACGT") == BarcodeMatch(0, 0, 3)
    finally:
        L.fqtk_b200_set_table_budget(32 << 20)
    with BarcodeMatcher(["NNNNNNNNNNNNNNNNNNNN", "ACGTACGTACGTACGTACGT"], 1, 0, use_cache=True) as m:
        assert m.mode == "brute"  # a catch-all barcode's neighbourhood (5^20) is over any budget
        assert m.assign(b"ACGTACGTACGTACGTACGT") == BarcodeMatch(0, 0, 0)  # tie: first index wins, delta 0 allows it


def test_out_of_alphabet_reads_take_the_warp_cooperative_path():
    """Reads with IUPAC / junk symbols are never in the memo table; every lane pattern of the slow path is hit."""
    rng = np.random.default_rng(8)
    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    reads = synth.reads_host(panel, cfg.seed_reads, 0, 40000)
    for frac in (0.01, 0.3, 1.0):
        r = reads.copy()
        rows = np.nonzero(rng.random(r.shape[0]) < frac)[0]
        r[rows, rng.integers(0, 16, rows.size)] = np.frombuffer(b"RYKMSWBDHVXacgtn.-", dtype=np.uint8)[
            rng.integers(0, 18, rows.size)]
        check_against_oracle(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, r, use_cache=True, expect_mode="table")


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs on their synthetic streams (reduced N against the oracle; full N by properties)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg_id,n", [(1, 10_000), (2, 300_000), (3, 300_000), (4, 200_000), (5, 120_000)])
@pytest.mark.parametrize("use_cache", MODES)
def test_configs_reduced_n(cfg_id, n, use_cache):
    cfg = synth.CONFIGS[cfg_id]
    panel = synth.panel(cfg)
    reads = synth.reads_host(panel, cfg.seed_reads, 0, n)
    check_against_oracle([bytes(r) for r in panel], cfg.max_mismatches, cfg.min_mismatch_delta, reads, use_cache,
                         expect_mode="table" if use_cache else "brute")


@pytest.mark.parametrize("arity", [0, 1, 2, 3, 5])
def test_packed_route_kernel_variants(arity):
    """The HBM-resident packed route has four kernels: k_probe3 (shared-memory cuckoo table of the pure A/C/G/T memo
    entries, 2 or 3 sub-tables, L <= 16), k_probe4 (L2-resident fingerprint table of every A/C/G/T/N candidate, verified
    against the panel, L <= 32; knob 1), k_probe5 (the exact entries in shared memory + the fingerprint table, L <= 16;
    knob 5) and k_probe2 (hot tier + global memo table; knob 0).  All bit-exact."""
    torch = torch_cuda()
    L = _lib.lib()
    rng = np.random.default_rng(4242 + arity)
    try:
        L.fqtk_b200_set_cuckoo_arity(arity)
        for cfg_id, n in [(2, 150_000), (3, 200_000)]:
            cfg = synth.CONFIGS[cfg_id]
            panel = synth.panel(cfg)
            bcs = [bytes(r) for r in panel]
            with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True) as m:
                if arity == 5:
                    assert int(m.info().cuckoo_probes) in (2, 3) and int(m.info().cuckoo_entries) > 0
                else:
                    assert int(m.info().cuckoo_probes) == (arity if arity >= 2 else 0)
                    assert (int(m.info().cuckoo_entries) > 0) == (arity >= 2)
                assert (int(m.info().l2_table_entries) > 0) == (arity in (1, 5))
            reads = synth.reads_host(panel, cfg.seed_reads, 31, n)
            reads[::41, 3] = ord("N")
            reads[::97, 1] = ord("r")
            check_against_oracle(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, reads, True, expect_mode="table")
        if arity in (0, 1, 5):  # the big-table configs on their own kernels, and on k_probe2
            for cfg_id, n in [(4, 120_000), (5, 80_000)]:
                cfg = synth.CONFIGS[cfg_id]
                panel = synth.panel(cfg)
                bcs = [bytes(r) for r in panel]
                reads = synth.reads_host(panel, cfg.seed_reads, 77, n)
                reads[::37, 2] = ord("N")
                reads[::101, 5] = ord("y")
                with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True) as m:
                    assert (int(m.info().l2_table_entries) > 0) == (arity in (1, 5))
                check_against_oracle(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, reads, True, expect_mode="table")
        for _ in range(16):  # pad nibbles (L % 8 != 0), one- to three-word keys, dirty reads, odd parameters
            Lb = int(rng.choice([1, 2, 3, 7, 8, 9, 12, 15, 16, 17, 20, 23, 24, 25, 29, 32]))
            S = int(rng.choice([1, 2, 5, 40, 300]))
            pa = ALPHABETS[["acgt", "acgtn", "iupac"][int(rng.integers(0, 3))]]
            ra = ALPHABETS[["acgt", "acgtn", "dirty"][int(rng.integers(0, 3))]]
            bcs = random_panel(rng, S, Lb, pa)
            mm, delta = int(rng.choice([0, 1, 2])), int(rng.choice([0, 1, 2, 3]))
            reads = random_reads(rng, bcs, Lb, 5000 + int(rng.integers(0, 300)), ra)
            check_against_oracle(bcs, mm, delta, reads, True)
    finally:
        L.fqtk_b200_set_cuckoo_arity(-1)
    del torch


def test_synth_device_equals_host():
    torch = torch_cuda()
    for cfg_id in (2, 3, 5):
        cfg = synth.CONFIGS[cfg_id]
        panel = synth.panel(cfg)
        n, first = 50_001, 123_456_789
        host = synth.reads_host(panel, cfg.seed_reads, first, n)
        d_ascii = torch.empty((n, cfg.barcode_len), dtype=torch.uint8, device="cuda")
        d_packed = torch.empty((n, cfg.words_per_read), dtype=torch.int32, device="cuda")
        synth.reads_device(panel, cfg.seed_reads, first, n, d_ascii.data_ptr(), d_packed.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_ascii.cpu().numpy(), host)
        assert np.array_equal(d_packed.cpu().numpy().view(np.uint32), synth.pack_host(host))
        # pack kernel = encode()
        d_packed2 = torch.zeros_like(d_packed)
        _lib.check(_lib.lib().fqtk_b200_pack_device(d_ascii.data_ptr(), n, cfg.barcode_len, cfg.barcode_len,
                                                    d_packed2.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(d_packed, d_packed2)


@pytest.mark.parametrize("cfg_id", [2, 3, 4, 5])
def test_configs_full_size_properties(cfg_id):
    """At BASELINE.json's full N: (i) table kernels == brute kernels read-for-read (checksum of the XOR and an
    exact equality), (ii) counts sum to N and equal the histogram of the result words, (iii) strided sub-ranges
    replayed on the host agree with the oracle, (iv) idempotence: a second pass doubles every count."""
    torch = torch_cuda()
    cfg = synth.CONFIGS[cfg_id]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    free, _ = torch.cuda.mem_get_info()
    n = cfg.n_reads
    W = cfg.words_per_read
    need = n * (4 * W + 8) + (2 << 30)
    assert need < free, "full-size config must fit one B200"
    stream = torch.cuda.current_stream().cuda_stream
    d_packed = torch.empty((n, W), dtype=torch.int32, device="cuda")
    synth.reads_device(panel, cfg.seed_reads, 0, n, 0, d_packed.data_ptr(), stream)
    # the rare paths of the table kernels at full N: windows of the stream are overwritten with
    #   (a) 3 000 consecutive reads that ALL carry a no-call: every lane of a tile is odd at once (k_probe3's resolve-on-the
    #       -spot branch, its stash overflowing; k_probe5's queue holding a whole tile),
    #   (b) reads with IUPAC codes / junk symbols sprinkled over a long stretch (the stash / slow path with few lanes busy),
    # and the batch ends 37 reads short of a tile boundary (the one-read-per-lane tail).
    n -= 37
    inject = {}
    for first, span, kind in ((n // 2 + 5, 3_000, "burst"), (n // 5 + 11, 60_000, "sprinkle"), (n - 2_000, 2_000, "sprinkle")):
        host = synth.reads_host(panel, cfg.seed_reads, first, span)
        if kind == "burst":
            host[np.arange(span), np.arange(span) % cfg.barcode_len] = ord("N")
        else:
            host[::7, 1] = ord("R")
            host[3::11, cfg.barcode_len - 1] = ord("-")
            host[5::13, 0] = ord("n")
        inject[first] = host
        d_packed[first:first + span] = torch.from_numpy(synth.pack_host(host).view(np.int32)).cuda()
    res_t = torch.empty(n, dtype=torch.int32, device="cuda")
    res_b = torch.empty(n, dtype=torch.int32, device="cuda")
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True) as mt, \
            BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, False) as mb:
        assert mt.mode == "table" and mb.mode == "brute"
        mt.assign_packed_device(d_packed.data_ptr(), n, res_t.data_ptr(), stream)
        mb.assign_packed_device(d_packed.data_ptr(), n, res_b.data_ptr(), stream)
        torch.cuda.synchronize()
        assert torch.equal(res_t, res_b)
        ct, cb = mt.counts(), mb.counts()
        assert np.array_equal(ct, cb)
        assert int(ct.sum()) == n
        # histogram of result words == counts
        idx = torch.where(res_t == -1, torch.full_like(res_t, cfg.n_samples), (res_t >> 16) & 0xFFFF)
        hist = torch.bincount(idx.to(torch.int64), minlength=cfg.n_samples + 1).cpu().numpy().astype(np.uint64)
        assert np.array_equal(hist, ct)
        del idx
        # strided sub-ranges against the oracle
        om = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, use_cache=True)
        span = 20_000
        for first in (0, n // 3 + 17):
            host_reads = synth.reads_host(panel, cfg.seed_reads, first, span)
            want, _ = om.assign_batch(host_reads)
            got = res_t[first:first + span].cpu().numpy().view(np.uint32)
            assert np.array_equal(got, want), (cfg_id, first)
        for first, host_reads in inject.items():  # the overwritten windows, against the oracle on what was written
            want, _ = om.assign_batch(host_reads)
            got = res_t[first:first + len(host_reads)].cpu().numpy().view(np.uint32)
            assert np.array_equal(got, want), (cfg_id, first)
            assert (want != _lib.NONE).any()
        # idempotence of the counters
        mt.assign_packed_device(d_packed.data_ptr(), n, res_t.data_ptr(), stream)
        torch.cuda.synchronize()
        assert np.array_equal(mt.counts(), 2 * ct)
        assert torch.equal(res_t, res_b)


def test_mostly_unmatched_stream_at_size():
    """Counter plumbing under a skewed stream: 300 M reads of which more than half match nothing (reads drawn from a
    DIFFERENT panel), so the unmatched bin of every CTA runs far past 2^16 between two histogram flushes of k_probe3,
    and the no-call reads among them are parked and come back unmatched.  Table kernels == brute kernels read for
    read, counts == histogram of the result words."""
    torch = torch_cuda()
    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    other = synth.make_panel(0xBEEF, cfg.n_samples, cfg.barcode_len, 3)
    bcs = [bytes(r) for r in panel]
    n = 300_000_000
    W = cfg.words_per_read
    stream = torch.cuda.current_stream().cuda_stream
    d_packed = torch.empty((n, W), dtype=torch.int32, device="cuda")
    k = n * 2 // 5
    synth.reads_device(panel, cfg.seed_reads, 0, k, 0, d_packed.data_ptr(), stream)
    synth.reads_device(other, cfg.seed_reads + 1, 0, n - k, 0, d_packed[k:].data_ptr(), stream)
    res_t = torch.empty(n, dtype=torch.int32, device="cuda")
    res_b = torch.empty(n, dtype=torch.int32, device="cuda")
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True) as mt, \
            BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, False) as mb:
        assert int(mt.info().cuckoo_probes) >= 2
        mt.assign_packed_device(d_packed.data_ptr(), n, res_t.data_ptr(), stream)
        mb.assign_packed_device(d_packed.data_ptr(), n, res_b.data_ptr(), stream)
        torch.cuda.synchronize()
        assert torch.equal(res_t, res_b)
        ct, cb = mt.counts(), mb.counts()
        assert np.array_equal(ct, cb)
        assert int(ct.sum()) == n
        assert int(ct[-1]) > n // 2
        idx = torch.where(res_t == -1, torch.full_like(res_t, cfg.n_samples), (res_t >> 16) & 0xFFFF)
        hist = torch.bincount(idx.to(torch.int64), minlength=cfg.n_samples + 1).cpu().numpy().astype(np.uint64)
        assert np.array_equal(hist, ct)


# ---------------------------------------------------------------------------------------------------------
# SURVEY 8f "next" #2: the batched host pipeline (fqtk_b200/demux.py) against the reference's end-to-end vectors
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("use_cache", MODES)
def test_batched_pipeline_reference_end_to_end_vectors(kats, use_cache):
    from fqtk_b200.demux import demux_batch

    for case in kats["demux_e2e"]:
        S = len(case["barcodes"])
        ids = [f"Sample{j:04d}" for j in range(S)]  # metadata_lines_from_barcodes (demux.rs:1046-1053)
        inputs = [[(f"ex_{i}".encode(), b.encode(), b";" * len(b)) for i, b in enumerate(col)] for col in case["inputs"]]
        with BarcodeMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"], use_cache) as m:
            out = demux_batch(m, ids, case["barcodes"], case["read_structures"], inputs, case["output_types"])
        got = {name: [[h.decode(), s.decode()] for h, s, _ in recs] for name, recs in out.files.items()}
        assert got == case["expect"], case["source"]
        for recs in out.files.values():
            assert all(q == b";" * len(s) for _, s, q in recs)
        if "expect_counts" in case:
            assert out.counts.tolist() == case["expect_counts"]
        assert int(out.counts.sum()) == len(case["inputs"][0])


def fastq_text(records):
    return "".join(f"@{h}\n{s}\n+\n{';' * len(s)}\n" for h, s in records).encode()


@pytest.mark.parametrize("use_cache", MODES)
def test_fastq_ingest_reference_end_to_end_vectors(kats, use_cache):
    """The reference's five end-to-end demux vectors again, this time from raw FASTQ TEXT: scanner -> per-read offsets ->
    B segments gathered and encoded on the GPU (fqtk_b200_matcher_assign_fastq) -> routing; no dense barcode rows and no
    per-record loop on the way in."""
    from fqtk_b200.fastq import demux_fastq_batch

    for case in kats["demux_e2e"]:
        S = len(case["barcodes"])
        ids = [f"Sample{j:04d}" for j in range(S)]
        texts = [fastq_text([(f"ex_{i}", b) for i, b in enumerate(col)]) for col in case["inputs"]]
        with BarcodeMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"], use_cache) as m:
            out = demux_fastq_batch(m, ids, case["barcodes"], case["read_structures"], texts, case["output_types"])
        got = {name: [[h.decode(), s.decode()] for h, s, _ in recs] for name, recs in out.files.items()}
        assert got == case["expect"], case["source"]
        for recs in out.files.values():
            assert all(q == b";" * len(s) for _, s, q in recs)
        if "expect_counts" in case:
            assert out.counts.tolist() == case["expect_counts"]
        assert int(out.counts.sum()) == len(case["inputs"][0])


def test_fastq_ingest_at_size_and_with_trailing_barcode_segments():
    """Dual-index layout of cfg 3 from two raw index FASTQ chunks (I1, I2: 8B each) + a read FASTQ, 300 k read sets with
    ragged headers: the device gather must give exactly the result words of dense barcode rows.  Then `+B` (the whole index
    read is the barcode, variable length): shorter barcodes are None, longer ones follow BarcodeMatcher::assign's panic /
    pre-filter rule, too-short reads fail like ReadSetIterator::next."""
    from fqtk_b200 import fastq
    from fqtk_b200.demux import parse_read_structure

    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    n = 300_007
    reads = synth.reads_host(panel, cfg.seed_reads, 1234, n)
    reads[::61, 2] = ord("n")
    want, want_counts = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta).assign_batch(reads)
    i1 = fastq_text([(f"r{i}:{'x' * (i % 13)} 1:N:0:0", bytes(reads[i, :8]).decode()) for i in range(n)])
    i2 = fastq_text([(f"r{i} 2:N:0:0", bytes(reads[i, 8:]).decode() + "AC"[: i % 3]) for i in range(n)])
    ix = [fastq.scan(i1), fastq.scan(i2)]
    segs = fastq.barcode_segments([parse_read_structure("8B"), parse_read_structure("8B+S")])
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, True) as m:
        got = fastq.assign_fastq(m, ix, segs)
        assert np.array_equal(got, want)
        assert np.array_equal(m.counts(), want_counts)
        # a read shorter than its segment: the reference panics in ReadSetIterator::next (demux.rs:309-315)
        short = fastq.scan(fastq_text([("a", "ACGTACG")]))
        with pytest.raises(_lib.Fqtk_b200Error, match="too few bases"):
            fastq.assign_fastq(m, [short, short], segs)
    # `+B`: variable-length barcodes
    L = 8
    bcs8 = [b"ACGTACGT", b"TTTTGGGG", b"CCCCAAAA"]
    seqs = ["ACGTACGT", "ACGTACG", "TTTTGGGG", "ACGTACGTA"[:8], "NNNNNNNNNN", "CCCCAAAA", "CC"]
    ix8 = [fastq.scan(fastq_text([(f"q{i}", s) for i, s in enumerate(seqs)]))]
    rest = fastq.barcode_segments([parse_read_structure("+B")])
    om = oracle.OracleMatcher(bcs8, 1, 1)
    with BarcodeMatcher(bcs8, 1, 1, True) as m:
        got = fastq.assign_fastq(m, ix8, rest)
        for i, sq in enumerate(seqs):
            w = int(got[i])
            exp = om.assign(sq.encode())  # the literal assign(): short -> None, all-N longer read -> None by the pre-filter
            assert (None if w == _lib.NONE else (w >> 16, (w >> 8) & 0xFF, w & 0xFF)) == \
                (None if exp is None else (exp.best_match, exp.best_mismatches, exp.next_best_mismatches)), (i, sq)
        assert int(m.counts().sum()) == len(seqs)
        long_ix = [fastq.scan(fastq_text([("z", "ACGTACGTAC")]))]
        with pytest.raises(MatcherPanic, match=r"length \(10\) differs from expected barcode \(ACGTACGT\) length \(8\)"):
            fastq.assign_fastq(m, long_ix, rest)
    assert L == 8


def test_device_fastq_scanner_equals_the_host_scanner():
    """fqtk_b200_fastq_scan_device against fqtk_b200_fastq_scan on the same chunks: LF and CRLF records, a record cut off by the
    chunk end, empty chunk, max_records, and the three malformed-record errors (same text, same first record)."""
    torch = torch_cuda()
    from fqtk_b200 import fastq
    rng = np.random.default_rng(5)

    def records(n, crlf=False, var=True):
        out = []
        for i in range(n):
            ln = int(rng.integers(1, 200)) if var else 8
            seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=ln))
            eol = b"\r\n" if crlf and i % 3 == 0 else b"\n"
            out.append(b"@r%d some comment" % i + eol + seq + eol + b"+" + (b"r%d" % i if i % 5 == 0 else b"") + eol + b"I" * ln + eol)
        return b"".join(out)

    def dev_scan(text, max_records=None):
        n_cap = max(1, len(text) // 7 + 1) if max_records is None else max_records
        d = torch.frombuffer(bytearray(text) if text else bytearray(1), dtype=torch.uint8).cuda()
        d_head = torch.zeros(n_cap, dtype=torch.int64, device="cuda")
        d_seq = torch.zeros(n_cap, dtype=torch.int64, device="cuda")
        d_len = torch.zeros(n_cap, dtype=torch.int32, device="cuda")
        n, used = fastq.scan_device(d.data_ptr(), len(text), n_cap, d_head.data_ptr(), d_seq.data_ptr(), d_len.data_ptr())
        return n, used, d_head[:n].cpu().numpy().astype(np.uint64), d_seq[:n].cpu().numpy().astype(np.uint64), \
            d_len[:n].cpu().numpy().astype(np.uint32)

    for text in (records(5000), records(3000, crlf=True), records(40000, var=False), records(7) + b"@cut\nACGT\n+\nII",
                 records(3)[:-1], b"", b"\n" * 3, records(70000) ):
        want = fastq.scan(text)
        n, used, head, seq, ln = dev_scan(text)
        assert (n, used) == (len(want), want.consumed)
        assert np.array_equal(head, want.head_offsets) and np.array_equal(seq, want.seq_offsets) and np.array_equal(ln, want.seq_lengths)
    text = records(1000)
    want = fastq.scan(text, 10)
    n, used, head, seq, ln = dev_scan(text, 10)
    assert (n, used) == (10, want.consumed) and np.array_equal(seq, want.seq_offsets)
    recs = [b"@r%d\nACGTACGT\n+\nIIIIIIII\n" % i for i in range(300)]

    def with_record(i, rec):
        return b"".join(recs[:i] + [rec] + recs[i + 1:])

    for bad, what in ((with_record(150, b"Xr150\nACGTACGT\n+\nIIIIIIII\n"), "record 150: header line"),
                      (with_record(200, b"@r200\nACGTACGT\n-\nIIIIIIII\n"), "record 200: separator line"),
                      (with_record(9, b"@r9\nACGTACGT\n+\nIIIIIII\n"), "record 9: sequence and quality lengths differ"),
                      (with_record(9, b"@r9\nACGTACGT\n+\nIIIIIII\n").replace(b"@r5\n", b"r5\n"), "record 5: header line")):
        with pytest.raises(_lib.Fqtk_b200Error) as host_err:
            fastq.scan(bad)
        with pytest.raises(_lib.Fqtk_b200Error) as dev_err:
            dev_scan(bad)
        assert what in dev_err.value.message
        assert dev_err.value.message == host_err.value.message


def test_assign_fastq_chunks_equals_scan_then_assign():
    """The one-call ingest (chunks -> device scan -> device rules -> gather / encode / match) against host scan +
    assign_fastq, on dual-index chunks with a trailing +B segment, carry-over at the chunk end, and the reference's two
    per-read errors (too few bases; barcode longer than the panel's, unless the no-call pre-filter gets there first)."""
    from fqtk_b200 import fastq
    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    n = 150_000
    reads = synth.reads_host(panel, cfg.seed_reads, 3, n)

    def fq_text(col_lo, col_hi, extra=b""):
        return b"".join(b"@q%d\n" % i + bytes(reads[i, col_lo:col_hi]) + extra + b"\n+\n" + b"F" * (col_hi - col_lo + len(extra)) + b"\n"
                        for i in range(n))

    i1, i2 = fq_text(0, 8), fq_text(8, 16, b"TTTT")  # I2 carries 4 template bases behind its 8 barcode bases
    segs_fixed = [(0, 0, 8), (1, 0, 8)]
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta) as m:
        want = fastq.assign_fastq(m, [fastq.scan(i1), fastq.scan(i2)], segs_fixed)
        m.reset_counts()
        got, used = fastq.assign_fastq_chunks(m, [i1, i2], segs_fixed)
        assert np.array_equal(got, want) and used == [len(i1), len(i2)]
        assert int(m.counts().sum()) == n
        # chunk ends in the middle of a record of one input: the records complete in BOTH inputs are matched
        cut1, cut2 = i1[: len(i1) // 2 + 5], i2[: len(i2) // 3 + 1]
        k = min(len(fastq.scan(cut1)), len(fastq.scan(cut2)))
        got, used = fastq.assign_fastq_chunks(m, [cut1, cut2], segs_fixed)
        assert got.shape[0] == k and np.array_equal(got, want[:k])
        assert used == [int(fastq.scan(cut1).head_offsets[k]) if k < len(fastq.scan(cut1)) else fastq.scan(cut1).consumed,
                        int(fastq.scan(cut2).head_offsets[k]) if k < len(fastq.scan(cut2)) else fastq.scan(cut2).consumed]
        got, _ = fastq.assign_fastq_chunks(m, [i1, i2], segs_fixed, max_reads=1000)
        assert np.array_equal(got, want[:1000])
        # errors: the table form's, for the first offending read
        short = i1.replace(b"@q777\n" + bytes(reads[777, 0:8]), b"@q777\n" + bytes(reads[777, 0:5]), 1).replace(
            b"@q777\n" + bytes(reads[777, 0:5]) + b"\n+\nFFFFFFFF", b"@q777\n" + bytes(reads[777, 0:5]) + b"\n+\nFFFFF", 1)
        with pytest.raises(Exception) as e1:
            fastq.assign_fastq_chunks(m, [short, i2], segs_fixed)
        with pytest.raises(Exception) as e2:
            fastq.assign_fastq(m, [fastq.scan(short), fastq.scan(i2)], segs_fixed)
        assert "Read 777 had too few bases to demux 5 vs. 8" in str(e1.value) and str(e1.value) == str(e2.value)
    # trailing +B: I2 = 8B then the rest as barcode too (12 bases) against a 20-base panel would differ; use a 12-base panel
    panel12 = np.concatenate([panel[:, 8:16], np.full((panel.shape[0], 4), ord("T"), dtype=np.uint8)], axis=1)
    with BarcodeMatcher([bytes(r) for r in panel12], 1, 2) as m:
        rest = [(0, 0, _lib.SEGMENT_REST)]
        want = fastq.assign_fastq(m, [fastq.scan(i2)], rest)
        got, _ = fastq.assign_fastq_chunks(m, [i2], rest)
        assert np.array_equal(got, want) and (want != _lib.NONE).mean() > 0.5
        longer = i2.replace(b"TTTT\n+\nFFFFFFFFFFFF\n", b"TTTTA\n+\nFFFFFFFFFFFFF\n", 1)  # read 0: 13 bases > 12
        with pytest.raises(Exception) as e1:
            fastq.assign_fastq_chunks(m, [longer], rest)
        with pytest.raises(Exception) as e2:
            fastq.assign_fastq(m, [fastq.scan(longer)], rest)
        assert "length (13) differs from expected barcode" in str(e1.value) and str(e1.value) == str(e2.value)
        nn = longer.replace(bytes(reads[0, 8:16]) + b"TTTTA", b"NNNNNNNNTTTTA", 1)  # the no-call pre-filter fires first: None
        got, _ = fastq.assign_fastq_chunks(m, [nn], rest)
        want = fastq.assign_fastq(m, [fastq.scan(nn)], rest)
        assert got[0] == _lib.NONE and np.array_equal(got, want)


def test_batched_pipeline_at_cfg1_scale():
    """cfg 1 (10 k single-end reads, 8B+T, 4 samples): every record lands in the file of the sample the oracle assigns,
    in input order, with the rewritten header; too-short reads are skipped and counted nowhere (demux.rs:2023-2073)."""
    from fqtk_b200.demux import TooFewBases, demux_batch

    cfg = synth.CONFIGS[1]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    n = cfg.n_reads
    reads = synth.reads_host(panel, cfg.seed_reads, 0, n)
    recs = [(f"r{i} 1:N:0:0".encode(), bytes(reads[i]) + b"A" * 50, b";" * 58) for i in range(n)]
    recs[17] = (b"short", b"ACGT", b";;;;")  # 4 bases < 8 + 1
    ids = [f"Sample{j:04d}" for j in range(cfg.n_samples)]
    om = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta)
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta) as m:
        with pytest.raises(TooFewBases):
            demux_batch(m, ids, [b.decode() for b in bcs], ["8B+T"], [recs])
        m.reset_counts()
        out = demux_batch(m, ids, [b.decode() for b in bcs], ["8B+T"], [recs], skip_too_few_bases=True)
    assert out.skipped == 1 and int(out.counts.sum()) == n - 1
    want_files = {}
    for i in range(n):
        if i == 17:
            continue
        r = om.assign_closed(bytes(reads[i]))
        prefix = "unmatched" if r is None else ids[r.best_match]
        want_files.setdefault(f"{prefix}.R1.fq.gz", []).append(
            (f"r{i} 1:N:0:".encode() + bytes(reads[i]), b"A" * 50, b";" * 50))
    assert out.files == want_files
    assert [mm.templates for mm in out.metrics] == out.counts.tolist()


def test_kernel_launch_counter_moves():
    from fqtk_b200.barcode_matching import kernel_launches

    before = kernel_launches()
    with BarcodeMatcher(["ACGT", "TTTT"], 1, 1) as m:
        m.assign(b"ACGT")
    assert kernel_launches() > before


# ---------------------------------------------------------------------------------------------------------
# SURVEY 8f "next" #1: B-segment gather + encode on the device (ReadSet::sample_barcode_sequence, demux.rs:121-123)
# ---------------------------------------------------------------------------------------------------------
def test_segments_reference_weird_read_structures(kats):
    """demux.rs:1739-1800: 4B4M8S / 4B100T / 100S3B / 6B1S1M1T -> AAAA+AAAA+GAT+TACAGA -> Sample0000."""
    case = [c for c in kats["demux_caller"] if "segments" in c][0]
    rows = [b"AAAACCCCGGGGTTTT", b"A" * 104, b"T" * 100 + b"GAT", b"TACAGAAAT"]
    arrays = [np.frombuffer(r, dtype=np.uint8).reshape(1, -1) for r in rows]
    segs = [(arrays[0], 0, 4), (arrays[1], 0, 4), (arrays[2], 100, 3), (arrays[3], 0, 6)]
    for use_cache in MODES:
        with BarcodeMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"], use_cache) as m:
            res = m.assign_segments(segs, 1)
            assert int(res[0]) >> 16 == 0 and res[0] != _lib.NONE
            assert m.counts().tolist() == case["expect_counts"]
            with pytest.raises(MatcherPanic, match="differs from expected barcode length"):
                m.assign_segments(segs[:3], 1)  # 11 bases for a 17-base panel


@pytest.mark.parametrize("use_cache", MODES)
def test_segments_equal_concatenated_rows(use_cache):
    torch = torch_cuda()
    rng = np.random.default_rng(21)
    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    n = 150_003
    reads = synth.reads_host(panel, cfg.seed_reads, 777, n)
    reads[::53, 5] = ord("n")
    reads[::211, 11] = ord("Y")
    # dual index: I1 = first 8 bases in its own 11-byte rows (offset 2), I2 = last 8 bases inside 151-byte R2 rows;
    # plus a three-piece split that takes two pieces from the same source
    i1 = np.full((n, 11), ord("#"), dtype=np.uint8)
    i1[:, 2:10] = reads[:, :8]
    r2 = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n, 151))
    r2[:, 100:108] = reads[:, 8:]
    three = np.full((n, 40), ord("-"), dtype=np.uint8)
    three[:, 30:35] = reads[:, :5]
    three[:, 1:4] = reads[:, 5:8]
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, use_cache) as m:
        want = m.assign_batch(reads)
        want_counts = m.counts()
        for segs in ([(i1, 2, 8), (r2, 100, 8)], [(three, 30, 5), (three, 1, 3), (r2, 100, 8)]):
            m.reset_counts()
            got = m.assign_segments(segs, n)
            assert np.array_equal(got, want)
            assert np.array_equal(m.counts(), want_counts)
        # device form
        m.reset_counts()
        d_i1, d_r2 = torch.from_numpy(i1).cuda(), torch.from_numpy(r2).cuda()
        d_res = torch.empty(n, dtype=torch.int32, device="cuda")
        m.assign_segments_device([(d_i1.data_ptr(), 11, 2, 8), (d_r2.data_ptr(), 151, 100, 8)], n, d_res.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_res.cpu().numpy().view(np.uint32), want)
        assert np.array_equal(m.counts(), want_counts)
    om = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta)
    assert np.array_equal(om.assign_batch(reads)[0], want)


# ---------------------------------------------------------------------------------------------------------
# SURVEY 8f "next" #3: per-sample routing = stable partition of read indices by assignment
# ---------------------------------------------------------------------------------------------------------
def expected_route(results, S):
    bucket = np.where(results == _lib.NONE, S, results >> 16).astype(np.int64)
    order = np.argsort(bucket, kind="stable").astype(np.uint32)
    offsets = np.zeros(S + 2, dtype=np.uint64)
    offsets[1:] = np.cumsum(np.bincount(bucket, minlength=S + 1))
    return order, offsets


@pytest.mark.parametrize("cfg_id,n", [(1, 10_000), (3, 400_037), (5, 150_001)])
def test_route_is_the_stable_partition(cfg_id, n):
    cfg = synth.CONFIGS[cfg_id]
    panel = synth.panel(cfg)
    reads = synth.reads_host(panel, cfg.seed_reads, 5, n)
    with BarcodeMatcher([bytes(r) for r in panel], cfg.max_mismatches, cfg.min_mismatch_delta) as m:
        res = m.assign_batch(reads)
        order, offsets = m.route(res)
        want_order, want_offsets = expected_route(res, cfg.n_samples)
        assert np.array_equal(offsets, want_offsets)
        assert np.array_equal(order, want_order)
        assert np.array_equal(np.diff(offsets.astype(np.int64)), m.counts().astype(np.int64))
        # every sample's run is in input order (what the reference's per-read writes guarantee, demux.rs:1505-1523)
        for j in (0, cfg.n_samples // 2, cfg.n_samples):
            run = order[int(offsets[j]):int(offsets[j + 1])]
            assert np.all(np.diff(run.astype(np.int64)) > 0)
        # tiny and empty batches
        o1, f1 = m.route(res[:1])
        assert o1.tolist() == [0] and int(f1[-1]) == 1
        o0, f0 = m.route(res[:0])
        assert o0.size == 0 and not f0.any()


@pytest.mark.parametrize("S", [1, 5, 384, 1000, 1536, 3000])
def test_route_kernel_versions_and_edge_sizes(S):
    """Routing only looks at result words, so any word stream drives it: skewed buckets, None words, sizes around the
    8 192-read tile and the 32-read step, a result pointer that is not 16-byte aligned (scalar histogram loads).
    S = 1 ... 1 000 take the mask-rank tile kernel, 1 536 the ballot tile kernel, 3 000 the per-warp version."""
    torch = torch_cuda()
    rng = np.random.default_rng(1000 + S)
    panel = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(S, 12))
    panel = np.unique(panel, axis=0)
    assert panel.shape[0] >= min(S, 3000) - 8
    S = panel.shape[0]
    stream = torch.cuda.current_stream().cuda_stream
    with BarcodeMatcher([bytes(r) for r in panel], 0, 0, use_cache=False) as m:
        for n in (1, 31, 32, 33, 511, 8191, 8192, 8193, 3 * 8192 + 5, 300 * 8192 + 17, 2_500_003):
            b = rng.integers(0, S, size=n)
            b[rng.random(n) < 0.15] = 0          # one heavy sample
            none = rng.random(n) < 0.1
            res = ((b.astype(np.uint32) << 16) | rng.integers(0, 3, size=n).astype(np.uint32) << 8 | 2).astype(np.uint32)
            res[none] = _lib.NONE
            want_order, want_offsets = expected_route(res, S)
            order, offsets = m.route(res)
            assert np.array_equal(offsets, want_offsets), (S, n)
            assert np.array_equal(order, want_order), (S, n)
            # device entry with a pointer 4 bytes past a 16-byte boundary
            d = torch.empty(n + 1, dtype=torch.int32, device="cuda")
            d[1:] = torch.from_numpy(res.view(np.int32)).cuda()
            d_order = torch.empty(n, dtype=torch.int32, device="cuda")
            d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
            m.route_device(d.data_ptr() + 4, n, d_order.data_ptr(), d_off.data_ptr(), stream)
            torch.cuda.synchronize()
            assert np.array_equal(d_order.cpu().numpy().view(np.uint32), want_order), (S, n, "unaligned")
            assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want_offsets)


def test_route_full_size_properties():
    """cfg 3 at 500 M reads on the device: order is a permutation, every bucket run is ascending and homogeneous."""
    torch = torch_cuda()
    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    n = cfg.n_reads
    stream = torch.cuda.current_stream().cuda_stream
    d_packed = torch.empty((n, cfg.words_per_read), dtype=torch.int32, device="cuda")
    synth.reads_device(panel, cfg.seed_reads, 0, n, 0, d_packed.data_ptr(), stream)
    d_res = torch.empty(n, dtype=torch.int32, device="cuda")
    d_order = torch.empty(n, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(cfg.n_samples + 2, dtype=torch.int64, device="cuda")
    with BarcodeMatcher([bytes(r) for r in panel], cfg.max_mismatches, cfg.min_mismatch_delta) as m:
        m.assign_packed_device(d_packed.data_ptr(), n, d_res.data_ptr(), stream)
        del d_packed
        m.route_device(d_res.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr(), stream)
        torch.cuda.synchronize()
        counts = m.counts().astype(np.int64)
    off = d_off.cpu().numpy()
    assert np.array_equal(np.diff(off), counts) and off[0] == 0 and off[-1] == n
    # homogeneous runs: the bucket of results[order[k]] is non-decreasing and matches the offsets table
    gathered = d_res[d_order.long()]
    bucket = torch.where(gathered == -1, torch.full_like(gathered, cfg.n_samples), (gathered >> 16) & 0xFFFF)
    assert bool((bucket[1:] >= bucket[:-1]).all())
    # ascending inside every run  <=>  order[k+1] > order[k] wherever the bucket does not change
    same = bucket[1:] == bucket[:-1]
    assert bool((d_order[1:][same] > d_order[:-1][same]).all())
    del gathered, bucket, same
    # permutation: a stable partition of [0, n) sums to n(n-1)/2 and has no repeats inside runs (checked above)
    assert int(d_order.long().sum().item()) == n * (n - 1) // 2


# ---------------------------------------------------------------------------------------------------------
# edges of the device entry points: tile boundaries, tiny batches, unaligned buffers, empty calls
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("use_cache", MODES)
def test_batch_sizes_around_the_tile_boundaries(use_cache):
    torch = torch_cuda()
    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    bcs = [bytes(r) for r in panel]
    reads = synth.reads_host(panel, cfg.seed_reads, 99, 1200)
    reads[::7, 2] = ord("K")  # out-of-alphabet reads in every tile, including the tail warp
    packed = synth.pack_host(reads)
    om = oracle.OracleMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta)
    want_all, _ = om.assign_batch(reads)
    stream = torch.cuda.current_stream().cuda_stream
    with BarcodeMatcher(bcs, cfg.max_mismatches, cfg.min_mismatch_delta, use_cache) as m:
        d_all = torch.from_numpy(packed.view(np.int32)).cuda()
        for n in (0, 1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 1023, 1199):
            m.reset_counts()
            d_res = torch.full((max(n, 1),), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
            m.assign_packed_device(d_all.data_ptr(), n, d_res.data_ptr(), stream)
            torch.cuda.synchronize()
            got = d_res.cpu().numpy().view(np.uint32)[:n]
            assert np.array_equal(got, want_all[:n]), n
            c = m.counts()
            assert int(c.sum()) == n
        # result buffer only 4-byte aligned: the one-read-per-thread kernel takes over, same answers
        n = 1000
        d_res = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
        m.reset_counts()
        m.assign_packed_device(d_all.data_ptr(), n, d_res.data_ptr() + 4, stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_res.cpu().numpy().view(np.uint32)[1:], want_all[:n])
        # packed input must be 16-byte aligned: refused, nothing counted
        with pytest.raises(_lib.Fqtk_b200Error):
            m.assign_packed_device(d_all.data_ptr() + 8, 10, d_res.data_ptr(), stream)
        assert int(m.counts().sum()) == n
        # empty host batch is a no-op
        out = m.assign_batch(np.zeros((0, cfg.barcode_len), dtype=np.uint8))
        assert out.shape == (0,)


def test_single_sample_and_single_base_panels():
    for bcs, mm, delta, reads in [
        (["A"], 0, 0, [b"A", b"C", b"N", b"a", b"-"]),
        (["N"], 0, 0, [b"A", b"N", b"."]),
        (["ACGTACGT"], 1, 2, [b"ACGTACGT", b"ACGTACGA", b"TTTTTTTT", b"NNNNNNNN"]),
        (["AC", "AG", "AT", "AA"], 0, 1, [b"AC", b"AG", b"AN", b"NN", b"TT"]),
    ]:
        om = oracle.OracleMatcher(bcs, mm, delta)
        for use_cache in MODES:
            with BarcodeMatcher(bcs, mm, delta, use_cache) as m:
                for r in reads:
                    got, want = m.assign(r), om.assign_closed(r)
                    as_tuple = lambda x: None if x is None else (x.best_match, x.best_mismatches, x.next_best_mismatches)
                    assert as_tuple(got) == as_tuple(want), (bcs, r, use_cache)
