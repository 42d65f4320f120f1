"""Pins the CPU oracle (oracle/fqtk_oracle.c) against every known-answer test the reference holds for the
demux matcher path (tests/golden/reference_kats.json, transcribed from the reference's #[cfg(test)] modules),
then fuzzes the literal restatement against the closed form the GPU computes.  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import BarcodeMatch, OracleMatcher, OraclePanic


def test_iupac_mask_table(kats):
    table = kats["iupac_masks"]["table"]
    for b in range(256):
        expect = table.get(chr(b), 0)
        assert oracle.lib().fqo_iupac_mask(b) == expect, chr(b)


def test_byte_is_nocall(kats):
    for c in kats["nocall_bytes"]["nocall"]:
        assert oracle.lib().fqo_byte_is_nocall(ord(c)) == 1
    for c in kats["nocall_bytes"]["not_nocall"]:
        assert oracle.lib().fqo_byte_is_nocall(ord(c)) == 0


def test_is_valid_iupac(kats):
    for c in kats["valid_iupac"]["valid"]:
        assert oracle.lib().fqo_is_valid_iupac(ord(c)) == 1, c
    for c in kats["valid_iupac"]["invalid"]:
        assert oracle.lib().fqo_is_valid_iupac(ord(c)) == 0, c


def test_encode_decode_roundtrip(kats):
    table = kats["iupac_masks"]["table"]
    for s in kats["encode_decode_roundtrip"]["strings"]:
        blocks, n = oracle.encode(s.encode())
        assert n == len(s)
        for i, ch in enumerate(s):
            assert (blocks[i // 8] >> (4 * (i % 8))) & 0xF == table[ch]
        assert oracle.decode_blocks(blocks, n) == s
    nc = kats["encode_decode_roundtrip"]["nocalls"]
    blocks, n = oracle.encode(nc["input"].encode())
    assert all(((blocks[0] >> (4 * i)) & 0xF) == nc["mask"] for i in range(n))
    assert oracle.decode_blocks(blocks, n) == nc["decoded"]


def test_bitenc_layout_lsb_first():
    # bitenc.rs:311-322: value i at bits 4*(i%8) of block i/8; unused high nibbles stay 0
    blocks, n = oracle.encode(b"ACGTACGTAC")
    assert n == 10 and len(blocks) == 2
    assert blocks[0] == 0x84218421
    assert blocks[1] == 0x00000021


def test_bitenc_hamming(kats):
    seqs = kats["bitenc_hamming"]["sequences"]
    for a, b, cap, want in kats["bitenc_hamming"]["cases"]:
        assert oracle.hamming_nibbles(seqs[a], seqs[b], cap) == want, (a, b, cap)


def test_hamming_is_min_of_exact_and_cap():
    rng = np.random.default_rng(7)
    for _ in range(2000):
        n = int(rng.integers(0, 40))
        a = rng.integers(0, 16, n).tolist()
        b = rng.integers(0, 16, n).tolist()
        exact = sum(1 for x, y in zip(a, b) if x & ~y & 0xF)
        cap = int(rng.integers(0, 12))
        assert oracle.hamming_nibbles(a, b, cap) == min(exact, cap)
        assert oracle.hamming_nibbles(a, b, 255) == exact


def test_count_mismatches(kats):
    for case in kats["count_mismatches"]:
        got = oracle.count_mismatches(case["observed"].encode(), case["expected"].encode())
        assert got == case["mismatches"], case


def test_count_mismatches_length_panics(kats):
    for case in kats["count_mismatches_panics"]:
        with pytest.raises(OraclePanic) as ei:
            oracle.count_mismatches(case["observed"].encode(), case["expected"].encode())
        assert str(ei.value) == case["message"]


def test_matcher_instantiation(kats):
    ok = kats["matcher_new"]["ok"]
    OracleMatcher(ok["barcodes"], ok["max_mismatches"], ok["min_mismatch_delta"])
    with pytest.raises(OraclePanic, match=kats["matcher_new"]["panic_empty"]["message"]):
        OracleMatcher([], 2, 1)
    with pytest.raises(OraclePanic, match="Sample barcode cannot be empty string"):
        OracleMatcher([""], 2, 1)


@pytest.mark.parametrize("use_cache", [True, False])
def test_assign_kats(kats, use_cache):
    for case in kats["assign"]:
        m = OracleMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"], use_cache)
        want = None if case["expect"] is None else BarcodeMatch(*case["expect"])
        read = case["read"].encode()
        assert m.assign(read) == want, case
        assert m.assign(read) == want, "second call goes through the memo cache"
        assert m.assign_internal(read) == want
        assert m.assign_closed(read) == want


@pytest.mark.parametrize("use_cache", [True, False])
def test_demux_caller_vectors(kats, use_cache):
    for case in kats["demux_caller"]:
        m = OracleMatcher(case["barcodes"], case["max_mismatches"], case["min_mismatch_delta"], use_cache)
        if "segments" in case:  # demux.rs:121-123 concatenates B segments in input order
            assert "".join(case["segments"]) == case["reads"][0]
        reads = np.frombuffer("".join(case["reads"]).encode(), dtype=np.uint8).reshape(len(case["reads"]), -1)
        for mode in (0, 1):
            res, counts = m.assign_batch(reads, mode=mode)
            got = [None if r == oracle.NONE else int(r) >> 16 for r in res]
            assert got == case["expect_sample"], case["source"]
            assert counts.tolist() == case["expect_counts"], case["source"]
    # single-sample panel keeps the 255 sentinel (barcode_matching.rs:122)
    m = OracleMatcher(["NNNNNNN"], 0, 2)
    assert m.assign(b"NNNNNNN") == BarcodeMatch(0, 0, 255)


def test_short_and_long_reads():
    m = OracleMatcher(["ACGTAC", "TTTTTT"], 1, 1)
    assert m.assign(b"ACGTA") is None  # barcode_matching.rs:167-169
    with pytest.raises(OraclePanic):  # :95-106
        m.assign(b"ACGTACG")
    # the no-call pre-filter (:170-172) fires before the length panic
    assert m.assign(b"NNNNNNN") is None
    assert m.assign_closed(b"NNNNNNN") is None
    with pytest.raises(OraclePanic):
        m.assign_closed(b"ACGTACG")


def test_cache_only_stores_some():
    m = OracleMatcher(["AAAA", "CCCC"], 1, 1, use_cache=True)
    assert m.assign(b"GGGG") is None
    assert m.cache_len == 0
    assert m.assign(b"AAAA") == BarcodeMatch(0, 0, 4)
    assert m.assign(b"AAAA") == BarcodeMatch(0, 0, 4)
    assert m.cache_len == 1


ALPHABETS = {
    "acgt": b"ACGT",
    "acgtn": b"ACGTN",
    "iupac": b"ACGTUMRWSYKVHDBNn.",
    "dirty": b"ACGTNn.acgtRYKMXx-*0 ",
}


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_literal_equals_closed_form(seed):
    """cap, pre-filter and cache never change the answer."""
    rng = np.random.default_rng(1000 + seed)
    cases = 0
    for _ in range(250):
        L = int(rng.choice([1, 2, 4, 7, 8, 9, 15, 16, 17, 20, 24, 31, 32, 33, 40]))
        S = int(rng.choice([1, 2, 3, 8, 40, 100]))
        panel_alpha = ALPHABETS[["acgt", "acgtn", "iupac"][int(rng.integers(0, 3))]]
        seen, bcs = set(), []
        while len(bcs) < S:
            if rng.random() < 0.5 and bcs:  # near-duplicates make ties and small deltas common
                b = bytearray(bcs[int(rng.integers(0, len(bcs)))])
                b[int(rng.integers(0, L))] = panel_alpha[int(rng.integers(0, len(panel_alpha)))]
                b = bytes(b)
            else:
                b = bytes(panel_alpha[i] for i in rng.integers(0, len(panel_alpha), L))
            if b not in seen:
                seen.add(b)
                bcs.append(b)
            elif len(seen) >= len(panel_alpha) ** L:
                break
        mm = int(rng.choice([0, 1, 2, 3, 100]))
        delta = int(rng.choice([0, 1, 2, 3, 100]))
        lit_c = OracleMatcher(bcs, mm, delta, use_cache=True)
        lit_n = OracleMatcher(bcs, mm, delta, use_cache=False)
        read_alpha = ALPHABETS[["acgt", "acgtn", "dirty"][int(rng.integers(0, 3))]]
        for _ in range(40):
            if rng.random() < 0.7:
                r = bytearray(bcs[int(rng.integers(0, len(bcs)))])
                for _ in range(int(rng.integers(0, 4))):
                    r[int(rng.integers(0, L))] = read_alpha[int(rng.integers(0, len(read_alpha)))]
                r = bytes(r)
            else:
                r = bytes(read_alpha[i] for i in rng.integers(0, len(read_alpha), L))
            want = lit_n.assign_closed(r)
            assert lit_n.assign(r) == want
            assert lit_c.assign(r) == want
            assert lit_c.assign(r) == want
            assert lit_n.assign_internal(r) == want
            cases += 1
    assert cases == 250 * 40


def test_batch_mt_equals_single_thread():
    rng = np.random.default_rng(5)
    S, L = 24, 10
    panel = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(S, L))
    panel = np.unique(panel, axis=0)
    S = panel.shape[0]
    reads = panel[rng.integers(0, S, 20000)].copy()
    flip = rng.random(reads.shape) < 0.05
    reads[flip] = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=int(flip.sum()))
    m = OracleMatcher([bytes(r) for r in panel], 1, 2)
    res1, c1 = m.assign_batch(reads)
    res2, c2, used = oracle.assign_batch_mt(panel, 1, 2, reads, threads=3)
    assert used >= 1
    assert np.array_equal(res1, res2) and np.array_equal(c1, c2)
    assert int(c1.sum()) == reads.shape[0]


def test_oracle_workload_generator_is_the_products_generator():
    """bench.py's CPU legs (cpu_baseline, --impl reference) take panel and reads from the oracle's own copy of the
    synthetic workload generator (oracle/fqtk_synth.c), so that the reference arm loads nothing of the product; it must
    produce, byte for byte, the stream of fqtk_b200/csrc/synth.cu that the GPU arm measures."""
    from fqtk_b200 import synth

    for cid in (1, 2, 3, 5):
        cfg = synth.CONFIGS[cid]
        want_panel = synth.panel(cfg)
        got_panel = oracle.synth_panel(cfg.seed_panel, cfg.n_samples, cfg.barcode_len, cfg.min_distance, cfg.n_degenerate)
        assert np.array_equal(got_panel, want_panel), cid
        for first, n in ((0, 5000), (cfg.n_reads - 777, 777), (123_456_789 % cfg.n_reads, 3000)):
            assert np.array_equal(oracle.synth_reads(got_panel, cfg.seed_reads, first, n),
                                  synth.reads_host(want_panel, cfg.seed_reads, first, n)), (cid, first)
