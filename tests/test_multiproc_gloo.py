"""N > 1 host logic on CPU: world_size 2 over gloo.  Each rank takes its shard of one synthetic read stream, the
per-shard counts (computed here by the oracle, standing in for the GPU kernel there is no device for) are summed with
the product's all_reduce_counts, and the total must equal the single-process counts."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, weak, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from fqtk_b200 import synth
    from fqtk_b200.distributed import all_reduce_counts, shard_bounds, weak_shard_first_read

    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    if weak:
        lo, hi = weak_shard_first_read(n_total, rank), weak_shard_first_read(n_total, rank) + n_total
    else:
        lo, hi = shard_bounds(n_total, rank, world)
    reads = synth.reads_host(panel, cfg.seed_reads, lo, hi - lo)
    m = oracle.OracleMatcher([bytes(r) for r in panel], cfg.max_mismatches, cfg.min_mismatch_delta)
    _, counts = m.assign_batch(reads, want_results=False)
    t = torch.from_numpy(counts.astype(np.int64))
    all_reduce_counts(t)
    if rank == 0:
        np.save(os.path.join(out_dir, "counts.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("weak", [False, True])
def test_two_rank_count_reduce_matches_single_process(tmp_path, weak):
    import oracle
    from fqtk_b200 import synth

    n = 30_001
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, weak, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "counts.npy")
    cfg = synth.CONFIGS[3]
    panel = synth.panel(cfg)
    total = n * world if weak else n
    reads = synth.reads_host(panel, cfg.seed_reads, 0, total)
    m = oracle.OracleMatcher([bytes(r) for r in panel], cfg.max_mismatches, cfg.min_mismatch_delta)
    _, want = m.assign_batch(reads, want_results=False)
    assert np.array_equal(got, want.astype(np.int64))
    assert int(got.sum()) == total


def test_shard_bounds_tile_the_range():
    from fqtk_b200.distributed import shard_bounds

    for n in (0, 1, 7, 1000, 500_000_000, 10**9 + 7):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
