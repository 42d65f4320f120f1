"""Multi-GPU plumbing for the matcher: the path shards embarrassingly — rank r owns a contiguous read
range, the panel is replicated, and the ONLY exchange is one all-reduce(sum) of the per-sample count table
(u64[S+1], 3 KB at S = 384) after the last batch.  torch.distributed is plumbing here, nothing more."""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n_reads: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank` of `world` over `n_reads` reads (strong-scaling split).
    Ranges tile [0, n_reads) exactly and differ in size by at most one read."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n_reads * rank // world, n_reads * (rank + 1) // world


def weak_shard_first_read(reads_per_rank: int, rank: int) -> int:
    """First read index of a rank when every rank processes `reads_per_rank` reads of one stream (weak scaling)."""
    return reads_per_rank * rank


def tensor_from_device_ptr(ptr: int, n: int, device):
    """int64 torch view (no copy) of a raw device pointer — the matcher's own u64[S+1] count table."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=device)


def all_reduce_counts(counts):
    """Sum the per-sample count table over all ranks, in place (integer sum: order-independent, bit-exact).
    `counts`: int64 tensor [S+1] on the rank's device (cuda + NCCL in production, cpu + gloo in the CPU tests)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
