"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for collectives.

Walks shard with NO data-path collective: walkers are independent (the reference itself
partitions rows arbitrarily, fugue.py:146-150) and the Philox stream is keyed by the global
walk id, so the union of the shards equals the single-GPU result.  SGNS is data-parallel over
walk shards with replicated tables: token counts are sum-reduced once, the two embedding tables
are averaged (sum-allreduce over NVLink, then x 1/G) at a fixed interval.
Everything here is backend-agnostic so the bookkeeping is tested with gloo on CPU.
"""
from typing import List, Sequence, Tuple

import torch


def world(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_bounds(n: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [lo, hi) ranges of n items over world_size ranks."""
    return [(n * r // world_size, n * (r + 1) // world_size) for r in range(world_size)]


def shard_start_vertices(start: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Rank's share of the (ascending) start-vertex list: "walks shard by start vertex"."""
    lo, hi = shard_bounds(int(start.numel()), world_size)[rank]
    return start[lo:hi]


def shard_layout(n_local_walks: int, n_local_rows: int, group=None) -> Tuple[int, int, int]:
    """(global index of this rank's first walk, total walks, table rows = max id + 1 over ranks)."""
    import torch.distributed as dist
    rank, size = world(group)
    if size == 1:
        return 0, n_local_walks, n_local_rows
    sizes = [None] * size
    dist.all_gather_object(sizes, (int(n_local_walks), int(n_local_rows)), group=group)
    return sum(s[0] for s in sizes[:rank]), sum(s[0] for s in sizes), max(s[1] for s in sizes)


def reduce_vocab(counts: torch.Tensor, first_pos: torch.Tensor, group=None) -> None:
    """In place: global token counts (sum) and global first-appearance positions (min)."""
    import torch.distributed as dist
    if world(group)[1] == 1:
        return
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(first_pos, op=dist.ReduceOp.MIN, group=group)


def average_tables(tables: Sequence[torch.Tensor], group=None, scale=None) -> None:
    """Model averaging in place: sum-allreduce then multiply by 1/G.  `scale(t, f)` lets the
    caller run the multiply in the CUDA library (n2v_scale); default is torch."""
    import torch.distributed as dist
    size = world(group)[1]
    if size == 1:
        return
    for t in tables:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if scale is not None and t.is_cuda:
            scale(t, 1.0 / size)
        else:
            t.mul_(1.0 / size)
