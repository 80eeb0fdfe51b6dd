"""Sharding of a mesh collection across the GPUs of one box.

Every mesh is independent end to end, so ranks take contiguous blocks of mesh indices and never exchange
activations.  The only collective is the sum of the per-rank codebook-usage histograms
(int64 [num_quantizers, codebook_size], 256 KiB at the MeshGPT shape) - one NCCL all-reduce over NVLink."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first n % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_contiguous_shards(face_counts: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous cuts that balance the NUMBER OF FACES (the cost driver) rather than the number of meshes -
    for ragged collections such as BASELINE config 4."""
    total = sum(face_counts)
    cuts, start, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        end = start
        while end < len(face_counts) and (acc + face_counts[end] <= target or end == start) and len(face_counts) - end > world - r - 1:
            acc += face_counts[end]
            end += 1
        if r == world - 1:
            acc += sum(face_counts[end:])
            end = len(face_counts)
        cuts.append((start, end))
        start = end
    return cuts


def allreduce_histogram(hist: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of the codebook-usage histogram (no-op without an initialised process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist
