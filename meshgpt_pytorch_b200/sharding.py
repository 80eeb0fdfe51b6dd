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
    for ragged collections such as BASELINE config 4.  Cut r is placed where the running face count is closest to
    r / world of the total (from either side), every rank keeping at least one mesh while there are enough."""
    n = len(face_counts)
    prefix = [0]
    for c in face_counts:
        prefix.append(prefix[-1] + int(c))
    total = prefix[-1]
    cuts, start = [], 0
    for r in range(world):
        if r == world - 1:
            end = n
        else:
            target = total * (r + 1) / world
            lo = min(n, start + 1) if n - start > world - r - 1 else start        # at least one mesh, if the others can still get one
            hi = max(lo, n - (world - r - 1))                                     # leave one mesh for each remaining rank
            end = min(range(lo, hi + 1), key=lambda e: (abs(prefix[e] - target), e))
        cuts.append((start, end))
        start = end
    return cuts


def allreduce_histogram(hist: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of the codebook-usage histogram (no-op without an initialised process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist
