"""Multi-GPU partitioning of the forward pass: one process per GPU, no data-path collective.

The reference has no distributed code at all (only ``nn.DataParallel`` in its trainer, trainer.py:375);
its batch driver iterates ``n`` frames as ``n-1`` sequential pairs (src/datasets.py:456-463, run.py:159).
Independent pairs shard across ranks with zero communication: rank r takes a
contiguous block of the pair list.  In sequence mode pair i = (frame i, frame i+1), so a rank's block
of pairs [a, b) needs frames [a, b] INCLUSIVE of the shared boundary frame.
"""
from __future__ import annotations

from typing import List, Tuple


def pair_range(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced block [lo, hi) of pair indices owned by ``rank`` (sizes differ by at most 1)."""
    if not (0 <= rank < world) or n_pairs < 0:
        raise ValueError("bad rank/world/n_pairs")
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Frames [lo, hi] (inclusive) a rank must read to process its block of the n_frames-1 sequential pairs."""
    lo, hi = pair_range(max(n_frames - 1, 0), rank, world)
    return (lo, hi) if hi > lo else (lo, lo - 1)


def row_slabs(H: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """Row slabs [r0, r1) of a frame of height H for spatial tiling; boundaries are multiples of ``align``
    (32 = 2**5 keeps every pyramid level's slab boundary on a whole row)."""
    if H % align:
        raise ValueError("H must be a multiple of align")
    units = H // align
    out = []
    for r in range(world):
        lo, hi = pair_range(units, r, world)
        out.append((lo * align, hi * align))
    return out


def gather_counts(count: int, elapsed_ms: float):
    """All-reduce (sum of units, max of time) across ranks for throughput reporting; identity when
    torch.distributed is not initialised.  This is the ONLY collective of the pair-sharded mode."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return count, elapsed_ms
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(count)], dtype=torch.float64, device=dev)
    m = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return int(round(t.item())), float(m.item())
