"""Multi-GPU host logic.  The volumetric path shards by stereo pair: every frame is independent
("batch: fully independent per sample"), so ranks own disjoint pairs and the data
path needs NO collective; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used for the
barrier and for the max-over-ranks timing only."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def pairs_for_rank(total_pairs: int, rank: int, world: int) -> List[int]:
    """Round-robin partition of stereo pairs: rank r owns pairs r, r+world, ..."""
    return list(range(rank, total_pairs, world))


def max_over_ranks(ms: torch.Tensor) -> torch.Tensor:
    """Job time = slowest rank (device-timed milliseconds)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms


def whole_job_voxels_per_s(voxels_per_pair: int, pairs_per_rank: int, world: int, steps: int, ms_total: float) -> float:
    """Aggregate throughput over all ranks (weak scaling: per-rank work is fixed)."""
    return voxels_per_pair * pairs_per_rank * world * steps / (ms_total * 1e-3)
