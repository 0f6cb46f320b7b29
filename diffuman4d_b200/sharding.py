"""Multi-GPU host logic.

Round 1 strategy = the reference's own: tasks of one alternation round are independent units
(src/samplers/sliding_iterative_sampler.py:192-199) pulled by one worker per GPU (src/samplers/sampling_runner.py:26-43)
with a barrier per round.  Here: one *process* per GPU (torch.distributed), a static balanced partition of the round's
tasks, and one all-gather of the updated latent/timestep-index grid entries per round (replaces the CPU dict + Lock of
SAMP:91-97,181-185).  No collective on the data path of a window step.

``frame_shard`` is the partition used by the frame-sharded window design.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_tasks(n_tasks: int, rank: int, world: int) -> List[int]:
    """Balanced contiguous partition: sizes differ by at most one (44 tasks / 8 ranks -> 6,6,6,6,5,5,5,5)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_tasks, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def frame_shard(num_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Frames [lo, hi) of a window owned by ``rank`` in the frame-sharded design; F must divide evenly so every rank
    launches the same plan."""
    if num_frames % world != 0:
        raise ValueError(f"num_frames ({num_frames}) must be divisible by the number of ranks ({world})")
    per = num_frames // world
    return rank * per, (rank + 1) * per


def exchange_grid_updates(keys: Sequence[Tuple[int, int]], latents: torch.Tensor, timestep_indices: torch.Tensor,
                          group=None) -> Dict[Tuple[int, int], Tuple[torch.Tensor, int]]:
    """All-gather the (spa, tem) grid cells this rank updated in the round.

    keys: the grid coordinates of the rows of ``latents`` [n, 4, h, w] / ``timestep_indices`` [n].  Returns the union over
    all ranks.  Ranks may contribute different counts (padded to the max for the collective)."""
    world = dist.get_world_size(group)
    n = torch.tensor([len(keys)], dtype=torch.int64, device=latents.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    nmax = int(max(c.item() for c in counts))
    shape = latents.shape[1:]
    pad_lat = torch.zeros((nmax, *shape), dtype=latents.dtype, device=latents.device)
    pad_key = torch.full((nmax, 3), -1, dtype=torch.int64, device=latents.device)
    if len(keys):
        pad_lat[:len(keys)] = latents
        pad_key[:len(keys), :2] = torch.tensor(list(keys), dtype=torch.int64, device=latents.device)
        pad_key[:len(keys), 2] = timestep_indices.to(torch.int64)
    all_lat = [torch.empty_like(pad_lat) for _ in range(world)]
    all_key = [torch.empty_like(pad_key) for _ in range(world)]
    dist.all_gather(all_lat, pad_lat, group=group)
    dist.all_gather(all_key, pad_key, group=group)
    out: Dict[Tuple[int, int], Tuple[torch.Tensor, int]] = {}
    for r in range(world):
        for i in range(int(counts[r].item())):
            s, t, ti = (int(v) for v in all_key[r][i])
            out[(s, t)] = (all_lat[r][i], ti)
    return out
