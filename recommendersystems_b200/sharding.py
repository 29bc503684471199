"""Seed sharding across ranks (mode "seed batches"): the graph is replicated on every GPU, rank j takes the
contiguous block [j*S/P, (j+1)*S/P) of the seed list, no collective runs during the iteration, and the per-seed top-k
lists are concatenated in seed order at the end -- the GPU analogue of the reference's one-thread-per-ego-network
(`Program.cs:61-66`).  Pure host logic: works with any torch.distributed backend (NCCL on the GPUs, gloo in the tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_bounds(n_seeds: int, world: int) -> List[int]:
    """First seed index of every rank's block (+ the end): block j = [b[j], b[j+1])."""
    return [(j * n_seeds) // world for j in range(world + 1)]


def shard_seeds(seeds: Sequence[int], rank: int, world: int) -> np.ndarray:
    b = shard_bounds(len(seeds), world)
    return np.ascontiguousarray(np.asarray(seeds, np.int32)[b[rank]:b[rank + 1]])


def recommend_sharded(recommend_batch: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray, np.ndarray]], seeds: Sequence[int],
                      k: int, dist=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Runs `recommend_batch(local_seeds) -> (ids[m,k], scores[m,k], counts[m])` on this rank's block and returns the
    lists of ALL seeds, in seed order, on every rank.  `dist` is torch.distributed (initialised) or None for one rank."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    local = shard_seeds(seeds, rank, world)
    if len(local):
        ids, sc, cnt = recommend_batch(local)
    else:
        ids, sc, cnt = np.zeros((0, k), np.int64), np.zeros((0, k), np.float64), np.zeros(0, np.int32)
    if world == 1:
        return ids, sc, cnt
    parts: List = [None] * world
    dist.all_gather_object(parts, (np.asarray(ids), np.asarray(sc), np.asarray(cnt)))
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
            np.concatenate([p[2] for p in parts]))
