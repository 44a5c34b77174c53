"""Episode sharding across GPUs: episodes are independent, frames within an episode are serial,
so rank r owns episodes {e : e % world_size == r} with no collective on the hot path.  The only exchange
is one all_reduce of a few counters after the loop (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_episodes(n_episodes: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment ``episode % world_size == rank`` (stable under growing n_episodes)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_episodes, world_size))


def shard_scenes(sequence_names: Sequence[str], rank: int, world_size: int) -> List[int]:
    """For TEST_TYPE default/longterm the memory persists across the sequences of a scene (reset only at
    seq 0, SMNet/loader.py:289-291), so whole scenes (``name[:13]``) are kept on one rank."""
    scenes: Dict[str, int] = {}
    for n in sequence_names:
        scenes.setdefault(n[:13], len(scenes))
    return [i for i, n in enumerate(sequence_names) if scenes[n[:13]] % world_size == rank]


def init_from_env(backend: str = "nccl") -> tuple:
    """(rank, local_rank, world_size) from torchrun's env; initialises the process group when world > 1."""
    rank, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def gather_counters(counters: Dict[str, float], device: torch.device) -> Dict[str, float]:
    """Sum a small dict of counters over all ranks (frames processed, touched cells, index checksums, eval
    TP/FP...).  Bytes << 1 KB; latency bound.  Identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(counters)
    keys = sorted(counters)
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(value: float, device: torch.device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
