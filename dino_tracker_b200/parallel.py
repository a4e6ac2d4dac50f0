"""Single-node multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed``
(NCCL on GPUs; the host logic below is backend-agnostic and is tested with ``gloo`` on CPU).

* video-parallel (configs 2/3): videos dealt to ranks by longest-processing-time-first on T * N_q; no data-path
  collective.
* frame-sharded long video (config 4): rank r owns a contiguous block of frames, runs ViT + delta-DINO for them
  writing straight into its slice of the full ``[T][P][C]`` buffer, then ONE in-place all-gather of the refined
  features (each (query, frame) correlation map needs only that frame + one descriptor, so any frame sharding is
  exact); query points are then sharded across ranks and the results gathered.
* best-buddies (config 5): features replicated, unordered frame pairs dealt round-robin
  (``best_buddies.best_buddies(rank=, world=)``), results gathered as objects.
"""
from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def frame_shard(T: int, world: int, rank: int):
    """Contiguous block of frames of ``rank``: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(T, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def query_shard(N: int, world: int, rank: int):
    return frame_shard(N, world, rank)


def lpt_assign(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of videos to ranks (cost = T * N_q)."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: loads[k])
        out[r].append(i)
        loads[r] += costs[i]
    return out


def allgather_frames(full: torch.Tensor, T: int, world: int, rank: int, group=None) -> torch.Tensor:
    """In-place all-gather of a frame-sharded ``[T][...]`` buffer: every rank has filled its own block
    ``frame_shard(T, world, rank)`` of ``full``; afterwards all blocks are valid everywhere.
    Equal blocks go through one ``all_gather_into_tensor`` directly on ``full`` (NCCL: NVLink/NVSwitch ring or
    NVLS); ragged blocks (T % world != 0) are padded to the largest block in a staging buffer."""
    if world == 1:
        return full
    base, extra = divmod(T, world)
    if extra == 0:
        s, e = frame_shard(T, world, rank)
        dist.all_gather_into_tensor(full, full[s:e].contiguous() if not full[s:e].is_contiguous() else full[s:e], group=group)
        return full
    blk = base + 1
    stage = torch.zeros((world * blk,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    s, e = frame_shard(T, world, rank)
    mine = torch.zeros((blk,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    mine[: e - s] = full[s:e]
    dist.all_gather_into_tensor(stage, mine, group=group)
    for r in range(world):
        rs, re = frame_shard(T, world, r)
        if r != rank:
            full[rs:re] = stage[r * blk: r * blk + (re - rs)]
    return full


def gather_rows(local: torch.Tensor, N: int, world: int, rank: int, group=None) -> torch.Tensor:
    """All ranks end up with the ``[N][...]`` concatenation of their row shards (``query_shard``)."""
    if world == 1:
        return local
    blk = -(-N // world)
    mine = torch.zeros((blk,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    mine[: local.shape[0]] = local
    stage = torch.zeros((world * blk,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(stage, mine, group=group)
    parts = []
    for r in range(world):
        rs, re = query_shard(N, world, r)
        parts.append(stage[r * blk: r * blk + (re - rs)])
    return torch.cat(parts, dim=0)


def sharded_long_video_infer(T: int, N: int, world: int, rank: int, refine_block: Callable, infer_rows: Callable,
                             full_features: torch.Tensor, group=None):
    """Config-4 driver, generic over the compute callables so that it runs under gloo/CPU in tests:
      refine_block(s, e)      -> fills full_features[s:e] (ViT + delta-DINO of this rank's frames)
      infer_rows(qs, qe)      -> (traj [qe-qs][T][2], occ [qe-qs][T]) for this rank's query rows, using full_features
    Returns (traj [N][T][2], occ [N][T]) on every rank."""
    s, e = frame_shard(T, world, rank)
    refine_block(s, e)
    allgather_frames(full_features, T, world, rank, group)
    qs, qe = query_shard(N, world, rank)
    traj, occ = infer_rows(qs, qe)
    traj = gather_rows(traj, N, world, rank, group)
    occ = gather_rows(occ.to(torch.uint8), N, world, rank, group).bool()
    return traj, occ
