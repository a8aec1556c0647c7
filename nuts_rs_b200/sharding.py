"""Chain sharding across GPUs: chains are independent, so a run of `num_chains` chains is split into contiguous
blocks, one per rank; the RNG stream of a chain is keyed by its GLOBAL id (reference `set_stream(chain_id + 1)`,
src/sampler.rs:1105-1106), which makes the draws invariant to the number of GPUs.  The only exchange of the path is the
gather of draws / statistics at the end (or per batch of draws) — no collective inside the sampling loop.

These helpers are backend-agnostic `torch.distributed` code: NCCL over NVLink on the GPU box, gloo in the CPU tests."""
from typing import Dict, List, Tuple

import numpy as np


def shard_range(num_chains: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [offset, offset + count) of global chain ids owned by `rank` (first ranks get the remainder)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, rem = divmod(num_chains, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def all_shard_ranges(num_chains: int, world_size: int) -> List[Tuple[int, int]]:
    return [shard_range(num_chains, world_size, r) for r in range(world_size)]


def gather_draws(local_draws, num_chains: int, group=None):
    """All-gather per-rank draws [n_draws, local_chains, dim] into [n_draws, num_chains, dim] (global chain order) on every rank.

    `local_draws` is a torch tensor on the device that matches the process group's backend (cuda for nccl, cpu for gloo).
    Shards may have unequal chain counts (num_chains not divisible by the world size): they are padded to the largest shard
    for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    ranges = all_shard_ranges(num_chains, world)
    max_count = max(c for _, c in ranges)
    n_draws, local_chains, dim = local_draws.shape
    rank = dist.get_rank(group)
    if local_chains != ranges[rank][1]:
        raise ValueError(f"rank {rank} holds {local_chains} chains, expected {ranges[rank][1]}")
    send = local_draws
    if local_chains != max_count:
        send = torch.zeros((n_draws, max_count, dim), dtype=local_draws.dtype, device=local_draws.device)
        send[:, :local_chains] = local_draws
    send = send.contiguous()
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return torch.cat([r[:, :c] for r, (_, c) in zip(recv, ranges)], dim=1)


def gather_stats(local_stats: Dict[str, np.ndarray], num_chains: int, device="cpu", group=None) -> Dict[str, np.ndarray]:
    """All-gather the per-draw statistics ([n_draws, local_chains] each) into global chain order."""
    import torch

    out = {}
    for name in sorted(local_stats):
        a = local_stats[name]
        t = torch.from_numpy(np.ascontiguousarray(a)[:, :, None].astype(np.float64)).to(device)
        out[name] = gather_draws(t, num_chains, group)[:, :, 0].cpu().numpy().astype(a.dtype)
    return out


def total_leapfrogs(local_count: int, device="cpu", group=None) -> int:
    """Sum of the per-rank leapfrog counters (the throughput numerator)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(local_count)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())
