"""Multi-GPU sharding of a batch of env instances.

Envs are independent, so the batch shards trivially: rank r owns the contiguous global env ids
``[offset_r, offset_r + count_r)``; the static network and the profile store are replicated per
GPU. There is no per-step communication. The only collective is one ``all_gather`` of per-env
episode returns (and optionally an ``all_reduce`` of the info sums) at episode end - KB-sized and
latency-bound. The device RNG is keyed by *global* env id, so results do not depend on how the
batch is sharded.

One process per GPU (``torch.distributed``, backend ``nccl``; ``gloo`` in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "all_gather_varlen", "ShardedVoltageControl", "bind_to_gpu_numa_node"]


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Pin the calling process to the host cores next to its GPU (8-GPU boxes: GPU0-3 <-> NUMA0, GPU4-7 <-> NUMA1).
    Call it BEFORE the first ``step_host`` (the pinned host buffers are allocated lazily, first touch decides their NUMA
    node): with eight ranks writing ~10 MB of observations per step into host memory, remote-node buffers were what
    held the round-1 end-to-end scaling at 0.64. Returns a short description; never raises."""
    import os
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} cores of the GPU's NUMA node"
        return "not bound (no usable cores listed)"
    except Exception as ex:
        return f"not bound ({type(ex).__name__})"


def shard_range(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(offset, count) of rank's contiguous shard; the first ``global_batch % world_size`` ranks
    get one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(global_batch), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def all_gather_varlen(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """Gather per-env values (dim 0 = local envs) from all ranks into global env order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_range(global_batch, r, world)[1] for r in range(world)]
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank}: expected {counts[rank]} local envs, got {local.shape[0]}")
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


class ShardedVoltageControl:
    """This rank's shard of a global batch of envs + the episode-end collectives."""

    def __init__(self, net, profiles, env_args: Optional[dict], global_batch: int, device: Optional[int] = None,
                 group=None, **kw):
        from .env import BatchedVoltageControl
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.global_batch = int(global_batch)
        self.offset, self.count = shard_range(global_batch, self.rank, self.world)
        self.env = BatchedVoltageControl(net, profiles, env_args, batch=self.count, device=device,
                                         env_id_offset=self.offset, **kw)

    def __getattr__(self, name):
        return getattr(self.env, name)

    def gather_episode_returns(self) -> torch.Tensor:
        """``sum_rewards`` of every env of the global batch, in global env order (one all_gather)."""
        local = self.env.get_field("sum_rewards")[:, 0].contiguous()
        if self.world == 1:
            return local
        return all_gather_varlen(local, self.global_batch, self.group)

    def reduce_info_mean(self, info_sum: torch.Tensor) -> torch.Tensor:
        """Mean over the global batch of per-env accumulated info ``[count, 11]`` (one all_reduce)."""
        tot = info_sum.sum(dim=0)
        if self.world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
        return tot / self.global_batch
