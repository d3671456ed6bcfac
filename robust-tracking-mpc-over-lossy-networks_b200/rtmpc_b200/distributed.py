"""Multi-GPU plumbing: closed-loop instances are independent, so the batch is cut into contiguous
shards (one process per GPU) and nothing is exchanged on the hot path.  The device RNG is keyed by the
GLOBAL instance id (``id_offset`` of ``rtmpc_loop_step`` / ``rtmpc_loop_rollout``), so a run gives the
same trajectories for any number of ranks.  Only the per-instance statistics / trajectories are
all-gathered once at the end (NCCL over NVLink on the GPU box; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard(total, rank, world):
    """Contiguous shard of ``total`` instances for ``rank``: returns (id_offset, count).  The first
    ``total % world`` ranks take one instance more."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def all_gather_instances(local, total=None):
    """Concatenate the per-instance tensors [count, ...] of all ranks in rank order (ragged shards
    allowed).  Without an initialised process group the input is returned unchanged."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    counts[dist.get_rank()] = local.shape[0]
    dist.all_reduce(counts)
    cmax = int(counts.max().item())
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.cat([p[:int(c)] for p, c in zip(parts, counts.tolist())])
    if total is not None and out.shape[0] != total:
        raise RuntimeError(f"gathered {out.shape[0]} instances, expected {total}")
    return out


def all_reduce_sum(counts):
    """Element-wise sum of a small statistics tensor over all ranks (status counts, iteration totals)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        counts = counts.clone()
        dist.all_reduce(counts)
    return counts


def all_reduce_max(value):
    """Max over ranks of a scalar tensor (device time of the slowest rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        value = value.clone()
        dist.all_reduce(value, op=dist.ReduceOp.MAX)
    return value
