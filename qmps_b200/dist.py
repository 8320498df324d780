"""Multi-GPU: shard the flat batch over ranks, one process per GPU, no data-path
collective; the only exchange is the final cost reduction / argmin:
an all-gather of one (cost, index) pair per rank -- 16 bytes -- over NCCL (gloo on CPU
for the host-logic tests).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "global_argmin", "global_sum"]


def shard_range(n_total, rank=None, world=None):
    """Contiguous [lo, hi) slice of range(n_total) owned by `rank` (ceil split)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    per = -(-n_total // world)
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def global_argmin(local_cost, local_index):
    """(min cost, its global index) over all ranks.  local_cost: 1-element float64 tensor,
    local_index: 1-element int64 tensor (already offset to global numbering; -1 = empty
    shard).  Ties resolve to the smallest index, like numpy.argmin on the unsharded batch."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local_cost.reshape(()), local_index.reshape(())
    world = dist.get_world_size()
    # pack both 64-bit words into one int64 pair so a single all_gather moves 16 B per rank
    packed = torch.stack([local_cost.reshape(()).view(torch.int64), local_index.reshape(())])
    gathered = torch.empty((world, 2), dtype=torch.int64, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed[None].contiguous())
    costs = gathered[:, 0].contiguous().view(torch.float64)
    idx = gathered[:, 1]
    big = torch.iinfo(torch.int64).max
    costs = torch.where(idx < 0, torch.full_like(costs, float("inf")), costs)
    best = costs.min()
    cand = torch.where(costs == best, idx, torch.full_like(idx, big))
    return best, cand.min()


def global_sum(x):
    if dist.is_initialized() and dist.get_world_size() > 1:
        x = x.clone()
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
    return x
