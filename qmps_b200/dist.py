"""Multi-GPU: shard the flat batch over ranks, one process per GPU, no data-path
collective; the only exchange is the final cost reduction / argmin:
an all-gather of one (cost, index) pair per rank -- 16 bytes -- over NCCL (gloo on CPU
for the host-logic tests).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "global_argmin", "global_sum", "nccl_comm_ptr", "argmin_allreduce"]


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


def nccl_comm_ptr(device=None, group=None):
    """The raw ``ncclComm_t`` of torch's NCCL process group for ``device`` (as an int), or 0 when there is no
    multi-rank NCCL group.  Passed to ``qmps_argmin_allreduce`` so that the C ABI issues the collective itself,
    on the caller's stream, with the communicator torch already built."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return 0
    pg = group if group is not None else dist.distributed_c10d._get_default_group()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    backend = pg._get_backend(device)
    if not hasattr(backend, "_comm_ptr"):
        raise RuntimeError("the process group's backend does not expose an NCCL communicator")
    ptr = int(backend._comm_ptr())
    if ptr == 0:                     # communicators are created lazily: one tiny collective builds it
        dist.all_reduce(torch.zeros(1, device=device), group=group)
        ptr = int(backend._comm_ptr())
    return ptr


def argmin_allreduce(cost, index_offset=0, comm=None):
    """(min cost, global argmin) over the shards of all ranks in ONE C-ABI call (``qmps_argmin_allreduce``):
    local reduction, a 16-byte-per-rank ncclAllGather and the final pass, stream-ordered on the current stream.
    ``cost``: float64 CUDA vector (this rank's shard); returns two 1-element CUDA tensors."""
    from . import _lib as L
    cost = cost.reshape(-1)
    if not (cost.is_cuda and cost.dtype == torch.float64 and cost.is_contiguous()):
        raise ValueError("cost must be a contiguous float64 CUDA tensor")
    bc = torch.empty((1,), dtype=torch.float64, device=cost.device)
    bi = torch.empty((1,), dtype=torch.int64, device=cost.device)
    if comm is None:
        comm = nccl_comm_ptr(cost.device)
    with torch.cuda.device(cost.device):
        L.check(L.load().qmps_argmin_allreduce(comm or None, cost.numel(), cost.data_ptr() if cost.numel() else None,
                                               int(index_offset), bc.data_ptr(), bi.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "argmin_allreduce")
    return bc, bi
