"""Walker / cosmology sharding over the GPUs of one box.

The reference distributes one likelihood evaluation per MPI rank through ``emcee.utils.MPIPool``
(``pyRSD/rsdfit/util/mpi_manager.py``, ``rsdfit/engines/mcmc.py``): walkers are independent, so here every rank
(one process per GPU, ``torch.distributed``) owns a contiguous block of walkers, runs the same batched kernels on
its block, and the per-walker results are gathered once per ensemble step -- NCCL over NVLink on GPUs, gloo on
CPU tensors (tests).  There is no collective on the data path itself.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """contiguous block [lo, hi) of rank `rank`: the first n % world ranks hold one extra item"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank %d of %d" % (rank, world))
    base, extra = divmod(int(n), world)
    lo = rank*base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world):
    return [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]


def gather_rows(local, n_total, group=None):
    """every rank contributes its block of rows ``local`` [n_local, ...]; returns [n_total, ...] in walker order on
    every rank.  Blocks may be ragged (n_total not a multiple of the world size): they are padded to the largest
    block for the fixed-size all-gather and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_total:
            raise ValueError("single process holds %d of %d rows" % (local.shape[0], n_total))
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, its block has %d" % (rank, local.shape[0], sizes[rank]))
    nmax = max(sizes)
    tail = tuple(local.shape[1:])
    send = local
    if local.shape[0] != nmax:
        send = torch.zeros((nmax,) + tail, dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
    recv = torch.empty((world*nmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(s == nmax for s in sizes):
        return recv
    recv = recv.view((world, nmax) + tail)
    return torch.cat([recv[r, :sizes[r]] for r in range(world)], dim=0)


def max_over_ranks(value, device=None, group=None):
    """max of a python float over ranks (timings are reported as the max over ranks)"""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
