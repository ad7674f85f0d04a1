"""Image sharding across ranks: one process per GPU, parameters replicated, each rank owns a contiguous
slice of the global minibatch; the only exchange of the forward ELBO is one scalar sum (and, with the backward pass, one
all-reduce of the flat gradient).  torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def shard_range(n_global, rank, world_size):
    """Images [lo, hi) of the global minibatch owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(int(n_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t):
    """In-place sum over ranks; identity for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def elbo_from_partials(sum_varexp, S, num_data, n_global, kls):
    """DS/dgp.py:92-98 with the data term summed over ranks: `sum_varexp` is this rank's sum of the S*n_local expected
    log-likelihoods (tensor, reduced in place), `kls` the per-layer KLs (replicated, NOT reduced)."""
    allreduce_sum_(sum_varexp)
    return sum_varexp / S * (float(num_data) / float(n_global)) - torch.as_tensor(kls, dtype=sum_varexp.dtype).sum()
