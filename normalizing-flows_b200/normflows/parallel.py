"""Batch-sharded data parallelism for the density pass (the reference has none).

Every op on the path is per-sample, parameters are replicated, so the batch shards across ranks with
no data-path collective; `forward_kld` needs exactly one all-reduce of (sum log_q, count)."""
import torch
import torch.distributed as dist


def shard_rows(n_rows, rank, world_size):
    """Contiguous row range [lo, hi) of rank `rank`; remainders go to the first ranks."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_kld(local_sum, local_count, group=None):
    """-(sum_r local_sum_r) / (sum_r local_count_r) with ONE collective on a 2-element fp64 tensor.
    local_sum: 0-dim tensor (sum of log_q over the local shard)."""
    buf = torch.stack([local_sum.to(torch.float64).reshape(()),
                       torch.full((), float(local_count), dtype=torch.float64, device=local_sum.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return -(buf[0] / buf[1])


def forward_kld_dp(model, x_local, group=None):
    """Global forward KL over all ranks' shards: each rank runs the fused stack on its rows, then one
    NCCL all-reduce of the partial sums."""
    h = model._stack()
    if h is None or h.base is None:
        raise NotImplementedError("forward_kld_dp needs an all-native stack with a DiagGaussian base")
    _, s = h.forward_kld(x_local, want_sum=True)
    return all_reduce_kld(s, x_local.shape[0], group).to(torch.float32)
