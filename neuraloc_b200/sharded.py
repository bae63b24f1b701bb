"""Multi-GPU rollout: samples are independent (src/OCflow.py:45-55 has no cross-sample op before the means at
:80-86), so the batch is sharded by rows over the ranks of one box, one process per GPU, with NO data-path
collective; the only exchange is one all-reduce (sum) of the 8-double vector
[sum L, G, HJt, HJfin, HJgrad, Q, W, count], after which every rank forms the same means and Jc.

`noMean` / `intermediates` outputs need no collective: each rank keeps its own rows.
"""
import torch

from .ocflow import costs_from_sums, ocflow_sums


def shard_rows(n, world_size, rank):
    """Contiguous row block of rank `rank`: rows [lo, hi) with blocks of ceil(n / world_size)."""
    per = -(-n // world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def OCflow_sharded(x_local, Phi, prob, tspan, nt, stepper="rk4", alph=(1.0,) * 6, group=None, local_sums=None):
    """Mean-mode OCflow over the union of every rank's `x_local` rows.

    Each rank runs the fused rollout on its own rows (which may be empty on some ranks when n < world_size)
    and contributes cost sums + its row count; one torch.distributed all-reduce (NCCL for CUDA tensors, gloo for
    CPU tensors) combines them.  `local_sums` lets the host-side logic be exercised without a GPU (tests inject
    a CPU evaluator); the default is the CUDA rollout.  Returns (Jc, cs) like OCflow, identical on all ranks.
    """
    import torch.distributed as dist
    alph = [float(a) for a in alph]
    fn = local_sums if local_sums is not None else ocflow_sums
    if x_local.shape[0] > 0:
        sums = fn(x_local, Phi, prob, tspan, nt, stepper, alph)
    else:
        sums = torch.zeros(8, dtype=torch.float64, device=x_local.device)
    sums = sums.to(torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return costs_from_sums(sums, alph, x_local.dtype)
