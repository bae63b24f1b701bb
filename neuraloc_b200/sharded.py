"""Multi-GPU rollout: samples are independent (src/OCflow.py:45-55 has no cross-sample op before the means at
:80-86), so the batch is sharded by rows over the ranks of one box, one process per GPU, with NO data-path
collective; the only exchange is one all-reduce (sum) of the 8-double vector
[sum L, G, HJt, HJfin, HJgrad, Q, W, count], after which every rank forms the same means and Jc.

`noMean` / `intermediates` outputs need no collective: each rank keeps its own rows.
"""
import torch

from .ocflow import costs_from_sums, ocflow_grad_sums, ocflow_sums, split_param_grads, _phi_tensors


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


def ocflow_grad_sharded(x_local, Phi, prob, tspan, nt, alph=(1.0,) * 6, group=None, local_eval=None, assign=True):
    """Data-parallel training evaluation (trainOC.py:172-173 over the union of every rank's rows): each rank runs the fused
    rollout + adjoint kernel on its own rows, and ONE all-reduce (sum) of the float64 vector [8 cost sums | P parameter-gradient
    sums] follows; every rank then holds the same Jc, cs and mean gradient.  With `assign` the gradients are written to the
    parameters' `.grad` (replacing them), so that `optim.step()` can follow as in the reference's loop.

    `local_eval(x, Phi, prob, tspan, nt, alph) -> (sums [8], grad [P])` lets the host logic run without a GPU (tests inject
    autograd through the CPU oracle); the default is noc_ocflow_grad.  Returns (Jc, cs, grads) with grads ordered like
    [A, c.weight, c.bias, w.weight, N.layers.0.weight, N.layers.1.weight, N.layers.0.bias, N.layers.1.bias]."""
    import torch.distributed as dist
    alph = [float(a) for a in alph]
    params = _phi_tensors(Phi)
    P = sum(p.numel() for p in params)
    dev = x_local.device
    if x_local.shape[0] > 0:
        if local_eval is not None:
            sums, grad = local_eval(x_local, Phi, prob, tspan, nt, alph)
        else:
            sums, grad, _ = ocflow_grad_sums(x_local, Phi, prob, tspan, nt, alph)
        dev = sums.device
        buf = torch.cat((sums.to(torch.float64), grad.to(torch.float64)))
    else:
        if local_eval is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        buf = torch.zeros(8 + P, dtype=torch.float64, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    Jc, cs = costs_from_sums(buf[:8], alph, x_local.dtype)
    grads = split_param_grads(Phi, (buf[8:] / buf[7]).to(x_local.dtype))
    if assign:
        for p, g in zip(params, grads):
            p.grad = g.to(device=p.device, dtype=p.dtype).clone()
    return Jc, cs, grads
