"""N > 1 host logic on CPU: world_size-2 gloo, row shards (uneven / empty), one all-reduce of the 8-double vector.
The per-shard evaluator is injected (the CPU oracle in noMean mode) because there is no GPU here; on the box the
default evaluator is the CUDA rollout and bench.py --gpus N exercises the same function over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_ckpt


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_sums(x, Phi, prob, tspan, nt, stepper, alph):
    from oracle import ocflow_oracle as orc
    P, D = Phi, prob            # the test passes oracle descriptors straight through
    with torch.no_grad():
        _, cs = orc.ocflow(x, P, D, tspan, nt, stepper, alph, noMean=True)
    return torch.cat([c.double().sum().view(1) for c in cs] + [torch.tensor([float(x.shape[0])], dtype=torch.float64)])


def _worker(rank, world, port, n, out_q):
    import neuraloc_b200 as nb
    from oracle import ocflow_oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sd, meta = load_ckpt("softcorridor")
    P = orc.params_from_state_dict(sd, torch.float64)
    D, xinit = orc.make_problem("softcorridor", meta["alph"], torch.float64)
    g = torch.Generator().manual_seed(11)
    x = xinit + torch.randn(n, 4, generator=g, dtype=torch.float64)
    lo, hi = nb.shard_rows(n, world, rank)
    Jc, cs = nb.OCflow_sharded(x[lo:hi], P, D, [0.0, 1.0], 6, "rk4", meta["alph"], local_sums=_oracle_sums)
    out_q.put((rank, float(Jc), [float(c) for c in cs], hi - lo))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 5, 8])
def test_two_rank_shards_reproduce_single_process_means(n):
    from oracle import ocflow_oracle as orc
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    sd, meta = load_ckpt("softcorridor")
    P = orc.params_from_state_dict(sd, torch.float64)
    D, xinit = orc.make_problem("softcorridor", meta["alph"], torch.float64)
    g = torch.Generator().manual_seed(11)
    x = xinit + torch.randn(n, 4, generator=g, dtype=torch.float64)
    with torch.no_grad():
        Jr, cr = orc.ocflow(x, P, D, [0.0, 1.0], 6, "rk4", meta["alph"])
    assert sum(r[3] for r in res) == n
    for _, Jc, cs, _ in res:                      # identical on both ranks and equal to the unsharded means
        assert abs(Jc - float(Jr)) <= 1e-10 * abs(float(Jr))
        assert np.allclose(cs, [float(c) for c in cr], rtol=1e-10, atol=1e-12)
    assert res[0][1:3] == res[1][1:3]


# ---- data-parallel training evaluation: one all-reduce of [8 cost sums | P gradient sums] ------------------------------------
def _oracle_grad_eval(x, Phi, prob, tspan, nt, alph):
    """(sums [8], flat gradient sums in state_dict order) by autograd through the CPU oracle, from the live module's weights."""
    from oracle import ocflow_oracle as orc
    sd = {k: v.detach().clone() for k, v in Phi.state_dict().items()}
    P = orc.params_from_state_dict(sd, x.dtype)
    leaves = [t.clone().requires_grad_(True) for t in (P.A, P.c_w, P.c_b, P.w, P.K[0], P.b[0], P.K[1], P.b[1])]
    Pg = orc.PhiParams(leaves[0], leaves[1], leaves[2], leaves[3], [leaves[4], leaves[6]], [leaves[5], leaves[7]], P.h)
    Jn, cs = orc.ocflow(x, Pg, prob, tspan, nt, "rk4", alph, noMean=True)
    Jn.sum().backward()
    sums = torch.cat([c.detach().double().sum().view(1) for c in cs] + [torch.tensor([float(x.shape[0])], dtype=torch.float64)])
    return sums, torch.cat([t.grad.reshape(-1) for t in leaves])


def _grad_worker(rank, world, port, n, out_q):
    import dataclasses
    import neuraloc_b200 as nb
    from oracle import ocflow_oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sd, meta = load_ckpt("softcorridor")
    net = nb.Phi(nTh=meta["nTh"], m=meta["m"], d=4, alph=meta["alph"]).double()
    net.load_state_dict(sd)
    D, xinit = orc.make_problem("softcorridor", meta["alph"], torch.float64)
    D = dataclasses.replace(D, training=True)
    g = torch.Generator().manual_seed(13)
    x = xinit + torch.randn(n, 4, generator=g, dtype=torch.float64)
    lo, hi = nb.shard_rows(n, world, rank)
    Jc, cs, grads = nb.ocflow_grad_sharded(x[lo:hi], net, D, [0.0, 1.0], 4, meta["alph"], local_eval=_oracle_grad_eval)
    out_q.put((rank, float(Jc), [g_.numpy() for g_ in grads], [p.grad.numpy() for p in (net.A, net.N.layers[1].weight)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 7])
def test_two_rank_training_gradient_equals_single_process(n):
    import dataclasses
    import neuraloc_b200 as nb
    from oracle import ocflow_oracle as orc
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, n, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    sd, meta = load_ckpt("softcorridor")
    net = nb.Phi(nTh=meta["nTh"], m=meta["m"], d=4, alph=meta["alph"]).double()
    net.load_state_dict(sd)
    D, xinit = orc.make_problem("softcorridor", meta["alph"], torch.float64)
    D = dataclasses.replace(D, training=True)
    g = torch.Generator().manual_seed(13)
    x = xinit + torch.randn(n, 4, generator=g, dtype=torch.float64)
    sums, flat = _oracle_grad_eval(x, net, D, [0.0, 1.0], 4, meta["alph"])
    want = [t.numpy() for t in nb.split_param_grads(net, flat / n)]
    for _, Jc, grads, assigned in res:
        for a, b in zip(grads, want):
            assert np.allclose(a, b, rtol=1e-10, atol=1e-12 * np.abs(b).max())
        assert np.allclose(assigned[0], want[0], rtol=1e-10) and np.allclose(assigned[1], want[5], rtol=1e-10)
    assert res[0][1] == res[1][1]
