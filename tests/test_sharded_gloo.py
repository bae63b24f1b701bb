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
