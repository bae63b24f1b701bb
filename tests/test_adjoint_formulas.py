"""CPU: the hand-derived discrete adjoint (tests/adjoint_ref.py — the formulas `rollout_grad_kernel` implements) against
reverse-mode autograd of the oracle rollout, i.e. against what trainOC.py:172-173 (`Jc.backward()`) computes."""
import dataclasses

import pytest
import torch

import adjoint_ref as ar
from helpers import PROBLEMS, oracle_setup
from oracle import ocflow_oracle as orc

NAMES = ["A", "c_w", "c_b", "w", "K0", "K1", "b0", "b1"]


def autograd_of_oracle(x, P, D, tspan, nt, alph):
    leaves = [t.clone().requires_grad_(True) for t in (P.A, P.c_w, P.c_b, P.w, P.K[0], P.K[1], P.b[0], P.b[1])]
    Pg = orc.PhiParams(leaves[0], leaves[1], leaves[2], leaves[3], [leaves[4], leaves[5]], [leaves[6], leaves[7]], P.h)
    xg = x.clone().requires_grad_(True)
    Jc, _ = orc.ocflow(xg, Pg, D, tspan, nt, "rk4", alph, noMean=True)
    Jc.sum().backward()
    return Jc.sum().detach(), {k: t.grad for k, t in zip(NAMES, leaves)}, xg.grad


def adversarial_batch(name, D, xinit, var0, n, seed=3):
    """Seeded batch with rows that switch on the pair interaction and the (train-mode) terrain terms."""
    g = torch.Generator().manual_seed(seed)
    d = xinit.shape[1]
    x = xinit.double() + var0 * torch.randn(n, d, generator=g, dtype=torch.float64)
    if name == "singlequad":
        x[:, 3:] = 0.1 * torch.randn(n, 9, generator=g, dtype=torch.float64)
        return x
    dim = D.agentDim
    x[0, dim:2 * dim] = x[0, 0:dim] + 0.3 * D.r
    if D.nAgents > 2:
        x[1, 2 * dim:3 * dim] = x[1, 0:dim] + 1.5 * D.r
    if name == "swap2":
        x[2, 0:2] = torch.tensor([0.5, 2.5]); x[3, 2:4] = torch.tensor([0.2, -2.0])
    if name == "swarm50":
        x[2, 0:3] = torch.tensor([0.0, 0.0, 3.0]); x[3, 3:6] = torch.tensor([3.0, 0.5, 2.0])
    return x


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("name", PROBLEMS)
def test_manual_adjoint_matches_autograd(name, training):
    P, D, xinit, meta = oracle_setup(name, torch.float64)
    D = dataclasses.replace(D, training=training)
    n, nt = (5, 3) if name == "swarm50" else (12, 5)
    x = adversarial_batch(name, D, xinit, meta["var0"], n)
    Ja, Ga, xa = autograd_of_oracle(x, P, D, [0.0, 1.0], nt, meta["alph"])
    Jm, Gm, xm = ar.manual_grad(x, P, D, [0.0, 1.0], nt, meta["alph"])
    assert abs(float(Ja - Jm)) <= 1e-12 * abs(float(Ja))
    assert float((xm - xa).abs().max()) <= 1e-11 * float(xa.abs().max())
    for k in NAMES:
        scale = float(Ga[k].abs().max())
        assert float((Gm[k] - Ga[k]).abs().max()) <= 1e-10 * max(scale, 1e-300), k
