"""GPU: the batched baseline objective (noc_baseline_loss; baseline2D.py:42-63, baselineQuad.py:40-72) and its gradient with
respect to the controls, through the C ABI, against the reference's own outputs (tests/golden/baseline.npz) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import ocflow_oracle as orc
from test_oracle_train_golden import BASE_CASES, baseline_case, baseline_fixture, ref_grad, sub_rows, train_fixture

pytestmark = pytest.mark.gpu


def product_problem(name, alph, dtype, training):
    import neuraloc_b200 as nb
    prob, _, _, _ = nb.initProb(name, 2, 2, var0=1.0, alph=alph, cvt=lambda v: v.to(dtype).cuda())
    (prob.train if training else prob.eval)()
    return prob


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name,mode", BASE_CASES)
def test_baseline_loss_and_gradient_match_reference(name, mode, dtype):
    import neuraloc_b200 as nb
    z = baseline_fixture()
    key, D, alph, U, z0 = baseline_case(z, name, mode)
    prob = product_problem(name, alph, dtype, mode == "train")
    loss, gU = nb.baseline_loss(U.to(dtype).cuda(), z0.to(dtype).cuda(), prob, alph[0], want_grad=True)
    ref, gref = z[key + "/loss"], z[key + "/gradU"]
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    assert np.abs(loss.double().cpu().numpy() - ref).max() <= tol * np.abs(ref).max()
    assert np.abs(gU.double().cpu().numpy() - gref).max() <= tol * np.abs(gref).max()


def test_baseline_reference_signatures_and_backward():
    """loss_fun(U, Z_0, prob, nt, alphG) on ONE state as baseline2D.py:97-102 uses it (err.backward(), Adam step), and a large batch
    against the oracle (many warps per CTA, grid-stride)."""
    import neuraloc_b200 as nb
    z = baseline_fixture()
    key, D, alph, U, z0 = baseline_case(z, "swap12", "train")
    prob = product_problem("swap12", alph, torch.float64, True)
    u = torch.nn.Parameter(U[0].clone().cuda())
    optim = torch.optim.Adam([{"params": u}], lr=0.1)
    optim.zero_grad()
    err = nb.loss_fun(u, z0[0].cuda(), prob, U.shape[1], alph[0])
    assert err.shape == (1, 1) and abs(err.item() - z[key + "/loss"][0]) <= 1e-11 * abs(z[key + "/loss"][0])
    err.backward()
    assert np.abs(u.grad.cpu().numpy() - z[key + "/gradU"][0]).max() <= 1e-11 * np.abs(z[key + "/gradU"][0]).max()
    optim.step()
    # quadcopter, reference signature
    keyq, Dq, alq, Uq, zq = baseline_case(z, "singlequad", None)
    probq = product_problem("singlequad", alq, torch.float64, True)
    J = nb.compute_loss(Uq[1].cuda(), zq[1].cuda(), probq, alq[0])
    assert J.dim() == 0 and abs(J.item() - z[keyq + "/loss"][1]) <= 1e-11 * abs(z[keyq + "/loss"][1])
    # 5000 states of the 50-agent swarm against the oracle
    g = torch.Generator().manual_seed(3)
    Ds, xi = orc.make_problem("swarm50", [1800.0, 1e7, 25000.0, 0, 0, 0], torch.float64)
    import dataclasses
    Ds = dataclasses.replace(Ds, training=True)
    n, nt = 5000, 6
    z0b = xi + 0.3 * torch.randn(n, 150, generator=g, dtype=torch.float64)
    Ub = (Ds.xtarget - z0b).unsqueeze(1) * torch.ones(n, nt, 150, dtype=torch.float64) + 0.5 * torch.randn(n, nt, 150, generator=g, dtype=torch.float64)
    Ub.requires_grad_(True)
    lo = orc.baseline_loss(Ds, Ub, z0b, 1800.0)
    lo.sum().backward()
    probs = product_problem("swarm50", [1800.0, 1e7, 25000.0, 0.0, 0.0, 0.0], torch.float64, True)
    lg, gg = nb.baseline_loss(Ub.detach().cuda(), z0b.cuda(), probs, 1800.0, want_grad=True)
    assert float((lg.cpu() - lo.detach()).abs().max()) <= 1e-11 * float(lo.abs().max())
    assert float((gg.cpu() - Ub.grad).abs().max()) <= 1e-10 * float(Ub.grad.abs().max())


@pytest.mark.parametrize("name", ["softcorridor", "swap2", "swap12", "singlequad", "swarm50"])
def test_training_gradient_matches_reference_backward_fixture(name):
    """noc_ocflow_grad against the gradients the UNMODIFIED reference's Jc.backward() produced (tests/golden/train_grads.npz)."""
    import neuraloc_b200 as nb
    from helpers import product_setup
    z = train_fixture()
    net, prob, _, meta = product_setup(name, torch.float64)
    prob.train()
    x, nt = torch.from_numpy(z[name + "/x"]).cuda(), int(z[name + "/nt"])
    sums, grad, _ = nb.ocflow_grad_sums(x, net, prob, [0.0, 1.0], nt, meta["alph"])
    n = x.shape[0]
    got = dict(zip(["A", "c_w", "c_b", "w", "K0", "K1", "b0", "b1"], nb.split_param_grads(net, grad)))
    costs = (sums[:7] / sums[7]).cpu().numpy()
    assert np.abs(costs[:5] - z[name + "/costs"][1:6]).max() <= 1e-10 * np.abs(z[name + "/costs"][1:6]).max()
    for k, gk in got.items():
        ref = ref_grad(z, name, k)
        mine = sub_rows(name, k, gk.cpu() / n).reshape(ref.shape)
        assert float((mine - ref).abs().max()) <= 1e-8 * max(float(ref.abs().max()), 1e-300), k
