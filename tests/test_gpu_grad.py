"""GPU: the fused rollout + discrete-adjoint kernel (noc_ocflow_grad, trainOC.py:172-173) through the C ABI against reverse-mode
autograd of the fp64 oracle rollout — what `Jc.backward()` computes in the reference."""
import dataclasses

import numpy as np
import pytest
import torch

from helpers import PROBLEMS, oracle_setup, product_setup
from test_adjoint_formulas import NAMES, adversarial_batch, autograd_of_oracle

pytestmark = pytest.mark.gpu

ORDER = ["A", "c_w", "c_b", "w", "K0", "K1", "b0", "b1"]          # split_param_grads order


def _setup(name, dtype, training):
    import neuraloc_b200 as nb  # noqa: F401
    net, prob, xinit, meta = product_setup(name, dtype)
    (prob.train if training else prob.eval)()
    P, D, xi64, _ = oracle_setup(name, torch.float64)
    D = dataclasses.replace(D, training=training)
    return net, prob, P, D, xi64, meta


def _compare(name, dtype, training, n, nt, monkeypatch=None, ts=None):
    import neuraloc_b200 as nb
    if ts is not None:
        monkeypatch.setenv("NOC_GRAD_TS", str(ts))
    net, prob, P, D, xi64, meta = _setup(name, dtype, training)
    x64 = adversarial_batch(name, D, xi64, meta["var0"], max(n, 4))[:n].contiguous()      # the builder edits rows 0..3
    Ja, Ga, xa = autograd_of_oracle(x64, P, D, [0.0, 1.0], nt, meta["alph"])
    sums, grad, gx = nb.ocflow_grad_sums(x64.to(dtype).cuda(), net, prob, [0.0, 1.0], nt, meta["alph"], want_xgrad=True)
    with torch.no_grad():
        Jn, csn = nb.OCflow(x64.to(dtype).cuda(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
    assert float(sums[7]) == n
    al = meta["alph"]
    Jsum = float(sums[0] + al[0] * sums[1] + al[3] * sums[2] + al[4] * sums[3] + al[5] * sums[4])
    # fp32: the gradient is compared with the fp64 gradient, so the tolerance is fp32's own conditioning on these batches —
    # torch's fp32 autograd is 1e-5 ... 7e-4 away from the fp64 gradient on them (scripts/grad_accuracy.py prints both distances);
    # the kernel measures 2e-6 ... 9e-4, except on swarm50's SHORT rollouts (nt = 3 with a net trained for nt = 26: fp32 loses 1e-5
    # of the state, kernel and torch alike, and the gradient amplifies it to 3e-3 ... 5e-3) — gated at the training
    # length in test_grad_fp32_swarm50_at_training_nt
    tolJ, tolg = (1e-10, 1e-8) if dtype == torch.float64 else (2e-5, 1.5e-2 if name == "swarm50" else 5e-3)
    assert abs(Jsum - float(Ja)) <= tolJ * abs(float(Ja)), (Jsum, float(Ja))
    assert abs(Jsum - float(Jn.double().sum())) <= tolJ * abs(float(Ja))            # same objective as the forward-only kernels
    got = dict(zip(ORDER, nb.split_param_grads(net, grad)))
    worst = {}
    for k in NAMES:
        ref = Ga[k]
        scale = float(ref.abs().max())
        err = float((got[k].double().cpu().reshape(ref.shape) - ref).abs().max())
        worst[k] = err / max(scale, 1e-300)
        assert err <= tolg * max(scale, 1e-30), "%s %s: d/d%s off by %.3e of its largest entry (all: %s)" % (name, dtype, k, worst[k], worst)
    xerr = float((gx.double().cpu() - xa).abs().max()) / float(xa.abs().max())
    assert xerr <= tolg, xerr
    return worst


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("name", PROBLEMS)
def test_grad_fp64_matches_autograd(name, training):
    n, nt = (6, 3) if name == "swarm50" else (13, 5)
    _compare(name, torch.float64, training, n, nt)


@pytest.mark.parametrize("name", ["softcorridor", "singlequad", "swarm50"])
def test_grad_single_sample(name):
    """n = 1 (one valid row in its tile, the rest padding), as a one-state fine-tuning step would call it"""
    _compare(name, torch.float64, True, 1, 3)


@pytest.mark.parametrize("name", PROBLEMS)
def test_grad_fp32_matches_autograd(name):
    n, nt = (6, 3) if name == "swarm50" else (13, 5)
    _compare(name, torch.float32, True, n, nt)


def test_grad_fp32_swarm50_at_training_nt():
    """swarm50 at the reference's training length (README.md:118: nt = 26), a plain batch from rho_0: the fp32 gradient within 2e-4
    of the fp64 autograd gradient (measured 8e-6 ... 3.5e-5; torch's own fp32 autograd: 6e-6 ... 2.6e-5)."""
    import neuraloc_b200 as nb
    net, prob, P, D, xi64, meta = _setup("swarm50", torch.float32, True)
    g = torch.Generator().manual_seed(5)
    n, nt = 12, 26
    x64 = xi64 + meta["var0"] * torch.randn(n, 150, generator=g, dtype=torch.float64)
    Ja, Ga, xa = autograd_of_oracle(x64, P, D, [0.0, 1.0], nt, meta["alph"])
    sums, grad, gx = nb.ocflow_grad_sums(x64.float().cuda(), net, prob, [0.0, 1.0], nt, meta["alph"], want_xgrad=True)
    got = dict(zip(ORDER, nb.split_param_grads(net, grad)))
    for k in NAMES:
        ref = Ga[k]
        err = float((got[k].double().cpu().reshape(ref.shape) - ref).abs().max())
        assert err <= 2e-4 * max(float(ref.abs().max()), 1e-30), (k, err / float(ref.abs().max()))
    assert float((gx.double().cpu() - xa).abs().max()) <= 2e-4 * float(xa.abs().max())


@pytest.mark.parametrize("name,ts", [("swap12", 2), ("swap12", 4), ("swap12", 8), ("softcorridor", 2), ("swap2", 2), ("singlequad", 4),
                                     ("singlequad", 8), ("swarm50", 4), ("swarm50", 8)])
def test_grad_tile_widths(name, ts, monkeypatch):
    """every tile width (2: staged-weight nets only), more tiles than one CTA wave for the narrow nets, a ragged last tile"""
    n, nt = (21, 2) if name == "swarm50" else (301, 4)
    fits64 = not (name == "swarm50" and ts == 8)          # 8-sample fp64 panels of the m = 512 net exceed shared memory
    _compare(name, torch.float64 if fits64 else torch.float32, True, n, nt, monkeypatch, ts)


def test_backward_through_ocflow_like_trainOC():
    """trainOC.py:169-174: zero_grad, Jc, cs = OCflow(...), Jc.backward(), optimizer step — on the reference-layout module."""
    import neuraloc_b200 as nb
    name, nt, n = "softcorridor", 8, 64
    net, prob, P, D, xi64, meta = _setup(name, torch.float64, True)
    net.train()
    x64 = adversarial_batch(name, D, xi64, meta["var0"], n)
    Ja, Ga, _ = autograd_of_oracle(x64, P, D, [0.0, 1.0], nt, meta["alph"])
    optim = torch.optim.Adam(net.parameters(), lr=0.01)
    optim.zero_grad()
    Jc, cs = nb.OCflow(x64.cuda(), net, prob, tspan=[0.0, 1.0], nt=nt, stepper="rk4", alph=meta["alph"])
    assert Jc.requires_grad and not cs[0].requires_grad
    Jc.backward()
    assert abs(float(Jc) - float(Ja) / n) <= 1e-10 * abs(float(Ja) / n)
    named = {"A": net.A, "c_w": net.c.weight, "c_b": net.c.bias, "w": net.w.weight, "K0": net.N.layers[0].weight,
             "K1": net.N.layers[1].weight, "b0": net.N.layers[0].bias, "b1": net.N.layers[1].bias}
    for k, p in named.items():
        ref = Ga[k] / n
        assert p.grad is not None and p.grad.shape == p.shape
        assert float((p.grad.cpu().reshape(ref.shape) - ref).abs().max()) <= 1e-8 * max(float(ref.abs().max()), 1e-30), k
    before = net.N.layers[1].weight.detach().clone()
    optim.step()
    assert not torch.equal(before, net.N.layers[1].weight.detach())
    # the next evaluation sees the updated weights (the packed-weight cache is keyed on the parameter versions)
    with torch.no_grad():
        J2, _ = nb.OCflow(x64.cuda(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
    assert float(J2) != float(Jc)


def test_grad_modes_that_are_not_differentiable_raise():
    import neuraloc_b200 as nb
    net, prob, P, D, xi64, meta = _setup("softcorridor", torch.float32, True)
    x = xi64.float().cuda().repeat(4, 1)
    with pytest.raises(RuntimeError):
        nb.OCflow(x, net, prob, [0.0, 1.0], 4, "rk4", meta["alph"], noMean=True)
    with pytest.raises(RuntimeError):
        nb.OCflow(x, net, prob, [0.0, 1.0], 4, "rk1", meta["alph"])


def test_short_training_run_matches_cpu_autograd():
    """trainOC.py:169-174 for 25 iterations from the same random initialisation, fp64: Adam on the GPU with the fused
    forward + adjoint kernel, and Adam on the CPU with autograd through the oracle rollout — the two loss histories and the final
    weights must coincide (the reference's training loop, transplanted)."""
    import copy
    import neuraloc_b200 as nb
    from oracle import ocflow_oracle as orc
    _, meta0 = __import__("helpers").load_ckpt("softcorridor")
    alph = meta0["alph"]
    torch.manual_seed(3)
    net_gpu = nb.Phi(nTh=2, m=32, d=4, alph=alph).double()
    net_cpu = copy.deepcopy(net_gpu)
    net_gpu = net_gpu.cuda()
    prob, _, _, xinit = nb.initProb("softcorridor", 2, 2, var0=1.0, alph=alph, cvt=lambda v: v.double().cuda())
    prob.train()
    D, xi = orc.make_problem("softcorridor", alph, torch.float64)
    D = dataclasses.replace(D, training=True)
    g = torch.Generator().manual_seed(9)
    x = xi + torch.randn(256, 4, generator=g, dtype=torch.float64)
    nt, iters = 8, 25
    hist = {}
    for tag, net, run in (("gpu", net_gpu, None), ("cpu", net_cpu, None)):
        optim = torch.optim.Adam(net.parameters(), lr=0.01)
        losses = []
        for _ in range(iters):
            optim.zero_grad()
            if tag == "gpu":
                Jc, cs = nb.OCflow(x.cuda(), net, prob, tspan=[0.0, 1.0], nt=nt, stepper="rk4", alph=alph)
            else:
                Pg = orc.PhiParams(net.A, net.c.weight, net.c.bias, net.w.weight, [l.weight for l in net.N.layers],
                                   [l.bias for l in net.N.layers], net.N.h)
                Jc, cs = orc.ocflow(x, Pg, D, [0.0, 1.0], nt, "rk4", alph)
            Jc.backward()
            optim.step()
            losses.append(float(Jc.detach()))
        hist[tag] = losses
    a, b = np.array(hist["gpu"]), np.array(hist["cpu"])
    assert b[-1] < b[0]                                                   # it trains
    assert np.abs(a - b).max() <= 1e-7 * np.abs(b).max(), (a, b)
    for pg, pc in zip(net_gpu.parameters(), net_cpu.parameters()):
        assert float((pg.detach().cpu() - pc.detach()).abs().max()) <= 1e-7 * max(float(pc.detach().abs().max()), 1e-30)
