#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own modules (src.OCflow / src.Phi / src.initProb, read-only) and writes

  tests/golden/ckpt/<problem>.npz     the pretrained value-network tensors (state_dict layout of
                                      src/Phi.py:77-87) + the `args` the checkpoint carries
  tests/golden/cases_<problem>.npz    inputs and reference outputs of OCflow in its three return modes,
                                      in fp32 and fp64, at xInit and on a small seeded batch
  tests/golden/functors.npz           problem-functor known answers on adversarial inputs
  tests/golden/phi_random.npz         Phi.forward / Phi.getGrad on random-weight nets incl. nTh=3,4
  tests/golden/config5.npz            BASELINE.json config 5 (random-init swarm50-shape Phi, fp64)
  tests/golden/initprob.npz           xtarget / xInit tables of src/initProb.py for every problem name

Nothing here is product code; the fixtures pin the oracle (oracle/) and the CUDA path.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("NOC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from src.OCflow import OCflow  # noqa: E402
from src.Phi import Phi  # noqa: E402
from src.initProb import initProb  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PROBLEMS = {  # name -> nt used by README.md:66-95
    "softcorridor": 50, "swap2": 50, "swap12": 50, "singlequad": 50, "swarm50": 80,
}
ALL_NAMES = ["softcorridor", "swarm", "swarm50", "singlequad", "midcross2", "midcross4", "midcross20",
             "midcross30", "swap2", "swap12", "swap12_5pair", "swap12_4pair", "swap12_3pair",
             "swap12_2pair", "swap12_1pair"]


def np_(t):
    return t.detach().cpu().numpy()


def load_ckpt(name):
    return torch.load(os.path.join(REF, "experiments/oc/pretrained", name + "_nn_checkpt.pth"),
                      map_location="cpu", weights_only=False)


def build(name, prec):
    """Same construction order as evalOC.py:43-64."""
    torch.set_default_dtype(prec)
    ck = load_ckpt(name)
    a = ck["args"]
    cvt = lambda x: x.type(prec)
    prob, x0, _, xInit = initProb(a.data, 10, 11, var0=1.0, alph=a.alph, cvt=cvt)
    prob.eval()
    net = Phi(nTh=a.nTh, m=a.m, d=x0.size(1), alph=a.alph)
    net.load_state_dict(ck["state_dict"])
    net = net.to(prec)
    net.eval()
    return ck, prob, net, xInit


def run_modes(x, net, prob, tspan, nt, stepper="rk4"):
    with torch.no_grad():
        Jc, cs = OCflow(x, net, prob, tspan=tspan, nt=nt, stepper=stepper, alph=net.alph)
        Jn, cn = OCflow(x, net, prob, tspan=tspan, nt=nt, stepper=stepper, alph=net.alph, noMean=True)
        zF, cF = OCflow(x, net, prob, tspan=tspan, nt=nt, stepper=stepper, alph=net.alph, intermediates=True)
    mean = np.array([float(Jc)] + [float(c) for c in cs])
    nomean = np.concatenate([np_(Jn)] + [np_(c.to(x.dtype)) for c in cn], axis=1)
    return mean, nomean, np_(zF), np_(cF)


def batch_inputs(name, xInit, var0, n, seed):
    g = torch.Generator().manual_seed(seed)
    T = lambda v: torch.tensor(v, dtype=torch.float32)
    d = xInit.shape[1]
    if name == "singlequad":
        x = torch.zeros(n, d, dtype=torch.float32)
        x[:, :3] = -1.5 + var0 * torch.randn(n, 3, generator=g, dtype=torch.float32)
    else:
        x = xInit.float() + var0 * torch.randn(n, d, generator=g, dtype=torch.float32)
    # adversarial rows: force early interactions / obstacle hits so W and Q paths are exercised
    if name == "swap12":
        x[0, 0:2] = x[0, 2:4] + T([0.3, 0.1])      # agents 0,1 closer than 2r = 1
        x[1, 4:6] = x[1, 20:22] + T([0.0, 0.6])
    if name == "softcorridor":
        x[0] = T([-2.4, -0.3, 2.4, -0.3])             # starts next to the Gaussians
        x[1] = T([-0.2, -2.0, 0.2, -2.0])             # agents within 2r of each other
    if name == "swap2":
        x[0] = T([-0.5, 2.5, 0.3, 2.8])               # inside the upper disc, and colliding
        x[1] = T([-1.0, -2.0, 1.0, -2.5])             # inside the lower disc
    if name == "swarm50":
        x[0, 0:3] = T([0.0, 0.0, 3.0])                # inside block 1
        x[0, 3:6] = T([3.0, 0.5, 2.0])                # inside block 2
        x[1, 6:9] = x[1, 9:12] + T([0.05, 0.0, 0.1])  # two agents within 2r = 0.2
    return x


def gen_ckpts():
    for name in PROBLEMS:
        ck = load_ckpt(name)
        a = ck["args"]
        meta = dict(data=a.data, m=a.m, nTh=a.nTh, alph=list(map(float, a.alph)), var0=float(a.var0),
                    nt=int(a.nt), nt_val=int(getattr(a, "nt_val", a.nt)))
        arrs = {k: np_(v) for k, v in ck["state_dict"].items()}
        np.savez(os.path.join(HERE, "ckpt", name + ".npz"), meta_json=np.array(json.dumps(meta)), **arrs)
        print("ckpt", name, meta)


def gen_cases():
    for name, nt in PROBLEMS.items():
        out = {}
        nb, ntb = (6, 20) if name == "swarm50" else (16, nt)
        for tag, prec in (("f32", torch.float32), ("f64", torch.float64)):
            ck, prob, net, xInit = build(name, prec)
            var0 = float(ck["args"].var0)
            out["xinit"] = np_(xInit).astype(np.float64)
            mean, nomean, zF, cF = run_modes(xInit, net, prob, [0.0, 1.0], nt)
            out["xinit_mean_" + tag], out["xinit_z_" + tag], out["xinit_ctrl_" + tag] = mean, zF, cF
            xb = batch_inputs(name, xInit, var0, nb, seed=7).to(prec)
            out["xb"] = np_(xb).astype(np.float32)
            mean, nomean, zF, cF = run_modes(xb, net, prob, [0.0, 1.0], ntb)
            out["b_mean_" + tag], out["b_nomean_" + tag] = mean, nomean
            out["b_z_" + tag], out["b_ctrl_" + tag] = zF, cF
            # forward Euler (stepRK1, OCflow.py:143-155) and a tspan != [0,1] restart (plotter.py:817-823)
            mean, nomean, zF, cF = run_modes(xb[:4].clone(), net, prob, [0.0, 1.0], 8, stepper="rk1")
            out["rk1_mean_" + tag], out["rk1_z_" + tag], out["rk1_ctrl_" + tag] = mean, zF, cF
            if name == "softcorridor":
                nS = int(0.1 * nt)
                m1, _, z1, c1 = run_modes(xInit, net, prob, [0.0, 0.1], nS)
                shock = torch.tensor([[-0.2, -0.7, -0.0, -0.6]], dtype=prec)
                xs = torch.from_numpy(z1[:, :4, -1]).to(prec) + shock
                m2, _, z2, c2 = run_modes(xs, net, prob, [0.1, 1.0], 1 + nt - nS)
                out["shock1_mean_" + tag], out["shock1_z_" + tag], out["shock1_ctrl_" + tag] = m1, z1, c1
                out["shock2_x_" + tag] = np_(xs)
                out["shock2_mean_" + tag], out["shock2_z_" + tag], out["shock2_ctrl_" + tag] = m2, z2, c2
            # "unknown stepper integrates nothing" quirk (OCflow.py:46-49)
            mean, _, zF, cF = run_modes(xb[:2].clone(), net, prob, [0.0, 1.0], 3, stepper="none")
            out["nostep_mean_" + tag], out["nostep_z_" + tag], out["nostep_ctrl_" + tag] = mean, zF, cF
        out["nt"], out["nt_batch"] = np.array(nt), np.array(ntb)
        np.savez_compressed(os.path.join(HERE, "cases_%s.npz" % name), **out)
        print("cases", name, "Jc(xInit) f64 = %.10e  f32 = %.10e" % (out["xinit_mean_f64"][0], out["xinit_mean_f32"][0]))
    torch.set_default_dtype(torch.float32)


def gen_functors():
    """calcLHQW / calcGradpH / calcCtrls on adversarial inputs, train and eval mode, fp32 and fp64."""
    out = {}
    g = torch.Generator().manual_seed(11)
    for name in ("softcorridor", "swap2", "swap12", "singlequad", "swarm50", "midcross4", "swap12_1pair", "swarm"):
        for tag, prec in (("f32", torch.float32), ("f64", torch.float64)):
            torch.set_default_dtype(prec)
            alph = [300.0, 2.5, 7.0, 1.0, 1.0, 1.0] if name != "singlequad" else [5000.0, 0.0, 0.0, 0.1, 0.0, 0.0]
            prob, _, _, xInit = initProb(name, 4, 4, var0=1.0, alph=alph, cvt=lambda x: x.type(prec))
            d = xInit.shape[1]
            n = 24
            gg = torch.Generator().manual_seed(11)
            x = xInit.double() + 0.8 * torch.randn(n, d, generator=gg, dtype=torch.float64)
            p = 2.0 * torch.randn(n, d, generator=gg, dtype=torch.float64)
            A, dim = prob.nAgents, prob.agentDim
            r = prob.r
            if name != "singlequad":
                xa = x.view(n, A, dim)
                if A >= 2:
                    # pairs at controlled distances around the 2r / 2.2r / 3.2r thresholds; coincident agents
                    for row, fac in enumerate([0.0, 1e-5, 0.5, 1.0, 1.9, 1.999, 2.001, 2.1, 2.199, 2.201, 3.1, 3.199, 3.201]):
                        xa[row, 1] = xa[row, 0]
                        xa[row, 1, 0] += fac * r
                    if A > 2:
                        xa[13, 2] = xa[13, 0]; xa[13, 2, 1] += 0.7 * r
                        xa[13, 1] = xa[13, 0]; xa[13, 1, 0] -= 1.2 * r
                # obstacle hits
                if name == "softcorridor":
                    xa[14, 0] = torch.tensor([-2.5, 0.05]); xa[14, 1] = torch.tensor([1.4, -0.1])
                if name == "swap2":
                    xa[14, 0] = torch.tensor([0.3, 3.0]); xa[15, 1] = torch.tensor([-0.2, -2.2])
                    xa[16, 0] = torch.tensor([0.0, 4.0 + 2.0 + 0.5 * r]); xa[17, 0] = torch.tensor([1.99, 4.0])
                    xa[18, 1] = torch.tensor([0.0, -3.5 - 2.0 - 1.1 * r])
                if name in ("swarm50", "swarm"):
                    xa[14, 0] = torch.tensor([0.0, 0.0, 3.0]); xa[14, 3] = torch.tensor([3.0, 0.5, 2.0])
                    xa[15, 1] = torch.tensor([2.0 + 0.5 * r, 0.0, 1.0]); xa[16, 2] = torch.tensor([1.0, 0.5 + 0.5 * r, 6.9])
                    xa[17, 4] = torch.tensor([3.9, -0.9, 3.9]); xa[18, 5] = torch.tensor([0.0, 0.0, 7.5])
                    xa[19, 6] = torch.tensor([4.0 + 0.9 * r, 1.0 + 0.9 * r, 4.0 + 0.9 * r])
                x = xa.reshape(n, d)
            else:
                x[:, 3:6] = 1.5 * torch.randn(n, 3, generator=gg, dtype=torch.float64)
            x = x.to(prec); p = p.to(prec)
            key = "%s_%s" % (name, tag)
            out[key + "_x"], out[key + "_p"] = np_(x), np_(p)
            out[key + "_xtarget"] = np_(prob.xtarget)
            out[key + "_meta"] = np.array(json.dumps(dict(
                cls=type(prob).__name__, obstacle=prob.obstacle, alph_Q=float(prob.alph_Q), alph_W=float(prob.alph_W),
                r=float(prob.r), nAgents=int(prob.nAgents), agentDim=int(prob.agentDim),
                mass=float(getattr(prob, "mass", 1.0)), grav=float(getattr(prob, "grav", 9.81)))))
            for mode in ("eval", "train"):
                getattr(prob, mode)()
                L, H, Q, W = prob.calcLHQW(x, p)
                gp = prob.calcGradpH(x, p)
                cc = prob.calcCtrls(x, p)
                out["%s_%s_LHQW" % (key, mode)] = np.stack(
                    [np_(v.to(prec)).reshape(-1) if torch.is_tensor(v) else np.zeros(n) for v in (L, H, Q, W)], axis=1)
                out["%s_%s_gradpH" % (key, mode)] = np_(gp)
                out["%s_%s_ctrls" % (key, mode)] = np_(cc)
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, "functors.npz"), **out)
    print("functors", len(out), "arrays")


def gen_phi_random():
    out = {}
    cfgs = [(2, 16, 4, 9), (3, 24, 7, 5), (4, 8, 3, 6), (2, 40, 30, 4), (5, 12, 2, 3)]
    for idx, (nTh, m, d, n) in enumerate(cfgs):
        torch.set_default_dtype(torch.float32)
        torch.manual_seed(100 + idx)
        net = Phi(nTh=nTh, m=m, d=d)
        # de-trivialise the defaults (w=1, c=0) and decouple the deep-copied residual layers
        with torch.no_grad():
            net.w.weight.normal_(); net.c.weight.normal_(); net.c.bias.normal_()
            for lay in net.N.layers[1:]:
                lay.weight.normal_(std=0.3); lay.bias.normal_()
        x = torch.randn(n, d + 1, dtype=torch.float32) * 1.5
        for k, v in net.state_dict().items():
            out["net%d_%s" % (idx, k)] = np_(v)
        out["net%d_dims" % idx] = np.array([nTh, m, d])
        out["net%d_x" % idx] = np_(x)
        for tag, prec in (("f32", torch.float32), ("f64", torch.float64)):
            netp = Phi(nTh=nTh, m=m, d=d).to(prec)
            netp.load_state_dict({k: v.to(prec) for k, v in net.state_dict().items()})
            with torch.no_grad():
                out["net%d_fwd_%s" % (idx, tag)] = np_(netp(x.to(prec)))
                out["net%d_grad_%s" % (idx, tag)] = np_(netp.getGrad(x.to(prec)))
    np.savez_compressed(os.path.join(HERE, "phi_random.npz"), **out)
    print("phi_random", len(out), "arrays")


def gen_config5():
    """BASELINE.json configs[4]: swarm50-shape random-init Phi, full validation loss, nt=50, fp64."""
    torch.set_default_dtype(torch.float32)
    alph = [1800.0, 1e7, 25000.0, 2.0, 1.0, 3.0]
    torch.manual_seed(0)
    net = Phi(nTh=2, m=512, d=150, alph=alph)
    sd = net.state_dict()
    checks = {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in sd.items()}
    net = net.double()
    torch.set_default_dtype(torch.float64)
    prob, _, _, xInit = initProb("swarm50", 4, 4, var0=0.1, alph=alph, cvt=lambda x: x.double())
    prob.eval()
    g = torch.Generator().manual_seed(5)
    x = xInit + 0.1 * torch.randn(12, 150, generator=g, dtype=torch.float64)
    mean, nomean, zF, cF = run_modes(x, net, prob, [0.0, 1.0], 50)
    out = dict(x=np_(x), mean_f64=mean, nomean_f64=nomean, z_last_f64=zF[:, :, -1], ctrl_last_f64=cF[:, :, -1],
               alph=np.array(alph))
    for k, v in checks.items():
        out["chk_" + k] = v
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, "config5.npz"), **out)
    print("config5 Jc = %.12e" % mean[0])


def gen_initprob():
    out = {}
    torch.set_default_dtype(torch.float64)
    for name in ALL_NAMES:
        prob, x0, x0v, xInit = initProb(name, 4, 4, var0=1.0, alph=[1.0, 2.0, 3.0, 1.0, 1.0, 1.0], cvt=lambda x: x.double())
        out[name + "_xtarget"] = np_(prob.xtarget)
        out[name + "_xinit"] = np_(xInit)
        out[name + "_meta"] = np.array(json.dumps(dict(
            cls=type(prob).__name__, obstacle=prob.obstacle, alph_Q=float(prob.alph_Q), alph_W=float(prob.alph_W),
            r=float(prob.r), nAgents=int(prob.nAgents), agentDim=int(prob.agentDim))))
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, "initprob.npz"), **out)
    print("initprob", len(ALL_NAMES), "problems")


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "ckpt"), exist_ok=True)
    gen_ckpts()
    gen_initprob()
    gen_functors()
    gen_phi_random()
    gen_cases()
    gen_config5()
