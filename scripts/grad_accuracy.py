"""fp32 training gradient of the CUDA kernel vs the fp64 oracle autograd, next to torch-CPU fp32 autograd's own distance from it,
on the adversarial test batch and on a plain batch from rho_0.   python scripts/grad_accuracy.py [problem ...]"""
import dataclasses, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuraloc_b200 as nb
from helpers import oracle_setup, product_setup
from test_adjoint_formulas import NAMES, adversarial_batch, autograd_of_oracle
ORDER = ["A", "c_w", "c_b", "w", "K0", "K1", "b0", "b1"]
for name in (sys.argv[1:] or ["swarm50", "swap12", "singlequad", "softcorridor", "swap2"]):
    net, prob, xinit, meta = product_setup(name, torch.float32)
    if os.environ.get("NOC_ACC_NOHJ"):
        meta["alph"] = list(meta["alph"][:3]) + [0.0, 0.0, 0.0]      # no |.| residual terms: is the fp32 gap their sign flips?
    prob.train()
    P, D, xi, _ = oracle_setup(name, torch.float64)
    D = dataclasses.replace(D, training=True)
    if os.environ.get("NOC_ACC_NOQW"):                                   # network-only: no terrain / interaction terms
        prob.alph_Q = prob.alph_W = 0.0
        D = dataclasses.replace(D, alph_Q=0.0, alph_W=0.0)
    print("alph", meta["alph"], "alph_Q/W", prob.alph_Q, prob.alph_W)
    for kind in ("adversarial", "plain"):
        n, nt = (6, 3) if name == "swarm50" else (13, 5)
        if kind == "adversarial":
            x = adversarial_batch(name, D, xi, meta["var0"], n)
        else:
            n = 64 if name == "swarm50" else 256
            if os.environ.get("NOC_ACC_NT"):                         # the reference's training nt: the rollout the net was trained for
                nt, n = int(os.environ["NOC_ACC_NT"]), min(n, 32)
            g = torch.Generator().manual_seed(5)
            x = xi + meta["var0"] * torch.randn(n, xi.shape[1], generator=g, dtype=torch.float64)
            if name == "singlequad":
                x[:, 3:] = 0
        J64, G64, X64 = autograd_of_oracle(x, P, D, [0.0, 1.0], nt, meta["alph"])
        J32, G32, X32 = autograd_of_oracle(x.float(), P.to(torch.float32), D.to(torch.float32), [0.0, 1.0], nt, meta["alph"])
        for ts in ("4", "8"):
            os.environ["NOC_GRAD_TS"] = ts
            try:
                sums, grad, gx = nb.ocflow_grad_sums(x.float().cuda(), net, prob, [0.0, 1.0], nt, meta["alph"], want_xgrad=True)
            except Exception as e:
                print(name, kind, "ts", ts, "failed:", str(e)[:80]); continue
            got = dict(zip(ORDER, nb.split_param_grads(net, grad)))
            rel = lambda a, b: float((a.double().cpu().reshape(b.shape) - b).abs().max() / b.abs().max().clamp_min(1e-300))
            print("%-12s %-11s n=%3d TS=%s  kernel fp32 vs fp64 oracle: %s   | torch fp32 vs fp64: %s" % (
                name, kind, n, ts, " ".join("%s %.1e" % (k, rel(got[k], G64[k])) for k in NAMES if k != "c_b"),
                " ".join("%.1e" % rel(G32[k], G64[k]) for k in NAMES if k != "c_b")))
            rms = lambda a, b: float(((a.double().cpu().reshape(b.shape) - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt())
            al = meta["alph"]
            Jk = float(sums[0] + al[0] * sums[1] + al[3] * sums[2] + al[4] * sums[3] + al[5] * sums[4])
            with torch.no_grad():
                from oracle import ocflow_oracle as orc
                c64 = [float(c.sum()) for c in orc.ocflow(x, P, D, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)[1]]
                c32 = [float(c.double().sum()) for c in orc.ocflow(x.float(), P.to(torch.float32), D.to(torch.float32), [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)[1]]
                zk, _ = nb.OCflow(x.float().cuda(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
                z64, _ = orc.ocflow(x, P, D, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
                z32, _ = orc.ocflow(x.float(), P.to(torch.float32), D.to(torch.float32), [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
            dd = x.shape[1]
            serr = lambda z: float(((z.double().cpu()[:, :dd, -1] - z64[:, :dd, -1]).norm(dim=1) / z64[:, :dd, -1].norm(dim=1)).max())
            print("      cost sums rel err [L G HJt HJfin HJgrad]: grad kernel %s | torch fp32 %s | final-state err: forward kernel %.1e torch fp32 %.1e" % (
                " ".join("%.1e" % (abs(float(sums[i]) - c64[i]) / max(abs(c64[i]), 1e-300)) for i in range(5)),
                " ".join("%.1e" % (abs(c32[i] - c64[i]) / max(abs(c64[i]), 1e-300)) for i in range(5)), serr(zk), serr(z32)))
            print("      grad_x max %.1e (torch %.1e) rms %.1e (torch %.1e) | K1 rms %.1e (torch %.1e) | J rel %.1e (torch %.1e)" % (
                rel(gx, X64), rel(X32, X64), rms(gx, X64), rms(X32, X64), rms(got["K1"], G64["K1"]), rms(G32["K1"], G64["K1"]),
                abs(Jk - float(J64)) / abs(float(J64)), abs(float(J32) - float(J64)) / abs(float(J64))))
