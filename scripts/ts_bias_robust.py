"""How stable is the optimal accumulate-bias compensation of the streamed swarm kernel?  For several input distributions /
step counts / weights: the mean terminal-state error vector (vs the fp64 oracle) at two compensation factors, and the
factor at which its projection crosses zero (linear interpolation)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuraloc_b200 as nb
from helpers import oracle_setup, product_setup, rel_state_err, mean_vec
from oracle import ocflow_oracle as orc
name = "swarm50"
net, prob, xinit, meta = product_setup(name, torch.float32)
d = xinit.shape[1]; alph = meta["alph"]
P64, D64, _, _ = oracle_setup(name, torch.float64)
torch.set_num_threads(os.cpu_count() or 1)
os.environ["NOC_FORCE_PATH"] = "tc"

def case(tag, x, nt, netg, P):
    with torch.no_grad():
        z64, _ = orc.ocflow(x.double(), P, D64, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
        m64 = mean_vec(orc.ocflow(x.double(), P, D64, [0.0, 1.0], nt, "rk4", alph))
    z64 = z64.numpy()
    res = {}
    for f in (1.0, 2.0):
        os.environ["NOC_TS_BIAS"] = str(f)
        with torch.no_grad():
            zs, _ = nb.OCflow(x.cuda(), netg, prob, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
            ms = mean_vec(nb.OCflow(x.cuda(), netg, prob, [0.0, 1.0], nt, "rk4", alph))
        dx = (zs.cpu().numpy()[:, :d, -1] - z64[:, :d, -1]).mean(0)
        res[f] = (dx, ms)
    d1, d2 = res[1.0][0], res[2.0][0]
    u = d1 - d2
    fstar = 1.0 + float(np.dot(d1, u) / np.dot(u, u))          # zero of the component along the response direction
    resid = d1 + (fstar - 1.0) * (d2 - d1)
    os.environ["NOC_TS_BIAS"] = "1.64"
    with torch.no_grad():
        zs, _ = nb.OCflow(x.cuda(), netg, prob, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
        ms = mean_vec(nb.OCflow(x.cuda(), netg, prob, [0.0, 1.0], nt, "rk4", alph))
    ce = np.abs(ms - m64) / np.maximum(np.abs(m64), 1e-12)
    print("%-28s f* = %.3f  |mean dx| at f=1: %.2e  residual at f*: %.2e | at f=1.64: state %.2e  G %.1e HJg %.1e L %.1e"
          % (tag, fstar, np.linalg.norm(d1), np.linalg.norm(resid), rel_state_err(zs.cpu().numpy(), z64, d), ce[2], ce[5], ce[1]), flush=True)

g = torch.Generator().manual_seed(5)
n = 256
case("bench dist nt=80", xinit.cpu() + 0.1 * torch.randn(n, d, generator=g), 80, net, P64)
case("bench dist nt=40", xinit.cpu() + 0.1 * torch.randn(n, d, generator=g), 40, net, P64)
case("bench dist nt=20", xinit.cpu() + 0.1 * torch.randn(n, d, generator=g), 20, net, P64)
case("var 0.5 nt=80", xinit.cpu() + 0.5 * torch.randn(n, d, generator=g), 80, net, P64)
case("var 1.0 nt=80", xinit.cpu() + 1.0 * torch.randn(n, d, generator=g), 80, net, P64)
# random-init weights of the same shape (config 5's net in fp32)
torch.manual_seed(0)
net2 = nb.Phi(nTh=2, m=512, d=150, alph=alph)
P2 = orc.params_from_state_dict({k: v.detach().clone() for k, v in net2.state_dict().items()}, torch.float64)
net2 = net2.float().cuda()
case("random-init net nt=50", xinit.cpu() + 0.1 * torch.randn(n, d, generator=g), 50, net2, P2)
