"""Least-squares fit of the three accumulate-bias compensation factors (GEMM-1, GEMM-2/3, GEMM-4) of the streamed swarm
kernel: the mean terminal-state error vector vs the fp64 oracle responds linearly to each factor."""
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
n, nt = 512, 80
g = torch.Generator().manual_seed(11)
x = xinit.cpu() + 0.1 * torch.randn(n, d, generator=g)
with torch.no_grad():
    z64, _ = orc.ocflow(x.double(), P64, D64, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
    m64 = mean_vec(orc.ocflow(x.double(), P64, D64, [0.0, 1.0], nt, "rk4", alph))
z64 = z64.numpy()
def run(f):
    os.environ["NOC_TS_BIAS"] = ",".join("%.6f" % v for v in f)
    with torch.no_grad():
        zs, _ = nb.OCflow(x.cuda(), net, prob, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
        ms = mean_vec(nb.OCflow(x.cuda(), net, prob, [0.0, 1.0], nt, "rk4", alph))
    zs = zs.cpu().numpy()
    # error signature: mean over samples of the whole trajectory error (all steps), flattened
    return (zs[:, :d, :] - z64[:, :d, :]).mean(0).ravel(), zs, ms
d0, _, _ = run((0, 0, 0))
cols = []
for i in range(3):
    f = [0, 0, 0]; f[i] = 2.0
    cols.append((run(f)[0] - d0) / 2.0)
Amat = np.stack(cols, 1)
fit, *_ = np.linalg.lstsq(Amat, -d0, rcond=None)
print("fit f1, f2, f4 =", fit, " residual %.3e of %.3e" % (np.linalg.norm(d0 + Amat @ fit), np.linalg.norm(d0)))
for f in (tuple(fit), (1.6, 1.6, 1.6)):
    dd, zs, ms = run(f)
    ce = np.abs(ms - m64) / np.maximum(np.abs(m64), 1e-12)
    dx = zs[:, :d, -1] - z64[:, :d, -1]
    print("f=%s: state %.2e  costs %s  |mean dx| %.2e rms %.2e" % (np.round(f, 3), rel_state_err(zs, z64, d), " ".join("%.1e" % v for v in ce[:6]),
          np.linalg.norm(dx.mean(0)), np.sqrt((dx ** 2).sum(1).mean())), flush=True)
