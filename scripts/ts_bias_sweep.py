"""Sweep of the accumulate-bias compensation factor of the streamed swarm kernel against the fp64 oracle."""
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
n, nt = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 80
g = torch.Generator().manual_seed(1234)
x = xinit.cpu() + 0.1 * torch.randn(n, d, generator=g)
P64, D64, _, _ = oracle_setup(name, torch.float64)
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    z64, _ = orc.ocflow(x.double(), P64, D64, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
    m64 = mean_vec(orc.ocflow(x.double(), P64, D64, [0.0, 1.0], nt, "rk4", alph))
z64 = z64.numpy()
def report(tag):
    with torch.no_grad():
        zs, _ = nb.OCflow(x.cuda(), net, prob, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
        ms = mean_vec(nb.OCflow(x.cuda(), net, prob, [0.0, 1.0], nt, "rk4", alph))
    zs = zs.cpu().numpy()
    dx = zs[:, :d, -1] - z64[:, :d, -1]
    # systematic part: the mean error vector over samples vs the rms per-sample error
    print("%-12s state err %.3e | mean-cost rel err [Jc L G HJt HJf HJg] %s | |mean dx| %.2e  rms |dx| %.2e" %
          (tag, rel_state_err(zs, z64, d), " ".join("%.1e" % v for v in (np.abs(ms - m64) / np.maximum(np.abs(m64), 1e-12))[:6]),
           np.linalg.norm(dx.mean(0)), np.sqrt((dx ** 2).sum(1).mean())), flush=True)
os.environ["NOC_FORCE_PATH"] = "tile"; report("tile")
os.environ["NOC_FORCE_PATH"] = "tc"
for f in sys.argv[2:] or ["0", "0.5", "1", "1.5", "2", "3"]:
    os.environ["NOC_TS_BIAS"] = f; report("ts f=" + f)
