"""Prints, per problem, how far the CUDA fp32 rollout and the fp32 CPU oracle each are from the fp64 CPU oracle
(per-step state, mean costs).  Run on the GPU box; NOC_LIB selects an alternative build of the library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import neuraloc_b200 as nb
from helpers import oracle_setup, product_setup, rel_state_err, mean_vec
from oracle import ocflow_oracle as orc

torch.set_num_threads(os.cpu_count())
print("library:", nb._cabi.LIB_PATH)
ONLY = os.environ.get("PROBE_ONLY")
NT = int(os.environ.get("PROBE_NT", "0"))
SEED = int(os.environ.get("PROBE_SEED", "77"))
NS = int(os.environ.get("PROBE_N", "0"))
for name, n, nt in (("softcorridor", 512, 50), ("swap2", 512, 50), ("swap12", 512, 50), ("singlequad", 512, 50), ("swarm50", 256, 80)):
    if ONLY and name != ONLY:
        continue
    nt = NT or nt
    n = NS or n
    net, prob, xinit, meta = product_setup(name, torch.float32)
    P32, D32, _, _ = oracle_setup(name, torch.float32)
    P64, D64, _, _ = oracle_setup(name, torch.float64)
    d = xinit.shape[1]
    g = torch.Generator().manual_seed(SEED)
    if name == "singlequad":
        x = torch.zeros(n, d); x[:, :3] = -1.5 + meta["var0"] * torch.randn(n, 3, generator=g)
    else:
        x = xinit.cpu() + meta["var0"] * torch.randn(n, d, generator=g)
    with torch.no_grad():
        z64, _ = orc.ocflow(x.double(), P64, D64, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
        m64 = mean_vec(orc.ocflow(x.double(), P64, D64, [0.0, 1.0], nt, "rk4", meta["alph"]))
        z32, _ = orc.ocflow(x, P32, D32, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
        m32 = mean_vec(orc.ocflow(x, P32, D32, [0.0, 1.0], nt, "rk4", meta["alph"]))
        zg, _ = nb.OCflow(x.cuda(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
        mg = mean_vec(nb.OCflow(x.cuda(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"]))
    rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    print("%-12s state: cuda-f64 %.2e  cpu32-f64 %.2e  cuda-cpu32 %.2e" % (
        name, rel_state_err(zg.cpu().numpy(), z64.numpy(), d), rel_state_err(z32.numpy(), z64.numpy(), d),
        rel_state_err(zg.cpu().numpy(), z32.numpy(), d)))
    print("   costs [Jc L G HJt HJf HJg Q W] cuda-f64:", " ".join("%.1e" % v for v in rel(mg, m64)))
    print("                                   cpu32-f64:", " ".join("%.1e" % v for v in rel(m32, m64)))
