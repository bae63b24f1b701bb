"""Per-step state error of the FMA tile kernel and the tensor-core kernel against the reference's fp64 and fp32 golden
trajectories (tests/golden/cases_*.npz).  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import neuraloc_b200 as nb
from helpers import load_cases, product_setup, rel_state_err

for name in ("softcorridor", "swap2", "swap12", "singlequad"):
    c = load_cases(name)
    net, prob, xinit, meta = product_setup(name, torch.float32)
    d = xinit.shape[1]
    xb = torch.from_numpy(c["xb"]).float().cuda()
    cases = [("xinit", xinit, int(c["nt"]), "rk4", "xinit_z_"), ("batch", xb, int(c["nt_batch"]), "rk4", "b_z_"), ("rk1", xb[:4], 8, "rk1", "rk1_z_")]
    for tag, x, nt, stp, key in cases:
        row = []
        for path in ("tile", "tc"):
            os.environ["NOC_FORCE_PATH"] = path
            with torch.no_grad():
                z, _ = nb.OCflow(x, net, prob, [0.0, 1.0], nt, stp, meta["alph"], intermediates=True)
            z = z.cpu().numpy()
            row.append("%s: vs f64 %.2e vs f32 %.2e" % (path, rel_state_err(z, c[key + "f64"], d), rel_state_err(z, c[key + "f32"], d)))
        print("%-12s %-6s ref32 vs f64 %.2e | %s" % (name, tag, rel_state_err(c[key + "f32"], c[key + "f64"], d), " | ".join(row)))
