"""Batch-1 (small batch) rollout through the one-CTA-per-sample kernel (profiling driver): python scripts/vec_run.py name [n] [reps]"""
import os, sys, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuraloc_b200 as nb
from helpers import product_setup
name = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 1; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
net, prob, xinit, meta = product_setup(name, torch.float32)
nt = 80 if name == "swarm50" else 50
x = xinit.repeat(n, 1).contiguous()
ms = []
with torch.no_grad():
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s = nb.ocflow_sums(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"]); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print("%s n=%d nt=%d path=%s: median %.3f ms  (%.2f us per evaluation)  Jsum %.6e" % (name, n, nt, nb._cabi.last_path(), statistics.median(ms), statistics.median(ms) * 1e3 / (4 * nt + 1), float(s[0])))
