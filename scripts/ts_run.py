"""One swarm50 rollout through the streamed tensor-core kernel (profiling driver): python scripts/ts_run.py n nt [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuraloc_b200 as nb
from helpers import product_setup
n, nt = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
net, prob, xinit, meta = product_setup("swarm50", torch.float32)
g = torch.Generator(device="cuda").manual_seed(1234)
x = xinit + 0.1 * torch.randn(n, xinit.shape[1], generator=g, device="cuda")
os.environ.setdefault("NOC_FORCE_PATH", "tc")
with torch.no_grad():
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s = nb.ocflow_sums(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("path %s n=%d nt=%d: %.3f ms -> %.3e sample-steps/s (Jsum %.6e)" % (nb._cabi.last_path(), n, nt, ms, n * nt / ms * 1e3, float(s[0])), flush=True)
