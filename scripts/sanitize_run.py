"""One small rollout (three return modes) for compute-sanitizer: python scripts/sanitize_run.py name n nt f32|f64 [path]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuraloc_b200 as nb
from helpers import product_setup
name, n, nt, dt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
if len(sys.argv) > 5:
    os.environ["NOC_FORCE_PATH"] = sys.argv[5]
dtype = torch.float32 if dt == "f32" else torch.float64
net, prob, xinit, meta = product_setup(name, dtype)
g = torch.Generator().manual_seed(3)
x = (xinit.cpu() + 0.1 * torch.randn(n, xinit.shape[1], generator=g).to(dtype)).cuda()
with torch.no_grad():
    Jc, cs = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
    path = nb._cabi.last_path()
    nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
    nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
torch.cuda.synchronize()
print("sanitize_run %s n=%d nt=%d %s kernel=%s Jc=%.6e" % (name, n, nt, dt, path, float(Jc)))
