"""One small training evaluation (noc_ocflow_grad) and one baseline objective (noc_baseline_loss) for compute-sanitizer:
python scripts/sanitize_train.py name n nt f32|f64"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuraloc_b200 as nb
from helpers import product_setup
name, n, nt, dt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dtype = torch.float32 if dt == "f32" else torch.float64
net, prob, xinit, meta = product_setup(name, dtype)
prob.train()
g = torch.Generator().manual_seed(3)
x = (xinit.cpu() + 0.1 * torch.randn(n, xinit.shape[1], generator=g).to(dtype)).cuda()
sums, grad, gx = nb.ocflow_grad_sums(x, net, prob, [0.0, 1.0], nt, meta["alph"], want_xgrad=True)
nc = 4 if name == "singlequad" else x.shape[1]
U = torch.randn(n, nt, nc, generator=g).to(dtype).cuda()
loss, gU = nb.baseline_loss(U, x, prob, meta["alph"][0], want_grad=True)
torch.cuda.synchronize()
print("sanitize_train %s n=%d nt=%d %s sumL=%.6e |grad|=%.4e loss0=%.6e" % (name, n, nt, dt, float(sums[0]), float(grad.norm()), float(loss[0])))
