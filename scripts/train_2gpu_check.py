"""Data-parallel training evaluation over NCCL: torchrun --nproc-per-node N scripts/train_2gpu_check.py [problem]
Every rank runs noc_ocflow_grad on its row shard, one all-reduce of [8 cost sums | P gradient sums] follows (ocflow_grad_sharded);
rank 0 also evaluates the whole batch on its own GPU and prints the largest difference."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuraloc_b200 as nb
from helpers import product_setup
name = sys.argv[1] if len(sys.argv) > 1 else "swap12"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
net, prob, xinit, meta = product_setup(name, torch.float64, device="cuda:%d" % local)
prob.train()
n, nt = 1003, 8
g = torch.Generator().manual_seed(7)
x = (xinit.cpu() + meta["var0"] * torch.randn(n, xinit.shape[1], generator=g, dtype=torch.float64)).cuda()
lo, hi = nb.shard_rows(n, world, rank)
Jc, cs, grads = nb.ocflow_grad_sharded(x[lo:hi].contiguous(), net, prob, [0.0, 1.0], nt, meta["alph"])
if rank == 0:
    sums, flat, _ = nb.ocflow_grad_sums(x, net, prob, [0.0, 1.0], nt, meta["alph"])
    full = nb.split_param_grads(net, flat / sums[7])
    Jfull, _ = nb.costs_from_sums(sums, [float(a) for a in meta["alph"]], torch.float64)
    err = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-300)) for a, b in zip(grads, full))
    print("train_2gpu_check %s world=%d n=%d nt=%d: Jc sharded %.12e full %.12e  max rel gradient difference %.2e  (p.grad set: %s)"
          % (name, world, n, nt, float(Jc), float(Jfull), err, all(p.grad is not None for p in net.parameters())))
dist.destroy_process_group()
