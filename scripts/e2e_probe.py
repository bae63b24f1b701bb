"""Wall-clock split of the host entry point: python scripts/e2e_probe.py [samples] [workload]
(device-resident call vs host-buffer call, same process, alternating, so that both see the same thermal state)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import neuraloc_b200 as nb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
wl = sys.argv[2] if len(sys.argv) > 2 else "swarm50"
NT = bench.WORKLOADS[wl]["nt"]
dev = torch.device("cuda", 0)
net, prob, xinit, meta = bench.build_case(wl, dev, torch.float32)
x = bench.sample_x(wl, xinit, meta["var0"], n, 1234, dev, torch.float32)
xh = x.cpu().pin_memory()
with torch.no_grad():
    for _ in range(2):
        nb.ocflow_sums(x, net, prob, [0.0, 1.0], NT, "rk4", meta["alph"]); nb.ocflow_sums(xh, net, prob, [0.0, 1.0], NT, "rk4", meta["alph"])
    torch.cuda.synchronize()
    td, th = [], []
    for _ in range(int(os.environ.get("NOC_PROBE_REPS", "4"))):
        t0 = time.perf_counter(); nb.ocflow_sums(x, net, prob, [0.0, 1.0], NT, "rk4", meta["alph"]); torch.cuda.synchronize(); td.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); nb.ocflow_sums(xh, net, prob, [0.0, 1.0], NT, "rk4", meta["alph"]); torch.cuda.synchronize(); th.append(time.perf_counter() - t0)
print("e2e_probe %s n=%d pool_keep=%s: device-resident %s s | host buffers %s s | ratio %.4f" % (
    wl, n, os.environ.get("NOC_POOL_KEEP_MB", "default"), ["%.3f" % t for t in td], ["%.3f" % t for t in th], sum(td) / sum(th)))
