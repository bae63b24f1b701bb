"""Development check of the streamed tensor-core swarm kernel (noc_ts_rollout.cuh) on a B200:
parity of the three return modes against the CPU oracle (fp32 and fp64) and against the FMA tile kernel, then timing.
Usage: python scripts/ts_check.py [n_time] [nt_time]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuraloc_b200 as nb                                   # noqa: E402
from helpers import oracle_setup, product_setup, rel_state_err, mean_vec   # noqa: E402
from oracle import ocflow_oracle as orc                      # noqa: E402


def run(x, net, prob, nt, alph, path, mode):
    os.environ["NOC_FORCE_PATH"] = path
    with torch.no_grad():
        if mode == "mean":
            out = mean_vec(nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", alph))
        elif mode == "nomean":
            J, cs = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", alph, noMean=True)
            out = torch.cat([J] + list(cs), 1).cpu().numpy()
        else:
            z, c = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
            out = (z.cpu().numpy(), c.cpu().numpy())
    ran = nb._cabi.last_path()
    torch.cuda.synchronize()
    return out, ran


def main():
    n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
    nt_time = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    name = "swarm50"
    net, prob, xinit, meta = product_setup(name, torch.float32)
    d = xinit.shape[1]
    alph = meta["alph"]
    g = torch.Generator().manual_seed(7)
    for n, nt in ((300, 6), (129, 3)):
        x = xinit.cpu() + 0.1 * torch.randn(n, d, generator=g)
        xc = x.cuda()
        zt, ran_t = run(xc, net, prob, nt, alph, "tile", "inter")
        zs, ran_s = run(xc, net, prob, nt, alph, "tc", "inter")
        print("paths:", ran_t, ran_s, flush=True)
        print("n=%d nt=%d  state err ts vs tile: %.3e   ctrl abs err %.3e   cost-integral abs err %.3e" %
              (n, nt, rel_state_err(zs[0], zt[0], d), np.abs(zs[1] - zt[1]).max(), np.abs(zs[0][:, d:, :] - zt[0][:, d:, :]).max()), flush=True)
        mt, _ = run(xc, net, prob, nt, alph, "tile", "mean")
        ms, _ = run(xc, net, prob, nt, alph, "tc", "mean")
        print("  mean tile", mt)
        print("  mean ts  ", ms)
        print("  rel diff ", np.abs(ms - mt) / np.maximum(np.abs(mt), 1e-12), flush=True)
        nt_, _ = run(xc, net, prob, nt, alph, "tile", "nomean")
        ns_, _ = run(xc, net, prob, nt, alph, "tc", "nomean")
        sc = np.maximum(np.abs(nt_).max(axis=0, keepdims=True), 1e-12)
        print("  noMean max rel-to-column-scale diff", (np.abs(ns_ - nt_) / sc).max(axis=0), flush=True)
    # oracle (fp32 and fp64) on a small batch with the documented nt
    n, nt = 64, 80
    x = xinit.cpu() + 0.1 * torch.randn(n, d, generator=g)
    P32, D32, _, _ = oracle_setup(name, torch.float32)
    P64, D64, _, _ = oracle_setup(name, torch.float64)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        z32, _ = orc.ocflow(x, P32, D32, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
        z64, _ = orc.ocflow(x.double(), P64, D64, [0.0, 1.0], nt, "rk4", alph, intermediates=True)
        m64 = mean_vec(orc.ocflow(x.double(), P64, D64, [0.0, 1.0], nt, "rk4", alph))
    for path in ("tile", "tc"):
        zs, ran = run(x.cuda(), net, prob, nt, alph, path, "inter")
        ms, _ = run(x.cuda(), net, prob, nt, alph, path, "mean")
        print("%s (%s): state err vs fp32 oracle %.3e, vs fp64 oracle %.3e (fp32 oracle vs fp64: %.3e)" %
              (path, ran, rel_state_err(zs[0], z32.numpy(), d), rel_state_err(zs[0], z64.numpy(), d),
               rel_state_err(z32.numpy(), z64.numpy(), d)))
        print("   mean-cost rel err vs fp64 oracle:", np.abs(ms - m64) / np.maximum(np.abs(m64), 1e-12), flush=True)
    # timing
    gg = torch.Generator(device="cuda").manual_seed(1234)
    x = xinit + 0.1 * torch.randn(n_time, d, generator=gg, device="cuda")
    for path in ("tc", "tile"):
        os.environ["NOC_FORCE_PATH"] = path
        with torch.no_grad():
            nb.ocflow_sums(x, net, prob, [0.0, 1.0], nt_time, "rk4", alph)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s = nb.ocflow_sums(x, net, prob, [0.0, 1.0], nt_time, "rk4", alph)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("time %s: n=%d nt=%d  %.2f ms  -> %.3e sample-steps/s   Jsum=%s" % (path, n_time, nt_time, ms, n_time * nt_time / ms * 1e3, s[:2].tolist()), flush=True)


if __name__ == "__main__":
    main()
