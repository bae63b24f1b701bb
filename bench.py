#!/usr/bin/env python
"""bench.py — RK4 sample-steps/sec of the closed-loop rollout (BASELINE.json metric) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload swap12|singlequad|swarm50|...] [--impl reference]

A "step" is one OCflow call (mean mode) over one batch of synthetic initial states drawn from the problem's
initial distribution; the default workload is BASELINE.json configs[1]: the swap12 pretrained
checkpoint, nt = 50, 2^20 samples per GPU, fp32.  Under torchrun (N > 1) every rank rolls out its own shard
(weak scaling: per-GPU work fixed) and the step ends with the single all-reduce of the 8-double cost vector.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same metric through the public
API with host (pinned) buffers, copies inside the timed region; `roofline` = algorithmic FLOP/s of the rollout
kernel against the FP32 (or FP64) FMA peak measured live on the same device; `cpu_baseline` = the CPU oracle port
(torch CPU, all host threads) timed on a bounded sample of the same workload.
`--impl reference` times only that CPU arm (the reference is pure Python/torch and cannot travel to the GPU box;
oracle/ocflow_oracle.py is its restatement, pinned against the reference's outputs in tests/).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np   # noqa: E402
import torch         # noqa: E402

# per-GPU samples, nt, dtype, CPU-sample size; flops/sample-step = 4 (4 m D + 4 m^2 (nTh-1) + min(2 D^2, 4 r D))
WORKLOADS = {
    "softcorridor": dict(ckpt="softcorridor", n=1 << 20, nt=50, dtype="f32", n_cpu=65536),
    "swap2": dict(ckpt="swap2", n=1 << 20, nt=50, dtype="f32", n_cpu=65536),
    "swap12": dict(ckpt="swap12", n=1 << 20, nt=50, dtype="f32", n_cpu=65536),
    "singlequad": dict(ckpt="singlequad", n=1 << 22, nt=50, dtype="f32", n_cpu=32768),
    "swarm50": dict(ckpt="swarm50", n=1 << 24, nt=80, dtype="f32", n_cpu=2048),
    "config5": dict(ckpt=None, n=1 << 20, nt=50, dtype="f64", n_cpu=1024),
}


# DRAM bytes per sample of the rollout kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
# capture, divided by the samples of that launch): profiles/r01_<workload>_<kernel>_ncu_summary.txt.  The algorithmic
# figure is 4 d bytes per sample (the initial state is read once; mean mode writes nothing per sample); the excess is
# register-spill / scratch-state write-back, negligible against the kernel's duration (compute-bound).
DRAM_BYTES_PER_SAMPLE = {
    ("swap12", "tensor"): (25.347328e6 + 15.872e6) / 262144,
    ("singlequad", "tensor"): (12.783872e6 + 1.024e6) / 262144,
    ("swap12", "tile"): (12.819456e6 + 256.0e6) / 131072,
    ("swarm50", "tile"): (14.542336e6 + 1.271296e6) / 16384,
}


def tensor_flops_per_sample_step(d, m, quadcopter):
    """bf16 tensor-core flops the tcgen05 kernel EXECUTES per sample-step (noc_tc_rollout.cuh): 4 evaluations x 6 split
    products x 2 flops x the padded GEMM volumes  KS*mp + 2*mp*mp + mp*KS + KS*KS  (KS = d+2 rounded up to 16; the width is
    padded to the epilogue chunk x threads per sample of the shape: 64 for the quadcopter shape, 16 otherwise)."""
    ks = -(-(d + 2) // 16) * 16
    pad = 64 if quadcopter else 16
    mp = -(-m // pad) * pad
    return 4 * 6 * 2 * (ks * mp + 2 * mp * mp + mp * ks + ks * ks)


def roofline(workload, W, d, meta, n, nt, fl, step_s, fma_peak, path):
    """Roofline of the rollout kernel of one launch.  FMA kernels: algorithmic fp32/fp64 flops against the FMA peak measured
    live.  Tensor-core kernel: the bf16 flops it executes against MEASURED_PEAKS.json's sustained dense bf16 figure; the
    algorithmic (fp32-equivalent) rate and the FMA peak stay alongside, since that ratio is what the kernel replaces."""
    alg = n * nt * fl / step_s / 1e12
    traffic = DRAM_BYTES_PER_SAMPLE[(workload, path)] * n if (workload, path) in DRAM_BYTES_PER_SAMPLE else None
    note = ("DRAM bytes per launch = measured bytes per sample of the ncu capture in profiles/ x samples; "
            "algorithmic = %d bytes (4 d per sample)" % (4 * d * n))
    if path == "tensor":
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tpeak = float(peaks.get("bf16_tflops_sustained", 0) or 0) or 2250.0
        src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks.get("bf16_tflops_sustained")
               else "B200_PROFILING.md fallback: nominal dense bf16 2250 TFLOP/s")
        ach = n * nt * tensor_flops_per_sample_step(d, meta["m"], meta.get("data") == "singlequad") / step_s / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": traffic,
                "traffic_note": note, "peak_source": src, "kernel": "rollout_tc_kernel (tcgen05, 3-way bf16 split: 6 MMAs per fp32 product)",
                "algorithmic_fp32_tflops": alg, "fp32_fma_peak_tflops": fma_peak, "algorithmic_over_fma_peak": alg / fma_peak}
    return {"bound": "fp32_fma" if W["dtype"] == "f32" else "fp64_fma", "achieved": alg, "peak": fma_peak, "unit": "TFLOP/s",
            "frac": alg / fma_peak, "traffic": traffic, "traffic_note": note, "kernel": "rollout_kernel (FMA sample tiles)",
            "peak_source": "measured live: register-resident FMA micro-benchmark on all SMs (noc_measure_fma_peak); "
                           "MEASURED_PEAKS.json has no FP32/FP64 FMA figure"}


def flops_per_sample_step(d, m, nTh, r):
    D = d + 1
    return 4 * (4 * m * D + 4 * m * m * (nTh - 1) + min(2 * D * D, 4 * r * D))


def build_case(workload, device, dtype):
    """-> net, prob, xInit (on device), meta, var0 — from committed fixtures only (no /root/reference)."""
    import neuraloc_b200 as nb
    from helpers import load_ckpt
    w = WORKLOADS[workload]
    cvt = lambda v: v.to(dtype).to(device)
    if w["ckpt"] is None:       # BASELINE.json configs[4]: random-init swarm50-shape Phi
        alph = [1800.0, 1e7, 25000.0, 2.0, 1.0, 3.0]
        torch.manual_seed(0)
        net = nb.Phi(nTh=2, m=512, d=150, alph=alph)
        meta = dict(data="swarm50", alph=alph, var0=0.1, m=512, nTh=2)
    else:
        sd, meta = load_ckpt(w["ckpt"])
        net = None
    prob, x0, _, xinit = nb.initProb(meta["data"], 2, 2, var0=1.0, alph=meta["alph"], cvt=cvt)
    prob.eval()
    if net is None:
        net = nb.Phi(nTh=meta["nTh"], m=meta["m"], d=x0.shape[1], alph=meta["alph"])
        net.load_state_dict(sd)
    net = net.to(dtype).to(device)
    net.eval()
    return net, prob, xinit, meta


def sample_x(workload, xinit, var0, n, seed, device, dtype):
    """x ~ rho_0 of the problem: xInit + var0 N(0,I); singlequad perturbs the position only."""
    g = torch.Generator(device=device).manual_seed(seed)
    d = xinit.shape[1]
    if workload == "singlequad":
        x = torch.zeros(n, d, device=device, dtype=dtype)
        x[:, :3] = -1.5 + var0 * torch.randn(n, 3, generator=g, device=device, dtype=dtype)
        return x
    x = torch.randn(n, d, generator=g, device=device, dtype=dtype)
    x.mul_(var0).add_(xinit.to(device))
    return x


def cpu_rate(workload, n_cpu, nt, dtype, threads):
    """sample-steps/s of the CPU oracle port on a bounded sample (one small warm-up call, one timed call)."""
    from oracle import ocflow_oracle as orc
    from helpers import load_ckpt
    w = WORKLOADS[workload]
    torch.set_num_threads(threads)
    if w["ckpt"] is None:
        import neuraloc_b200 as nb
        alph = [1800.0, 1e7, 25000.0, 2.0, 1.0, 3.0]
        torch.manual_seed(0)
        sd = nb.Phi(nTh=2, m=512, d=150, alph=alph).state_dict()
        meta = dict(data="swarm50", alph=alph, var0=0.1)
    else:
        sd, meta = load_ckpt(w["ckpt"])
    P = orc.params_from_state_dict(sd, dtype)
    D, xinit = orc.make_problem(meta["data"], meta["alph"], dtype)
    x = sample_x(workload, xinit, meta["var0"], n_cpu, 1234, "cpu", dtype)
    with torch.no_grad():
        orc.ocflow(x[: max(1, n_cpu // 16)], P, D, [0.0, 1.0], nt, "rk4", meta["alph"])
        t0 = time.perf_counter()
        orc.ocflow(x, P, D, [0.0, 1.0], nt, "rk4", meta["alph"])
        dt = time.perf_counter() - t0
    return n_cpu * nt / dt, dt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            f = [v.strip() for v in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            smax = mx
            if t_begin - 0.05 <= t <= t_end + 0.05:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:       # region shorter than the sampling period: use whatever was seen
            sm = [float(l.split(",")[0]) for _, l in self.rows if l and l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def latency_mode(args, W, dtype, threads):
    """Batch-1 rollout latency: ONE OCflow(xInit) call, wall clock, CPU tensors in and out — the protocol of
    timeDeployment/timeOC.py:76-81 (nt = 50; 80 for swarm50 as in README.md:85) — beside the CPU oracle port timed on
    one thread (the published numbers were taken under `taskset -c 0`, README.md:152-155)."""
    import neuraloc_b200 as nb
    from oracle import ocflow_oracle as orc
    from helpers import load_ckpt
    nb._cabi.lib()
    torch.cuda.set_device(0)
    nt = W["nt"]
    net, prob, xinit, meta = build_case(args.workload, torch.device("cpu"), dtype)      # CPU tensors, like timeOC.py
    alph = meta["alph"]
    with torch.no_grad():
        for _ in range(5):
            nb.OCflow(xinit, net, prob, [0.0, 1.0], nt, "rk4", alph)
        wall = []
        for _ in range(200):
            t0 = time.perf_counter()
            Jc, cs = nb.OCflow(xinit, net, prob, [0.0, 1.0], nt, "rk4", alph)        # H2D + rollout + D2H + sync inside
            wall.append(time.perf_counter() - t0)
        xd = xinit.cuda()
        netd = net.cuda()
        ev = []
        for _ in range(100):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); nb.ocflow_sums(xd, netd, prob, [0.0, 1.0], nt, "rk4", alph); e1.record()
            ev.append((e0, e1))
        torch.cuda.synchronize()
        dev_ms = statistics.median(a.elapsed_time(b) for a, b in ev)
        # CPU oracle port, one thread
        torch.set_num_threads(1)
        if W["ckpt"] is None:
            sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        else:
            sd, _ = load_ckpt(W["ckpt"])
        P = orc.params_from_state_dict(sd, dtype)
        D, xi = orc.make_problem(meta["data"], alph, dtype)
        orc.ocflow(xi, P, D, [0.0, 1.0], nt, "rk4", alph)
        cpu = []
        for _ in range(5):
            t0 = time.perf_counter()
            Jr, _ = orc.ocflow(xi, P, D, [0.0, 1.0], nt, "rk4", alph)
            cpu.append(time.perf_counter() - t0)
    line = {"metric": "batch1_rollout_latency", "value": statistics.median(wall) * 1e3, "unit": "ms", "n_gpus": 1,
            "higher_is_better": False, "dtype": W["dtype"], "data": "xInit of the problem",
            "config": {"workload": "%s, one OCflow(xInit) call, nt=%d, CPU tensors in/out (timeOC.py:76-81)" % (args.workload, nt)},
            "host_wall_ms": {"median": statistics.median(wall) * 1e3, "p10": sorted(wall)[20] * 1e3, "p90": sorted(wall)[180] * 1e3},
            "device_ms_median": dev_ms, "ms_per_rk4_step": statistics.median(wall) * 1e3 / nt,
            "cpu_baseline": {"value": statistics.median(cpu) * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                             "sample": "median of 5 warm calls of the torch-CPU oracle port, 1 thread"},
            "Jc": float(Jc), "Jc_cpu": float(Jr)}
    return line


def intermediates_mode(args, W, dtype):
    """OCflow(..., intermediates=True) on a device-resident batch: zFull [n,d+4,nt+1] and ctrlFull [n,nCtrl,nt+1] are written
    to HBM (5 grad-Phi evaluations per step instead of 4).  Reports sample-steps/s and the output bandwidth."""
    import neuraloc_b200 as nb
    nb._cabi.lib()
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    n, nt = (args.n or min(W["n"], 1 << 18)), W["nt"]
    net, prob, xinit, meta = build_case(args.workload, device, dtype)
    x = sample_x(args.workload, xinit, meta["var0"], n, 1234, device, dtype)
    with torch.no_grad():
        for _ in range(max(1, args.warmup)):
            zf, cf = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            zf, cf = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
    out_bytes = (zf.numel() + cf.numel()) * zf.element_size()
    t = statistics.mean(ms) * 1e-3
    return {"metric": "rk4_sample_steps_per_sec_intermediates", "value": n * nt / t, "unit": "sample-steps/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": statistics.mean(ms), "dtype": W["dtype"],
            "config": {"workload": "%s, intermediates=True, %d samples, nt=%d" % (args.workload, n, nt)},
            "output_bytes": out_bytes, "output_GBps": out_bytes / t / 1e9,
            "roofline": {"bound": "hbm", "achieved": out_bytes / t / 1e9, "peak": 6539.2, "unit": "GB/s", "frac": out_bytes / t / 1e9 / 6539.2,
                         "note": "peak = MEASURED_PEAKS.json hbm_gbs; this mode is still compute-bound at these sizes"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="swap12", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", dest="n", type=int, default=0, help="samples per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency", action="store_true", help="batch-1 rollout latency (timeDeployment/timeOC.py protocol) instead of throughput")
    ap.add_argument("--intermediates", action="store_true", help="time OCflow(..., intermediates=True): trajectories + controls written to HBM (SURVEY.md 8f N2)")
    args = ap.parse_args()
    W = WORKLOADS[args.workload]
    n = args.n or W["n"]
    nt = W["nt"]
    dtype = torch.float32 if W["dtype"] == "f32" else torch.float64
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm: CPU oracle port on host cores
    if args.impl == "reference":
        if rank != 0:
            return
        rates = []
        for _ in range(max(1, args.steps)):
            r, dt = cpu_rate(args.workload, W["n_cpu"], nt, dtype, threads)
            rates.append((r, dt))
        rate = W["n_cpu"] * nt * len(rates) / sum(dt for _, dt in rates)
        sample = "%d samples x nt=%d per step (of %d), torch CPU %s, %d threads" % (W["n_cpu"], nt, n, torch.__version__, threads)
        line = {"impl": "reference", "metric": "rk4_sample_steps_per_sec", "value": rate, "unit": "sample-steps/s",
                "n_gpus": args.gpus, "steps": len(rates), "warmup": 1, "ms_per_step": 1e3 * sum(dt for _, dt in rates) / len(rates),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": W["dtype"], "data": "synthetic",
                "config": {"workload": "%s pretrained checkpoint, nt=%d, x ~ xInit + var0*N(0,I)" % (args.workload, nt),
                           "samples_per_step": W["n_cpu"]},
                "cpu_baseline": {"value": rate, "unit": "sample-steps/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": rate, "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if args.latency:
        print(json.dumps(latency_mode(args, W, dtype, threads)))
        return
    if args.intermediates:
        print(json.dumps(intermediates_mode(args, W, dtype)))
        return

    # ------------------------------------------------------------------ our arm
    import neuraloc_b200 as nb
    lib = nb._cabi.lib()      # raises if the CUDA library is missing: no fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    net, prob, xinit, meta = build_case(args.workload, device, dtype)
    d = xinit.shape[1]
    x = sample_x(args.workload, xinit, meta["var0"], n, 1234 + rank, device, dtype)
    alph = meta["alph"]
    fl = flops_per_sample_step(d, meta["m"], meta["nTh"], min(10, d + 1))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # > L2 (126 MB)

    def step():
        sums = nb.ocflow_sums(x, net, prob, [0.0, 1.0], nt, "rk4", alph)
        if dist is not None:
            dist.all_reduce(sums)
        return sums

    with torch.no_grad():
        # the clock sampler starts before the warm-up: nvidia-smi needs a few hundred ms before its first row, longer than
        # a short timed region; rows are filtered by timestamp to the timed region afterwards
        sampler = ClockSampler(local_rank) if rank == 0 else None
        for _ in range(args.warmup):
            sums = step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        l0 = nb._cabi.launch_count()
        t_begin = time.perf_counter()
        for e0, e1 in ev:
            flush.zero_()                 # L2 flush between timed iterations (not timed)
            e0.record()
            sums = step()
            e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        launches = nb._cabi.launch_count() - l0
        clocks = sampler.stop(t_begin, t_end) if sampler else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    tot_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    tot_s = float(tot_ms) * 1e-3
    value = world * n * nt * args.steps / tot_s
    Jc, cs = nb.costs_from_sums(sums, alph, dtype)
    path = nb._cabi.last_path()

    # ---- e2e: public API with HOST buffers (H2D of the step's inputs + D2H of its result inside the timed region)
    xh = x.cpu().pin_memory()
    with torch.no_grad():
        for _ in range(max(1, min(2, args.warmup))):
            nb.ocflow_sums(xh, net, prob, [0.0, 1.0], nt, "rk4", alph)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sh = nb.ocflow_sums(xh, net, prob, [0.0, 1.0], nt, "rk4", alph)     # noc_ocflow_host: H2D + rollout + D2H + sync
            if dist is not None:
                sd_ = sh.to(device)
                dist.all_reduce(sd_)
                sh = sd_.cpu()
        torch.cuda.synchronize()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = world * n * nt * args.steps / float(e2e_s)
    e2e_Jc = float(nb.costs_from_sums(sh, alph, dtype)[0])

    if rank == 0:
        pk = __import__("ctypes").c_double(0.0)
        nb._cabi.check(lib.noc_measure_fma_peak(0 if W["dtype"] == "f32" else 1, pk))
        peak = pk.value
        ach = n * nt * fl / (statistics.mean(step_ms) * 1e-3) / 1e12        # per GPU: the rollout kernel of one launch
        line = {
            "metric": "rk4_sample_steps_per_sec", "value": value, "unit": "sample-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(tot_ms) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": W["dtype"], "data": "synthetic",
            "config": {"workload": "%s %s, nt=%d, %d samples per GPU, x ~ xInit + var0*N(0,I) seed 1234+rank, eval mode"
                                   % (args.workload, "pretrained checkpoint" if W["ckpt"] else "random-init swarm50-shape Phi", nt, n),
                       "samples_per_gpu": n, "nt": nt, "d": d, "m": meta["m"], "nTh": meta["nTh"], "parallelism": "dp%d (row shards, one all-reduce of 8 doubles)" % world,
                       "l2": "flushed between timed steps (256 MiB write, untimed)", "flops_per_sample_step": fl,
                       "Jc": float(Jc), "Jc_e2e": e2e_Jc},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "sample-steps/s", "h2d_bytes_per_step": int(xh.numel() * xh.element_size()),
                    "d2h_bytes_per_step": 64},
            "gpu_launches": int(launches),
            "roofline": roofline(args.workload, W, d, meta, n, nt, fl, statistics.mean(step_ms) * 1e-3, peak, path),
        }
        if world == 1 and not args.no_cpu_baseline:
            lat = latency_mode(args, W, dtype, threads)     # the other half of BASELINE.json's metric
            line["batch1_latency"] = {"value": lat["value"], "unit": "ms", "device_ms": lat["device_ms_median"],
                                      "ms_per_rk4_step": lat["ms_per_rk4_step"], "protocol": lat["config"]["workload"],
                                      "cpu_port_1thread_ms": lat["cpu_baseline"]["value"]}
            r, dt = cpu_rate(args.workload, W["n_cpu"], nt, dtype, threads)
            line["cpu_baseline"] = {"value": r, "unit": "sample-steps/s", "cores": threads, "kind": "port",
                                    "sample": "%d of the %d samples, nt=%d, one call, %.1f s; torch CPU %s oracle port (the Python reference cannot travel)"
                                              % (W["n_cpu"], n, nt, dt, torch.__version__)}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
