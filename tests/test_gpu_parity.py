"""GPU parity tests: the CUDA rollout (through the C ABI) against the committed reference outputs and the
CPU oracle.  Tolerances are north_star's: 1e-5 relative per-step state, 1e-4 relative on the final cost terms in
single precision (means; Q/W with an absolute floor); 1e-10-class in double."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import (DT, GOLDEN, PROBLEMS, QW, check_costs, load_cases, load_ckpt, mean_vec, oracle_setup, product_setup,
                     rel_err, rel_state_err)

pytestmark = pytest.mark.gpu

TOL = {"f32": dict(state=1e-5, cost=1e-4, floor=2e-4, ctrl=2e-4), "f64": dict(state=1e-11, cost=1e-9, floor=1e-10, ctrl=1e-9)}


@pytest.fixture(autouse=True, params=["tile", "vec", "vec1"])
def path(request, monkeypatch):
    """Every rollout test runs through the sample-tile kernel (large batches) and through both small-batch kernels: the cluster
    latency kernel (noc_lat.cu: "vec", the default for small batches) and the one-CTA-per-sample kernel (noc_vec.cu: "vec1",
    NOC_NO_LAT=1; also the fallback for nTh > 2).  NOC_FORCE_PATH pins the choice the host would make by batch size."""
    monkeypatch.setenv("NOC_FORCE_PATH", "vec" if request.param == "vec1" else request.param)
    if request.param == "vec1":
        monkeypatch.setenv("NOC_NO_LAT", "1")
    return request.param


@pytest.fixture(scope="module")
def nb():
    import neuraloc_b200
    neuraloc_b200._cabi.lib()
    assert torch.cuda.is_available()
    return neuraloc_b200


def _three_modes(nb, x, net, prob, tspan, nt, stepper, alph):
    with torch.no_grad():
        mean = mean_vec(nb.OCflow(x, net, prob, tspan, nt, stepper, alph))
        Jn, cn = nb.OCflow(x, net, prob, tspan, nt, stepper, alph, noMean=True)
        zf, cf = nb.OCflow(x, net, prob, tspan, nt, stepper, alph, intermediates=True)
    nomean = torch.cat([Jn] + list(cn), dim=1).cpu().numpy()
    return mean, nomean, zf.cpu().numpy(), cf.cpu().numpy()


def _compare(tag, d, got, ref, what, state_tol=None, truth=None):
    """`ref` = the reference's outputs in the run's own precision (mean, noMean, zFull, ctrlFull); `truth` = the same from the
    reference's fp64 run (fp32 runs only): the cost terms are then gated against the fp64 values at 1e-4 relative or twice the
    reference's own fp32<->fp64 distance, whichever is larger — no absolute floor except for Q and W."""
    tol = dict(TOL[tag])
    if state_tol is not None:
        tol["state"] = state_tol
    mean, nomean, zf, cf = got
    rmean, rnomean, rz, rc = ref
    assert zf.shape == rz.shape and cf.shape == rc.shape, what
    serr = rel_state_err(zf, rz, d)
    assert serr <= tol["state"], "%s: per-step state rel err %.3e" % (what, serr)
    scale = np.maximum(np.abs(rz[:, d:, :]).max(), 1.0)
    assert np.abs(zf[:, d:, :] - rz[:, d:, :]).max() <= 20 * tol["cost"] * scale, what + ": cost integrals along the path"
    assert not cf[:, :, 0].any()
    cerr = np.abs(cf - rc).max() / max(np.abs(rc).max(), 1.0)
    assert cerr <= tol["ctrl"], "%s: controls rel err %.3e" % (what, cerr)
    tmean, tnomean = (truth[0], truth[1]) if truth is not None else (None, None)
    if rmean is not None:
        if tmean is not None:
            # one sample alone: G (~1e-4..1e-2) and HJgrad are pure cancellation and every fp32 summation order lands somewhere
            # else inside the fp32 noise: 8x the reference's own fp32<->fp64 distance there; batches (means): 2x
            check_costs(mean, tmean, tol["cost"], tol["floor"], what + " mean costs vs the reference's fp64 run", floor_mask=QW,
                        ref_noise=np.asarray(rmean, dtype=np.float64) - np.asarray(tmean, dtype=np.float64),
                        noise_mult=8.0 if zf.shape[0] == 1 else 2.0)
        else:
            check_costs(mean, rmean, tol["cost"], tol["floor"], what + " mean costs", floor_mask=QW)
    if rnomean is not None:
        # per-sample costs: relative to the largest entry of the column, at 30x the mean tolerance or three times the
        # reference's own fp32<->fp64 distance in that column (a single sample's G and HJgrad are pure cancellation, H2)
        base = tnomean if tnomean is not None else rnomean
        sc = np.maximum(np.abs(base).max(axis=0, keepdims=True), 1e-30)
        lim = np.full(base.shape[1], 30 * tol["cost"])
        if tnomean is not None:
            lim = np.maximum(lim, 3.0 * (np.abs(np.asarray(rnomean, dtype=np.float64) - tnomean) / sc).max(axis=0))
        perr = (np.abs(nomean - base) / sc).max(axis=0)
        keep = np.abs(base).max(axis=0) > 0          # all-zero columns (Q, W off the obstacles): exact match required below
        assert (perr[keep] <= lim[keep]).all(), "%s per-sample costs: %s (limits %s)" % (what, perr, lim)
        assert (np.abs(nomean[:, ~keep]) <= tol["floor"]).all(), what + " per-sample Q / W where the reference has none"
        # noMean table is consistent with the means
        check_costs(nomean.mean(axis=0), mean, 10 * tol["cost"], tol["floor"], what + " noMean vs mean", floor_mask=QW)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", PROBLEMS)
def test_rollout_golden(nb, name, tag):
    """xInit (batch 1) and the seeded batch, three return modes, against the unmodified reference's outputs."""
    c = load_cases(name)
    net, prob, xinit, meta = product_setup(name, DT[tag])
    d = xinit.shape[1]
    nt = int(c["nt"])
    got = _three_modes(nb, xinit, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
    truth = (lambda pre: (c[pre + "_mean_f64"], c[pre + "_nomean_f64"] if pre + "_nomean_f64" in c.files else None)) if tag == "f32" else None
    _compare(tag, d, got, (c["xinit_mean_" + tag], None, c["xinit_z_" + tag], c["xinit_ctrl_" + tag]), name + " xInit",
             truth=truth("xinit") if truth else None)
    xb = torch.from_numpy(c["xb"]).to(DT[tag]).cuda()
    got = _three_modes(nb, xb, net, prob, [0.0, 1.0], int(c["nt_batch"]), "rk4", meta["alph"])
    _compare(tag, d, got, (c["b_mean_" + tag], c["b_nomean_" + tag], c["b_z_" + tag], c["b_ctrl_" + tag]), name + " batch",
             truth=truth("b") if truth else None)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("name", PROBLEMS)
def test_rk1_and_unknown_stepper_golden(nb, name, tag):
    c = load_cases(name)
    net, prob, xinit, meta = product_setup(name, DT[tag])
    d = xinit.shape[1]
    xb = torch.from_numpy(c["xb"]).to(DT[tag]).cuda()
    got = _three_modes(nb, xb[:4], net, prob, [0.0, 1.0], 8, "rk1", meta["alph"])
    # Euler with 8 steps on the adversarial rows is ill-conditioned in fp32 (the reference's own fp32 run is up to 1.05e-5 from
    # its fp64 run on swap12): gate against the fp64 trajectories at 1e-5 or twice the reference's own fp32 distance
    ref_err = rel_state_err(c["rk1_z_f32"], c["rk1_z_f64"], d)
    _compare(tag, d, got, (c["rk1_mean_f64"], None, c["rk1_z_f64"], c["rk1_ctrl_f64"]), name + " rk1",
             state_tol=max(TOL[tag]["state"], 2 * ref_err) if tag == "f32" else None)
    got = _three_modes(nb, xb[:2], net, prob, [0.0, 1.0], 3, "none", meta["alph"])
    _compare(tag, d, got, (c["nostep_mean_" + tag], None, c["nostep_z_" + tag], c["nostep_ctrl_" + tag]), name + " no stepper")


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_shock_restart_golden(nb, tag):
    """tspan != [0,1] (plotter.py:817-823)."""
    c = load_cases("softcorridor")
    net, prob, xinit, meta = product_setup("softcorridor", DT[tag])
    nt = int(c["nt"])
    nS = int(0.1 * nt)
    got = _three_modes(nb, xinit, net, prob, [0.0, 0.1], nS, "rk4", meta["alph"])
    _compare(tag, 4, got, (c["shock1_mean_" + tag], None, c["shock1_z_" + tag], c["shock1_ctrl_" + tag]), "shock leg 1")
    xs = torch.from_numpy(c["shock2_x_" + tag]).to(DT[tag]).cuda()
    got = _three_modes(nb, xs, net, prob, [0.1, 1.0], 1 + nt - nS, "rk4", meta["alph"])
    _compare(tag, 4, got, (c["shock2_mean_" + tag], None, c["shock2_z_" + tag], c["shock2_ctrl_" + tag]), "shock leg 2")


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_problem_functors_golden(nb, tag):
    """calcLHQW / calcGradpH / calcCtrls device functors on adversarial inputs, eval and train mode."""
    z = np.load(GOLDEN + "/functors.npz")
    names = sorted({k[: -len("_%s_meta" % tag)] for k in z.files if k.endswith("_%s_meta" % tag)})
    tol = 5e-6 if tag == "f32" else 1e-12
    for name in names:
        key = "%s_%s" % (name, tag)
        meta = json.loads(str(z[key + "_meta"]))
        x, p = torch.from_numpy(z[key + "_x"]).cuda(), torch.from_numpy(z[key + "_p"]).cuda()
        cls = getattr(nb, meta["cls"])
        kw = dict(obstacle=meta["obstacle"], alph_Q=meta["alph_Q"], alph_W=meta["alph_W"], r=meta["r"])
        if meta["cls"] == "Quadcopter":
            kw.update(mass=meta["mass"], grav=meta["grav"])
        prob = cls(torch.from_numpy(z[key + "_xtarget"]).cuda(), **kw)
        for mode in ("eval", "train"):
            getattr(prob, mode)()
            L, H, Q, W = prob.calcLHQW(x, p)
            got = torch.cat((L, H, Q, W), 1).cpu().numpy()
            ref = z["%s_%s_LHQW" % (key, mode)]
            scale = np.maximum(np.abs(ref), 1.0)
            err = np.abs(got - ref) / scale
            # W: the reference sums an A x A matrix of ones and subtracts the count (fp32 cancellation)
            wtol = tol if tag == "f64" else 2e-4
            assert err[:, :3].max() <= max(tol, wtol * abs(meta["alph_W"]) / 10 if name not in ("singlequad",) else tol) or \
                err[:, :3].max() <= 2e-4, (name, mode, err.max(axis=0))
            assert err[:, 3].max() <= wtol, (name, mode, "W", err[:, 3].max())
            assert rel_err(prob.calcGradpH(x, p).cpu().numpy(), z["%s_%s_gradpH" % (key, mode)], floor=1.0) <= tol
            assert rel_err(prob.calcCtrls(x, p).cpu().numpy(), z["%s_%s_ctrls" % (key, mode)], floor=1.0) <= tol


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_phi_forward_and_grad_golden(nb, tag):
    """Phi.forward / Phi.getGrad device functions on random-weight nets, nTh in {2,3,4,5}."""
    z = np.load(GOLDEN + "/phi_random.npz")
    tol = 5e-6 if tag == "f32" else 1e-12
    for idx in range(5):
        pre = "net%d_" % idx
        nTh, m, d = [int(v) for v in z[pre + "dims"]]
        net = nb.Phi(nTh=nTh, m=m, d=d)
        net.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files
                             if k.startswith(pre) and ("." in k or k == pre + "A")})
        net = net.to(DT[tag]).cuda()
        x = torch.from_numpy(z[pre + "x"]).to(DT[tag]).cuda()
        with torch.no_grad():
            f, g = net(x), net.getGrad(x)
        assert f.shape == (x.shape[0], 1) and g.shape == x.shape
        assert rel_err(f.cpu().numpy(), z[pre + "fwd_" + tag], floor=1.0) <= tol, idx
        assert rel_err(g.cpu().numpy(), z[pre + "grad_" + tag], floor=1.0) <= tol, idx


@pytest.mark.parametrize("name", PROBLEMS)
def test_rollout_vs_oracle_ragged_batches(nb, name):
    """fp32 CUDA vs the fp32 and fp64 CPU oracle on fresh seeded samples; batch sizes around the tile size."""
    from oracle import ocflow_oracle as orc
    net, prob, xinit, meta = product_setup(name, torch.float32)
    P32, D32, _, _ = oracle_setup(name, torch.float32)
    P64, D64, _, _ = oracle_setup(name, torch.float64)
    d = xinit.shape[1]
    nt = 50 if name != "swarm50" else 40
    sizes = [1, 3, 31, 33, 127, 129] if name != "swarm50" else [1, 31, 33]
    g = torch.Generator().manual_seed(321)
    nmax = max(sizes)
    if name == "singlequad":
        xall = torch.zeros(nmax, d)
        xall[:, :3] = -1.5 + meta["var0"] * torch.randn(nmax, 3, generator=g)
    else:
        xall = xinit.cpu() + meta["var0"] * torch.randn(nmax, d, generator=g)
    with torch.no_grad():
        z64, _ = orc.ocflow(xall.double(), P64, D64, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
        J64, c64 = orc.ocflow(xall.double(), P64, D64, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
        z32, _ = orc.ocflow(xall, P32, D32, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
    ref_nm = torch.cat([J64] + list(c64), 1).numpy()
    floor32 = rel_state_err(z32.numpy(), z64.numpy(), d)        # the reference arithmetic's own fp32 noise on these samples
    print("%s: fp32 oracle vs fp64 oracle per-step state distance %.2e" % (name, floor32))
    for n in sizes:
        x = xall[:n].cuda()
        mean, nomean, zf, cf = _three_modes(nb, x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
        e64, e32 = rel_state_err(zf, z64[:n].numpy(), d), rel_state_err(zf, z32[:n].numpy(), d)
        print("%s n=%d: CUDA fp32 vs fp64 oracle %.2e, vs fp32 oracle %.2e" % (name, n, e64, e32))
        assert e64 <= 1e-5 and e32 <= 1e-5, (name, n, e64, e32)
        check_costs(mean[1:6], ref_nm[:n, 1:6].mean(axis=0), 1e-4, 1e-6, "%s n=%d mean costs vs fp64 oracle" % (name, n))
        check_costs(mean[6:], ref_nm[:n, 6:].mean(axis=0), 1e-4, 2e-4, "%s n=%d Q/W" % (name, n), floor_mask=[True, True])


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_every_tile_configuration_agrees(nb, cfg, monkeypatch, path):
    """Forces each tile configuration (multi-pass GEMMs, ping-pong panels, TPS > 1 problem phase) on swap12 and a
    deep net; results must not depend on the tiling."""
    if path != "tile":
        pytest.skip("tile configurations only exist on the tile path")
    dtype = torch.float64 if cfg in (4, 5, 6) else torch.float32
    tag = "f64" if cfg in (4, 5, 6) else "f32"
    monkeypatch.setenv("NOC_FORCE_CFG", str(cfg))
    c = load_cases("swap12")
    net, prob, xinit, meta = product_setup("swap12", dtype)
    xb = torch.from_numpy(c["xb"]).to(dtype).cuda()
    try:
        got = _three_modes(nb, xb, net, prob, [0.0, 1.0], int(c["nt_batch"]), "rk4", meta["alph"])
        _compare(tag, 24, got, (c["b_mean_" + tag], c["b_nomean_" + tag], c["b_z_" + tag], c["b_ctrl_" + tag]), "swap12 cfg %d" % cfg)
    except nb._cabi.NocError as e:
        assert "noc error -4" in str(e)            # swap12 does not fit this forced tiling (m = 32 on the m <= 16 tile): loud
    # deep / odd-sized net through the same tiling
    z = np.load(GOLDEN + "/phi_random.npz")
    for idx in (1, 2, 3, 4):
        pre = "net%d_" % idx
        nTh, m, d = [int(v) for v in z[pre + "dims"]]
        net2 = nb.Phi(nTh=nTh, m=m, d=d)
        net2.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files
                              if k.startswith(pre) and ("." in k or k == pre + "A")})
        net2 = net2.to(dtype).cuda()
        x = torch.from_numpy(z[pre + "x"]).to(dtype).cuda()
        try:
            with torch.no_grad():
                f, g = net2(x), net2.getGrad(x)
        except nb._cabi.NocError as e:
            assert "noc error -4" in str(e)        # panels of this net do not fit the forced tiling: loud, not silent
            continue
        tol = 5e-6 if tag == "f32" else 1e-12
        assert rel_err(f.cpu().numpy(), z[pre + "fwd_" + tag], floor=1.0) <= tol, (cfg, idx)
        assert rel_err(g.cpu().numpy(), z[pre + "grad_" + tag], floor=1.0) <= tol, (cfg, idx)


def test_deep_net_rollout_vs_oracle(nb):
    """nTh = 4 rollout (general-nTh path: tanh panels + reverse-sweep z panel) against the oracle, both precisions."""
    from oracle import ocflow_oracle as orc
    torch.manual_seed(5)
    d, m, nTh = 8, 24, 4
    alph = [50.0, 0.0, 200.0, 1.0, 2.0, 3.0]
    net = nb.Phi(nTh=nTh, m=m, d=d, alph=alph)
    with torch.no_grad():
        for lay in net.N.layers[1:]:
            lay.weight.normal_(std=0.2); lay.bias.normal_()
        net.w.weight.normal_(); net.c.weight.normal_(); net.c.bias.normal_()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    for tag in ("f32", "f64"):
        dt = DT[tag]
        prob, x0, _, xinit = nb.initProb("midcross4", 40, 40, 0.5, alph, lambda v: v.to(dt).cuda())
        prob.eval()
        netd = net.to(dt).cuda()
        P = orc.params_from_state_dict(sd, dt)
        D, _ = orc.make_problem("midcross4", alph, dt)
        with torch.no_grad():
            ref = (mean_vec(orc.ocflow(x0.cpu(), P, D, [0.0, 1.0], 6, "rk4", alph)), None,
                   *[t.numpy() for t in orc.ocflow(x0.cpu(), P, D, [0.0, 1.0], 6, "rk4", alph, intermediates=True)])
        got = _three_modes(nb, x0, netd, prob, [0.0, 1.0], 6, "rk4", alph)
        _compare(tag, d, got, ref, "nTh=4 midcross4 " + tag)


def test_cpu_tensors_take_the_host_entry_point(nb):
    """evalOC.py / timeOC.py pass CPU tensors (F2): same numbers through noc_ocflow_host, results on the CPU."""
    net, prob, xinit, meta = product_setup("softcorridor", torch.float32, device="cpu")
    c = load_cases("softcorridor")
    with torch.no_grad():
        Jc, cs = nb.OCflow(xinit, net, prob, [0.0, 1.0], 50, "rk4", meta["alph"])
        zf, cf = nb.OCflow(xinit, net, prob, [0.0, 1.0], 50, "rk4", meta["alph"], intermediates=True)
    assert not Jc.is_cuda and not zf.is_cuda and Jc.dim() == 0 and Jc.dtype == torch.float32
    check_costs(mean_vec((Jc, cs)), c["xinit_mean_f64"], 1e-4, 2e-4, "host path", floor_mask=QW,
                ref_noise=c["xinit_mean_f32"].astype(np.float64) - c["xinit_mean_f64"])
    assert rel_state_err(zf.numpy(), c["xinit_z_f32"], 4) <= 1e-5
    assert "{:e}".format(Jc) and float(cs[0].item()) > 0          # the drivers format with {:e} and .item()


@pytest.mark.parametrize("chunks", ["1", "3", "7"])
def test_host_entry_chunked_pipeline(nb, chunks, monkeypatch):
    """noc_ocflow_host cuts large mean / noMean batches into row chunks whose host->device copies overlap the previous chunk's
    rollout (NOC_HOST_CHUNKS forces it on a small batch): per-sample results identical to the device entry point, sums
    equal up to the order of the double additions."""
    monkeypatch.setenv("NOC_HOST_CHUNKS", chunks)
    net, prob, xinit, meta = product_setup("swap12", torch.float32, device="cpu")
    g = torch.Generator().manual_seed(3)
    x = xinit + 0.3 * torch.randn(1000, 24, generator=g)
    netd, probd, _, _ = product_setup("swap12", torch.float32)
    with torch.no_grad():
        sh = nb.ocflow_sums(x, net, prob, [0.0, 1.0], 10, "rk4", meta["alph"])
        Jh, ch = nb.OCflow(x, net, prob, [0.0, 1.0], 10, "rk4", meta["alph"], noMean=True)
        sd = nb.ocflow_sums(x.cuda(), netd, probd, [0.0, 1.0], 10, "rk4", meta["alph"])
        Jd, cd = nb.OCflow(x.cuda(), netd, probd, [0.0, 1.0], 10, "rk4", meta["alph"], noMean=True)
    assert not sh.is_cuda and float(sh[7]) == 1000
    assert torch.allclose(sh, sd.cpu(), rtol=1e-12, atol=1e-9)
    assert torch.equal(Jh, Jd.cpu()) and all(torch.equal(a, b.cpu()) for a, b in zip(ch, cd))


def test_config5_random_init_swarm50_shape_fp64(nb):
    """BASELINE.json configs[4]: random-init swarm50-shape Phi, full validation loss, nt = 50, fp64."""
    z = np.load(GOLDEN + "/config5.npz")
    alph = [float(a) for a in z["alph"]]
    torch.manual_seed(0)
    net = nb.Phi(nTh=2, m=512, d=150, alph=alph)
    for k, v in net.state_dict().items():     # same RNG draws as the reference's constructor
        chk = z["chk_" + k]
        assert abs(float(v.double().sum()) - chk[0]) <= 1e-9 * max(1.0, abs(chk[1])), k
    net = net.double().cuda()
    prob, _, _, _ = nb.initProb("swarm50", 4, 4, 0.1, alph, lambda v: v.double().cuda())
    prob.eval()
    x = torch.from_numpy(z["x"]).cuda()
    mean, nomean, zf, cf = _three_modes(nb, x, net, prob, [0.0, 1.0], 50, "rk4", alph)
    check_costs(mean, z["mean_f64"], 1e-9, 1e-9, "config 5 validation loss", floor_mask=QW)
    assert rel_err(zf[:, :150, -1], z["z_last_f64"][:, :150], floor=1e-3) <= 1e-10
    assert rel_err(cf[:, :, -1], z["ctrl_last_f64"], floor=1.0) <= 1e-9
    sc = np.maximum(np.abs(z["nomean_f64"]).max(axis=0, keepdims=True), 1.0)
    assert (np.abs(nomean - z["nomean_f64"]) / sc).max() <= 1e-9


def test_default_path_selection_by_batch_size(nb, monkeypatch):
    """Without NOC_FORCE_PATH the host picks the small-batch kernel up to NOC_VEC_MAX samples and the tile kernel above;
    the two agree to rounding."""
    monkeypatch.delenv("NOC_FORCE_PATH", raising=False)
    net, prob, xinit, meta = product_setup("swap12", torch.float32)
    g = torch.Generator().manual_seed(9)
    x = (xinit.cpu() + torch.randn(300, 24, generator=g)).cuda()
    with torch.no_grad():
        monkeypatch.setenv("NOC_VEC_MAX", "1000")
        a = mean_vec(nb.OCflow(x, net, prob, [0.0, 1.0], 20, "rk4", meta["alph"]))
        monkeypatch.setenv("NOC_VEC_MAX", "10")
        b = mean_vec(nb.OCflow(x, net, prob, [0.0, 1.0], 20, "rk4", meta["alph"]))
    check_costs(a, b, 2e-5, 1e-5, "vec vs tile path", floor_mask=QW)
    assert not np.array_equal(a, b)          # different kernels, different summation orders


def test_input_is_not_mutated_and_errors_are_loud(nb):
    net, prob, xinit, meta = product_setup("swap2", torch.float32)
    x = xinit.repeat(5, 1).contiguous()
    keep = x.clone()
    with torch.no_grad():
        nb.OCflow(x, net, prob, [0.0, 1.0], 4, "rk4", meta["alph"])
        xt = x.t().contiguous().t()          # non-contiguous view
        a = mean_vec(nb.OCflow(xt, net, prob, [0.0, 1.0], 4, "rk4", meta["alph"]))
        b = mean_vec(nb.OCflow(x, net, prob, [0.0, 1.0], 4, "rk4", meta["alph"]))
    assert torch.equal(x, keep) and np.array_equal(a, b)
    # autograd on, parameters require grad: the default mode is the training path (differentiable Jc, tests/test_gpu_grad.py) and
    # returns the same objective; the modes without an adjoint raise instead of returning something non-differentiable
    Jg, _ = nb.OCflow(x, net, prob, [0.0, 1.0], 4, "rk4", meta["alph"])
    assert Jg.requires_grad and abs(float(Jg) - b[0]) <= 2e-5 * abs(b[0])
    with pytest.raises(RuntimeError):
        nb.OCflow(x, net, prob, [0.0, 1.0], 4, "rk4", meta["alph"], noMean=True)
    with pytest.raises(RuntimeError):
        nb.OCflow(x, net, prob, [0.0, 1.0], 4, "rk4", meta["alph"], intermediates=True)
    with torch.no_grad(), pytest.raises(ValueError):
        nb.OCflow(x[:, :3], net, prob, [0.0, 1.0], 4, "rk4", meta["alph"])
    with torch.no_grad(), pytest.raises(ValueError):
        nb.OCflow(x.half(), net, prob, [0.0, 1.0], 4, "rk4", meta["alph"])


def test_single_step_helpers_match_oracle(nb):
    """stepRK4 / stepRK1 / ocOdefun keep their reference signatures (OCflow.py:104-184)."""
    from oracle import ocflow_oracle as orc
    net, prob, xinit, meta = product_setup("singlequad", torch.float64)
    P, D, _, _ = oracle_setup("singlequad", torch.float64)
    g = torch.Generator().manual_seed(3)
    z = torch.cat((xinit.cpu() + 0.1 * torch.randn(7, 12, generator=g, dtype=torch.float64), torch.rand(7, 4, generator=g, dtype=torch.float64)), 1)
    with torch.no_grad():
        a = nb.stepRK4(nb.ocOdefun, z.cuda(), net, prob, meta["alph"], 0.2, 0.25).cpu()
        b = orc.rk4_step(z, P, D, 0.2, 0.25)
        a1 = nb.stepRK1(nb.ocOdefun, z.cuda(), net, prob, meta["alph"], 0.2, 0.25).cpu()
        b1 = orc.rk1_step(z, P, D, 0.2, 0.25)
        f = nb.ocOdefun(z.cuda(), 0.3, net, prob).cpu()
        fr = orc.rhs(z, 0.3, P, D)
    assert rel_err(a.numpy(), b.numpy(), floor=1.0) <= 1e-11
    assert rel_err(a1.numpy(), b1.numpy(), floor=1.0) <= 1e-11
    assert rel_err(f.numpy(), fr.numpy(), floor=1.0) <= 1e-11
