"""GPU parity of the streamed tensor-core rollout for the wide swarm network (neuraloc_b200/csrc/noc_ts_rollout.cuh: CTA pairs,
TMA-streamed weights, fp16-split operands) against the unmodified reference's outputs (tests/golden) and the CPU oracle, at
north_star's tolerances: 1e-5 relative per-step state, 1e-4 relative on the final cost terms (G included, no absolute floor)."""
import os

import numpy as np
import pytest
import torch

from helpers import DT, QW, check_costs, load_cases, mean_vec, oracle_setup, product_setup, rel_state_err
from test_gpu_parity import _compare, _three_modes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import neuraloc_b200
    neuraloc_b200._cabi.lib()
    return neuraloc_b200


@pytest.fixture(autouse=True)
def tc_path(monkeypatch):
    monkeypatch.setenv("NOC_FORCE_PATH", "tc")


def test_ts_rollout_golden(nb):
    """xInit (batch 1) and the 6 adversarial rows (agents inside the blocks / within interaction range), three return modes."""
    c = load_cases("swarm50")
    net, prob, xinit, meta = product_setup("swarm50", DT["f32"])
    d = xinit.shape[1]
    got = _three_modes(nb, xinit, net, prob, [0.0, 1.0], int(c["nt"]), "rk4", meta["alph"])
    assert nb._cabi.last_path() == "tensor"
    _compare("f32", d, got, (c["xinit_mean_f32"], None, c["xinit_z_f32"], c["xinit_ctrl_f32"]), "ts swarm50 xInit",
             truth=(c["xinit_mean_f64"], None))
    xb = torch.from_numpy(c["xb"]).float().cuda()
    got = _three_modes(nb, xb, net, prob, [0.0, 1.0], int(c["nt_batch"]), "rk4", meta["alph"])
    # nt = 20 on the adversarial rows is ill-conditioned in fp32 (the reference's own fp32 run is far from its fp64 run):
    # gate against the fp64 trajectories at 1e-5 or twice the reference's own fp32 distance
    ref_err = rel_state_err(c["b_z_f32"], c["b_z_f64"], d)
    _compare("f32", d, got, (c["b_mean_f64"], c["b_nomean_f64"], c["b_z_f64"], c["b_ctrl_f64"]), "ts swarm50 batch",
             state_tol=max(1e-5, 2 * ref_err))


def test_ts_rk1_and_unknown_stepper_golden(nb):
    c = load_cases("swarm50")
    net, prob, xinit, meta = product_setup("swarm50", DT["f32"])
    d = xinit.shape[1]
    xb = torch.from_numpy(c["xb"]).float().cuda()
    got = _three_modes(nb, xb[:4], net, prob, [0.0, 1.0], 8, "rk1", meta["alph"])
    ref_err = rel_state_err(c["rk1_z_f32"], c["rk1_z_f64"], d)
    _compare("f32", d, got, (c["rk1_mean_f64"], None, c["rk1_z_f64"], c["rk1_ctrl_f64"]), "ts rk1", state_tol=max(1e-5, 2 * ref_err))
    got = _three_modes(nb, xb[:2], net, prob, [0.0, 1.0], 3, "none", meta["alph"])
    _compare("f32", d, got, (c["nostep_mean_f32"], None, c["nostep_z_f32"], c["nostep_ctrl_f32"]), "ts no stepper")


@pytest.mark.parametrize("n,nt,tspan", [(300, 40, (0.0, 1.0)), (129, 40, (0.1, 1.0)), (64, 80, (0.0, 1.0))])
def test_ts_vs_oracle_ragged(nb, n, nt, tspan):
    """Ragged batches (a partly filled CTA, a CTA with no valid sample, several tiles), tspan != [0,1]: all three return modes
    against the fp32 and fp64 oracle."""
    from oracle import ocflow_oracle as orc
    net, prob, xinit, meta = product_setup("swarm50", torch.float32)
    P32, D32, _, _ = oracle_setup("swarm50", torch.float32)
    P64, D64, _, _ = oracle_setup("swarm50", torch.float64)
    d = xinit.shape[1]
    g = torch.Generator().manual_seed(100 + n)
    x = xinit.cpu() + 0.1 * torch.randn(n, d, generator=g)
    torch.set_num_threads(os.cpu_count() or 1)
    ts = list(tspan)
    with torch.no_grad():
        mean, nomean, zf, cf = _three_modes(nb, x.cuda(), net, prob, ts, nt, "rk4", meta["alph"])
        assert nb._cabi.last_path() == "tensor"
        z32, c32 = orc.ocflow(x, P32, D32, ts, nt, "rk4", meta["alph"], intermediates=True)
        z64, c64 = orc.ocflow(x.double(), P64, D64, ts, nt, "rk4", meta["alph"], intermediates=True)
        m64 = mean_vec(orc.ocflow(x.double(), P64, D64, ts, nt, "rk4", meta["alph"]))
        J64, cs64 = orc.ocflow(x.double(), P64, D64, ts, nt, "rk4", meta["alph"], noMean=True)
    ref = rel_state_err(z32.numpy(), z64.numpy(), d)
    e64 = rel_state_err(zf, z64.numpy(), d)
    assert e64 <= max(1e-5, 2 * ref), "state vs fp64 oracle %.2e (fp32 oracle itself %.2e)" % (e64, ref)
    assert not cf[:, :, 0].any()
    assert np.abs(cf - c64.numpy()).max() <= 2e-4 * max(1.0, float(c64.abs().max()))
    check_costs(mean[:6], m64[:6], 1e-4, 0.0, "ts mean costs (Jc, L, G, HJt, HJfin, HJgrad) vs fp64 oracle, n=%d" % n)
    check_costs(mean[6:], m64[6:], 1e-4, 2e-4, "ts mean Q, W", floor_mask=[True, True])
    n64 = torch.cat([J64] + list(cs64), 1).numpy()
    sc = np.maximum(np.abs(n64).max(axis=0, keepdims=True), 1e-30)
    perr = (np.abs(nomean - n64) / sc).max(axis=0)
    assert (perr[[0, 1, 3, 4]] <= 1e-4).all() and perr[2] <= 3e-3 and perr[5] <= 3e-3, "per-sample costs vs fp64 oracle: %s" % perr
    check_costs(nomean.mean(axis=0)[:6], mean[:6], 1e-5, 0.0, "noMean vs mean")
    # cost integrals along the trajectory
    scale = np.maximum(np.abs(z64.numpy()[:, d:, :]).max(), 1.0)
    assert np.abs(zf[:, d:, :] - z64.numpy()[:, d:, :]).max() <= 1e-4 * scale


def test_ts_many_tiles_subsampled_oracle(nb):
    """A benchmark-size batch (more tiles than CTA pairs, ragged tail): 256 rows sub-sampled from it are checked per sample
    against the fp64 oracle; sums add over row shards; the result does not depend on which tile a sample lands in."""
    from oracle import ocflow_oracle as orc
    net, prob, xinit, meta = product_setup("swarm50", torch.float32)
    P64, D64, _, _ = oracle_setup("swarm50", torch.float64)
    d = xinit.shape[1]
    n, nt = 128 * 74 * 2 + 77, 80
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = xinit + 0.1 * torch.randn(n, d, generator=g, device="cuda")
    with torch.no_grad():
        J, cs = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
        assert nb._cabi.last_path() == "tensor"
        s_all = nb.ocflow_sums(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
        lo = nb.ocflow_sums(x[:7777].contiguous(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
        hi = nb.ocflow_sums(x[7777:].contiguous(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
    tab = torch.cat([J] + list(cs), 1).double().cpu().numpy()
    assert float(s_all[7]) == n
    assert torch.allclose(s_all, lo + hi, rtol=1e-9, atol=1e-6)
    check_costs((s_all[:7] / n).cpu().numpy(), tab[:, 1:].mean(axis=0), 1e-6, 1e-7, "mean mode vs mean of noMean", floor_mask=QW[1:])
    idx = torch.linspace(0, n - 1, 256).long()
    xs = x[idx.cuda()].cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        J64, cs64 = orc.ocflow(xs.double(), P64, D64, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
    n64 = torch.cat([J64] + list(cs64), 1).numpy()
    got = tab[idx.numpy()]
    sc = np.maximum(np.abs(n64).max(axis=0, keepdims=True), 1e-30)
    perr = (np.abs(got - n64) / sc).max(axis=0)
    assert (perr[[0, 1, 3, 4]] <= 1e-4).all() and perr[2] <= 3e-3 and perr[5] <= 3e-3, "per-sample costs vs fp64 oracle: %s" % perr
    check_costs(got.mean(axis=0)[:6], n64.mean(axis=0)[:6], 1e-4, 0.0, "means over the 256 sub-sampled rows vs fp64 oracle")


def test_ts_intermediates_staged_equals_direct(nb, monkeypatch):
    """intermediates=True: the tile-major staging buffer + transpose (coalesced) gives bit-identical zFull / ctrlFull to the
    direct strided writes (the fallback when the staging buffer cannot be allocated); ragged batch, several tiles per CTA pair."""
    net, prob, xinit, meta = product_setup("swarm50", torch.float32)
    d = xinit.shape[1]
    g = torch.Generator().manual_seed(8)
    x = (xinit.cpu() + 0.1 * torch.randn(128 * 74 + 301, d, generator=g)).cuda()
    with torch.no_grad():
        za, ca = nb.OCflow(x, net, prob, [0.0, 1.0], 3, "rk4", meta["alph"], intermediates=True)
        monkeypatch.setenv("NOC_TS_NOSTAGE", "1")
        zb, cb = nb.OCflow(x, net, prob, [0.0, 1.0], 3, "rk4", meta["alph"], intermediates=True)
    assert nb._cabi.last_path() == "tensor"
    assert za.shape == (x.shape[0], d + 4, 4) and ca.shape == (x.shape[0], d, 4)
    assert torch.equal(za, zb) and torch.equal(ca, cb)
    assert torch.equal(za[:, :d, 0], x) and not ca[:, :, 0].any() and not za[:, d:, 0].any()


def test_ts_train_mode_and_narrower_net(nb):
    """prob.train() (boxes inflated by r, Gaussian + 999 terrain, the 3.2 r interaction cut-off: SwarmTraj.py:101-119,140-156) on
    the adversarial rows, and a random-init network narrower than the kernel's 512 columns (m = 320: zero-padded units), both
    against the fp64 oracle."""
    import dataclasses
    from oracle import ocflow_oracle as orc
    c = load_cases("swarm50")
    net, prob, xinit, meta = product_setup("swarm50", torch.float32)
    P64, D64, _, _ = oracle_setup("swarm50", torch.float64)
    d = xinit.shape[1]
    g = torch.Generator().manual_seed(31)
    xb = torch.cat((torch.from_numpy(c["xb"]).float(), xinit.cpu() + 0.1 * torch.randn(130, d, generator=g)), 0)
    nt = 40
    prob.train()
    Dtr = dataclasses.replace(D64, training=True)
    with torch.no_grad():
        zf, cf = nb.OCflow(xb.cuda(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
        assert nb._cabi.last_path() == "tensor"
        Jn, cn = nb.OCflow(xb.cuda(), net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
        zr, cr = orc.ocflow(xb.double(), P64, Dtr, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
        Jr, csr = orc.ocflow(xb.double(), P64, Dtr, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
    prob.eval()
    assert rel_state_err(zf.cpu().numpy(), zr.numpy(), d) <= 1e-5
    got, ref = torch.cat([Jn] + list(cn), 1).double().cpu().numpy(), torch.cat([Jr] + list(csr), 1).numpy()
    sc = np.maximum(np.abs(ref).max(axis=0, keepdims=True), 1e-30)
    perr = (np.abs(got - ref) / sc).max(axis=0)
    assert (perr[[0, 1, 3, 4, 6, 7]] <= 2e-4).all(), "train-mode per-sample costs vs fp64 oracle: %s" % perr
    assert ref[:, 6].max() > 100.0                              # the +999 terrain of agents inside the inflated boxes is exercised
    # a narrower random-init net through the same kernel
    torch.manual_seed(3)
    net2 = nb.Phi(nTh=2, m=320, d=d, alph=meta["alph"])
    P2 = orc.params_from_state_dict({k: v.detach().clone() for k, v in net2.state_dict().items()}, torch.float64)
    net2 = net2.float().cuda()
    x = xinit.cpu() + 0.1 * torch.randn(200, d, generator=g)
    with torch.no_grad():
        z2, _ = nb.OCflow(x.cuda(), net2, prob, [0.0, 1.0], 20, "rk4", meta["alph"], intermediates=True)
        assert nb._cabi.last_path() == "tensor"
        m2 = mean_vec(nb.OCflow(x.cuda(), net2, prob, [0.0, 1.0], 20, "rk4", meta["alph"]))
        zr2, _ = orc.ocflow(x.double(), P2, D64, [0.0, 1.0], 20, "rk4", meta["alph"], intermediates=True)
        mr2 = mean_vec(orc.ocflow(x.double(), P2, D64, [0.0, 1.0], 20, "rk4", meta["alph"]))
    assert rel_state_err(z2.cpu().numpy(), zr2.numpy(), d) <= 1e-5
    check_costs(m2[:6], mr2[:6], 1e-4, 0.0, "m = 320 random-init net through the streamed kernel")
