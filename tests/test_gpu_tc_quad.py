"""GPU parity of the tensor-core rollout kernel (neuraloc_b200/csrc/noc_tc_quad.cu; tcgen05 MMAs on 3-way bf16 splits of
the fp32 operands) against the unmodified reference's outputs for singlequad, at the same tolerances as the FMA kernels
(1e-5 relative per-step state, 1e-4 relative final cost terms), plus agreement with the FMA tile kernel sample by sample."""
import numpy as np
import pytest
import torch

from helpers import DT, check_costs, load_cases, mean_vec, product_setup, rel_state_err
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


def test_tc_rollout_golden(nb):
    c = load_cases("singlequad")
    net, prob, xinit, meta = product_setup("singlequad", DT["f32"])
    nt = int(c["nt"])
    got = _three_modes(nb, xinit, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"])
    _compare("f32", 12, got, (c["xinit_mean_f32"], None, c["xinit_z_f32"], c["xinit_ctrl_f32"]), "tc singlequad xInit")
    xb = torch.from_numpy(c["xb"]).float().cuda()
    got = _three_modes(nb, xb, net, prob, [0.0, 1.0], int(c["nt_batch"]), "rk4", meta["alph"])
    _compare("f32", 12, got, (c["b_mean_f32"], c["b_nomean_f32"], c["b_z_f32"], c["b_ctrl_f32"]), "tc singlequad batch")


def test_tc_rk1_and_unknown_stepper_golden(nb):
    c = load_cases("singlequad")
    net, prob, xinit, meta = product_setup("singlequad", DT["f32"])
    xb = torch.from_numpy(c["xb"]).float().cuda()
    got = _three_modes(nb, xb[:4], net, prob, [0.0, 1.0], 8, "rk1", meta["alph"])
    _compare("f32", 12, got, (c["rk1_mean_f32"], None, c["rk1_z_f32"], c["rk1_ctrl_f32"]), "tc rk1")
    got = _three_modes(nb, xb[:2], net, prob, [0.0, 1.0], 3, "none", meta["alph"])
    _compare("f32", 12, got, (c["nostep_mean_f32"], None, c["nostep_z_f32"], c["nostep_ctrl_f32"]), "tc no stepper")


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 40_000])
def test_tc_matches_fma_tile_kernel(nb, n, monkeypatch):
    """Ragged tile counts (one CTA runs several tiles at n = 40000): per-sample costs and trajectories agree with the
    FMA tile kernel to fp32 rounding, and the mean is the mean of the per-sample table."""
    net, prob, xinit, meta = product_setup("singlequad", torch.float32)
    g = torch.Generator().manual_seed(n)
    x = xinit.cpu() + 0.4 * torch.randn(n, 12, generator=g)
    x[:, 3:] = 0 if n % 2 else x[:, 3:] * 0.2
    x = x.cuda()
    nt = 20
    with torch.no_grad():
        Jt, ct = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
        mt = mean_vec(nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"]))
        if n <= 1000:
            zt, ut = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
        monkeypatch.setenv("NOC_FORCE_PATH", "tile")
        Jf, cf = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], noMean=True)
        if n <= 1000:
            zf, uf = nb.OCflow(x, net, prob, [0.0, 1.0], nt, "rk4", meta["alph"], intermediates=True)
    tt, tf = torch.cat([Jt] + list(ct), 1).double().cpu().numpy(), torch.cat([Jf] + list(cf), 1).double().cpu().numpy()
    sc = np.maximum(np.abs(tf).max(axis=0, keepdims=True), 1.0)
    assert (np.abs(tt - tf) / sc).max() <= 1e-4, "per-sample costs tc vs fma"
    check_costs(mt, tt.mean(axis=0), 1e-6, 1e-7, "tc mean vs mean of tc noMean")
    if n <= 1000:
        assert rel_state_err(zt.cpu().numpy(), zf.cpu().numpy(), 12) <= 5e-6
        assert (ut - uf).abs().max() <= 1e-4 * max(1.0, float(uf.abs().max()))


def test_tc_is_selected_only_where_it_applies(nb, monkeypatch):
    """NOC_FORCE_PATH=tc on a shape the kernel is not written for falls back to the normal choice (still CUDA)."""
    net, prob, xinit, meta = product_setup("softcorridor", torch.float32)
    before = nb._cabi.lib().noc_launch_count()
    with torch.no_grad():
        Jc, cs = nb.OCflow(xinit, net, prob, [0.0, 1.0], 10, "rk4", meta["alph"])
    assert nb._cabi.lib().noc_launch_count() > before and np.isfinite(float(Jc))
