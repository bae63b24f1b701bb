"""On-device rho_0 sampler (noc_sample_rho0): the raw Philox words and the normals against the numpy oracle,
shard reproducibility, and the distribution of the reference's host-side draw (src/initProb.py:107-120,132-140,252-262:
xInit + var0 * randn; the quadcopter perturbs its position only)."""
import numpy as np
import pytest
import torch

from helpers import product_setup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import neuraloc_b200
    neuraloc_b200._cabi.lib()
    return neuraloc_b200


def test_raw_philox_words_match_the_oracle(nb):
    from oracle import philox_oracle as po
    L = nb._cabi.lib()
    for seed, g0, ng in ((0, 0, 1), (0xA4093822299F31D0, 5, 1000), (1234, (1 << 32) - 3, 16)):
        out = torch.empty(4 * ng, dtype=torch.int32, device="cuda")
        nb._cabi.check(L.noc_philox_raw(seed, g0, ng, out.data_ptr(), None))
        got = out.cpu().numpy().view(np.uint32).reshape(ng, 4)
        assert np.array_equal(got, po.raw_groups(seed, g0, ng)), (seed, g0)
    out = torch.empty(4, dtype=torch.int32, device="cuda")
    nb._cabi.check(L.noc_philox_raw(0, 0, 1, out.data_ptr(), None))
    assert [hex(int(v)) for v in out.cpu().numpy().view(np.uint32)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float64, 1e-13)])
def test_samples_match_the_oracle_and_shards_reproduce_rows(nb, dtype, tol):
    from oracle import philox_oracle as po
    center = torch.linspace(-3, 5, 13, dtype=dtype)
    n, d, var0, seed = 3001, 13, 0.37, 777
    x = nb.sample_rho0(center, var0, n, seed=seed)
    assert x.is_cuda and x.shape == (n, d) and x.dtype == dtype
    ref = center.double().numpy()[None, :] + var0 * po.normals(seed, 0, n, d)
    assert np.abs(x.double().cpu().numpy() - ref).max() <= tol * 10
    part = nb.sample_rho0(center, var0, 200, seed=seed, row0=1500)
    assert torch.equal(part, x[1500:1700])                       # a shard draws exactly its rows of the full batch
    assert torch.equal(nb.sample_rho0(center, var0, n, seed=seed), x) and not torch.equal(nb.sample_rho0(center, var0, n, seed=seed + 1), x)


@pytest.mark.parametrize("name", ["swarm50", "swap12", "singlequad"])
def test_distribution_of_rho0(nb, name):
    """Moments and a Kolmogorov-Smirnov test per problem at its own var0 (swarm50 0.1, swap12 1.0, singlequad 0.1 on the position
    columns only, the other nine columns exactly xInit)."""
    from scipy import stats
    net, prob, xinit, meta = product_setup(name, torch.float32)
    d = xinit.shape[1]
    n, var0 = 400000 if d < 100 else 60000, meta["var0"]
    cols = 3 if name == "singlequad" else None
    x = nb.sample_rho0(xinit, var0, n, seed=1234, noise_cols=cols)
    z = ((x - xinit) / var0).double()
    k = cols or d
    if cols:
        assert torch.equal(x[:, cols:], xinit[:, cols:].expand(n, d - cols))
    zz = z[:, :k]
    assert zz.mean(0).abs().max() < 5.0 / np.sqrt(n) and (zz.std(0) - 1).abs().max() < 5.0 / np.sqrt(2 * n)
    cov = (zz.T @ zz / n - torch.eye(k, dtype=torch.float64, device=zz.device)).abs().max()
    assert cov < 6.0 / np.sqrt(n)                                 # columns are independent
    flat = zz.reshape(-1)[:2000000].cpu().numpy()
    assert stats.kstest(flat, "norm").pvalue > 1e-3
    assert np.abs(flat).max() > 4.0                                # tails are populated
    # the rollout accepts the device-generated batch directly (no host copy anywhere)
    with torch.no_grad():
        s = nb.ocflow_sums(x[:1024].contiguous(), net, prob, [0.0, 1.0], 8, "rk4", meta["alph"])
    assert float(s[7]) == 1024 and torch.isfinite(s).all()


def test_resample_device_has_the_reference_signature(nb):
    net, prob, xinit, meta = product_setup("softcorridor", torch.float32)
    x0 = torch.zeros(64, 4, device="cuda")
    x1 = nb.resample_device(x0, xinit, 1.0, seed=3)
    assert x1.shape == x0.shape and x1.device == x0.device and not torch.equal(x1, nb.resample_device(x0, xinit, 1.0, seed=4))
