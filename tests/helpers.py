"""Shared helpers for the test-suite: fixture loading and oracle set-up (CPU only)."""
import json
import os

import numpy as np
import torch

from oracle import ocflow_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PROBLEMS = ["softcorridor", "swap2", "swap12", "singlequad", "swarm50"]
DT = {"f32": torch.float32, "f64": torch.float64}


def load_ckpt(name):
    z = np.load(os.path.join(GOLDEN, "ckpt", name + ".npz"))
    meta = json.loads(str(z["meta_json"]))
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta_json"}
    return sd, meta


def load_cases(name):
    return np.load(os.path.join(GOLDEN, "cases_%s.npz" % name))


def oracle_setup(name, dtype):
    sd, meta = load_ckpt(name)
    P = orc.params_from_state_dict(sd, dtype)
    D, xinit = orc.make_problem(meta["data"], meta["alph"], dtype)
    return P, D, xinit, meta


def rel_state_err(z, zref, d):
    """max over steps and samples of ||x_k - x_k^ref||_2 / ||x_k^ref||_2 (SURVEY.md §8d parity gate)."""
    z = np.asarray(z, dtype=np.float64)[:, :d, :]
    zr = np.asarray(zref, dtype=np.float64)[:, :d, :]
    num = np.linalg.norm(z - zr, axis=1)
    den = np.maximum(np.linalg.norm(zr, axis=1), 1e-30)
    return float((num / den).max())


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), max(floor, 1e-300))))


# ---------------------------------------------------------------------------------------------------------
# product-side set-up (neuraloc_b200 host mirrors); importing is CPU-safe, calling needs the GPU
# ---------------------------------------------------------------------------------------------------------
def product_setup(name, dtype, device="cuda"):
    """-> (net, prob, xInit, meta) built exactly like evalOC.py:51-64 but from the committed checkpoint fixtures."""
    import neuraloc_b200 as nb
    sd, meta = load_ckpt(name)
    cvt = lambda v: v.to(dtype).to(device)
    prob, x0, _, xinit = nb.initProb(meta["data"], 4, 4, var0=1.0, alph=meta["alph"], cvt=cvt)
    prob.eval()
    net = nb.Phi(nTh=meta["nTh"], m=meta["m"], d=x0.shape[1], alph=meta["alph"])
    net.load_state_dict(sd)
    net = net.to(dtype).to(device)
    net.eval()
    return net, prob, xinit, meta


def mean_vec(out):
    Jc, cs = out
    return np.array([float(Jc)] + [float(c) for c in cs])


QW = np.array([False] * 6 + [True, True])      # positions of Q, W in the 8-vector [Jc, L, G, HJt, HJfin, HJgrad, Q, W]


def check_costs(got, ref, rel, abs_floor, what="", floor_mask=None, ref_noise=None, noise_mult=2.0):
    """|got - ref| <= rel * |ref| per entry.  `abs_floor` applies ONLY to the entries selected by `floor_mask` (Q and W, which
    are 0 or tiny on most inputs: SURVEY.md H2/H3) — L, G, HJt, HJfin, HJgrad and Jc are gated relatively, G included.
    `ref_noise` (same shape) widens an entry's tolerance to `noise_mult` (2) times the reference's own fp32<->fp64 distance there."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    tol = rel * np.abs(ref)
    if ref_noise is not None:
        tol = np.maximum(tol, noise_mult * np.abs(np.asarray(ref_noise, dtype=np.float64)))
    ok = err <= tol
    if floor_mask is not None:
        ok = ok | (np.asarray(floor_mask, dtype=bool) & (err <= abs_floor))
    assert ok.all(), "%s: got %s ref %s relerr %s (tolerance %s)" % (what, got, ref, err / np.maximum(np.abs(ref), 1e-300),
                                                                   tol / np.maximum(np.abs(ref), 1e-300))
