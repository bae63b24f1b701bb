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
