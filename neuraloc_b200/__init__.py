"""neuraloc_b200 — B200-native closed-loop rollout for NeuralOC value networks.

The package mirrors the reference's interface for ONE hot path (donken/NeuralOC, src/OCflow.py:7):

    from neuraloc_b200 import OCflow, Phi, initProb
    Jc, cs = OCflow(x, net, prob, tspan=[0.0, 1.0], nt=50, stepper="rk4", alph=net.alph)

Everything numeric runs in hand-written CUDA (libnoc_b200.so, C ABI in include/noc_b200.h) on an
sm_100a device.  There is no CPU implementation: importing works anywhere, calling needs the built
library and a GPU and fails loudly otherwise.
"""
from .ocflow import OCflow, ocG, ocOdefun, stepRK1, stepRK4, ocflow_sums, costs_from_sums, invalidate_cache, OCflow_shock, ocflow_grad_sums, split_param_grads   # noqa: F401
from .phi import Phi, ResNN   # noqa: F401
from .problems import Cross2D, SwarmTraj, Quadcopter   # noqa: F401
from .init_prob import initProb, resample   # noqa: F401
from .sharded import OCflow_sharded, shard_rows, ocflow_grad_sharded   # noqa: F401
from .sampler import sample_rho0, resample_device   # noqa: F401
from .baseline import baseline_loss, loss_fun, compute_loss   # noqa: F401
from . import _cabi   # noqa: F401

__all__ = ["OCflow", "ocG", "ocOdefun", "stepRK1", "stepRK4", "Phi", "ResNN", "Cross2D", "SwarmTraj", "Quadcopter",
           "initProb", "resample", "OCflow_sharded", "shard_rows", "ocflow_sums", "costs_from_sums", "invalidate_cache", "sample_rho0", "resample_device", "OCflow_shock", "ocflow_grad_sums", "split_param_grads", "baseline_loss", "loss_fun", "compute_loss", "ocflow_grad_sharded"]
