"""On-device sampler of the initial distribution rho_0: the device-side counterpart of
`xInit + cvt(var0 * torch.randn(n, d))` in src/initProb.py:27-28,107-120,132-140,196-203 and of `resample` (:252-262).

    x = sample_rho0(xInit, var0, n, seed=1234)            # [n, d] on the GPU, never on the host
    x = resample_device(x0, xInit, var0, seed=it)         # same shape as x0 (trainOC.py:255-256)

The generator is Philox4x32-10 + Box-Muller inside libnoc_b200.so (noc_sample_rho0); the random stream is not torch's, so the
parity with the reference is distributional (tests/test_gpu_sampler.py); shards reproduce their rows of the full batch
(`row0`).  For the quadcopter only the position columns are perturbed (`noise_cols=3`, initProb.py:132-140)."""
import torch

from . import _cabi


def sample_rho0(xInit, var0, n, seed=0, row0=0, noise_cols=None, device=None, dtype=None):
    if not torch.cuda.is_available():
        raise RuntimeError("neuraloc_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    L = _cabi.lib()
    dtype = dtype or xInit.dtype
    code = {torch.float32: _cabi.F32, torch.float64: _cabi.F64}.get(dtype)
    if code is None:
        raise ValueError("sample_rho0 supports float32 and float64")
    device = device or (xInit.device if xInit.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    center = xInit.detach().reshape(-1).to(device=device, dtype=dtype).contiguous()
    d = center.numel()
    with torch.cuda.device(device):
        x = torch.empty(int(n), d, dtype=dtype, device=device)
        rc = L.noc_sample_rho0(center.data_ptr(), d, d if noise_cols is None else int(noise_cols), float(var0), int(seed) & (2 ** 64 - 1),
                               int(row0), int(n), code, x.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _cabi.check(rc)
    return x


def resample_device(x0, xInit, var0, seed=0, noise_cols=None):
    """src/initProb.py:252-262 on the device: a fresh batch of x0's shape around xInit."""
    return sample_rho0(xInit, var0, x0.shape[0], seed=seed, noise_cols=noise_cols, device=x0.device if x0.is_cuda else None,
                       dtype=x0.dtype)
