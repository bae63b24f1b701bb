"""The baseline method's discrete-control objective on the GPU, behind the reference's own function names.

    loss_fun(U, Z_0, prob, nt, alphG)        baseline2D.py:42-63 / timeBaseline.py:50-70   (Cross2D, SwarmTraj problems)
    compute_loss(ctrls, x0, prob, alphG)     baselineQuad.py:47-72                          (one quadcopter)

Both accept the reference's single-sample arguments (U [nt, nc], Z_0 [d]) or a batch (U [n, nt, nc], Z_0 [n, d]) and are
differentiable with respect to the controls (`err.backward()`, baseline2D.py:97-102): forward and backward are one launch of
`noc_baseline_loss` (one warp per sample).  No torch arithmetic, no CPU path.
"""
import ctypes as C

import torch
from torch.autograd.function import once_differentiable

from . import _cabi
from .ocflow import _dtype_code, _prob_struct, _require_cuda


def baseline_loss(U, z0, prob, alphG, want_grad=False):
    """-> (loss [n], gradU [n, nt, nc] or None) on the CUDA device, for U [n, nt, nc] and z0 [n, d]."""
    _require_cuda()
    L = _cabi.lib()
    if U.dim() != 3 or z0.dim() != 2 or U.shape[0] != z0.shape[0]:
        raise ValueError("U must be [n, nt, nc] and z0 [n, d]")
    code = _dtype_code(U.dtype)
    n, nt, nc = U.shape
    d = z0.shape[1]
    device = U.device if U.is_cuda else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(device):
        pst, _keep = _prob_struct(prob, device, U.dtype)
        want_nc = 4 if pst.kind == _cabi.PROB_KINDS["Quadcopter"] else d
        if nc != want_nc:
            raise ValueError("U has %d control channels, the problem needs %d" % (nc, want_nc))
        Ud = U.detach().to(device).contiguous()
        zd = z0.detach().to(device=device, dtype=U.dtype).contiguous()
        loss = torch.empty(n, dtype=U.dtype, device=device)
        gU = torch.empty_like(Ud) if want_grad else None
        rc = L.noc_baseline_loss(C.byref(pst), Ud.data_ptr(), zd.data_ptr(), n, d, nt, float(alphG), code, loss.data_ptr(),
                                 None if gU is None else gU.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _cabi.check(rc)
    return loss, gU


class _BaselineLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, z0, prob, alphG):
        loss, gU = baseline_loss(U, z0, prob, alphG, want_grad=U.requires_grad)
        ctx.gU = None if gU is None else gU.to(U.device)
        return loss.to(U.device)

    @staticmethod
    @once_differentiable          # the kernels produce first derivatives only: a double backward raises instead of returning zeros
    def backward(ctx, gout):
        return (None if ctx.gU is None else ctx.gU * gout.reshape(-1, 1, 1), None, None, None)


def _batched(U, z0):
    single = U.dim() == 2
    return (U.unsqueeze(0), z0.reshape(1, -1), True) if single else (U, z0, False)


def loss_fun(U, Z_0, prob, nt, alphG):
    """baseline2D.py:42-63.  U nt-by-d (-> loss [1,1], as the reference) or n-by-nt-by-d (-> [n])."""
    Ub, zb, single = _batched(U, Z_0)
    if Ub.shape[1] != int(nt):
        raise ValueError("U has %d time steps, nt = %d" % (Ub.shape[1], int(nt)))
    out = _BaselineLoss.apply(Ub, zb, prob, float(alphG))
    return out.reshape(1, 1) if single else out


def compute_loss(ctrls, x0, prob, alphG=5000):
    """baselineQuad.py:47-72.  ctrls nt-by-4 (-> 0-dim loss) or n-by-nt-by-4 (-> [n])."""
    Ub, zb, single = _batched(ctrls, x0)
    out = _BaselineLoss.apply(Ub, zb, prob, float(alphG))
    return out.reshape(()) if single else out
