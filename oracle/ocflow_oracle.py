"""CPU oracle for the NeuralOC closed-loop rollout — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this module, and only as the checker (or as the timed CPU baseline). The product path
(neuraloc_b200/) never imports it and has no CPU fallback.

What it restates (reference = donken/NeuralOC, read-only at /root/reference in the build container):

  phi_forward / phi_grad   src/Phi.py:91-96, 99-138   (ResNN forward src/Phi.py:40-52, act :8-9)
  cross2d_* / swarm_* / quad_*   src/problem/Cross2D.py:69-165, SwarmTraj.py:68-167, Quadcopter.py:65-197
  gauss_pdf                src/utils.py:70-86
  rhs                      src/OCflow.py:104-140 (ocOdefun)
  rk4_step / rk1_step      src/OCflow.py:157-184, 143-155
  ocflow                   src/OCflow.py:7-95
  make_problem             src/initProb.py:9-249 (targets / initial centres / radii)
  baseline_loss            baseline2D.py:42-63 (loss_fun), baselineQuad.py:40-72 (dyn, compute_loss), batched over samples

The arithmetic lives in PyTorch (CPU aten kernels; the reference pins torch==1.7.0, this image has
2.11): this file is a functional restatement on plain tensors (no nn.Module, no problem classes), in
the same dtype the caller passes (fp32 or fp64), with the same op order where rounding could matter.

Parity pinning: the reference ships NO tests or golden vectors, so the oracle is pinned
against outputs of the unmodified reference itself, generated in the build container by
tests/golden/make_golden.py and committed under tests/golden/ (tests/test_oracle_golden.py), and
against the live reference when /root/reference is present (tests/test_oracle_vs_reference.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

__all__ = ["PhiParams", "ProbDesc", "phi_forward", "phi_grad", "lhqw", "grad_p_hamiltonian", "controls",
           "rhs", "ocflow", "stage_time_table", "make_problem", "params_from_state_dict", "baseline_loss"]


# ----------------------------------------------------------------------------------------------
# value network
# ----------------------------------------------------------------------------------------------
@dataclass
class PhiParams:
    """Flat view of a Phi state_dict (layout: src/Phi.py:77-87, 32-36)."""
    A: torch.Tensor            # [r, D]
    c_w: torch.Tensor          # [1, D]
    c_b: torch.Tensor          # [1]
    w: torch.Tensor            # [1, m]
    K: List[torch.Tensor]      # K[0]: [m, D]; K[i>=1]: [m, m]
    b: List[torch.Tensor]      # b[i]: [m]
    h: float = field(default=0.0)  # ResNet step 1/(nTh-1), src/Phi.py:38

    def __post_init__(self):
        if self.h == 0.0:
            self.h = 1.0 / (len(self.K) - 1)

    @property
    def nTh(self):
        return len(self.K)

    def to(self, dtype):
        return PhiParams(self.A.to(dtype), self.c_w.to(dtype), self.c_b.to(dtype), self.w.to(dtype),
                         [k.to(dtype) for k in self.K], [v.to(dtype) for v in self.b], self.h)


def params_from_state_dict(sd, dtype=None) -> PhiParams:
    g = lambda k: torch.as_tensor(sd[k]) if dtype is None else torch.as_tensor(sd[k]).to(dtype)
    n = 0
    while "N.layers.%d.weight" % n in sd:
        n += 1
    return PhiParams(g("A"), g("c.weight"), g("c.bias"), g("w.weight"),
                     [g("N.layers.%d.weight" % i) for i in range(n)], [g("N.layers.%d.bias" % i) for i in range(n)])


def _act(v):  # antiderivative of tanh, src/Phi.py:8-9
    a = v.abs()
    return a + torch.log(1 + torch.exp(-2.0 * a))


def _resnet(P: PhiParams, s):
    """u_{nTh-1}; src/Phi.py:40-52."""
    u = _act(torch.addmm(P.b[0], s, P.K[0].t()))
    for i in range(1, P.nTh):
        u = u + P.h * _act(torch.addmm(P.b[i], u, P.K[i].t()))
    return u


def phi_forward(P: PhiParams, s):
    """Phi(s) = w.N(s) + 0.5 s'A'A s + c_w.s + c_b  -> [n,1]; src/Phi.py:91-96."""
    sym = P.A.t() @ P.A
    return _resnet(P, s) @ P.w.t() + 0.5 * ((s @ sym) * s).sum(dim=1, keepdim=True) + (s @ P.c_w.t() + P.c_b)


def phi_grad(P: PhiParams, s):
    """grad_s Phi(s) -> [n, D] (last column = d/dt); src/Phi.py:99-138."""
    sym = P.A.t() @ P.A
    pre0 = torch.addmm(P.b[0], s, P.K[0].t())
    us = [_act(pre0)]
    for i in range(1, P.nTh):
        us.append(us[-1] + P.h * _act(torch.addmm(P.b[i], us[-1], P.K[i].t())))
    zrev = P.w.t()                                            # [m,1], broadcasts over samples
    for i in range(P.nTh - 1, 0, -1):
        th = torch.tanh(torch.addmm(P.b[i], us[i - 1], P.K[i].t()))   # the reference recomputes this
        zrev = zrev + P.h * (P.K[i].t() @ (th.t() * zrev))
    z0 = P.K[0].t() @ (torch.tanh(pre0).t() * zrev)
    return (z0 + sym @ s.t() + P.c_w.t()).t()


# ----------------------------------------------------------------------------------------------
# problems
# ----------------------------------------------------------------------------------------------
@dataclass
class ProbDesc:
    kind: str                  # 'Cross2D' | 'SwarmTraj' | 'Quadcopter'
    xtarget: torch.Tensor      # [d]
    obstacle: Optional[str] = None
    alph_Q: float = 1.0
    alph_W: float = 1.0
    r: float = 0.5
    nAgents: int = 1
    agentDim: int = 2
    mass: float = 1.0
    grav: float = 9.81
    training: bool = False

    def to(self, dtype):
        return ProbDesc(self.kind, self.xtarget.to(dtype), self.obstacle, self.alph_Q, self.alph_W, self.r,
                        self.nAgents, self.agentDim, self.mass, self.grav, self.training)


def gauss_pdf(x, mu, cov):
    """diagonal-covariance normal pdf, [k,dim] -> [k,1]; src/utils.py:70-86."""
    dim = x.shape[1]
    mu = torch.as_tensor(mu, dtype=x.dtype).view(1, dim)
    cov = torch.as_tensor(cov, dtype=x.dtype).view(1, dim)
    den = (2 * math.pi) ** (0.5 * dim) * torch.sqrt(torch.prod(cov))
    return torch.exp(-0.5 * torch.sum((x - mu) ** 2 / cov, 1, keepdim=True)) / den


def _norm_to(xa, mu):
    return torch.norm(xa - torch.as_tensor(mu, dtype=xa.dtype).view(1, -1), dim=1)


def _obstacle_cross2d(D: ProbDesc, xa):
    """per-agent terrain, src/problem/Cross2D.py:90-119. eval-mode hardcorridor returns the bool mask."""
    if D.obstacle == "softcorridor":
        cov = [0.2, 0.2]
        return sum(gauss_pdf(xa, mu, cov) for mu in ([-2.5, 0.0], [2.5, 0.0], [-1.5, 0.0], [1.5, 0.0]))
    if D.obstacle == "hardcorridor":
        mu1, mu2 = [0.0, 4.0], [0.0, -3.5]
        if not D.training:
            return (_norm_to(xa, mu1) < 2.0) | (_norm_to(xa, mu2) < 2.0)
        q = gauss_pdf(xa, mu1, [1.0, 1.0]) + gauss_pdf(xa, mu2, [1.0, 1.0])
        keep = (_norm_to(xa, mu1) < 2.0 + D.r) | (_norm_to(xa, mu2) < 2.0 + D.r)
        q[~keep] = 0.0
        return q
    return 0.0 * xa


def _obstacle_swarm(D: ProbDesc, xa):
    """src/problem/SwarmTraj.py:90-122."""
    if D.obstacle != "blocks":
        return 0.0 * xa
    px, py, pz = xa[:, 0], xa[:, 1], xa[:, 2]
    g = D.r if D.training else 0.0
    if D.training:
        inside = ((px < 2.0 + g) & (px > -2.0 - g) & (py < 0.5 + g) & (py > -0.5 - g) & (pz < 7.0 + g)) | \
                 ((px < 4.0 + g) & (px > 2.0 - g) & (py < 1.0 + g) & (py > -1.0 - g) & (pz < 4.0 + g))
    else:
        inside = ((px < 2.0) & (px > -2.0) & (py < 0.5) & (py > -0.5) & (pz < 7.0)) | \
                 ((px < 4.0) & (px > 2.0) & (py < 1.0) & (py > -1.0) & (pz < 4.0))
        return inside.unsqueeze(1)
    cov1 = 3.0 * torch.tensor([3.0, 1.0, 3.0], dtype=xa.dtype)
    cov2 = 3.0 * torch.tensor([3.0, 1.0, 1.0], dtype=xa.dtype)
    q = gauss_pdf(xa[:, 0:3], [0.0, 0.0, 2.0], cov1) + gauss_pdf(xa[:, 0:3], [2.5, 0.0, 2.0], cov2) + 999.0
    q[~inside] = 0.0
    return q


def _terrain(D: ProbDesc, x):
    """calcQ: sum of per-agent terrain; Cross2D.py:121-128 / SwarmTraj.py:124-131 / Quadcopter.py:124-131."""
    if D.obstacle is None:
        return 0.0 * x[:, 0].unsqueeze(1)
    xa = x.reshape(-1, D.agentDim)
    if D.kind == "Cross2D":
        q = _obstacle_cross2d(D, xa)
    elif D.kind == "SwarmTraj":
        q = _obstacle_swarm(D, xa)
    else:
        q = 0.0 * xa   # Quadcopter.py:115-122: no obstacle is implemented
    return torch.sum(q.reshape(x.shape[0], -1), dim=1, keepdim=True).to(x.dtype)   # bool/int counts -> real


def _interaction(D: ProbDesc, x):
    """calcW for Cross2D (:130-162) and SwarmTraj (:133-164): pairwise Gaussian repulsion inside a cut-off."""
    n, A, dim, r = x.shape[0], D.nAgents, D.agentDim, D.r
    if A == 1:
        return 0.0 * x[:, 0]
    if A == 2:
        dist = torch.norm(x[:, 0:dim] - x[:, dim:2 * dim], p=2, dim=1, keepdim=True)
        near = dist < (2.2 * r if D.training else 2 * r)
        return near * torch.exp(-dist ** 2 / (2 * r ** 2))
    cut = 2 * r
    if D.training:
        cut = 2.2 * r if D.kind == "Cross2D" else 3.2 * r
    xa = x.view(n, A, dim)
    dist = torch.norm(xa.reshape(n, A, 1, dim) - xa.reshape(n, 1, A, dim), p=2, dim=3)
    e = torch.exp(-((dist < cut) * dist) ** 2 / (2 * r ** 2))
    ones = e == 1.0
    return ((e.sum(dim=[1, 2]) - ones.sum(dim=[1, 2])) / 2.0).view(-1, 1)


def _quad_f(ang):
    """rotation-matrix third column; Quadcopter.py:182-197."""
    s0, s1, s2 = torch.sin(ang[:, 0]), torch.sin(ang[:, 1]), torch.sin(ang[:, 2])
    c0, c1, c2 = torch.cos(ang[:, 0]), torch.cos(ang[:, 1]), torch.cos(ang[:, 2])
    return s0 * s2 + c0 * s1 * c2, -c0 * s2 + s0 * s1 * c2, c1 * c2


def _quad_thrust(D: ProbDesc, xa, pa):
    f7, f8, f9 = _quad_f(xa[:, 3:6])
    u = -1 / (2 * D.mass) * (f7 * pa[:, 6] + f8 * pa[:, 7] + f9 * pa[:, 8]).view(-1, 1)
    return u, f7, f8, f9


def _quad_interaction(D: ProbDesc, x):
    """Quadcopter.py:134-158 (nAgents<=2 only; the >2 branch of the reference is broken and unreachable)."""
    if D.nAgents == 1:
        return (0.0 * x[:, 0]).view(-1, 1)
    if D.nAgents == 2:
        dist = torch.norm(x[:, 0:3] - x[:, 12:15], p=2, dim=1, keepdim=True)
        return (dist < 2 * D.r) * torch.exp(-dist ** 2 / (2 * D.r ** 2))
    raise NotImplementedError("Quadcopter with more than two agents (broken in the reference)")


def lhqw(D: ProbDesc, x, p):
    """(L, H, Q, W), each [n,1]; Cross2D.py:73-87, SwarmTraj.py:71-87, Quadcopter.py:86-113."""
    if D.kind in ("Cross2D", "SwarmTraj"):
        if D.kind == "Cross2D":
            Q = D.alph_Q * _terrain(D, x)                       # returned PRE-SCALED
            L = 0.5 * torch.sum(p ** 2, dim=1, keepdim=True) + Q
        else:
            Q = _terrain(D, x).view(-1, 1) if D.alph_Q > 0 else 0.0 * x[:, 0].view(-1, 1)
            L = 0.5 * torch.sum(p ** 2, dim=1, keepdim=True) + D.alph_Q * Q
        if D.alph_W != 0.0:
            W = _interaction(D, x)
            L = L + D.alph_W * W
        else:
            W = 0.0 * L
        H = -L + torch.sum(p ** 2, dim=1, keepdim=True)
        return L, H, Q, W
    # Quadcopter
    H = 0.0
    Q = _terrain(D, x).view(-1, 1)
    L = D.alph_Q * Q
    if D.alph_W > 0.0:
        W = _quad_interaction(D, x).view(-1, 1)
        L = L + D.alph_W * W
    else:
        W = 0.0 * L
    for j in range(D.nAgents):
        xa, pa = x[:, 12 * j:12 * (j + 1)], p[:, 12 * j:12 * (j + 1)]
        sq = (pa[:, 9] ** 2 + pa[:, 10] ** 2 + pa[:, 11] ** 2).view(-1, 1)
        u, f7, f8, f9 = _quad_thrust(D, xa, pa)
        L = L + 2 + u ** 2 + 0.25 * sq
        H = H - L \
            - torch.sum(xa[:, 6:9] * pa[:, 0:3], dim=1, keepdim=True) \
            - torch.sum(xa[:, 9:12] * pa[:, 3:6], dim=1, keepdim=True) \
            - (u / D.mass) * (f7 * pa[:, 6] + f8 * pa[:, 7] + f9 * pa[:, 8]).unsqueeze(1) \
            + D.grav * pa[:, 8].unsqueeze(1) + 0.5 * sq
    return L, H, Q, W


def grad_p_hamiltonian(D: ProbDesc, x, p):
    """Cross2D.py:69-70, SwarmTraj.py:68-69 (= p); Quadcopter.py:65-84."""
    if D.kind != "Quadcopter":
        return p
    cols = []
    for j in range(D.nAgents):
        xa, pa = x[:, 12 * j:12 * (j + 1)], p[:, 12 * j:12 * (j + 1)]
        u, f7, f8, f9 = _quad_thrust(D, xa, pa)
        cols += [-xa[:, 6:], -(u / D.mass) * f7.view(-1, 1), -(u / D.mass) * f8.view(-1, 1),
                 -(u / D.mass) * f9.view(-1, 1) + D.grav, 0.5 * pa[:, 9:12]]
    return torch.cat(cols, dim=1)


def controls(D: ProbDesc, x, p):
    """Cross2D.py:164-165, SwarmTraj.py:166-167 (= -p); Quadcopter.py:165-174 ([u, -p[9:12]/2] per agent)."""
    if D.kind != "Quadcopter":
        return -p
    cols = []
    for j in range(D.nAgents):
        xa, pa = x[:, 12 * j:12 * (j + 1)], p[:, 12 * j:12 * (j + 1)]
        u = _quad_thrust(D, xa, pa)[0]
        cols += [u, -0.5 * pa[:, 9:12]]
    return torch.cat(cols, dim=1)


# ----------------------------------------------------------------------------------------------
# integrator + objective
# ----------------------------------------------------------------------------------------------
def _with_time(x, t):
    return torch.nn.functional.pad(x, (0, 1, 0, 0), value=t)


def rhs(z, t, P: PhiParams, D: ProbDesc):
    """d/dt [x, int L, int |Phi_t - H|, int Q, int W]; src/OCflow.py:104-140."""
    d = z.shape[1] - 4
    s = _with_time(z[:, :d], t)
    g = phi_grad(P, s)
    L, H, Q, W = lhqw(D, s[:, :d], g[:, :d])
    out = torch.zeros_like(z)
    out[:, :d] = -grad_p_hamiltonian(D, s[:, :d], g[:, :d])
    out[:, d] = L.squeeze()
    out[:, d + 1] = torch.abs(g[:, -1] - H.squeeze())
    out[:, d + 2] = Q.squeeze()
    out[:, d + 3] = W.squeeze()
    return out


def rk4_step(z, P, D, ta, tb):
    """src/OCflow.py:157-184 (h is recomputed from the two end points, in double)."""
    h = tb - ta
    k = h * rhs(z, ta, P, D)
    acc = z + (1.0 / 6.0) * k
    k = h * rhs(z + 0.5 * k, ta + (h / 2), P, D)
    acc = acc + (2.0 / 6.0) * k
    k = h * rhs(z + 0.5 * k, ta + (h / 2), P, D)
    acc = acc + (2.0 / 6.0) * k
    k = h * rhs(z + k, ta + h, P, D)
    return acc + (1.0 / 6.0) * k


def rk1_step(z, P, D, ta, tb):
    """src/OCflow.py:143-155."""
    return z + (tb - ta) * rhs(z, ta, P, D)


def stage_time_table(t0: float, t1: float, nt: int):
    """Per step: (t_a, t_a + h'/2, t_a + h', t_ctrl) in double, replaying OCflow.py:25,35,47,50,53 and :169.

    t_ctrl is the time the control at the NEW state is evaluated at: `tk - h` after `tk += h` (quirk 3)."""
    h = (t1 - t0) / nt
    tk = t0
    rows = []
    for _ in range(nt):
        ta, tb = tk, tk + h
        hh = tb - ta
        tk += h
        rows.append((ta, ta + (hh / 2), ta + hh, tk - h))
    return rows


def ocflow(x, P: PhiParams, D: ProbDesc, tspan: Sequence[float], nt: int, stepper: str = "rk4",
           alph: Sequence[float] = (1.0,) * 6, intermediates: bool = False, noMean: bool = False):
    """Restatement of src/OCflow.py:7-95 with the reference's return conventions."""
    n, d = x.shape
    h = (tspan[1] - tspan[0]) / nt
    z = torch.cat((x, torch.zeros(n, 4, dtype=x.dtype)), 1)
    tk = tspan[0]
    if intermediates:
        zfull = torch.zeros(n, d + 4, nt + 1, dtype=x.dtype)
        zfull[:, :, 0] = z
        nctrl = controls(D, x, phi_grad(P, _with_time(x, 0))[:, :d]).shape[1]
        cfull = torch.zeros(n, nctrl, nt + 1, dtype=x.dtype)
    for k in range(nt):
        if stepper == "rk4":
            z = rk4_step(z, P, D, tk, tk + h)
        elif stepper == "rk1":
            z = rk1_step(z, P, D, tk, tk + h)
        tk += h
        if intermediates:
            zfull[:, :, k + 1] = z
            pk = phi_grad(P, _with_time(z[:, :d], tk - h))[:, :d]
            cfull[:, :, k + 1] = controls(D, z[:, :d], pk)
    res = z[:, :d] - D.xtarget
    cG = 0.5 * torch.sum(res ** 2, 1, keepdim=True)
    sT = _with_time(z[:, :d], tspan[1])
    phiT = phi_forward(P, sT)
    gT = phi_grad(P, sT)[:, :d]
    hjf = torch.sum(torch.abs(phiT - alph[0] * cG), 1)
    hjg = torch.sum(torch.abs(gT - alph[0] * res), 1)
    if noMean:
        cs = [z[:, -4].view(-1, 1), cG.view(-1, 1), z[:, -3].view(-1, 1), hjf.view(-1, 1), hjg.view(-1, 1),
              z[:, -2].view(-1, 1), z[:, -1].view(-1, 1)]
        return cs[0] + alph[0] * cs[1] + alph[3] * cs[2] + alph[4] * cs[3] + alph[5] * cs[4], cs
    cs = [torch.mean(z[:, -4]), torch.mean(cG), torch.mean(z[:, -3]), torch.mean(hjf), torch.mean(hjg),
          torch.mean(z[:, -2]), torch.mean(z[:, -1])]
    if intermediates:
        return zfull, cfull
    return cs[0] + alph[0] * cs[1] + alph[3] * cs[2] + alph[4] * cs[3] + alph[5] * cs[4], cs


# ----------------------------------------------------------------------------------------------
# baseline: discrete-control objective
# ----------------------------------------------------------------------------------------------
def baseline_loss(D: ProbDesc, U, z0, alphG: float):
    """Per-sample objective of the baseline method for controls U [n, nt, nc] and initial states z0 [n, d] -> [n].

    Cross2D / SwarmTraj — loss_fun, baseline2D.py:42-63: Z += h U_i; loss += h L(Z, U_i) (L of calcLHQW with p := U_i, at the
    NEW state); + alphG * 0.5 |Z_nt - xtarget|^2.  Quadcopter — dyn / compute_loss, baselineQuad.py:40-72: x += h dyn(c_i, x)
    with dyn = [x[6:], (c0/mass) f(x[3:6]) - grav e_z, c[1:4]]; J += h (2 + |c_i|^2); + alphG * 0.5 |x_nt - xtarget|^2."""
    n, nt = U.shape[0], U.shape[1]
    h = 1.0 / nt
    Z = z0
    loss = torch.zeros(n, dtype=U.dtype)
    if D.kind == "Quadcopter":
        for i in range(nt):
            c = U[:, i, :]
            f7, f8, f9 = _quad_f(Z[:, 3:6])
            tmp = c[:, 0] / D.mass
            dx = torch.cat([Z[:, 6:], (tmp * f7).unsqueeze(1), (tmp * f8).unsqueeze(1), (tmp * f9 - D.grav).unsqueeze(1), c[:, 1:4]], 1)
            Z = Z + h * dx
            loss = loss + h * (2 + torch.norm(c, p=2, dim=1) ** 2)
        return loss + alphG * 0.5 * torch.norm(Z - D.xtarget, p=2, dim=1) ** 2
    for i in range(nt):
        Z = Z + h * U[:, i, :]
        L = lhqw(D, Z, U[:, i, :])[0]
        loss = loss + h * L.reshape(-1)
    cG = 0.5 * torch.sum((Z - D.xtarget) ** 2, 1)
    return loss + alphG * cG


# ----------------------------------------------------------------------------------------------
# problem table (src/initProb.py:9-249) — targets, initial centres, radii
# ----------------------------------------------------------------------------------------------
def _swarm_targets(rows, shift=(0.0, -0.5, -3.0)):
    top = torch.tensor(rows, dtype=torch.float64)
    return torch.cat((top, top + torch.tensor(shift, dtype=torch.float64)), 0)


_SWARM32 = [[-2, 2, 8], [-1, 2, 8], [0, 2, 8], [1, 2, 8], [2, 2, 8], [-2.5, 3, 8], [-1.5, 3, 8], [-.5, 3, 8],
            [.5, 3, 8], [1.5, 3, 8], [2.5, 3, 8], [-2, 4, 8], [-1, 4, 8], [0, 4, 8], [1, 4, 8], [2, 4, 8]]
_SWARM50 = [[-2, 2, 6], [-1, 2, 6], [0, 2, 6], [1, 2, 6], [2, 2, 6], [3, 2, 6], [4, 2, 6],
            [-2.5, 3, 7], [-1.5, 3, 7], [-.5, 3, 7], [.5, 3, 7], [1.5, 3, 7], [2.5, 3, 7], [3.5, 3, 7],
            [-2, 4, 8], [-1, 4, 8], [0, 4, 8], [1, 4, 8], [2, 4, 8], [3, 4, 8], [4, 4, 8],
            [-2, 3, 5], [-1, 3, 5], [1, 3, 5], [2, 3, 5]]
_SWAP12_T = [2, 2, 0, 0, 10, 0, -10, 0, 5, 5, -5, -5, -4, 2, -6, -1, 5, -5, -5, 5, 2, -2, -2, -2]
_SWAP12_I = [0, 0, 2, 2, -10, 0, 10, 0, -5, -5, 5, 5, -6, -1, -4, 2, -5, 5, 5, -5, -2, -2, 2, -2]


def make_problem(name: str, alph: Sequence[float], dtype=torch.float32):
    """-> (ProbDesc in eval mode, xInit [1,d]); constants of src/initProb.py (cited per branch)."""
    f = lambda v: torch.as_tensor(v, dtype=torch.float64).reshape(-1)
    aQ, aW = float(alph[1]), float(alph[2])
    if name == "softcorridor":                                   # initProb.py:25-31
        D = ProbDesc("Cross2D", f([2, 2, -2, 2]), "softcorridor", aQ, aW, 0.5, 2, 2)
        xi = f([-2, -2, 2, -2])
    elif name in ("swarm", "swarm50"):                           # initProb.py:33-74, 76-125
        tg = _swarm_targets(_SWARM32 if name == "swarm" else _SWARM50)
        xi = (torch.tensor([1.0, -1.0, -1.0], dtype=torch.float64) * tg + torch.tensor([0.0, 0.0, 10.0], dtype=torch.float64)).reshape(-1)
        D = ProbDesc("SwarmTraj", tg.reshape(-1), "blocks", aQ, aW, 0.2 if name == "swarm" else 0.1, tg.shape[0], 3)
    elif name == "singlequad":                                   # initProb.py:127-142
        D = ProbDesc("Quadcopter", f([2, 2, 2] + [0] * 9), None, 0.0, 0.0, 1.0, 1, 12)
        xi = f([-1.5] * 3 + [0] * 9)
    elif name == "midcross2":                                    # initProb.py:144-151
        D = ProbDesc("Cross2D", f([2, 2, -2, 2]), None, aQ, aW, 0.5, 2, 2)
        xi = f([-2, -2, 2, -2])
    elif name in ("midcross4", "midcross20"):                    # initProb.py:153-173
        A, lim, r = (4, 2.0, 0.4) if name == "midcross4" else (20, 6.0, 0.15)
        xx = torch.linspace(-lim, lim, A).double()
        tg = torch.stack((xx.flip(0), lim * torch.ones(A, dtype=torch.float64)), 1).reshape(-1)
        xi = torch.stack((xx, -lim * torch.ones(A, dtype=torch.float64)), 1).reshape(-1)
        D = ProbDesc("Cross2D", tg, None, aQ, aW, r, A, 2)
    elif name == "midcross30":                                   # initProb.py:174-187
        A = 30
        xx = torch.linspace(-6, 6, A).double()
        lvl = torch.tensor([6.0, 4.0, 2.0], dtype=torch.float64).repeat(A // 3)
        tg = torch.stack((xx.flip(0), lvl), 1).reshape(-1)
        xi = torch.stack((xx, -lvl), 1).reshape(-1)
        D = ProbDesc("Cross2D", tg, None, aQ, aW, 0.2, A, 2)
    elif name == "swap2":                                        # initProb.py:188-195
        D = ProbDesc("Cross2D", f([10, 0, -10, 0]), "hardcorridor", aQ, aW, 1.0, 2, 2)
        xi = f([-10, 0, 10, 0])
    elif name == "swap12" or name.startswith("swap12_"):        # initProb.py:196-243
        A = 12 if name == "swap12" else 2 * int(name[7])
        D = ProbDesc("Cross2D", f(_SWAP12_T[:2 * A]), None, aQ, aW, 0.5, A, 2)
        xi = f(_SWAP12_I[:2 * A])
    else:
        raise ValueError("incorrect value passed to --data: %r" % name)
    return D.to(dtype), xi.reshape(1, -1).to(dtype)
