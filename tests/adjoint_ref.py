"""Hand-derived discrete adjoint of the closed-loop rollout — TEST INFRASTRUCTURE (the formulas the CUDA kernel
`rollout_grad_kernel` implements, restated with torch CPU ops so that they can be checked against autograd of the oracle).

Reference path: trainOC.py:169-174 (`Jc, cs = OCflow(...); Jc.backward()`), i.e. reverse-mode differentiation of
src/OCflow.py:7-95 through stepRK4 (:157-184), ocOdefun (:104-140), Phi.getGrad (src/Phi.py:99-138) and the problems'
calcLHQW / calcGradpH.  Everything below is for nTh = 2 (every shipped configuration).

Notation: s = [x, t];  o = K0 s + b0, T0 = tanh(o), u0 = act(o);  a1 = K1 u0 + b1, T1 = tanh(a1);
y = T1*w;  z1 = w + h K1'y;  v = T0*z1;  g = grad Phi = K0'v + A'A s + c_w.
One "adjoint evaluation" gets the adjoint of the RHS outputs (abar on dx/dt, cL on L, cH on |Phi_t - H|), forms
gbar = d psi / d g and the direct d psi / d x from the problem functor, then differentiates the scalar gbar . grad Phi(s)
(+ beta Phi(s) in the terminal block) with respect to s and to every parameter:
    odot = K0 gbar, udot = T0*odot, adot = K1 udot
    bar_a1 = h w*adot*(1 - T1^2) + beta h y;   bar_u0 = K1' bar_a1 + beta w
    bar_o  = (1 - T0^2)*odot*z1 + T0*bar_u0
    sbar   = K0' bar_o + A'A gbar + beta (A'A s + c_w)
    dK1 += (h y) (x) udot + bar_a1 (x) u0;  db1 += bar_a1;  dK0 += v (x) gbar + bar_o (x) s;  db0 += bar_o
    dw  += udot + h T1*adot + beta u1;  dA += (A s) (x) gbar + (A gbar) (x) s + beta (A s) (x) s;  dc_w += gbar + beta s;  dc_b += beta
"""
import math

import torch


def _act(v):
    a = v.abs()
    return a + torch.log(1 + torch.exp(-2.0 * a))


def chain(P, s):
    K0, K1, b0, b1, w = P.K[0], P.K[1], P.b[0], P.b[1], P.w.reshape(-1)
    o = s @ K0.t() + b0
    T0, u0 = torch.tanh(o), _act(o)
    a1 = u0 @ K1.t() + b1
    T1 = torch.tanh(a1)
    y = T1 * w
    z1 = w + P.h * (y @ K1)
    v = T0 * z1
    S = P.A.t() @ P.A
    q = s @ S
    g = v @ K0 + q + P.c_w.reshape(-1)
    return dict(o=o, T0=T0, u0=u0, a1=a1, T1=T1, y=y, z1=z1, v=v, q=q, g=g, S=S)


def zero_grads(P):
    return dict(A=torch.zeros_like(P.A), c_w=torch.zeros_like(P.c_w), c_b=torch.zeros_like(P.c_b), w=torch.zeros_like(P.w),
                K0=torch.zeros_like(P.K[0]), K1=torch.zeros_like(P.K[1]), b0=torch.zeros_like(P.b[0]), b1=torch.zeros_like(P.b[1]))


def adjoint_eval(P, s, c, gbar, beta, G):
    """Accumulates d(gbar . g + beta . Phi)/d theta into G (sums over the batch); returns d/ds [n, D]."""
    K0, K1, w, h = P.K[0], P.K[1], P.w.reshape(-1), P.h
    T0, T1, u0, y, z1, v = c["T0"], c["T1"], c["u0"], c["y"], c["z1"], c["v"]
    odot = gbar @ K0.t()
    udot = T0 * odot
    adot = udot @ K1.t()
    bar_a1 = h * w * adot * (1 - T1 * T1) + beta * h * y
    bar_u0 = bar_a1 @ K1 + beta * w
    bar_o = (1 - T0 * T0) * odot * z1 + T0 * bar_u0
    sbar = bar_o @ K0 + gbar @ c["S"] + beta * (c["q"] + P.c_w.reshape(-1))
    G["K1"] += (h * y).t() @ udot + bar_a1.t() @ u0
    G["b1"] += bar_a1.sum(0)
    G["K0"] += v.t() @ gbar + bar_o.t() @ s
    G["b0"] += bar_o.sum(0)
    u1 = u0 + h * _act(c["a1"])
    G["w"] += (udot + h * T1 * adot + beta * u1).sum(0, keepdim=True)
    As, Ag = s @ P.A.t(), gbar @ P.A.t()
    G["A"] += As.t() @ gbar + Ag.t() @ s + (beta * As).t() @ s
    G["c_w"] += (gbar + beta * s).sum(0, keepdim=True)
    G["c_b"] += beta.sum().reshape(1) if torch.is_tensor(beta) else torch.zeros(1, dtype=s.dtype)
    return sbar


def _gauss(xa, mu, cov):
    mu = torch.as_tensor(mu, dtype=xa.dtype)
    cov = torch.as_tensor(cov, dtype=xa.dtype)
    den = (2 * math.pi) ** (0.5 * xa.shape[-1]) * torch.sqrt(torch.prod(cov))
    pdf = torch.exp(-0.5 * torch.sum((xa - mu) ** 2 / cov, -1, keepdim=True)) / den
    return pdf, -pdf * (xa - mu) / cov


def terrain_grad(D, x):
    """d calcQ / d x  (unscaled sum of the per-agent terrain), [n, d]."""
    n, d = x.shape
    if D.obstacle is None or D.kind == "Quadcopter":
        return torch.zeros_like(x)
    xa = x.reshape(n, D.nAgents, D.agentDim)
    if D.obstacle == "softcorridor":
        return sum(_gauss(xa, mu, [0.2, 0.2])[1] for mu in ([-2.5, 0.0], [2.5, 0.0], [-1.5, 0.0], [1.5, 0.0])).reshape(n, d)
    if not D.training:
        return torch.zeros_like(x)                       # eval mode: inside-counts, no gradient
    if D.obstacle == "hardcorridor":
        mu1, mu2 = torch.tensor([0.0, 4.0], dtype=x.dtype), torch.tensor([0.0, -3.5], dtype=x.dtype)
        keep = (torch.norm(xa - mu1, dim=-1) < 2.0 + D.r) | (torch.norm(xa - mu2, dim=-1) < 2.0 + D.r)
        gq = _gauss(xa, mu1, [1.0, 1.0])[1] + _gauss(xa, mu2, [1.0, 1.0])[1]
        return (gq * keep.unsqueeze(-1)).reshape(n, d)
    if D.obstacle == "blocks":
        px, py, pz = xa[..., 0], xa[..., 1], xa[..., 2]
        g = D.r
        inside = ((px < 2.0 + g) & (px > -2.0 - g) & (py < 0.5 + g) & (py > -0.5 - g) & (pz < 7.0 + g)) | \
                 ((px < 4.0 + g) & (px > 2.0 - g) & (py < 1.0 + g) & (py > -1.0 - g) & (pz < 4.0 + g))
        gq = _gauss(xa, [0.0, 0.0, 2.0], [9.0, 3.0, 9.0])[1] + _gauss(xa, [2.5, 0.0, 2.0], [9.0, 3.0, 3.0])[1]
        return (gq * inside.unsqueeze(-1)).reshape(n, d)
    raise ValueError(D.obstacle)


def interaction_grad(D, x):
    """d calcW / d x, [n, d]:  sum over pairs inside the cut-off of -e_ij (x_i - x_j) / r^2 on agent i (and + on agent j)."""
    n, d = x.shape
    A, dim, r = D.nAgents, D.agentDim, D.r
    if A < 2:
        return torch.zeros_like(x)
    cut = 2 * r
    if D.training:
        cut = 2.2 * r if (A == 2 or D.kind == "Cross2D") else 3.2 * r
    xa = x.reshape(n, A, dim)
    diff = xa.unsqueeze(2) - xa.unsqueeze(1)                 # [n, i, j, dim] = x_i - x_j
    dist = torch.norm(diff, dim=3)
    e = torch.exp(-dist ** 2 / (2 * r * r)) * (dist < cut)
    e = e * (1 - torch.eye(A, dtype=x.dtype))
    return (-(e.unsqueeze(-1) * diff).sum(2) / (r * r)).reshape(n, d)


def prob_adjoint(D, x, g, abar, cL, cH):
    """psi = abar . (-grad_p H) + cL L + cH |g_t - H|  ->  (d psi / d g [n, D], direct d psi / d x [n, d])."""
    n, d = x.shape
    p, gt = g[:, :d], g[:, d]
    gbar = torch.zeros_like(g)
    if D.kind in ("Cross2D", "SwarmTraj"):
        aQ = D.alph_Q if (D.kind == "Cross2D" or D.alph_Q > 0) else 0.0
        from oracle import ocflow_oracle as orc
        L, H, _, _ = orc.lhqw(D, x, p)
        sg = torch.sign(gt - H.reshape(-1))
        k = (cH * sg).reshape(-1, 1)
        gbar[:, :d] = -abar + (cL - k) * p
        gbar[:, d] = k.reshape(-1)
        gx = aQ * terrain_grad(D, x)
        if D.alph_W != 0.0:
            gx = gx + D.alph_W * interaction_grad(D, x)
        return gbar, (cL + k) * gx
    # Quadcopter, one agent (Quadcopter.py:65-113)
    assert D.nAgents == 1
    m_, grav = D.mass, D.grav
    sps, cps = torch.sin(x[:, 3]), torch.cos(x[:, 3])
    sth, cth = torch.sin(x[:, 4]), torch.cos(x[:, 4])
    sph, cph = torch.sin(x[:, 5]), torch.cos(x[:, 5])
    F = torch.stack([sps * sph + cps * sth * cph, -cps * sph + sps * sth * cph, cth * cph], 1)
    dF = torch.zeros(n, 3, 3, dtype=x.dtype)                 # dF[:, c, a] = d F_c / d angle_a
    dF[:, 0, 0] = cps * sph - sps * sth * cph; dF[:, 0, 1] = cps * cth * cph; dF[:, 0, 2] = sps * cph - cps * sth * sph
    dF[:, 1, 0] = sps * sph + cps * sth * cph; dF[:, 1, 1] = sps * cth * cph; dF[:, 1, 2] = -cps * cph - sps * sth * sph
    dF[:, 2, 1] = -sth * cph; dF[:, 2, 2] = -cth * sph
    Pm = p[:, 6:9]
    fp = (F * Pm).sum(1)
    u = -fp / (2 * m_)
    sq = (p[:, 9:12] ** 2).sum(1)
    L = 2 + u * u + 0.25 * sq
    H = -L - (x[:, 6:9] * p[:, 0:3]).sum(1) - (x[:, 9:12] * p[:, 3:6]).sum(1) - (u / m_) * fp + grav * p[:, 8] + 0.5 * sq
    k = cH * torch.sign(gt - H)
    cLk = cL + k
    aF = (abar[:, 6:9] * F).sum(1)
    cfp = -cLk * u / m_ + 2 * k * u / m_ - aF / (2 * m_ * m_)
    xbar = torch.zeros_like(x)
    xbar[:, 6:12] = abar[:, 0:6]
    xbar[:, 6:9] += k.unsqueeze(1) * p[:, 0:3]
    xbar[:, 9:12] += k.unsqueeze(1) * p[:, 3:6]
    xbar[:, 3:6] = cfp.unsqueeze(1) * torch.einsum("nca,nc->na", dF, Pm) + (u / m_).unsqueeze(1) * torch.einsum("nca,nc->na", dF, abar[:, 6:9])
    gbar[:, 0:3] = k.unsqueeze(1) * x[:, 6:9]
    gbar[:, 3:6] = k.unsqueeze(1) * x[:, 9:12]
    gbar[:, 6:9] = cfp.unsqueeze(1) * F
    gbar[:, 8] -= k * grav
    gbar[:, 9:12] = -0.5 * abar[:, 9:12] - k.unsqueeze(1) * p[:, 9:12] + (0.5 * cLk).unsqueeze(1) * p[:, 9:12]
    gbar[:, d] = k
    return gbar, xbar


def manual_grad(x, P, D, tspan, nt, alph):
    """(sum over samples of the per-sample Jc, dict of d(sum Jc)/d theta, d(sum Jc)/dx [n,d]) for stepper 'rk4', nTh = 2."""
    from oracle import ocflow_oracle as orc
    n, d = x.shape
    dt = x.dtype
    times = orc.stage_time_table(tspan[0], tspan[1], nt)
    xs, z = [], torch.cat((x, torch.zeros(n, 4, dtype=dt)), 1)
    pad = lambda xx, t: torch.nn.functional.pad(xx, (0, 1), value=t)
    for k in range(nt):                                     # forward, keeping the four stage inputs of every step
        ta, tm, tb, _ = times[k]
        h = tb - ta
        st = []
        X = z[:, :d]
        k1 = h * orc.rhs(z, ta, P, D); st.append((X, ta))
        z2 = z + 0.5 * k1
        k2 = h * orc.rhs(z2, tm, P, D); st.append((z2[:, :d], tm))
        z3 = z + 0.5 * k2
        k3 = h * orc.rhs(z3, tm, P, D); st.append((z3[:, :d], tm))
        z4 = z + k3
        k4 = h * orc.rhs(z4, tb, P, D); st.append((z4[:, :d], tb))
        z = z + (1.0 / 6.0) * k1 + (2.0 / 6.0) * k2 + (2.0 / 6.0) * k3 + (1.0 / 6.0) * k4
        xs.append((st, h))
    xT = z[:, :d]
    res = xT - D.xtarget
    cG = 0.5 * (res ** 2).sum(1)
    sT = pad(xT, tspan[1])
    c = chain(P, sT)
    phiT = orc.phi_forward(P, sT).reshape(-1)
    J = z[:, d] + alph[0] * cG + alph[3] * z[:, d + 1] + alph[4] * (phiT - alph[0] * cG).abs() + alph[5] * (c["g"][:, :d] - alph[0] * res).abs().sum(1)
    G = zero_grads(P)
    # terminal block
    sf = torch.sign(phiT - alph[0] * cG)
    sgn = torch.sign(c["g"][:, :d] - alph[0] * res)
    gbar = torch.zeros(n, d + 1, dtype=dt)
    gbar[:, :d] = alph[5] * sgn
    beta = (alph[4] * sf).reshape(-1, 1)
    sbar = adjoint_eval(P, sT, c, gbar, beta, G)
    lam = sbar[:, :d] + (alph[0] * (1 - alph[4] * sf)).unsqueeze(1) * res - alph[5] * alph[0] * sgn
    wts = (1.0 / 6.0, 2.0 / 6.0, 2.0 / 6.0, 1.0 / 6.0)
    zero = torch.zeros(n, 1, dtype=dt)
    for k in range(nt - 1, -1, -1):
        st, h = xs[k]
        Xbar = [None] * 4
        for i in (3, 2, 1, 0):
            kbar = wts[i] * lam                              # adjoint of K_i (x part)
            if i == 2:
                kbar = kbar + Xbar[3]
            elif i < 2:
                kbar = kbar + 0.5 * Xbar[i + 1]
            X, t = st[i]
            s = pad(X, t)
            c = chain(P, s)
            gb, xdir = prob_adjoint(D, X, c["g"], h * kbar, h * wts[i] * 1.0, h * wts[i] * alph[3])
            sbar = adjoint_eval(P, s, c, gb, zero, G)
            Xbar[i] = sbar[:, :d] + xdir
        lam = lam + Xbar[0] + Xbar[1] + Xbar[2] + Xbar[3]
    return J.sum(), G, lam
