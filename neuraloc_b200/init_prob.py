"""Problem factory with the reference's signature: initProb(sData, nTrain, nVal, var0, alph, cvt)
(src/initProb.py:9-249) and resample (:252-262).  Host-side and O(1); it builds the problem object,
the centre xInit of rho_0 and Gaussian samples around it.  Target / initial-centre tables are data
of the reference's experiments (cited per entry)."""
import torch
from torch.nn.functional import pad

from .problems import Cross2D, Quadcopter, SwarmTraj

_SWARM_ROWS = {   # initProb.py:37-52 (swarm) and :81-105 (swarm50): upper formation; the lower one is shifted by (0,-0.5,-3)
    "swarm": ([[-2, 2], [-1, 2], [0, 2], [1, 2], [2, 2], [-2.5, 3], [-1.5, 3], [-.5, 3], [.5, 3], [1.5, 3], [2.5, 3],
               [-2, 4], [-1, 4], [0, 4], [1, 4], [2, 4]], [8.0] * 16, 0.2),
    "swarm50": ([[-2, 2], [-1, 2], [0, 2], [1, 2], [2, 2], [3, 2], [4, 2],
                 [-2.5, 3], [-1.5, 3], [-.5, 3], [.5, 3], [1.5, 3], [2.5, 3], [3.5, 3],
                 [-2, 4], [-1, 4], [0, 4], [1, 4], [2, 4], [3, 4], [4, 4], [-2, 3], [-1, 3], [1, 3], [2, 3]],
                [6.0] * 7 + [7.0] * 7 + [8.0] * 7 + [5.0] * 4, 0.1),
}
_SWAP12 = dict(   # initProb.py:196-203
    target=[2, 2, 0, 0, 10, 0, -10, 0, 5, 5, -5, -5, -4, 2, -6, -1, 5, -5, -5, 5, 2, -2, -2, -2],
    init=[0, 0, 2, 2, -10, 0, 10, 0, -5, -5, 5, 5, -6, -1, -4, 2, -5, 5, 5, -5, -2, -2, 2, -2])


def _centres(sData):
    """-> (problem class, kwargs without alph, xtarget 1-D float64, xInit 1-D float64)."""
    t = lambda v: torch.tensor(v, dtype=torch.float64)
    if sData == "softcorridor":                                            # :25-31
        return Cross2D, dict(obstacle="softcorridor", r=0.5), t([2, 2, -2, 2]), t([-2, -2, 2, -2])
    if sData in _SWARM_ROWS:                                               # :33-125
        xy, zz, r = _SWARM_ROWS[sData]
        top = torch.cat((t(xy), t(zz).view(-1, 1)), 1)
        tg = torch.cat((top, top + t([0, -0.5, -3])), 0)
        xi = t([1, -1, -1]) * tg + t([0, 0, 10])
        return SwarmTraj, dict(obstacle="blocks", r=r), tg.reshape(-1), xi.reshape(-1)
    if sData == "singlequad":                                              # :127-142
        return Quadcopter, dict(obstacle=None), t([2.0] * 3 + [0.0] * 9), t([-1.5] * 3 + [0.0] * 9)
    if sData == "midcross2":                                               # :144-151 (default r)
        return Cross2D, dict(obstacle=None), t([2, 2, -2, 2]), t([-2, -2, 2, -2])
    if sData in ("midcross4", "midcross20", "midcross30"):                 # :153-187
        A, lim, r = {"midcross4": (4, 2.0, 0.4), "midcross20": (20, 6.0, 0.15), "midcross30": (30, 6.0, 0.2)}[sData]
        xx = torch.linspace(-lim, lim, A).double()
        lvl = t([6.0, 4.0, 2.0]).repeat(A // 3) if sData == "midcross30" else lim * torch.ones(A, dtype=torch.float64)
        return (Cross2D, dict(obstacle=None, r=r), torch.stack((xx.flip(0), lvl), 1).reshape(-1),
                torch.stack((xx, -lvl), 1).reshape(-1))
    if sData == "swap2":                                                   # :188-195
        return Cross2D, dict(obstacle="hardcorridor", r=1.0), t([10, 0, -10, 0]), t([-10, 0, 10, 0])
    if sData == "swap12" or (sData.startswith("swap12_") and sData.endswith("pair")):   # :196-243
        A = 12 if sData == "swap12" else 2 * int(sData[7:-4])
        return Cross2D, dict(obstacle=None, r=0.5), t(_SWAP12["target"][:2 * A]), t(_SWAP12["init"][:2 * A])
    raise ValueError("incorrect value passed to --data: %r" % (sData,))   # the reference prints and exit(1)s (:244-246)


def initProb(sData, nTrain, nVal, var0, alph, cvt):
    """-> prob, x0 [nTrain,d], x0v, xInit [1,d]; sampling follows src/initProb.py branch by branch."""
    cls, kw, tg, xi = _centres(sData)
    xtarget, xInit = cvt(tg), cvt(xi).view(1, -1)
    d = tg.numel()
    if cls is SwarmTraj:                                                   # half around xInit, half around the target (:107-123)
        half = nTrain // 2
        x0 = torch.cat((xInit + cvt(var0 * torch.randn(half, d)), xtarget + cvt(var0 * torch.randn(half, d))), 0)
        x0v = xInit + cvt(var0 * torch.randn(half, d))
        prob = SwarmTraj(xtarget, alph_Q=alph[1], alph_W=alph[2], **kw)
    elif cls is Quadcopter:                                                # noise on the position only (:132-140)
        x0 = pad(xInit[:, :3] + cvt(var0 * torch.randn(nTrain, 3)), [0, d - 3, 0, 0], value=0)
        x0v = pad(cvt(xi[:3] + var0 * torch.randn(nVal, 3)), [0, d - 3, 0, 0], value=0)
        prob = Quadcopter(xtarget, obstacle=None, alph_Q=0.0, alph_W=0.0)
    else:
        x0 = xInit + cvt(var0 * torch.randn(nTrain, d))
        nv = nTrain if sData in ("softcorridor", "midcross2") else nVal   # those two branches draw nTrain (:29,:149)
        x0v = xInit + cvt(var0 * torch.randn(nv, d))
        prob = Cross2D(xtarget, alph_Q=alph[1], alph_W=alph[2], **kw)
    return prob, x0, x0v, xInit


def resample(x0, xInit, var0, cvt):
    """src/initProb.py:252-262."""
    return xInit + cvt(var0 * torch.randn(*x0.shape))
