"""Host-side mirrors of the reference problem classes (src/problem/Cross2D.py, SwarmTraj.py, Quadcopter.py).

They carry exactly the attributes OCflow reads (`xtarget, d, nAgents, agentDim, alph_Q, alph_W, r,
obstacle, training`, plus `mass, grav`) and the `train()/eval()` switches.  calcLHQW / calcGradpH /
calcCtrls evaluate the device functors of the rollout kernel through noc_prob_eval — no torch
arithmetic.  The rollout accepts these objects or the reference's own (it duck-types by class name).
"""
import torch


class _Problem:
    agentDim = 0

    def __init__(self, xtarget, obstacle=None, alph_Q=1.0, alph_W=1.0, r=0.5):
        self.xtarget = xtarget.squeeze()
        self.d = xtarget.numel()
        self.obstacle = obstacle
        self.alph_Q = alph_Q
        self.alph_W = alph_W
        self.nAgents = self.d // self.agentDim
        self.r = r
        self.training = True

    def __repr__(self):
        return "%s Object" % type(self).__name__

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def _eval(self, x, p):
        from .ocflow import prob_eval
        return prob_eval(self, x, p)

    def calcLHQW(self, x, p):
        """-> L, H, Q, W, each [n,1]."""
        lhqw = self._eval(x, p)[0]
        return lhqw[:, 0:1], lhqw[:, 1:2], lhqw[:, 2:3], lhqw[:, 3:4]

    def calcGradpH(self, x, p):
        return self._eval(x, p)[1]

    def calcCtrls(self, x, p):
        return self._eval(x, p)[2]


class Cross2D(_Problem):
    """2-D agents; obstacles None | 'softcorridor' | 'hardcorridor' (src/problem/Cross2D.py:10-165)."""
    agentDim = 2


class SwarmTraj(_Problem):
    """3-D agents; obstacles None | 'blocks' (src/problem/SwarmTraj.py:12-167)."""
    agentDim = 3


class Quadcopter(_Problem):
    """12-D quadcopter agents (src/problem/Quadcopter.py:7-197)."""
    agentDim = 12

    def __init__(self, xtarget, obstacle=None, alph_Q=1.0, alph_W=1.0, mass=1.0, grav=9.81, r=1.0):
        super().__init__(xtarget, obstacle, alph_Q, alph_W, r)
        self.mass = mass
        self.grav = grav
