// noc_adjoint.cuh — derivatives of the problem functors shared by the training kernel (noc_grad.cu) and the baseline objective
// (noc_baseline.cu).
#pragma once
#include "noc_rollout.cuh"

namespace noc {

// gradient of the per-agent terrain cost (the value is terrain_agent's); zero where the reference's mask cuts the Gaussians off
// and in eval mode (inside-counts).  Cross2D.py:90-119, SwarmTraj.py:90-122, utils.py:70-86.
template <typename real>
__device__ void terrain_agent_grad(const ProbPack& pr, real x0, real x1, real x2, real (&gq)[3]) {
    gq[0] = gq[1] = gq[2] = real(0);
    if (pr.obstacle == 1) {
        const real c = real(0.2), ic = real(1) / c;
        const real mus[4] = {real(-2.5), real(2.5), real(-1.5), real(1.5)};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            real pdf = gauss2<real>(x0, x1, mus[q], real(0), c, c);
            gq[0] -= pdf * (x0 - mus[q]) * ic;
            gq[1] -= pdf * x1 * ic;
        }
        return;
    }
    if (!pr.training) return;
    if (pr.obstacle == 2) {
        real d1 = r_sqrt(x0 * x0 + (x1 - real(4)) * (x1 - real(4)));
        real d2 = r_sqrt(x0 * x0 + (x1 + real(3.5)) * (x1 + real(3.5)));
        real thr = real(2.0 + pr.r);
        if (!(d1 < thr || d2 < thr)) return;
        real p1 = gauss2<real>(x0, x1, real(0), real(4), real(1), real(1));
        real p2 = gauss2<real>(x0, x1, real(0), real(-3.5), real(1), real(1));
        gq[0] = -(p1 + p2) * x0;
        gq[1] = -p1 * (x1 - real(4)) - p2 * (x1 + real(3.5));
        return;
    }
    if (pr.obstacle == 3) {
        double r = pr.r;
        bool in = (x0 < real(2.0 + r) && x0 > real(-2.0 - r) && x1 < real(0.5 + r) && x1 > real(-0.5 - r) && x2 < real(7.0 + r)) ||
                  (x0 < real(4.0 + r) && x0 > real(2.0 - r) && x1 < real(1.0 + r) && x1 > real(-1.0 - r) && x2 < real(4.0 + r));
        if (!in) return;
        real p1 = gauss3<real>(x0, x1, x2, real(0), real(0), real(2), real(9), real(3), real(9));
        real p2 = gauss3<real>(x0, x1, x2, real(2.5), real(0), real(2), real(9), real(3), real(3));
        gq[0] = -p1 * x0 * real(1.0 / 9.0) - p2 * (x0 - real(2.5)) * real(1.0 / 9.0);
        gq[1] = -p1 * x1 * real(1.0 / 3.0) - p2 * x1 * real(1.0 / 3.0);
        gq[2] = -p1 * (x2 - real(2)) * real(1.0 / 9.0) - p2 * (x2 - real(2)) * real(1.0 / 3.0);
    }
}

}  // namespace noc
