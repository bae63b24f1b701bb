// noc_baseline.cu — the baseline's discrete-control objective, batched over many initial states, with its gradient.
//
// Replaces loss_fun (baseline2D.py:42-63; timeBaseline.py:50-70 is the same function) for the Cross2D / SwarmTraj problems,
//     Z_{i+1} = Z_i + h U_i,   loss = sum_i h L(Z_{i+1}, U_i) + alpha_G G(Z_nt),   L from prob.calcLHQW(Z, U)   (h = 1/nt)
// and compute_loss / dyn (baselineQuad.py:40-72) for the quadcopter,
//     x_{i+1} = x_i + h dyn(c_i, x_i),   J = sum_i h (2 + |c_i|^2) + alpha_G 0.5 |x_nt - x_target|^2,
// and `err.backward()` of the baselines' optimisation loops (baseline2D.py:97-102, baselineQuad.py:80-86): d loss / d U.
// The reference optimises ONE initial state at a time in a Python loop; here a warp owns a sample (lanes over the state
// components / agents / pairs, the state and its adjoint in shared memory), so a comparison study over thousands of x0 is one
// launch.  The backward sweep re-reads the stored states Z_1..Z_nt (Cross2D / SwarmTraj) or x_0..x_{nt-1} (quadcopter).
#include "noc_launch.cuh"
#include "noc_adjoint.cuh"

namespace noc {

template <typename real>
struct BaseArgs {
    ProbPack prob;
    const real* U;      // [n][nt][nc]
    const real* z0;     // [n][d]
    long long n;
    int d, nt, nc;
    real alphG;
    real* loss;         // [n]
    real* gradU;        // [n][nt][nc] or NULL
    real* zsave;        // [warps of the grid][nt][d]: state history of the sample a warp is working on (gradU only)
};

template <typename real>
__device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

template <typename real>
__global__ void __launch_bounds__(128) baseline_loss_kernel(const BaseArgs<real> A) {
    extern __shared__ __align__(16) unsigned char base_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const ProbPack& pr = A.prob;
    const int d = A.d, nt = A.nt, nc = A.nc;
    real* Z = reinterpret_cast<real*>(base_smem) + warp * 2 * d;     // state
    real* Lm = Z + d;                                                 // adjoint of the state
    const real h = real(1.0 / nt);
    const real* xt = static_cast<const real*>(pr.xtarget);
    const int Ag = pr.nAgents, dim = pr.agentDim;
    const bool needQ = (pr.obstacle != 0) && (pr.kind == 0 || pr.alph_Q > 0.0) && pr.kind != 2;
    const bool needW = (pr.alph_W != 0.0) && Ag >= 2 && pr.kind != 2;
    const real aQ = needQ ? real(pr.alph_Q) : real(0), aW = real(pr.alph_W);
    const real cut = real(pr.cutW), c2 = real(2 * pr.r * pr.r), inv_r2 = real(1.0 / (pr.r * pr.r));
    for (long long smp = (long long)blockIdx.x * nwarp + warp; smp < A.n; smp += (long long)gridDim.x * nwarp) {
        const real* Us = A.U + smp * nt * nc;
        real* zs = A.gradU ? A.zsave + ((size_t)blockIdx.x * nwarp + warp) * nt * d : nullptr;   // per warp, reused by its next sample
        for (int c = lane; c < d; c += 32) Z[c] = A.z0[smp * d + c];
        __syncwarp();
        real loss = real(0);
        if (pr.kind == 2) {
            // ---------------- quadcopter (one agent): explicit Euler on dyn (baselineQuad.py:40-62)
            const real im = real(1.0 / pr.mass);
            for (int i = 0; i < nt; ++i) {
                const real c0 = Us[i * 4], c1 = Us[i * 4 + 1], c2_ = Us[i * 4 + 2], c3 = Us[i * 4 + 3];
                if (zs) for (int c = lane; c < 12; c += 32) zs[i * 12 + c] = Z[c];
                real sps, cps, sth, cth, sph, cph;
                r_sincos(Z[3], &sps, &cps); r_sincos(Z[4], &sth, &cth); r_sincos(Z[5], &sph, &cph);
                const real F[3] = {sps * sph + cps * sth * cph, -cps * sph + sps * sth * cph, cth * cph};
                real dx = real(0);
                if (lane < 6) dx = Z[6 + lane];
                else if (lane < 9) dx = (c0 * im) * F[lane - 6] - (lane == 8 ? real(pr.grav) : real(0));
                else if (lane < 12) dx = (lane == 9) ? c1 : (lane == 10 ? c2_ : c3);
                __syncwarp();
                if (lane < 12) Z[lane] = Z[lane] + h * dx;
                __syncwarp();
                loss = loss + h * (real(2) + (c0 * c0 + c1 * c1 + c2_ * c2_ + c3 * c3));
            }
            real r2 = real(0);
            if (lane < 12) { real res = Z[lane] - xt[lane]; r2 = res * res; Lm[lane] = A.alphG * res; }
            r2 = warp_sum(r2);
            loss = loss + A.alphG * real(0.5) * r2;
            if (lane == 0) A.loss[smp] = loss;
            if (!A.gradU) { __syncwarp(); continue; }
            __syncwarp();
            for (int i = nt - 1; i >= 0; --i) {
                real* gu = A.gradU + (smp * nt + i) * 4;
                const real c0 = Us[i * 4];
                const real x3 = zs[i * 12 + 3], x4 = zs[i * 12 + 4], x5 = zs[i * 12 + 5];
                real sps, cps, sth, cth, sph, cph;
                r_sincos(x3, &sps, &cps); r_sincos(x4, &sth, &cth); r_sincos(x5, &sph, &cph);
                const real F[3] = {sps * sph + cps * sth * cph, -cps * sph + sps * sth * cph, cth * cph};
                const real dF[3][3] = {{cps * sph - sps * sth * cph, cps * cth * cph, sps * cph - cps * sth * sph},
                                       {sps * sph + cps * sth * cph, sps * cth * cph, -cps * cph - sps * sth * sph},
                                       {real(0), -sth * cph, -cth * sph}};
                const real a6 = Lm[6], a7 = Lm[7], a8 = Lm[8];
                real add = real(0);
                if (lane >= 6 && lane < 12) add = h * Lm[lane - 6];
                else if (lane >= 3 && lane < 6) add = h * (c0 * im) * (dF[0][lane - 3] * a6 + dF[1][lane - 3] * a7 + dF[2][lane - 3] * a8);
                if (lane == 0) gu[0] = real(2) * h * c0 + h * im * (F[0] * a6 + F[1] * a7 + F[2] * a8);
                if (lane >= 1 && lane < 4) gu[lane] = real(2) * h * Us[i * 4 + lane] + h * Lm[8 + lane];
                __syncwarp();
                if (lane < 12) Lm[lane] = Lm[lane] + add;
                __syncwarp();
            }
            continue;
        }
        // ---------------- Cross2D / SwarmTraj: Z += h U_i, running cost at the NEW state (baseline2D.py:54-58)
        for (int i = 0; i < nt; ++i) {
            real uu = real(0);
            for (int c = lane; c < d; c += 32) {
                const real u = Us[i * nc + c];
                const real z = Z[c] + h * u;
                Z[c] = z;
                if (zs) zs[i * d + c] = z;
                uu = r_fma(u, u, uu);
            }
            __syncwarp();
            real q = real(0), w = real(0);
            if (needQ)
                for (int a = lane; a < Ag; a += 32)
                    q += terrain_agent<real>(pr, Z[a * dim], Z[a * dim + 1], dim == 3 ? Z[a * dim + 2] : real(0));
            if (needW) {
                const int npairs = Ag * (Ag - 1) / 2;
                for (int p = lane; p < npairs; p += 32) {
                    int a = 0, rem = p;
                    while (rem >= Ag - 1 - a) { rem -= Ag - 1 - a; ++a; }
                    const int b = a + 1 + rem;
                    real d2 = real(0);
                    for (int c = 0; c < dim; ++c) { real df = Z[a * dim + c] - Z[b * dim + c]; d2 = r_fma(df, df, d2); }
                    const real dd = r_sqrt(d2);
                    if (dd < cut) {
                        const real e = r_exp(-(dd * dd) / c2);
                        if (Ag == 2 || e != real(1)) w += e;
                    }
                }
            }
            uu = warp_sum(uu); q = warp_sum(q); w = warp_sum(w);
            real L;
            if (pr.kind == 0) L = real(0.5) * uu + real(pr.alph_Q) * q;
            else L = real(0.5) * uu + real(pr.alph_Q) * ((pr.alph_Q > 0.0) ? q : real(0));
            if (pr.alph_W != 0.0) L = L + aW * w;
            loss = loss + h * L;
        }
        real r2 = real(0);
        for (int c = lane; c < d; c += 32) { real res = Z[c] - xt[c]; r2 = r_fma(res, res, r2); Lm[c] = A.alphG * res; }
        r2 = warp_sum(r2);
        loss = loss + A.alphG * (real(0.5) * r2);
        if (lane == 0) A.loss[smp] = loss;
        __syncwarp();
        if (!A.gradU) continue;
        for (int i = nt - 1; i >= 0; --i) {
            for (int c = lane; c < d; c += 32) Z[c] = zs[i * d + c];
            __syncwarp();
            for (int a = lane; a < Ag; a += 32) {
                const real xi[3] = {Z[a * dim], Z[a * dim + 1], dim == 3 ? Z[a * dim + 2] : real(0)};
                real gi[3] = {real(0), real(0), real(0)};
                if (needQ) {
                    real gq[3];
                    terrain_agent_grad<real>(pr, xi[0], xi[1], xi[2], gq);
                    gi[0] = aQ * gq[0]; gi[1] = aQ * gq[1]; gi[2] = aQ * gq[2];
                }
                if (needW)
                    for (int b = 0; b < Ag; ++b) {
                        if (b == a) continue;
                        const real df[3] = {xi[0] - Z[b * dim], xi[1] - Z[b * dim + 1], dim == 3 ? xi[2] - Z[b * dim + 2] : real(0)};
                        const real d2 = r_fma(df[2], df[2], r_fma(df[1], df[1], df[0] * df[0]));
                        const real dd = r_sqrt(d2);
                        if (dd < cut) {
                            const real ce = aW * r_exp(-(dd * dd) / c2) * inv_r2;
                            gi[0] -= ce * df[0]; gi[1] -= ce * df[1]; gi[2] -= ce * df[2];
                        }
                    }
                for (int c = 0; c < dim; ++c) {
                    const int row = a * dim + c;
                    const real adj = Lm[row] + h * gi[c];                 // adjoint of Z_{i+1}
                    Lm[row] = adj;
                    A.gradU[(smp * nt + i) * nc + row] = h * Us[i * nc + row] + h * adj;
                }
            }
            __syncwarp();
        }
    }
}

template <typename real>
int baseline_loss(const ProbPack& pr, const real* U, const real* z0, long long n, int d, int nt, double alphG, real* loss, real* gradU,
                  cudaStream_t st) {
    BaseArgs<real> A;
    memset(&A, 0, sizeof A);
    A.prob = pr; A.U = U; A.z0 = z0; A.n = n; A.d = d; A.nt = nt; A.nc = (pr.kind == NOC_PROB_QUADCOPTER) ? 4 : d;
    A.alphG = (real)alphG; A.loss = loss; A.gradU = gradU;
    const int warps = 4;
    const size_t smem = sizeof(real) * 2 * d * warps;
    const int grid = std::max(1, (int)std::min<long long>((n + warps - 1) / warps, 32LL * sm_count()));
    ScratchBuf b_zsave;                                    // freed (stream-ordered) on every exit path
    if (gradU) {
        NOC_CUDA(b_zsave.alloc(sizeof(real) * (size_t)grid * warps * nt * d, st));
        A.zsave = b_zsave.as<real>();
    }
    baseline_loss_kernel<real><<<grid, 32 * warps, smem, st>>>(A);
    count_launch();
    NOC_CUDA(cudaGetLastError());
    return NOC_OK;
}

template int baseline_loss<float>(const ProbPack&, const float*, const float*, long long, int, int, double, float*, float*, cudaStream_t);
template int baseline_loss<double>(const ProbPack&, const double*, const double*, long long, int, int, double, double*, double*, cudaStream_t);

}  // namespace noc
