// noc_tc_quad.cu — tensor-core (tcgen05 / TMEM) rollout kernel for the m = 128 class of value networks
// (singlequad: d = 12, m = 128, nTh = 2), fp32 in / fp32 out.
//
// One CTA of 128 threads owns a tile of 128 samples; THREAD r IS SAMPLE r: it holds the augmented state and the RK
// accumulators in registers for all nt steps, evaluates the quadcopter terms in registers, and is the epilogue thread of
// TMEM lane r.  The four contractions of a grad-Phi evaluation (Phi.py:99-138) run on the 5th-generation tensor cores
// with M = 128 samples:
//     GEMM-1  O  [128 x m]  = S [128 x 16] . K0b'        S = [x, t, 1, 0..]  (the 1-column folds the bias b0 in)
//     GEMM-2  A1 [128 x m]  = U0 [128 x m] . K1'         B = K1 read K-major
//     GEMM-3  Z1 [128 x m]  = Y  [128 x m] . K1          B = the SAME K1 buffer read MN-major
//     GEMM-4  G  [128 x 16] = V  [128 x m] . K0b (MN-major view of the GEMM-1 buffer)  +  S . symb'  (A'A and c_w)
// Accumulators live in TMEM (fp32); tanh(o) is parked in TMEM columns [128, 128+m) between GEMM-1 and GEMM-3.
// Precision: every fp32 operand is split into three bf16 terms (hi, mid, lo) and each logical product is six MMAs
// (hh, hm, mh, hl, mm, lh) -- the dropped terms are O(2^-24) -- because a single bf16 / tf32 pass breaks the 1e-5
// per-step-state tolerance.  Operands are written by the epilogue threads straight into the canonical
// no-swizzle UMMA layout (8-row x 16-byte core matrices), so no TMA and no extra pass is needed; one thread issues the
// MMAs and tcgen05.commit signals an mbarrier the 128 epilogue threads wait on.
#include "noc_launch.cuh"
#include "noc_tc.cuh"

namespace noc {

struct TcArgs {
    int d, m;                       // D = d + 1 <= 14, m % 16 == 0, m <= 128
    float h;
    const float *K0, *b0, *K1, *b1, *w, *A, *c_w, *c_b;   // reference layout (fp32, device)
    int r;
    ProbPack prob;
    const float* x;
    long long n;
    int nt, stepper, mode;
    const double* times;
    float alph0, alph3, alph4, alph5, t_end;
    double* partials;
    float* out_a; float* out_b; float* out_c;
    int ntiles;
};

// fp32 pair -> three packed bf16x2 terms, v ~ hi + mid + lo (exact to ~2^-24 |v|); registers only
__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);          // .x = a (low half), .y = b
    return *reinterpret_cast<unsigned*>(&v);
}
__device__ __forceinline__ float bf_lo(unsigned p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(unsigned p) { return __uint_as_float(p & 0xffff0000u); }
__device__ __forceinline__ void split3x2(float a, float b, unsigned& hi, unsigned& mid, unsigned& lo) {
    hi = pack_bf16x2(a, b);
    const float ra = a - bf_lo(hi), rb = b - bf_hi(hi);
    mid = pack_bf16x2(ra, rb);
    lo = pack_bf16x2(ra - bf_lo(mid), rb - bf_hi(mid));
}

// write 8 consecutive K-elements (one 16-byte chunk) of row `row` of an A/B operand, all three split planes
__device__ __forceinline__ void store_chunk3(unsigned char* base, int plane_bytes, int row, int k0, int K, const float (&v)[8]) {
    uint4 h, mi, l;
    split3x2(v[0], v[1], h.x, mi.x, l.x);
    split3x2(v[2], v[3], h.y, mi.y, l.y);
    split3x2(v[4], v[5], h.z, mi.z, l.z);
    split3x2(v[6], v[7], h.w, mi.w, l.w);
    const int off = il_off(row, k0, K);
    *reinterpret_cast<uint4*>(base + off) = h;
    *reinterpret_cast<uint4*>(base + plane_bytes + off) = mi;
    *reinterpret_cast<uint4*>(base + 2 * plane_bytes + off) = l;
}
__device__ __forceinline__ void load_chunk3(const unsigned char* base, int plane_bytes, int row, int k0, int K, float (&v)[8]) {
    const int off = il_off(row, k0, K);
    const uint4 a = *reinterpret_cast<const uint4*>(base + off), b = *reinterpret_cast<const uint4*>(base + plane_bytes + off),
                c = *reinterpret_cast<const uint4*>(base + 2 * plane_bytes + off);
    const unsigned pa[4] = {a.x, a.y, a.z, a.w}, pb[4] = {b.x, b.y, b.z, b.w}, pc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = (bf_lo(pa[i]) + bf_lo(pb[i])) + bf_lo(pc[i]);
        v[2 * i + 1] = (bf_hi(pa[i]) + bf_hi(pb[i])) + bf_hi(pc[i]);
    }
}

// one logical fp32 product block: six bf16 MMAs over the three planes of A and B (same descriptor geometry per plane).
// The hi*hi products go to `d_main`, the five correction products (2^-8 .. 2^-16 of the main term) to `d_corr`: the
// tensor core's fp32 accumulator truncates on every add, and keeping the small terms in their own accumulator keeps
// that (biased) error 2^-8 smaller for them -- the epilogue adds the two in round-to-nearest fp32.
__device__ __forceinline__ void mma6(unsigned d_main, unsigned d_corr, unsigned a_addr, int a_plane, unsigned a_lbo, unsigned a_sbo,
                                     unsigned b_addr, int b_plane, unsigned b_lbo, unsigned b_sbo, unsigned idesc, int accumulate) {
    umma_bf16(d_main, umma_desc(a_addr, a_lbo, a_sbo), umma_desc(b_addr, b_lbo, b_sbo), idesc, accumulate);                       // hh
    umma_bf16(d_corr, umma_desc(a_addr + 2 * a_plane, a_lbo, a_sbo), umma_desc(b_addr, b_lbo, b_sbo), idesc, accumulate);         // lh
    umma_bf16(d_corr, umma_desc(a_addr, a_lbo, a_sbo), umma_desc(b_addr + 2 * b_plane, b_lbo, b_sbo), idesc, 1);                  // hl
    umma_bf16(d_corr, umma_desc(a_addr + a_plane, a_lbo, a_sbo), umma_desc(b_addr + b_plane, b_lbo, b_sbo), idesc, 1);            // mm
    umma_bf16(d_corr, umma_desc(a_addr + a_plane, a_lbo, a_sbo), umma_desc(b_addr, b_lbo, b_sbo), idesc, 1);                      // mh
    umma_bf16(d_corr, umma_desc(a_addr, a_lbo, a_sbo), umma_desc(b_addr + b_plane, b_lbo, b_sbo), idesc, 1);                      // hm
}

__global__ void __launch_bounds__(128, 1) rollout_tc_quad_kernel(const TcArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int d = 12, D = 13;            // one quadcopter: s = [x(12), t]
    const int m = A.m;
    const ProbPack& pr = A.prob;
    // shared-memory map (bytes); every operand has three planes (hi, mid, lo)
    const int pK1 = m * m * 2, pK0 = m * 16 * 2, pSy = 16 * 16 * 2, pX = 128 * m * 2, pS = 128 * 16 * 2;
    unsigned char* sK1 = smem;
    unsigned char* sK0 = sK1 + 3 * pK1;
    unsigned char* sSy = sK0 + 3 * pK0;
    unsigned char* sX = sSy + 3 * pSy;
    unsigned char* sS = sX + 3 * pX;
    float* sb1 = reinterpret_cast<float*>(sS + 3 * pS);
    float* sw = sb1 + m;
    float* scw = sw + m;                                 // 16 floats
    float* sred = scw + 16;                              // 4 warps x 8
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ unsigned tmem_base_s;

    // ---- one-time per CTA: weights -> split bf16 operands in canonical layout
    for (int i = tid; i < m * (m / 8); i += 128) {       // K1[o][k0..k0+8)
        const int o = i / (m / 8), k0 = (i % (m / 8)) * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = A.K1[o * m + k0 + e];
        store_chunk3(sK1, pK1, o, k0, m, v);
    }
    for (int i = tid; i < m * 2; i += 128) {             // K0b[j][k]: K0 | b0 | 0
        const int j = i / 2, k0 = (i % 2) * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { const int k = k0 + e; v[e] = (k < D) ? A.K0[j * D + k] : (k == D ? A.b0[j] : 0.f); }
        store_chunk3(sK0, pK0, j, k0, 16, v);
    }
    for (int i = tid; i < 16 * 2; i += 128) {            // symb[n = k'][k]: (A'A)[k][k'] | c_w[k'] in column D | 0
        const int kp = i / 2, k0 = (i % 2) * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = k0 + e;
            float s = 0.f;
            if (kp < D && k < D) { for (int q = 0; q < A.r; ++q) s = fmaf(A.A[q * D + k], A.A[q * D + kp], s); }
            else if (kp < D && k == D) s = A.c_w[kp];
            v[e] = s;
        }
        store_chunk3(sSy, pSy, kp, k0, 16, v);
    }
    for (int i = tid; i < m; i += 128) { sb1[i] = A.b1[i]; sw[i] = A.w[i]; }
    if (tid < 16) scw[tid] = (tid < D) ? A.c_w[tid] : 0.f;
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    if (tid == 0) mbar_init(smem_u32(&mbar), 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tacc = tmem_base_s;                   // accumulator columns [0, m); GEMM-4 uses [0,16) and [16,32)
    const unsigned tT0 = tacc + 128;                     // tanh(o) columns
    const unsigned tcor = tacc + 256;                    // correction-term accumulator, same column map as tacc
    const unsigned lane_bits = (unsigned)(warp * 32) << 16;
    const unsigned mb = smem_u32(&mbar);
    int phase = 0;
    const unsigned idesc_m_k = umma_idesc_bf16(128, m, 0), idesc_m_mn = umma_idesc_bf16(128, m, 1);
    const unsigned idesc_16_mn = umma_idesc_bf16(128, 16, 1), idesc_16_k = umma_idesc_bf16(128, 16, 0);
    const unsigned aX = smem_u32(sX), aS = smem_u32(sS), aK1 = smem_u32(sK1), aK0 = smem_u32(sK0), aSy = smem_u32(sSy);
    const unsigned sboM = (unsigned)(m >> 3) * 128;      // 8-row-group stride of an operand with K = m

    // publish my operand writes, let thread 0 issue `issue`, wait for the tensor core
    auto run_mma = [&](auto issue) {
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue();
            umma_commit(mb);
        }
        mbar_wait(mb, phase);
        phase ^= 1;
        tc_fence_after();
    };

    // grad Phi at s = [xs, t]: returns g[0..16) in registers; terminal: also Phi(s)
    auto chain = [&](const float (&xs)[12], float t, float (&g)[16], bool terminal, float& phi_out) {
        {   // S operand row: [x(12), t, 1, 0, 0]
            float v0[8], v1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v0[e] = xs[e];
            v1[0] = xs[8]; v1[1] = xs[9]; v1[2] = xs[10]; v1[3] = xs[11]; v1[4] = t; v1[5] = 1.f; v1[6] = 0.f; v1[7] = 0.f;
            store_chunk3(sS, pS, tid, 0, 16, v0);
            store_chunk3(sS, pS, tid, 8, 16, v1);
        }
        run_mma([&] { mma6(tacc, tcor, aS, pS, 128, 256, aK0, pK0, 128, 256, idesc_m_k, 0); });                 // GEMM-1 (K = 16)
        for (int c0 = 0; c0 < m; c0 += 32) {             // u0 = act(o) -> X operand, tanh(o) -> TMEM
            float v[32], tt[32];
            tmem_ld32_sum(tacc + lane_bits + c0, tcor + lane_bits + c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) act_tanh(v[i], v[i], tt[i]);
            tmem_st32(tT0 + lane_bits + c0, tt);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float c8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) c8[e] = v[q * 8 + e];
                store_chunk3(sX, pX, tid, c0 + q * 8, m, c8);
            }
        }
        run_mma([&] {                                    // GEMM-2: A1 = U0 . K1'  (B K-major)
            for (int kb = 0; kb < m / 16; ++kb)
                mma6(tacc, tcor, aX + kb * 256, pX, 128, sboM, aK1 + kb * 256, pK1, 128, sboM, idesc_m_k, kb > 0);
        });
        float phiN = 0.f;
        for (int c0 = 0; c0 < m; c0 += 32) {             // y = tanh(a1 + b1) * w -> X operand
            float v[32];
            tmem_ld32_sum(tacc + lane_bits + c0, tcor + lane_bits + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float c8[8], u8[8];
                if (terminal) load_chunk3(sX, pX, tid, c0 + q * 8, m, u8);      // u0, before it is overwritten
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int col = c0 + q * 8 + e;
                    const float pre = v[q * 8 + e] + sb1[col], wv = sw[col];
                    if (terminal) {
                        float av, tv;
                        act_tanh(pre, av, tv);
                        phiN = fmaf(wv, u8[e] + A.h * av, phiN);
                        c8[e] = tv * wv;
                    } else {
                        c8[e] = tanh_only(pre) * wv;
                    }
                }
                store_chunk3(sX, pX, tid, c0 + q * 8, m, c8);
            }
        }
        run_mma([&] {                                    // GEMM-3: Z1 = Y . K1  (the same buffer, MN-major)
            for (int kb = 0; kb < m / 16; ++kb)
                mma6(tacc, tcor, aX + kb * 256, pX, 128, sboM, aK1 + kb * 2 * sboM, pK1, sboM, 128, idesc_m_mn, kb > 0);
        });
        for (int c0 = 0; c0 < m; c0 += 32) {             // v = tanh(o) * (w + h z1acc) -> X operand
            float v[32], tt[32];
            tmem_ld32_sum(tacc + lane_bits + c0, tcor + lane_bits + c0, v);
            tmem_ld32(tT0 + lane_bits + c0, tt);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float c8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { const int i = q * 8 + e; c8[e] = tt[i] * (sw[c0 + i] + A.h * v[i]); }
                store_chunk3(sX, pX, tid, c0 + q * 8, m, c8);
            }
        }
        run_mma([&] {                                    // GEMM-4: [0,16) = V . K0b (MN-major view), [16,32) = S . symb'
            for (int kb = 0; kb < m / 16; ++kb)
                mma6(tacc, tcor, aX + kb * 256, pX, 128, sboM, aK0 + kb * 512, pK0, 256, 128, idesc_16_mn, kb > 0);
            mma6(tacc + 16, tcor + 16, aS, pS, 128, 256, aSy, pSy, 128, 256, idesc_16_k, 0);
        });
        float gw[16], gq[16];
        {
            float v[32], c[32];                           // [0,16) = V.K0b, [16,32) = S.symb'; main and correction accumulators
            tmem_ld32(tacc + lane_bits, v);
            tmem_ld32(tcor + lane_bits, c);
#pragma unroll
            for (int i = 0; i < 16; ++i) { gw[i] = v[i] + c[i]; gq[i] = v[16 + i] + c[16 + i]; }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) g[i] = gw[i] + gq[i];                 // gq already carries c_w (column D of symb)
        if (terminal) {                                   // Phi = w.u1 + 0.5 s'A'A s + c_w.s + c_b  (Phi.py:96)
            float quad = 0.f, lin = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float sv = (k < d) ? xs[k < d ? k : 0] : t;
                quad = fmaf(sv, gq[k] - scw[k], quad);
                lin = fmaf(scw[k], sv, lin);
            }
            phi_out = phiN + 0.5f * quad + (lin + A.c_b[0]);
        }
    };

    // Quadcopter terms in registers (Quadcopter.py:65-113, 160-197), single agent
    auto quad_terms = [&](const float (&x)[12], const float (&g)[16], float (&dx)[12], float& L, float& HJ, float& uctrl) {
        float sps, cps, sth, cth, sph, cph;
        sincosf(x[3], &sps, &cps); sincosf(x[4], &sth, &cth); sincosf(x[5], &sph, &cph);
        const float f7 = sps * sph + cps * sth * cph, f8 = -cps * sph + sps * sth * cph, f9 = cth * cph;
        const float fp = f7 * g[6] + f8 * g[7] + f9 * g[8];
        const float u = float(-1.0 / (2.0 * pr.mass)) * fp;
        const float sq = g[9] * g[9] + g[10] * g[10] + g[11] * g[11];
        const float Q = 0.f;
        L = float(pr.alph_Q) * Q;
        L = L + 2.f + u * u + 0.25f * sq;
        const float um = u / float(pr.mass);
        const float xv = x[6] * g[0] + x[7] * g[1] + x[8] * g[2];
        const float xw = x[9] * g[3] + x[10] * g[4] + x[11] * g[5];
        const float H = 0.f - L - xv - xw - um * fp + float(pr.grav) * g[8] + 0.5f * sq;
        HJ = fabsf(g[12] - H);
#pragma unroll
        for (int c = 0; c < 6; ++c) dx[c] = x[6 + c];
        dx[6] = -(-um * f7); dx[7] = -(-um * f8); dx[8] = -(-um * f9 + float(pr.grav));
#pragma unroll
        for (int c = 9; c < 12; ++c) dx[c] = -(0.5f * g[c]);
        uctrl = u;
    };

    double csum[7] = {0, 0, 0, 0, 0, 0, 0};
    long long cnt = 0;
    const int nstage = (A.stepper == 4) ? 4 : (A.stepper == 1 ? 1 : 0);
    const bool inter = (A.mode == 2);
    const int ntp1 = A.nt + 1;
    for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
        const long long s0 = (long long)tile * 128;
        const int nvalid = (int)((A.n - s0 < 128) ? (A.n - s0) : 128);
        const bool valid = tid < nvalid;
        const long long gs = s0 + (valid ? tid : nvalid - 1);
        float z0[16], za[16];
#pragma unroll
        for (int c = 0; c < 12; ++c) z0[c] = A.x[gs * 12 + c];
#pragma unroll
        for (int c = 12; c < 16; ++c) z0[c] = 0.f;
        if (inter && valid) {
#pragma unroll
            for (int c = 0; c < 16; ++c) A.out_b[(gs * 16 + c) * ntp1] = z0[c];
#pragma unroll
            for (int c = 0; c < 4; ++c) A.out_c[(gs * 4 + c) * ntp1] = 0.f;
        }
        // ONE call site for the chain: the nt * (stages [+ 1 control evaluation]) + 1 terminal evaluations of a tile are a
        // flat sequence; `xs` always holds the next evaluation's input.
        float g[16], dx[12], xs[12], phi1 = 0.f;
#pragma unroll
        for (int c = 0; c < 12; ++c) xs[c] = z0[c];
        const int per = nstage + (inter ? 1 : 0);
        const int total = A.nt * per + 1;
        for (int it = 0; it < total; ++it) {
            const bool term = (it == total - 1);
            const int k = term ? 0 : it / per, st = term ? 0 : it % per;
            const bool ctl = !term && st == nstage;                 // control evaluation after the step (OCflow.py:51-55)
            const double* tt = A.times + 5 * k;
            const float hstep = float(tt[4]);                       // h = t1 - t0 recomputed per step (OCflow.py:169)
            float wgt = 1.f, cnext = 0.f, tcur = float(tt[0]);
            if (term) tcur = A.t_end;
            else if (ctl) tcur = float(tt[3]);                      // new state, OLD time (quirk 3)
            else if (nstage == 4) {                                 // RK4 weights (OCflow.py:172-182)
                if (st == 0) { wgt = float(1.0 / 6.0); cnext = 0.5f; }
                else if (st == 1) { wgt = float(2.0 / 6.0); cnext = 0.5f; tcur = float(tt[1]); }
                else if (st == 2) { wgt = float(2.0 / 6.0); cnext = 1.0f; tcur = float(tt[1]); }
                else { wgt = float(1.0 / 6.0); tcur = float(tt[2]); }
            }
            chain(xs, tcur, g, term, phi1);
            if (term) break;
            float L, HJ, uc;
            quad_terms(xs, g, dx, L, HJ, uc);
            if (ctl) {
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) A.out_b[(gs * 16 + c) * ntp1 + (k + 1)] = z0[c];
                    A.out_c[(gs * 4 + 0) * ntp1 + (k + 1)] = uc;
#pragma unroll
                    for (int c = 1; c < 4; ++c) A.out_c[(gs * 4 + c) * ntp1 + (k + 1)] = -0.5f * g[8 + c];
                }
                continue;
            }
            float kk[16];
#pragma unroll
            for (int c = 0; c < 12; ++c) kk[c] = hstep * dx[c];
            kk[12] = hstep * L; kk[13] = hstep * HJ; kk[14] = hstep * 0.f; kk[15] = hstep * 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) za[c] = ((st == 0) ? z0[c] : za[c]) + wgt * kk[c];
            if (st != nstage - 1) {
#pragma unroll
                for (int c = 0; c < 12; ++c) xs[c] = z0[c] + cnext * kk[c];
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) z0[c] = za[c];
#pragma unroll
                for (int c = 0; c < 12; ++c) xs[c] = z0[c];
            }
        }
        // terminal block (OCflow.py:58-90): xs = x(T), g = grad Phi(x(T), T), phi1 = Phi(x(T), T)
        float xT[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) xT[c] = xs[c];
        const float* xt = static_cast<const float*>(pr.xtarget);
        float cG = 0.f, hjg = 0.f;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const float res = xT[c] - xt[c];
            cG = fmaf(res, res, cG);
            hjg += fabsf(g[c] - A.alph0 * res);
        }
        cG *= 0.5f;
        const float cost[7] = {z0[12], cG, z0[13], fabsf(phi1 - A.alph0 * cG), hjg, z0[14], z0[15]};
        if (A.mode == 0) {
            // deterministic CTA sum: lanes -> warp (shuffle tree), warps -> thread 0 in fixed order
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                float v = valid ? cost[q] : 0.f;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if ((tid & 31) == 0) sred[warp * 8 + q] = v;
            }
            __syncthreads();
            if (tid == 0) {
                for (int q = 0; q < 7; ++q) csum[q] += (double)sred[q] + (double)sred[8 + q] + (double)sred[16 + q] + (double)sred[24 + q];
                cnt += nvalid;
            }
            __syncthreads();
        } else if (A.mode == 1 && valid) {
            float* o = A.out_a + gs * 8;
            o[0] = cost[0] + A.alph0 * cost[1] + A.alph3 * cost[2] + A.alph4 * cost[3] + A.alph5 * cost[4];
#pragma unroll
            for (int q = 0; q < 7; ++q) o[1 + q] = cost[q];
        }
    }
    if (A.mode == 0 && A.partials && tid == 0) {
        for (int q = 0; q < 7; ++q) A.partials[blockIdx.x * 8 + q] = csum[q];
        A.partials[blockIdx.x * 8 + 7] = (double)cnt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tacc, 512);
}

// Host side: eligibility is decided by the caller (noc_api.cu): fp32, Quadcopter with one agent, nTh = 2, m % 16 == 0, m <= 128.
int tc_quad_rollout(int d, int m, int r, double h, const PhiRaw<float>& raw, const ProbPack& pr, const float* x, long long n,
                    const double* dtimes, int nt, int stepper, int mode, const double* alph, double t_end, double* out_sums,
                    float* out_nomean, float* zFull, float* ctrlFull, int smem_limit, cudaStream_t st) {
    TcArgs A;
    memset(&A, 0, sizeof A);
    A.d = d; A.m = m; A.h = (float)h; A.r = r;
    A.K0 = raw.K[0]; A.b0 = raw.b[0]; A.K1 = raw.K[1]; A.b1 = raw.b[1]; A.w = raw.w; A.A = raw.A; A.c_w = raw.c_w; A.c_b = raw.c_b;
    A.prob = pr; A.x = x; A.n = n; A.nt = nt; A.stepper = stepper; A.mode = mode; A.times = dtimes;
    A.alph0 = (float)alph[0]; A.alph3 = (float)alph[3]; A.alph4 = (float)alph[4]; A.alph5 = (float)alph[5];
    A.t_end = (float)t_end;
    A.out_a = out_nomean; A.out_b = zFull; A.out_c = ctrlFull;
    A.ntiles = (int)((n + 127) / 128);
    const size_t smem = 3 * ((size_t)m * m * 2 + (size_t)m * 16 * 2 + 16 * 16 * 2 + (size_t)128 * m * 2 + 128 * 16 * 2) +
                        sizeof(float) * (2 * (size_t)m + 16 + 32) + 1024;
    if (smem > (size_t)smem_limit) return fail(NOC_ERR_NOMEM, "tensor-core rollout needs %zu B of shared memory", smem);
    NOC_CUDA(cudaFuncSetAttribute(rollout_tc_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::max(1, std::min(A.ntiles, sm_count()));
    double* partials = nullptr;
    if (mode == NOC_MODE_MEAN) {
        NOC_CUDA(cudaMallocAsync((void**)&partials, sizeof(double) * 8 * (size_t)grid, st));
        A.partials = partials;
    }
    rollout_tc_quad_kernel<<<grid, 128, smem, st>>>(A);
    count_launch();
    NOC_CUDA(cudaGetLastError());
    if (partials) {
        int frc = launch_finish(partials, grid, out_sums, st);
        if (frc) return frc;
        NOC_CUDA(cudaFreeAsync(partials, st));
    }
    return NOC_OK;
}

}  // namespace noc
