// noc_rollout.cuh — the persistent sample-tile rollout kernel (FP32 / FP64 FMA register tiles).
//
// One launch integrates a whole OCflow call (src/OCflow.py:7-95).  A CTA owns a tile of TS samples for
// all nt steps: every hidden vector of Phi lives in shared memory as [unit][sample] panels (sample
// contiguous), each Phi contraction is a register-tiled FMA GEMM (RO outputs x RS samples per thread;
// fp32 uses Blackwell's packed FFMA2) whose epilogue applies the activation in registers, and the problem
// terms (Cross2D / SwarmTraj / Quadcopter) are evaluated per sample between contractions.  Nothing but the
// initial state is read from HBM and nothing but the costs (or, with intermediates=True, the trajectory) is
// written.  Weights: the whole packed blob staged in shared memory (small nets) or streamed through
// warp-private cp.async rings (m >= 128).  DESIGN.md §3 has the configuration table and the measurements.
//
// Per ODE-function evaluation (ocOdefun, OCflow.py:104-140; Phi.getGrad, Phi.py:99-138), nTh = 2:
//   GEMM-1  o  = K0 s + b0        ->  u0 = act(o) (panel U), t0 = tanh(o) (panel T0)
//   GEMM-2  a1 = K1 u0 + b1       ->  y  = tanh(a1) * w           (U, in place)
//   GEMM-3  z1 = w + h K1' y      ->  v  = t0 * z1                (U, in place)
//   GEMM-4  g  = A'A s + K0' v + c_w                              (panel G)
// then L, H, Q, W from (x, p = g[:d]) and the RK combination.  General nTh keeps tanh(a_i) panels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "noc_types.cuh"

namespace noc {

// ------------------------------------------------------------------------------------------------
// tile configuration
// ------------------------------------------------------------------------------------------------
template <typename real_, int RO_, int RS_, int WO_, int NWO_, int NWS_, bool WSMEM_, bool ZGLOBAL_ = false>
struct Cfg {
    using real = real_;
    static constexpr int RO = RO_;            // outputs per thread
    static constexpr int RS = RS_;            // samples per thread
    static constexpr int WO = WO_;            // lanes along outputs
    static constexpr int WS = 32 / WO_;       // lanes along samples
    static constexpr int NWO = NWO_;          // warps along outputs
    static constexpr int NWS = NWS_;          // warps along samples
    static constexpr int WB = RO * WO;        // outputs per warp
    static constexpr int PB = WB * NWO;       // outputs per pass of the CTA
    static constexpr int TS = NWS * WS * RS;  // samples per tile
    static constexpr int NT = 32 * NWO * NWS; // threads per CTA
    static constexpr int TPS = NT / TS;       // threads per sample in the problem phase
    // row padding (in elements) chosen so that the epilogue's 16-byte panel stores of a quarter-warp
    // fall into distinct banks ("shared-memory panels")
    static constexpr int PAD = (sizeof(real_) == 4) ? (WO_ >= 8 ? 4 : 8) : (WO_ >= 8 ? 2 : 4);
    static constexpr int TSP = TS + PAD;
    static constexpr bool WSMEM = WSMEM_;     // whole weight blob staged in shared memory (small nets) ...
    // the augmented state [x, L, HJt, Q, W] (two panels of d+4 rows) lives in a per-CTA global scratch instead of
    // shared memory: frees a third of the per-sample footprint of small nets, i.e. almost twice the resident warps
    static constexpr bool ZGLOBAL = ZGLOBAL_;
    static constexpr bool WSTREAM = !WSMEM_;  // ... or streamed through warp-private cp.async rings (m >= 128)
    static constexpr int GRP = (sizeof(real_) == 4) ? 8 : 4;   // rows per cp.async group of a warp's weight stream
    // a warp owns all outputs of its own samples and one thread owns one sample in the problem phase:
    // tiles are warp-private and __syncwarp() replaces __syncthreads()
    // ... unless the CTA is one big 15-warp tile: there the warps are kept in lockstep with block barriers, because 15
    // warps drifting through ~40 KB of straight-line code thrash the instruction cache (measured: 28 % of stall
    // samples were instruction-fetch stalls when drifting, and lockstep is 20 % faster)
    static constexpr bool WARP_PRIVATE = (NWO_ == 1) && (TPS == 1) && !ZGLOBAL_;
    static_assert(NT % TS == 0, "threads per sample must be integral");
    static_assert(RS_ * sizeof(real_) % 16 == 0 && RO_ * sizeof(real_) % 16 == 0, "16-byte vector tiles");
};

// packed column of output `o` (see PhiPack): the RO outputs a thread owns are interleaved by WO in
// output space (so that epilogue stores of neighbouring lanes hit neighbouring panel rows) and stored
// as 16-byte chunks in the packed weight row (so that they load as vectors).
template <class C>
__host__ __device__ inline int pack_col(int o) {
    constexpr int VEC = 16 / (int)sizeof(typename C::real);      // elements per 16-byte chunk
    int pass = o / C::PB, rem = o % C::PB;
    int wo = rem / C::WB, r2 = rem % C::WB;
    int ro = r2 / C::WO, lo = r2 % C::WO;
    // chunk c = ro / VEC of every lane is stored lane-contiguous: the WO lanes of a warp read one conflict-free
    // run of WO * 16 bytes per vector load
    return pass * C::PB + wo * C::WB + (ro / VEC) * (C::WO * VEC) + lo * VEC + (ro % VEC);
}

// ------------------------------------------------------------------------------------------------
// math: act = antiderivative of tanh (Phi.py:8-9) and tanh, sharing e = exp(-2|x|)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float r_exp(float x) {
#ifdef NOC_PRECISE_MATH
    return expf(x);
#else
    return __expf(x);
#endif
}
__device__ __forceinline__ double r_exp(double x) { return exp(x); }
__device__ __forceinline__ float r_log1pe(float e) {   // log(1 + e), e in [0,1]
#ifdef NOC_PRECISE_MATH
    return logf(1.0f + e);
#else
    return __logf(1.0f + e);
#endif
}
__device__ __forceinline__ double r_log1pe(double e) { return log(1.0 + e); }
__device__ __forceinline__ float r_div(float a, float b) {
#ifdef NOC_PRECISE_MATH
    return a / b;
#else
    return __fdividef(a, b);
#endif
}
__device__ __forceinline__ double r_div(double a, double b) { return a / b; }
__device__ __forceinline__ float r_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double r_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float r_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double r_abs(double x) { return fabs(x); }
__device__ __forceinline__ float r_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double r_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ void r_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void r_sincos(double x, double* s, double* c) { sincos(x, s, c); }

template <typename real>
__device__ __forceinline__ void act_tanh(real pre, real& act, real& th) {
    real a = r_abs(pre);
    real e = r_exp(real(-2) * a);
    act = a + r_log1pe(e);
    th = copysign(r_div(real(1) - e, real(1) + e), pre);
}
template <typename real>
__device__ __forceinline__ real tanh_only(real pre) {
    real e = r_exp(real(-2) * r_abs(pre));
    return copysign(r_div(real(1) - e, real(1) + e), pre);
}

#ifndef NOC_PRECISE_MATH
// fp32 fast path: the three transcendentals as bare MUFU ops (ex2 / lg2 / rcp .approx.ftz).  The library intrinsics
// wrap each of them in range fix-ups (denormal results of exp, denormal arguments of log, huge divisors) that cannot
// occur here -- e = exp(-2|x|) is in [0,1], 1+e in [1,2] -- and those fix-ups doubled the size of the activation
// epilogues, which matters because one stage's code has to stay resident in the instruction cache.
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <>
__device__ __forceinline__ void act_tanh<float>(float pre, float& act, float& th) {
    const float a = fabsf(pre);
    const float e = mufu_ex2(a * -2.885390081777927f);               // exp(-2a) = 2^(-2 a log2(e))
    act = fmaf(mufu_lg2(1.0f + e), 0.6931471805599453f, a);          // a + ln(1 + e)
    th = copysignf((1.0f - e) * mufu_rcp(1.0f + e), pre);
}
template <>
__device__ __forceinline__ float tanh_only<float>(float pre) {
    const float e = mufu_ex2(fabsf(pre) * -2.885390081777927f);
    return copysignf((1.0f - e) * mufu_rcp(1.0f + e), pre);
}
#endif

// ------------------------------------------------------------------------------------------------
// 16-byte vector access
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void ld_panel(const float* p, float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
template <int N>
__device__ __forceinline__ void ld_panel(const double* p, double (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        double2 t = *reinterpret_cast<const double2*>(p + 2 * i);
        v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
}
template <int N>
__device__ __forceinline__ void st_panel(float* p, const float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i)
        *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <int N>
__device__ __forceinline__ void st_panel(double* p, const double (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) *reinterpret_cast<double2*>(p + 2 * i) = make_double2(v[2 * i], v[2 * i + 1]);
}

// my RO weights of one packed row: chunk c sits WO * VEC elements after chunk c-1 (see pack_col)
template <class C, typename real>
__device__ __forceinline__ void ld_wrow(const real* p, real (&w)[C::RO]) {
    constexpr int VEC = 16 / (int)sizeof(real);
#pragma unroll
    for (int c = 0; c < C::RO / VEC; ++c) {
        real t[VEC];
        ld_panel<VEC>(p + c * C::WO * VEC, t);
#pragma unroll
        for (int e = 0; e < VEC; ++e) w[c * VEC + e] = t[e];
    }
}

template <class C>
__device__ __forceinline__ void tile_sync() {
    if (C::WARP_PRIVATE) __syncwarp(); else __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// per-thread coordinates and panel pointers
// ------------------------------------------------------------------------------------------------
template <class C>
struct ThreadMap {
    int tid, wo, lo, scol, pcol, orow;
    __device__ ThreadMap() {
        tid = threadIdx.x;
        int warp = tid >> 5, lane = tid & 31;
        wo = warp % C::NWO;
        int ws = warp / C::NWO;
        lo = lane % C::WO;
        int ls = lane / C::WO;
        scol = (ws * C::WS + ls) * C::RS;          // first of my RS sample columns
        pcol = wo * C::WB + lo * (16 / (int)sizeof(typename C::real));   // my first 16-byte chunk in a packed weight row (within a pass)
        orow = wo * C::WB + lo;                    // output row of ro = 0 (within a pass); ro adds ro * WO
    }
};

// Element offsets of every panel from the dynamic shared-memory base.  Device functions rebuild their
// pointers as `smem_base<real>() + offset` so that the compiler keeps them in the shared address space
// (LDS/STS) even inside non-inlined functions; pointers fetched from a struct would decay to generic LD/ST.
struct Panels {
    int U, U2, T[MAXL], Zb, S, G, Qs, Z0, ZA, SC, RED, PN, QX, GP;
    int W;            // staged weight blob (WSMEM configurations) / slab ring (streamed configurations)
    int ring_slab, ring_ns;   // elements per ring slot (GRP rows of WB columns), slots per warp
    unsigned smem_u32;   // shared-window address of the dynamic shared-memory base (cp.async destinations)
};


__device__ __forceinline__ void cp_async16(unsigned dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {     // at most n groups still in flight
    if (n <= 0) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    else if (n == 1) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    else asm volatile("cp.async.wait_group 2;\n" ::: "memory");
}

// position of the weight stream of one warp: next group to issue = rows [row, row+GR) of my rows of matrix `seq`
// (element offset `src` into the blob, `step` elements between my consecutive rows, `cnt` rows in total) goes to
// ring slot `wslot`; the next group to consume sits in slot `cslot`.  Uniform within a warp; lives in registers.
struct WStream { int seq, row, wslot, cslot, src, step, cnt; };

// Everything the device functions need to know about the call, kept at the start of dynamic shared memory: a
// struct reached through a pointer parameter would live in local memory, and with ~220 KB of shared memory carved
// out of the L1 every such access is an L2 round trip (round-1 profile: long-scoreboard stalls on LDL).
constexpr int META_WARPS = 8, META_BYTES = 4096;
template <typename real>
struct Meta {
    PhiPack<real> P;
    ProbPack pr;
    Panels tp;
    int wtab[META_WARPS][2 * MAXL + 1][4];   // streamed configs, per warp and matrix: src offset, row step, rows, first | stride << 16
};

extern __shared__ __align__(16) unsigned char noc_smem_raw[];
template <typename real>
__device__ __forceinline__ real* smem_base() { return reinterpret_cast<real*>(noc_smem_raw + META_BYTES); }
template <typename real>
__device__ __forceinline__ Meta<real>& meta() { return *reinterpret_cast<Meta<real>*>(noc_smem_raw); }

// ------------------------------------------------------------------------------------------------
// contractions
// ------------------------------------------------------------------------------------------------
template <class C, typename real>
__device__ __forceinline__ void fma_tile(real (&acc)[C::RO][C::RS], const real (&w)[C::RO], const real (&a)[C::RS]) {
#pragma unroll
    for (int i = 0; i < C::RO; ++i)
#pragma unroll
        for (int j = 0; j < C::RS; ++j) acc[i][j] = r_fma(w[i], a[j], acc[i][j]);
}
// fp32: Blackwell's packed FFMA2 (fma.rn.f32x2, sm_100+) does the FMAs of two neighbouring samples in one instruction,
// with the weight as a scalar operand broadcast to both halves -- half the issue slots for the same (bit-identical)
// arithmetic.  Measured on the 8x8 inner loop: 68.6 vs 59.7 TFLOP/s (profiles/r01_fma_peaks.txt).
template <class C>
__device__ __forceinline__ void fma_tile(float (&acc)[C::RO][C::RS], const float (&w)[C::RO], const float (&a)[C::RS]) {
#pragma unroll
    for (int i = 0; i < C::RO; ++i)
#pragma unroll
        for (int j = 0; j < C::RS; j += 2) {
            const float2 r = __ffma2_rn(make_float2(w[i], w[i]), make_float2(a[j], a[j + 1]), make_float2(acc[i][j], acc[i][j + 1]));
            acc[i][j] = r.x;
            acc[i][j + 1] = r.y;
        }
}

// WSMEM configurations: acc[RO][RS] += sum_k W[k][pcol..pcol+RO) * in[k][scol..scol+RS), weights from the staged
// copy of the blob.  Groups of 4 k-steps run straight-line with register double-buffering (the next step's two
// vectors are loaded while the current step's FFMA block issues); the K % 4 tail is a plain loop.
template <class C, typename real>
__device__ __forceinline__ void gemm_acc(real (&acc)[C::RO][C::RS], int woff, int ldw, int in_off, int K) {
    const real* in = smem_base<real>() + in_off;
    const real* W = smem_base<real>() + woff;
    real wq[2][C::RO], a[2][C::RS];
    int k = 0;
    if (K >= 4) {
        ld_wrow<C>(W, wq[0]);
        ld_panel<C::RS>(in, a[0]);
        for (; k + 4 <= K; k += 4) {
            const real* Wk = W + k * ldw;
            const real* ik = in + k * C::TSP;
            const bool more = (k + 8 <= K);          // another full group follows: prefetch its first step
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u < 3) {
                    ld_wrow<C>(Wk + (u + 1) * ldw, wq[(u + 1) & 1]);
                    ld_panel<C::RS>(ik + (u + 1) * C::TSP, a[(u + 1) & 1]);
                } else if (more) {
                    ld_wrow<C>(Wk + 4 * ldw, wq[0]);
                    ld_panel<C::RS>(ik + 4 * C::TSP, a[0]);
                }
                fma_tile<C>(acc, wq[u & 1], a[u & 1]);
            }
        }
    }
    for (; k < K; ++k) {
        ld_wrow<C>(W + k * ldw, wq[0]);
        ld_panel<C::RS>(in + k * C::TSP, a[0]);
        fma_tile<C>(acc, wq[0], a[0]);
    }
}

// ---- streamed configurations: warp-private weight streams -----------------------------------------------------
// A warp only ever reads ITS OWN WB columns of a weight row, so every warp streams exactly those columns through a
// private ring in shared memory: GR rows per cp.async group, ring_ns groups deep.  All hand-over is warp-local
// (cp.async.wait_group + __syncwarp); the stream runs ahead across contraction and stage boundaries because the
// order in which one grad-Phi evaluation consumes the matrices is fixed (PhiPack::seq_*).

// load the cached descriptor of my rows of matrix ws.seq, skipping matrices I have no rows of (idle K-split warps)
template <typename real>
__device__ __forceinline__ void ws_load_matrix(WStream& ws, int warp) {
    const Meta<real>& M = meta<real>();
    while (M.wtab[warp][ws.seq][2] == 0) ws.seq = (ws.seq + 1 == M.P.nseq) ? 0 : ws.seq + 1;
    ws.src = M.wtab[warp][ws.seq][0];
    ws.step = M.wtab[warp][ws.seq][1];
    ws.cnt = M.wtab[warp][ws.seq][2];
    ws.row = 0;
}

// issue the next group of my stream (up to GRP rows of one matrix) into ring slot ws.wslot
template <class C, typename real>
__device__ __forceinline__ void ws_issue(WStream& ws, int warp, int lane) {
    constexpr int VEC = 16 / (int)sizeof(real), CPR = C::WB / VEC, GR = C::GRP;
    constexpr int RPI = 32 / CPR;                    // rows one warp-wide cp.async covers
    const Meta<real>& M = meta<real>();
    const int rows = (ws.cnt - ws.row < GR) ? (ws.cnt - ws.row) : GR;
    const int lr = lane / CPR, lc = (lane % CPR) * VEC;
    const real* src = M.P.blob + ws.src + lr * ws.step + lc;
    const int ring = M.tp.W + (warp * M.tp.ring_ns + ws.wslot) * (GR * C::WB);
    const unsigned dst = M.tp.smem_u32 + (unsigned)((ring + lr * C::WB + lc) * (int)sizeof(real));
    if (rows == GR) {
#pragma unroll
        for (int j = 0; j < GR / RPI; ++j)
            cp_async16(dst + (unsigned)(j * RPI * C::WB * (int)sizeof(real)), src + j * RPI * ws.step);
    } else {
        for (int r = lr; r < rows; r += RPI)
            cp_async16(dst + (unsigned)((r - lr) * C::WB * (int)sizeof(real)), src + (r - lr) * ws.step);
    }
    cp_async_commit();
    ws.row += rows;
    ws.src += rows * ws.step;
    ws.wslot = (ws.wslot + 1 == M.tp.ring_ns) ? 0 : ws.wslot + 1;
    if (ws.row >= ws.cnt) { ws.seq = (ws.seq + 1 == M.P.nseq) ? 0 : ws.seq + 1; ws_load_matrix<real>(ws, warp); }
}

// acc += my rows of matrix `seq` (the next one in my stream) times the matching rows of the input panel
template <class C, typename real>
__device__ __forceinline__ void gemm_wstream(real (&acc)[C::RO][C::RS], WStream& ws, int seq, int in_off, int lo, int warp, int lane) {
    constexpr int VEC = 16 / (int)sizeof(real), GR = C::GRP;
    const Meta<real>& M = meta<real>();
    const int cnt = M.wtab[warp][seq][2];
    if (cnt == 0) return;
    const int fs = M.wtab[warp][seq][3];
    const int first = fs & 0xffff, stride = fs >> 16;
    const int ns = M.tp.ring_ns;
    const real* sm = smem_base<real>();
    const real* ring = sm + M.tp.W + warp * ns * (GR * C::WB) + lo * VEC;
    const real* ins = sm + in_off + first * C::TSP;
    const int astep = stride * C::TSP;
    for (int g0 = 0; g0 < cnt; g0 += GR, ins += GR * astep) {
        const int rows = (cnt - g0 < GR) ? (cnt - g0) : GR;
        cp_async_wait_pending(ns - 2);              // my chunks of this group have landed ...
        __syncwarp();                               // ... and so have the other lanes'
        ws_issue<C, real>(ws, warp, lane);          // refill the slot consumed one group ago
        const real* Wsl = ring + ws.cslot * (GR * C::WB);
        real wq[2][C::RO], a[2][C::RS];
        ld_wrow<C>(Wsl, wq[0]);
        ld_panel<C::RS>(ins, a[0]);
        if (rows == GR) {                           // full group: straight-line, register double-buffered
#pragma unroll
            for (int r = 0; r < GR; ++r) {
                if (r + 1 < GR) {
                    ld_wrow<C>(Wsl + (r + 1) * C::WB, wq[(r + 1) & 1]);
                    ld_panel<C::RS>(ins + (r + 1) * astep, a[(r + 1) & 1]);
                }
                fma_tile<C>(acc, wq[r & 1], a[r & 1]);
            }
        } else {                                    // last group of a matrix
            for (int r = 0; r < rows; ++r) {
                ld_wrow<C>(Wsl + r * C::WB, wq[0]);
                ld_panel<C::RS>(ins + r * astep, a[0]);
                fma_tile<C>(acc, wq[0], a[0]);
            }
        }
        ws.cslot = (ws.cslot + 1 == ns) ? 0 : ws.cslot + 1;
    }
}

template <class C, typename real>
__device__ __forceinline__ void zero_acc(real (&acc)[C::RO][C::RS]) {
#pragma unroll
    for (int i = 0; i < C::RO; ++i)
#pragma unroll
        for (int j = 0; j < C::RS; ++j) acc[i][j] = real(0);
}

// m-wide contraction number `seq` of the evaluation, output pass `pass`
template <class C, typename real>
__device__ __forceinline__ void gemm_m(real (&acc)[C::RO][C::RS], const PhiPack<real>& P, const Panels& tp, WStream& ws,
                                       const ThreadMap<C>& tm, int seq, int pass, int in_off) {
    if (C::WSMEM) gemm_acc<C, real>(acc, tp.W + P.seq_off[seq] + pass * C::PB + tm.pcol, P.seq_N[seq], in_off + tm.scol, P.seq_K[seq]);
    else gemm_wstream<C, real>(acc, ws, seq, in_off + tm.scol, tm.lo, tm.tid >> 5, tm.tid & 31);
}

// scalar weight read (bias, w, c): staged copy or the global blob through the read-only path
template <class C, typename real>
__device__ __forceinline__ real wscalar(const PhiPack<real>& P, const Panels& tp, int idx) {
    if (C::WSMEM) return smem_base<real>()[tp.W + idx];
    return __ldg(P.blob + idx);
}

// ------------------------------------------------------------------------------------------------
// grad Phi (and, in the TERMINAL pass, the pieces of Phi itself) for the TS samples whose s = [x,t]
// sits in panel S.  On return panel G holds grad_s Phi (D rows).  TERMINAL additionally leaves
// partial sums of w . u_{nTh-1} in PN and A'A s in Qs.
// ------------------------------------------------------------------------------------------------
template <class C, typename real, bool TERMINAL>
__device__ __noinline__ WStream phi_chain(const ThreadMap<C> tm, WStream ws) {
    constexpr int RO = C::RO, RS = C::RS, TSP = C::TSP, PB = C::PB, WO = C::WO;
    const Meta<real>& M = meta<real>();
    const PhiPack<real>& P = M.P;
    const Panels& tp = M.tp;
    real* sm = smem_base<real>();
    const int npm = C::WSTREAM ? 1 : P.Npm / PB;
    const int nTh = P.nTh;
    real acc[RO][RS];
    real pn[RS];
#pragma unroll
    for (int j = 0; j < RS; ++j) pn[j] = real(0);
    int cur = tp.U, nxt = tp.U2;

    // The m-wide contractions of one evaluation, in stream order (one code path, so the unrolled FFMA blocks exist
    // once in the instruction cache):
    //   q = 0            opening layer (Phi.py:114-115):  u0 = act(K0 s + b0) -> U, tanh -> T[0]
    //   q = 1..nTh-1     forward layer i = q (Phi.py:118-120): u_i = u_{i-1} + h act(K_i u_{i-1} + b_i); the last one
    //                    leaves y = tanh(a_i) * w instead (u_{nTh-1} is only needed by Phi.forward, TERMINAL)
    //   q = nTh..2nTh-2  reverse layer i = 2nTh-1-q (Phi.py:124-131): z_i = z_{i+1} + h K_i'(tanh(a_i) * z_{i+1}),
    //                    z_nTh = w; the epilogue forms the next y with the tanh of the layer below
    for (int q = 0; q <= 2 * nTh - 2; ++q) {
        const int kind = (q == 0) ? 0 : (q < nTh ? 1 : 2);
        const int layer = (kind == 1) ? q : 2 * nTh - 1 - q;
        const bool last = (kind == 1) && (layer == nTh - 1);
        const int in_off = (q == 0) ? tp.S : cur;
        const int out_off = (q == 0) ? tp.U : nxt;
        for (int pass = 0; pass < npm; ++pass) {
            zero_acc<C>(acc);
            gemm_m<C>(acc, P, tp, ws, tm, q, pass, in_off);
            if (q > 0 && nxt == cur) tile_sync<C>();     // in place (single pass): every reader of `cur` is done
#pragma unroll
            for (int ro = 0; ro < RO; ++ro) {
                const int o = pass * PB + tm.orow + ro * WO;
                if (o < P.m) {
                    real out[RS];
                    real* po = sm + o * TSP + tm.scol;       // my RS samples of row o, relative to a panel offset
                    if (kind == 0) {
                        const real bb = wscalar<C>(P, tp, P.off_b[0] + o);
                        real tt[RS];
#pragma unroll
                        for (int j = 0; j < RS; ++j) act_tanh(acc[ro][j] + bb, out[j], tt[j]);
                        st_panel<RS>(po + tp.T[0], tt);
                    } else if (kind == 1) {
                        const real bb = wscalar<C>(P, tp, P.off_b[layer] + o);
                        if (!last) {
                            real uo[RS], tt[RS];
                            ld_panel<RS>(po + cur, uo);
#pragma unroll
                            for (int j = 0; j < RS; ++j) {
                                real av;
                                act_tanh(acc[ro][j] + bb, av, tt[j]);
                                out[j] = uo[j] + P.h * av;
                            }
                            st_panel<RS>(po + tp.T[layer], tt);
                        } else {
                            const real wv = wscalar<C>(P, tp, P.off_w + o);
                            if (TERMINAL) {          // Phi.forward needs u_{nTh-1} (Phi.py:96): accumulate w . u_last
                                real uo[RS];
                                ld_panel<RS>(po + cur, uo);
#pragma unroll
                                for (int j = 0; j < RS; ++j) {
                                    real av, tv;
                                    act_tanh(acc[ro][j] + bb, av, tv);
                                    pn[j] = r_fma(wv, uo[j] + P.h * av, pn[j]);
                                    out[j] = tv * wv;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < RS; ++j) out[j] = tanh_only(acc[ro][j] + bb) * wv;
                            }
                        }
                    } else {
                        real zi[RS], tt[RS];
                        if (layer == nTh - 1) {
                            const real wv = wscalar<C>(P, tp, P.off_w + o);
#pragma unroll
                            for (int j = 0; j < RS; ++j) zi[j] = wv + P.h * acc[ro][j];
                        } else {
                            ld_panel<RS>(po + tp.Zb, zi);
#pragma unroll
                            for (int j = 0; j < RS; ++j) zi[j] = zi[j] + P.h * acc[ro][j];
                        }
                        if (layer > 1) st_panel<RS>(po + tp.Zb, zi);
                        ld_panel<RS>(po + tp.T[layer - 1], tt);
#pragma unroll
                        for (int j = 0; j < RS; ++j) out[j] = tt[j] * zi[j];
                    }
                    st_panel<RS>(po + out_off, out);
                }
            }
        }
        tile_sync<C>();
        if (q > 0) { const int t = cur; cur = nxt; nxt = t; }
        if (TERMINAL && last) {                      // w . u_last: reduce over the WO lanes that share my samples
#pragma unroll
            for (int j = 0; j < RS; ++j)
#pragma unroll
                for (int off = 1; off < WO; off <<= 1) pn[j] += __shfl_xor_sync(0xffffffffu, pn[j], off);
            if (tm.lo == 0) st_panel<RS>(sm + tp.PN + tm.wo * TSP + tm.scol, pn);
        }
    }
    const int seq = 2 * nTh - 1;                     // the two D-wide matrices: sym, then W4

    // GEMM-4 (Phi.py:133-136): grad = K0' v + A'A s + c_w'.  The A'A s product is accumulated first so that
    // the terminal pass can keep it (Phi.forward's quadratic term, Phi.py:96) without a second register tile.
    if (C::WSMEM) {
        const int npd = P.Npd / PB;
        for (int pass = 0; pass < npd; ++pass) {
            zero_acc<C>(acc);
            for (int dq = 0; dq < 2; ++dq) {
                gemm_m<C>(acc, P, tp, ws, tm, seq + dq, pass, dq == 0 ? tp.S : cur);
                if (TERMINAL && dq == 0) {
#pragma unroll
                    for (int ro = 0; ro < RO; ++ro) {
                        int o = pass * PB + tm.orow + ro * WO;
                        if (o < P.D) st_panel<RS>(sm + tp.Qs + o * TSP + tm.scol, acc[ro]);
                    }
                }
            }
            if (tp.G == cur) tile_sync<C>();     // G aliases the hidden panel (single pass): readers are done
#pragma unroll
            for (int ro = 0; ro < RO; ++ro) {
                int o = pass * PB + tm.orow + ro * WO;
                if (o < P.D) {
                    real cw = wscalar<C>(P, tp, P.off_cw + o);
                    real g[RS];
#pragma unroll
                    for (int j = 0; j < RS; ++j) g[j] = acc[ro][j] + cw;
                    st_panel<RS>(sm + tp.G + o * TSP + tm.scol, g);
                }
            }
        }
        tile_sync<C>();
    } else {
        // streamed configurations: D is much narrower than the CTA's output span, so the warps are re-tiled as
        // ntile_d output tiles, every warp taking a K-slice of its tile's rows; slices >= 1 hand their partial sums
        // to slice 0 through panel GP, one slice per round.
        const int tile_d = tm.wo % P.ntile_d, kpart = tm.wo / P.ntile_d;
        const int orow = tile_d * C::WB + tm.lo;
        const int ks_t = (C::NWO - tile_d + P.ntile_d - 1) / P.ntile_d;      // K-slices of my output tile
        zero_acc<C>(acc);
        for (int dq = 0; dq < 2; ++dq) {
            gemm_wstream<C, real>(acc, ws, seq + dq, (dq == 0 ? tp.S : cur) + tm.scol, tm.lo, tm.tid >> 5, tm.tid & 31);
            if (dq == 1 || TERMINAL) {
                for (int round = 1; round < P.ksplit; ++round) {     // all warps walk the rounds
                    if (kpart == round) {
#pragma unroll
                        for (int ro = 0; ro < RO; ++ro) st_panel<RS>(sm + tp.GP + (orow + ro * WO) * TSP + tm.scol, acc[ro]);
                    }
                    __syncthreads();
                    if (kpart == 0 && round < ks_t) {
#pragma unroll
                        for (int ro = 0; ro < RO; ++ro) {
                            real t[RS];
                            ld_panel<RS>(sm + tp.GP + (orow + ro * WO) * TSP + tm.scol, t);
#pragma unroll
                            for (int j = 0; j < RS; ++j) acc[ro][j] += t[j];
                        }
                    }
                    __syncthreads();
                }
                if (dq == 0) {                   // TERMINAL: keep A'A s; the other slices restart from zero
                    if (kpart == 0) {
#pragma unroll
                        for (int ro = 0; ro < RO; ++ro) {
                            int o = orow + ro * WO;
                            if (o < P.D) st_panel<RS>(sm + tp.Qs + o * TSP + tm.scol, acc[ro]);
                        }
                    } else {
                        zero_acc<C>(acc);
                    }
                }
            }
        }
        __syncthreads();                         // every reader of `cur` (which G may alias) is done
        if (kpart == 0) {
#pragma unroll
            for (int ro = 0; ro < RO; ++ro) {
                int o = orow + ro * WO;
                if (o < P.D) {
                    real cw = wscalar<C>(P, tp, P.off_cw + o);
                    real g[RS];
#pragma unroll
                    for (int j = 0; j < RS; ++j) g[j] = acc[ro][j] + cw;
                    st_panel<RS>(sm + tp.G + o * TSP + tm.scol, g);
                }
            }
        }
        __syncthreads();
    }
    return ws;
}

// ------------------------------------------------------------------------------------------------
// problem terms
// ------------------------------------------------------------------------------------------------
// diagonal Gaussian pdf (src/utils.py:70-86) for a 2-D / 3-D point
template <typename real>
__device__ __forceinline__ real gauss2(real x0, real x1, real m0, real m1, real c0, real c1) {
    const double twopi = 6.283185307179586476925286766559;
    // the covariances are literals at every call site: the reciprocals fold at compile time (a true division is ~10
    // instructions; differs from the reference's quotient by <= 1 ulp)
    real iden = real(1) / (real(twopi) * r_sqrt(c0 * c1));
    real e = (x0 - m0) * (x0 - m0) * (real(1) / c0) + (x1 - m1) * (x1 - m1) * (real(1) / c1);
    return r_exp(real(-0.5) * e) * iden;
}
template <typename real>
__device__ __forceinline__ real gauss3(real x0, real x1, real x2, real m0, real m1, real m2, real c0, real c1, real c2) {
    const double twopi15 = 15.749609945722419;   // (2 pi)^(3/2)
    real iden = real(1) / (real(twopi15) * r_sqrt(c0 * c1 * c2));
    real e = (x0 - m0) * (x0 - m0) * (real(1) / c0) + (x1 - m1) * (x1 - m1) * (real(1) / c1) + (x2 - m2) * (x2 - m2) * (real(1) / c2);
    return r_exp(real(-0.5) * e) * iden;
}

// per-agent terrain cost (Cross2D.py:90-119, SwarmTraj.py:90-122) at the agent position (x0, x1[, x2]).
// Eval-mode hard obstacles return the 0/1 "inside" indicator (the reference returns the boolean mask, F6).
template <typename real>
__device__ real terrain_agent(const ProbPack& pr, real x0, real x1, real x2) {
    if (pr.obstacle == 1) {             // softcorridor: four Gaussians, cov 0.2
        real c = real(0.2);
        return ((gauss2<real>(x0, x1, real(-2.5), real(0), c, c) + gauss2<real>(x0, x1, real(2.5), real(0), c, c)) +
                gauss2<real>(x0, x1, real(-1.5), real(0), c, c)) + gauss2<real>(x0, x1, real(1.5), real(0), c, c);
    }
    if (pr.obstacle == 2) {             // hardcorridor: discs of radius 2 (+r in training) around (0,4), (0,-3.5)
        real d1 = r_sqrt(x0 * x0 + (x1 - real(4)) * (x1 - real(4)));
        real d2 = r_sqrt(x0 * x0 + (x1 + real(3.5)) * (x1 + real(3.5)));
        if (!pr.training) return (d1 < real(2.0) || d2 < real(2.0)) ? real(1) : real(0);
        real thr = real(2.0 + pr.r);
        if (!(d1 < thr || d2 < thr)) return real(0);
        return gauss2<real>(x0, x1, real(0), real(4), real(1), real(1)) + gauss2<real>(x0, x1, real(0), real(-3.5), real(1), real(1));
    }
    if (pr.obstacle == 3) {             // blocks: two boxes
        if (!pr.training) {
            bool in = (x0 < real(2.0) && x0 > real(-2.0) && x1 < real(0.5) && x1 > real(-0.5) && x2 < real(7.0)) ||
                      (x0 < real(4.0) && x0 > real(2.0) && x1 < real(1.0) && x1 > real(-1.0) && x2 < real(4.0));
            return in ? real(1) : real(0);
        }
        double r = pr.r;
        bool in = (x0 < real(2.0 + r) && x0 > real(-2.0 - r) && x1 < real(0.5 + r) && x1 > real(-0.5 - r) && x2 < real(7.0 + r)) ||
                  (x0 < real(4.0 + r) && x0 > real(2.0 - r) && x1 < real(1.0 + r) && x1 > real(-1.0 - r) && x2 < real(4.0 + r));
        if (!in) return real(0);
        return (gauss3<real>(x0, x1, x2, real(0), real(0), real(2), real(9), real(3), real(9)) +
                gauss3<real>(x0, x1, x2, real(2.5), real(0), real(2), real(9), real(3), real(3))) + real(999);
    }
    return real(0);
}

// Interaction cost of one sample for agents i = part, part+TPS, ... against every j > i (Cross2D.py:147-160,
// SwarmTraj.py:147-162).  Fast path: only the minimum squared distance of my pairs is tracked (DIM loads,
// DIM subs/fmas and one min per pair); the exact sqrt / cut-off / exp / "== 1" rule runs only when some pair is
// within the (slightly widened) cut-off, which on the benchmark distributions is < 1 % of the evaluations.
template <int DIM, int TSP, typename real>
__device__ __forceinline__ real interaction_pairs(const real* xs, int A, int part, int tps, real cut, real c2) {
    const real guard = cut * cut * real(1.0001);
    real dmin = guard;
    for (int i = part; i < A - 1; i += tps) {
        real xi[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) xi[c] = xs[(i * DIM + c) * TSP];
        const real* pj = xs + (i + 1) * DIM * TSP;
#pragma unroll 4
        for (int j = i + 1; j < A; ++j, pj += DIM * TSP) {
            real d2 = real(0);
#pragma unroll
            for (int c = 0; c < DIM; ++c) { real df = xi[c] - pj[c * TSP]; d2 = r_fma(df, df, d2); }
            dmin = fmin(dmin, d2);
        }
    }
    real w = real(0);
    if (dmin < guard) {                   // rare: redo my pairs exactly
        for (int i = part; i < A - 1; i += tps) {
            for (int j = i + 1; j < A; ++j) {
                real d2 = real(0);
#pragma unroll
                for (int c = 0; c < DIM; ++c) { real df = xs[(i * DIM + c) * TSP] - xs[(j * DIM + c) * TSP]; d2 = r_fma(df, df, d2); }
                if (d2 < guard) {
                    real dd = r_sqrt(d2);
                    if (dd < cut) {
                        real e = r_exp(-(dd * dd) / c2);
                        if (e != real(1)) w += e;   // pairs whose Gaussian rounds to 1 are dropped (mask2)
                    }
                }
            }
        }
    }
    return w;
}

// Fills, for every sample of the tile, SC rows L, HJ = |Phi_t - H|, Q, W (and H in row 7), from x in
// panel S and p = grad_x Phi in panel G (calcLHQW: Cross2D.py:73-87, SwarmTraj.py:71-87,
// Quadcopter.py:86-113).  Quadcopter also leaves u/mass, f7, f8, f9, u per agent in QX for the
// dynamics and the controls.
template <class C, typename real>
__device__ __noinline__ void problem_phase(int d, int tid) {
    constexpr int TS = C::TS, TSP = C::TSP, TPS = C::TPS;
    const Meta<real>& M = meta<real>();
    const ProbPack& pr = M.pr;
    const Panels& tp = M.tp;
    real* sm = smem_base<real>();
    const int s = tid % TS, part = tid / TS;
    const real* xs = sm + tp.S + s;
    const real* ps = sm + tp.G + s;
    real* sc = sm + tp.SC + s;

    if (pr.kind == 2) {                   // ---------------- Quadcopter
        if (part == 0) {
            real H = real(0), Q = real(0), W = real(0);
            real L = real(pr.alph_Q) * Q;
            if (pr.alph_W > 0.0) {
                if (pr.nAgents == 2) {    // Quadcopter.py:138-142
                    real d2 = real(0);
                    for (int c = 0; c < 3; ++c) { real df = xs[c * TSP] - xs[(12 + c) * TSP]; d2 = r_fma(df, df, d2); }
                    real dd = r_sqrt(d2);
                    if (dd < real(2 * pr.r)) W = r_exp(-(dd * dd) / real(2 * pr.r * pr.r));
                }
                L = L + real(pr.alph_W) * W;
            }
            for (int a = 0; a < pr.nAgents; ++a) {
                const real* x = xs + a * 12 * TSP;
                const real* p = ps + a * 12 * TSP;
                real sps, cps, sth, cth, sph, cph;
                r_sincos(x[3 * TSP], &sps, &cps);
                r_sincos(x[4 * TSP], &sth, &cth);
                r_sincos(x[5 * TSP], &sph, &cph);
                real f7 = sps * sph + cps * sth * cph;          // Quadcopter.py:190-195
                real f8 = -cps * sph + sps * sth * cph;
                real f9 = cth * cph;
                real p6 = p[6 * TSP], p7 = p[7 * TSP], p8 = p[8 * TSP];
                real fp = f7 * p6 + f8 * p7 + f9 * p8;
                real u = real(-1.0 / (2.0 * pr.mass)) * fp;     // calcU, Quadcopter.py:160-163
                real p9 = p[9 * TSP], p10 = p[10 * TSP], p11 = p[11 * TSP];
                real sq = p9 * p9 + p10 * p10 + p11 * p11;
                L = L + real(2) + u * u + real(0.25) * sq;
                real um = u / real(pr.mass);
                real xv = x[6 * TSP] * p[0] + x[7 * TSP] * p[1 * TSP] + x[8 * TSP] * p[2 * TSP];
                real xw = x[9 * TSP] * p[3 * TSP] + x[10 * TSP] * p[4 * TSP] + x[11 * TSP] * p[5 * TSP];
                H = H - L - xv - xw - um * fp + real(pr.grav) * p8 + real(0.5) * sq;
                real* qx = sm + tp.QX + (a * 5) * TSP + s;
                qx[0] = um; qx[TSP] = f7; qx[2 * TSP] = f8; qx[3 * TSP] = f9; qx[4 * TSP] = u;
            }
            sc[SC_L * TSP] = L;
            sc[SC_HJ * TSP] = r_abs(ps[d * TSP] - H);
            sc[SC_Q * TSP] = Q;
            sc[SC_W * TSP] = W;
            sc[7 * TSP] = H;
        }
        tile_sync<C>();
        return;
    }

    // ---------------- Cross2D / SwarmTraj
    const int A = pr.nAgents, dim = pr.agentDim;
    real pp = real(0), q = real(0), w = real(0);
    for (int r = part; r < d; r += TPS) { real v = ps[r * TSP]; pp = r_fma(v, v, pp); }
    const bool needQ = (pr.obstacle != 0) && (pr.kind == 0 || pr.alph_Q > 0.0);
    if (needQ)
        for (int a = part; a < A; a += TPS) {
            const real* xa = xs + a * dim * TSP;
            q += terrain_agent<real>(pr, xa[0], xa[TSP], dim == 3 ? xa[2 * TSP] : real(0));
        }
    if (pr.alph_W != 0.0 && A >= 2) {
        const real cut = real(pr.cutW);
        const real c2 = real(2 * pr.r * pr.r);
        if (A == 2) {                     // Cross2D.py:133-145 / SwarmTraj.py:136-145: no "== 1" rule here
            if (part == 0) {
                real d2 = real(0);
                for (int c = 0; c < dim; ++c) { real df = xs[c * TSP] - xs[(dim + c) * TSP]; d2 = r_fma(df, df, d2); }
                real dd = r_sqrt(d2);
                if (dd < cut) w = r_exp(-(dd * dd) / c2);
            }
        } else if (dim == 2) {
            w = interaction_pairs<2, TSP, real>(xs, A, part, TPS, cut, c2);
        } else {
            w = interaction_pairs<3, TSP, real>(xs, A, part, TPS, cut, c2);
        }
    }
    if (TPS > 1) {
        real* red = sm + tp.RED + s;
        red[(0 * TPS + part) * TSP] = pp;
        red[(1 * TPS + part) * TSP] = q;
        red[(2 * TPS + part) * TSP] = w;
        __syncthreads();
        if (part == 0) {
            pp = red[0]; q = red[(1 * TPS) * TSP]; w = red[(2 * TPS) * TSP];
            for (int k = 1; k < TPS; ++k) {
                pp += red[(0 * TPS + k) * TSP]; q += red[(1 * TPS + k) * TSP]; w += red[(2 * TPS + k) * TSP];
            }
        }
    }
    if (part == 0) {
        real Qret, L;
        if (pr.kind == 0) {               // Cross2D returns Q pre-scaled by alph_Q (quirk 6)
            Qret = real(pr.alph_Q) * q;
            L = real(0.5) * pp + Qret;
        } else {
            Qret = (pr.alph_Q > 0.0) ? q : real(0);
            L = real(0.5) * pp + real(pr.alph_Q) * Qret;
        }
        if (pr.alph_W != 0.0) L = L + real(pr.alph_W) * w; else w = real(0);
        real H = -L + pp;
        sc[SC_L * TSP] = L;
        sc[SC_HJ * TSP] = r_abs(ps[d * TSP] - H);
        sc[SC_Q * TSP] = Qret;
        sc[SC_W * TSP] = w;
        sc[7 * TSP] = H;
    }
    tile_sync<C>();
}

// dx/dt = -grad_p H for state row `row` of sample `s` (Cross2D.py:69-70, SwarmTraj.py:68-69, Quadcopter.py:65-84)
template <class C, typename real>
__device__ __forceinline__ real state_rate(const ProbPack& pr, const Panels& tp, int row, int s) {
    constexpr int TSP = C::TSP;
    const real* sm = smem_base<real>();
    if (pr.kind != 2) return -sm[tp.G + row * TSP + s];
    int a = row / 12, c = row % 12;
    if (c < 6) return sm[tp.S + (a * 12 + 6 + c) * TSP + s];
    if (c < 9) {
        real um = sm[tp.QX + (a * 5) * TSP + s];
        real f = sm[tp.QX + (a * 5 + 1 + (c - 6)) * TSP + s];
        real g = -um * f;
        if (c == 8) g = g + real(pr.grav);
        return -g;
    }
    return -(real(0.5) * sm[tp.G + row * TSP + s]);
}

// control channel `c` of sample `s` (calcCtrls: Cross2D.py:164-165, SwarmTraj.py:166-167, Quadcopter.py:165-174)
template <class C, typename real>
__device__ __forceinline__ real control_value(const ProbPack& pr, const Panels& tp, int c, int s) {
    constexpr int TSP = C::TSP;
    const real* sm = smem_base<real>();
    if (pr.kind != 2) return -sm[tp.G + c * TSP + s];
    int a = c / 4, q = c % 4;
    if (q == 0) return sm[tp.QX + (a * 5 + 4) * TSP + s];
    return real(-0.5) * sm[tp.G + (a * 12 + 8 + q) * TSP + s];
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void carve(const SmemPlan& sp, Panels& tp) {
    constexpr int TSP = C::TSP;
    tp.U = sp.U * TSP; tp.U2 = sp.U2 * TSP;
    for (int i = 0; i < MAXL; ++i) tp.T[i] = sp.T[i] * TSP;
    tp.Zb = sp.Zb * TSP; tp.S = sp.S * TSP; tp.G = sp.G * TSP; tp.Qs = sp.Qs * TSP;
    tp.Z0 = sp.Z0 * TSP; tp.ZA = sp.ZA * TSP; tp.SC = sp.SC * TSP; tp.RED = sp.RED * TSP;
    tp.PN = sp.PN * TSP; tp.QX = sp.QX * TSP; tp.GP = sp.GP * TSP;
    tp.W = sp.wsm_off; tp.ring_slab = sp.ring_slab; tp.ring_ns = sp.ring_ns;
    tp.smem_u32 = (unsigned)__cvta_generic_to_shared(noc_smem_raw + META_BYTES);   // = smem_base()
}

// S[0..d) <- rows of `src` (a [row][TSP] panel in shared memory or in the global scratch), S[d] <- t
template <class C, typename real>
__device__ __forceinline__ void stage_input_from(const Panels& tp, const real* src, int d, real t, int tid) {
    constexpr int TS = C::TS, TSP = C::TSP;
    real* sm = smem_base<real>();
    const int s = tid % TS;
    for (int row = tid / TS; row <= d; row += C::TPS) sm[tp.S + row * TSP + s] = (row < d) ? src[row * TSP + s] : t;
}

template <class C, typename real>
__global__ void __launch_bounds__(C::NT) rollout_kernel(const RolloutArgs<real> A, const int kmode) {
    constexpr int TS = C::TS, TSP = C::TSP, NT = C::NT;
    real* sm = smem_base<real>();
    static_assert(sizeof(Meta<real>) <= META_BYTES, "Meta must fit in its reserved shared-memory header");
    static_assert(!C::WSTREAM || C::NT / 32 <= META_WARPS, "wtab holds META_WARPS warps");
    Meta<real>& M = meta<real>();
    const ThreadMap<C> tm;
    const int tid = tm.tid;
    if (tid == 0) {
        M.P = A.phi;
        M.pr = A.prob;
        carve<C>(A.sp, M.tp);
    }
    __syncthreads();
    const PhiPack<real>& P = M.P;
    const ProbPack& pr = M.pr;
    const Panels& tp = M.tp;
    const int d = P.d, D = P.D;

    WStream ws = {0, 0, 0, 0, 0, 0, 0};
    if (C::WSMEM) {                       // stage the packed weights once per CTA
        for (int i = tid; i < P.blob_len; i += NT) sm[tp.W + i] = P.blob[i];
    } else {
        // my rows of every matrix of the stream: m-wide matrices -> all rows of my WB columns; D-wide matrices ->
        // rows kpart, kpart + ksplit, ... of output tile tile_d (or none, for warps beyond ntile_d * ksplit)
        const int warp = tid >> 5, lane = tid & 31;
        const int tile_d = tm.wo % P.ntile_d, kpart = tm.wo / P.ntile_d;
        for (int q = lane; q < P.nseq; q += 32) {
            const bool dwide = q >= P.nseq - 2;
            const int ks_t = (C::NWO - tile_d + P.ntile_d - 1) / P.ntile_d;   // every warp takes a K-slice of its tile
            const int first = dwide ? kpart : 0, stride = dwide ? ks_t : 1;
            const int cnt = (P.seq_K[q] - first + stride - 1) / stride;
            M.wtab[warp][q][0] = P.seq_off[q] + first * P.seq_N[q] + (dwide ? tile_d : tm.wo) * C::WB;
            M.wtab[warp][q][1] = stride * P.seq_N[q];
            M.wtab[warp][q][2] = cnt > 0 ? cnt : 0;
            M.wtab[warp][q][3] = first | (stride << 16);
        }
        __syncwarp();
        if (kmode != KMODE_PROB) {        // start my weight stream: ring_ns - 1 groups in flight
            ws_load_matrix<real>(ws, warp);
            for (int i = 0; i < tp.ring_ns - 1; ++i) ws_issue<C, real>(ws, warp, lane);
        }
    }
    __syncthreads();
    // augmented state panels: shared memory, or this CTA's slice of the global scratch (large nets)
    const bool zglob = C::ZGLOBAL || (C::WSTREAM && A.sp.z_global);
    real* zb = zglob ? (A.zscratch + (size_t)blockIdx.x * A.zstride) : sm;

    double csum = 0.0;                    // threads 0..6: this CTA's running sum of cost q (mean mode)
    long long cnt = 0;

    for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
        const long long s0 = (long long)tile * TS;
        const int nvalid = (int)((A.n - s0 < TS) ? (A.n - s0) : TS);

        // ---------------------------------------------------------------- evaluation-only modes
        if (kmode == KMODE_PHI) {         // Phi.forward / Phi.getGrad on rows of s = [x,t]  (Phi.py:91-138)
            for (int idx = tid; idx < TS * D; idx += NT) {
                int s = idx / D, c = idx % D;
                long long gs = s0 + (s < nvalid ? s : nvalid - 1);
                sm[tp.S + c * TSP + s] = A.x[gs * D + c];
            }
            __syncthreads();
            ws = phi_chain<C, real, true>(tm, ws);
            __syncthreads();
            for (int s = tid; s < nvalid; s += NT) {
                if (A.out_a) {
                    real phiN = real(0), quad = real(0), lin = real(0);
                    for (int k = 0; k < C::NWO; ++k) phiN += sm[tp.PN + k * TSP + s];
                    for (int o = 0; o < D; ++o) {
                        real sv = sm[tp.S + o * TSP + s];
                        quad = r_fma(sv, sm[tp.Qs + o * TSP + s], quad);
                        lin = r_fma(wscalar<C>(P, tp, P.off_cw + o), sv, lin);
                    }
                    A.out_a[s0 + s] = phiN + real(0.5) * quad + (lin + wscalar<C>(P, tp, P.off_cb));
                }
            }
            if (A.out_b)
                for (int idx = tid; idx < nvalid * D; idx += NT) {
                    int s = idx / D, c = idx % D;
                    A.out_b[(s0 + s) * D + c] = sm[tp.G + c * TSP + s];
                }
            __syncthreads();
            continue;
        }
        if (kmode == KMODE_PROB) {        // calcLHQW / calcGradpH / calcCtrls on rows (x, p)
            for (int idx = tid; idx < TS * d; idx += NT) {
                int s = idx / d, c = idx % d;
                long long gs = s0 + (s < nvalid ? s : nvalid - 1);
                sm[tp.S + c * TSP + s] = A.x[gs * d + c];
                sm[tp.G + c * TSP + s] = A.p_in[gs * d + c];
            }
            for (int s = tid; s < TS; s += NT) sm[tp.G + d * TSP + s] = real(0);
            __syncthreads();
            problem_phase<C, real>(d, tid);
            __syncthreads();
            if (A.out_a)
                for (int s = tid; s < nvalid; s += NT) {
                    real* o = A.out_a + (s0 + s) * 4;
                    o[0] = sm[tp.SC + SC_L * TSP + s]; o[1] = sm[tp.SC + 7 * TSP + s];
                    o[2] = sm[tp.SC + SC_Q * TSP + s]; o[3] = sm[tp.SC + SC_W * TSP + s];
                }
            if (A.out_b)
                for (int idx = tid; idx < nvalid * d; idx += NT) {
                    int s = idx / d, c = idx % d;
                    A.out_b[(s0 + s) * d + c] = -state_rate<C, real>(pr, tp, c, s);
                }
            if (A.out_c)
                for (int idx = tid; idx < nvalid * pr.nctrl; idx += NT) {
                    int s = idx / pr.nctrl, c = idx % pr.nctrl;
                    A.out_c[(s0 + s) * pr.nctrl + c] = control_value<C, real>(pr, tp, c, s);
                }
            __syncthreads();
            continue;
        }

        // ---------------------------------------------------------------- rollout (OCflow.py:7-95)
        int Z0 = tp.Z0, ZA = tp.ZA;
        if (zglob) { Z0 = 0; ZA = (d + 4) * TSP; }
        for (int idx = tid; idx < TS * d; idx += NT) {       // z = [x, 0, 0, 0, 0]  (OCflow.py:33)
            int s = idx / d, c = idx % d;
            long long gs = s0 + (s < nvalid ? s : nvalid - 1);   // padding samples replay the last valid one
            zb[Z0 + c * TSP + s] = A.x[gs * d + c];
        }
        for (int row = tid / TS; row < 4; row += C::TPS) zb[Z0 + (d + row) * TSP + tid % TS] = real(0);
        __syncthreads();

        const bool inter = (A.mode == 2);
        const int ntp1 = A.nt + 1;
        if (inter) {                                          // zFull[:,:,0] = z, ctrlFull[:,:,0] = 0 (OCflow.py:37-43)
            for (int idx = tid; idx < nvalid * (d + 4); idx += NT) {
                int s = idx / (d + 4), row = idx % (d + 4);
                A.out_b[((s0 + s) * (d + 4) + row) * ntp1] = zb[Z0 + row * TSP + s];
            }
            for (int idx = tid; idx < nvalid * pr.nctrl; idx += NT) {
                int s = idx / pr.nctrl, c = idx % pr.nctrl;
                A.out_c[((s0 + s) * pr.nctrl + c) * ntp1] = real(0);
            }
        }

        const int nstage = (A.stepper == 4) ? 4 : (A.stepper == 1 ? 1 : 0);
        for (int k = 0; k < A.nt; ++k) {
            const double* tt = A.times + 5 * k;
            const real hstep = real(tt[4]);                   // h = t1 - t0 recomputed per step (OCflow.py:169)
            if (nstage > 0) {
                stage_input_from<C, real>(tp, zb + Z0, d, real(tt[0]), tid);
                tile_sync<C>();
            }
            for (int st = 0; st < nstage; ++st) {
                // RK4 weights (OCflow.py:172-182); python doubles rounded to the tensor dtype
                real wgt, cnext, tnext;
                if (nstage == 1) { wgt = real(1); cnext = real(0); tnext = real(0); }
                else if (st == 0) { wgt = real(1.0 / 6.0); cnext = real(0.5); tnext = real(tt[1]); }
                else if (st == 1) { wgt = real(2.0 / 6.0); cnext = real(0.5); tnext = real(tt[1]); }
                else if (st == 2) { wgt = real(2.0 / 6.0); cnext = real(1.0); tnext = real(tt[2]); }
                else { wgt = real(1.0 / 6.0); cnext = real(0); tnext = real(0); }
                const bool lastst = (st == nstage - 1);

                ws = phi_chain<C, real, false>(tm, ws);        // G <- grad Phi([x_stage, t])
                problem_phase<C, real>(d, tid);            // SC <- L, |Phi_t - H|, Q, W

                if (pr.kind == 2) {                           // Quadcopter rates read other rows of S: K first, then update
                    for (int row = tid / TS; row < d; row += C::TPS) {
                        const int s = tid % TS;
                        real f = state_rate<C, real>(pr, tp, row, s);
                        sm[tp.G + row * TSP + s] = hstep * f;
                    }
                    tile_sync<C>();
                }
                // RK combination (OCflow.py:172-182).  Four independent (row, sample) items per trip with all loads
                // first: the augmented state may live in the global scratch, and its latency is paid once per batch.
                {
                    const int s = tid % TS;
                    for (int row0 = tid / TS; row0 < d + 4; row0 += 4 * C::TPS) {
                        real kk[4], z0v[4], zpv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int row = row0 + u * C::TPS;
                            if (row < d + 4) {
                                if (row >= d) kk[u] = hstep * sm[tp.SC + (row - d) * TSP + s];
                                else if (pr.kind == 2) kk[u] = sm[tp.G + row * TSP + s];
                                else kk[u] = hstep * (-sm[tp.G + row * TSP + s]);
                                z0v[u] = zb[Z0 + row * TSP + s];
                                zpv[u] = (st == 0) ? z0v[u] : zb[ZA + row * TSP + s];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int row = row0 + u * C::TPS;
                            if (row < d + 4) {
                                zb[ZA + row * TSP + s] = zpv[u] + wgt * kk[u];
                                if (!lastst && row < d) sm[tp.S + row * TSP + s] = z0v[u] + cnext * kk[u];
                            }
                        }
                    }
                }
                if (!lastst)
                    for (int s = tid; s < TS; s += NT) sm[tp.S + d * TSP + s] = tnext;
                tile_sync<C>();
            }
            if (nstage > 0) { int t = Z0; Z0 = ZA; ZA = t; }

            if (inter) {                                      // OCflow.py:51-55
                __syncthreads();
                for (int idx = tid; idx < nvalid * (d + 4); idx += NT) {
                    int s = idx / (d + 4), row = idx % (d + 4);
                    A.out_b[((s0 + s) * (d + 4) + row) * ntp1 + (k + 1)] = zb[Z0 + row * TSP + s];
                }
                stage_input_from<C, real>(tp, zb + Z0, d, real(tt[3]), tid);   // new state, OLD time (quirk 3)
                __syncthreads();
                ws = phi_chain<C, real, false>(tm, ws);
                if (pr.kind == 2) problem_phase<C, real>(d, tid);
                __syncthreads();
                for (int idx = tid; idx < nvalid * pr.nctrl; idx += NT) {
                    int s = idx / pr.nctrl, c = idx % pr.nctrl;
                    A.out_c[((s0 + s) * pr.nctrl + c) * ntp1 + (k + 1)] = control_value<C, real>(pr, tp, c, s);
                }
                __syncthreads();
            }
        }

        // ---------------------------------------------------------------- terminal block (OCflow.py:58-90)
        __syncthreads();
        stage_input_from<C, real>(tp, zb + Z0, d, A.t_end, tid);
        __syncthreads();
        ws = phi_chain<C, real, true>(tm, ws);
        __syncthreads();
        const real* xt = static_cast<const real*>(pr.xtarget);
        for (int s = tid; s < TS; s += NT) {
            real cG = real(0), hjg = real(0);
            for (int r = 0; r < d; ++r) {
                real res = zb[Z0 + r * TSP + s] - xt[r];
                cG = r_fma(res, res, cG);
                hjg += r_abs(sm[tp.G + r * TSP + s] - A.alph0 * res);
            }
            cG = real(0.5) * cG;
            real phiN = real(0), quad = real(0), lin = real(0);
            for (int k = 0; k < C::NWO; ++k) phiN += sm[tp.PN + k * TSP + s];
            for (int o = 0; o < D; ++o) {
                real sv = sm[tp.S + o * TSP + s];
                quad = r_fma(sv, sm[tp.Qs + o * TSP + s], quad);
                lin = r_fma(wscalar<C>(P, tp, P.off_cw + o), sv, lin);
            }
            real phi1 = phiN + real(0.5) * quad + (lin + wscalar<C>(P, tp, P.off_cb));
            real* sc = sm + tp.SC + s;
            sc[0 * TSP] = zb[Z0 + d * TSP + s];               // L
            sc[1 * TSP] = cG;                                 // G
            sc[2 * TSP] = zb[Z0 + (d + 1) * TSP + s];         // HJt
            sc[3 * TSP] = r_abs(phi1 - A.alph0 * cG);         // HJfin
            sc[4 * TSP] = hjg;                                // HJgrad
            sc[5 * TSP] = zb[Z0 + (d + 2) * TSP + s];         // Q
            sc[6 * TSP] = zb[Z0 + (d + 3) * TSP + s];         // W
        }
        __syncthreads();
        if (A.mode == 0) {
            if (tid < 7) for (int s = 0; s < nvalid; ++s) csum += (double)sm[tp.SC + tid * TSP + s];
            cnt += nvalid;
        } else if (A.mode == 1) {
            for (int s = tid; s < nvalid; s += NT) {
                const real* sc = sm + tp.SC + s;
                real L = sc[0], Gc = sc[TSP], HJt = sc[2 * TSP], HJf = sc[3 * TSP], HJg = sc[4 * TSP];
                real* o = A.out_a + (s0 + s) * 8;
                o[0] = L + A.alph0 * Gc + A.alph3 * HJt + A.alph4 * HJf + A.alph5 * HJg;   // OCflow.py:75
                o[1] = L; o[2] = Gc; o[3] = HJt; o[4] = HJf; o[5] = HJg; o[6] = sc[5 * TSP]; o[7] = sc[6 * TSP];
            }
        }
        __syncthreads();
    }

    if (C::WSTREAM) cp_async_wait_pending(0);     // prefetched slabs nobody will consume
    if (kmode == KMODE_ROLLOUT && A.mode == 0 && A.partials) {
        if (tid < 7) A.partials[blockIdx.x * 8 + tid] = csum;
        if (tid == 7) A.partials[blockIdx.x * 8 + 7] = (double)cnt;
    }
}

// ------------------------------------------------------------------------------------------------
// weight packing: reference layout (nn.Linear, y = x W' + b) -> K-major, output-permuted, zero-padded
// ------------------------------------------------------------------------------------------------
// (the host zero-fills the blob with cudaMemsetAsync first: padding columns are zero weights)
template <class C, typename real>
__global__ void pack_phi_kernel(const PhiRaw<real> R, const PhiPack<real> P, real* __restrict__ blob) {
    const int D = P.D, m = P.m, nTh = P.nTh, r = P.r;
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    // W1[k][pc(o)] = K0[o][k]            W4[j][pc(o)] = K0[j][o]
    for (int i = t0; i < m * D; i += stride) {
        int o = i / D, k = i % D;
        real v = R.K[0][i];
        blob[P.off_W1 + k * P.Npm + pack_col<C>(o)] = v;
        blob[P.off_W4 + o * P.Npd + pack_col<C>(k)] = v;
    }
    // Kf_i[k][pc(o)] = K_i[o][k]         Kr_i[j][pc(k)] = K_i[j][k]
    for (int l = 1; l < nTh; ++l)
        for (int i = t0; i < m * m; i += stride) {
            int o = i / m, k = i % m;
            real v = R.K[l][i];
            blob[P.off_Kf[l] + k * P.Npm + pack_col<C>(o)] = v;
            blob[P.off_Kr[l] + o * P.Npm + pack_col<C>(k)] = v;
        }
    // sym[k][pc(o)] = sum_q A[q][k] A[q][o]   (A'A, Phi.py:110)
    for (int i = t0; i < D * D; i += stride) {
        int k = i / D, o = i % D;
        real s = real(0);
        for (int q = 0; q < r; ++q) s = r_fma(R.A[q * D + k], R.A[q * D + o], s);
        blob[P.off_sym + k * P.Npd + pack_col<C>(o)] = s;
    }
    for (int l = 0; l < nTh; ++l)
        for (int i = t0; i < m; i += stride) blob[P.off_b[l] + i] = R.b[l][i];
    for (int i = t0; i < m; i += stride) blob[P.off_w + i] = R.w[i];
    for (int i = t0; i < D; i += stride) blob[P.off_cw + i] = R.c_w[i];
    if (t0 == 0) blob[P.off_cb] = R.c_b[0];
}

}  // namespace noc
