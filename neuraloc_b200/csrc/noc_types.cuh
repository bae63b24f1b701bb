// noc_types.cuh — plain-old-data descriptors shared by the host API (noc_api.cu) and the kernels.
//
// Vocabulary follows the reference: Phi (value network, src/Phi.py), prob (src/problem/*.py),
// z = [x, L, HJt, Q, W] (the augmented state OCflow integrates, src/OCflow.py:33,104-140).
#pragma once
#include <cstdint>

namespace noc {

constexpr int MAXL = 8;        // max nTh (ResNet layers) the packed descriptor carries

// Packed value-network weights ("blob") as the kernels consume them.  All offsets are in elements of
// `real` from `blob`.  Weight matrices are stored K-major ([in][out_packed]) so that one thread reads
// its RO output columns for a fixed input k with one vector load; `out_packed` is the permutation
// described in noc_rollout.cuh (pack_col), padded to Npm / Npd columns with zeros.
template <typename real>
struct PhiPack {
    int d, D, m, nTh, r;
    int Npm, Npd;              // padded packed widths for m-wide and D-wide outputs
    real h;                    // ResNet step (Phi.py:38)
    const real* blob;
    int off_W1;                // [D][Npm]   opening layer, forward:  o = K0 s
    int off_Kf[MAXL];          // [m][Npm]   layer i forward:         a_i = K_i u_{i-1}
    int off_Kr[MAXL];          // [m][Npm]   layer i reverse:         K_i' y
    int off_W4;                // [m][Npd]   opening layer reverse:   K0' v
    int off_sym;               // [D][Npd]   A'A (Phi.py:110)
    int off_b[MAXL];           // [m] each
    int off_w, off_cw, off_cb; // [m], [D], [1]
    int blob_len;
    // The weight matrices in the order one grad-Phi evaluation consumes them (W1, Kf_1.., Kr_nTh-1..1, sym, W4):
    // they are laid out contiguously in that order, so a streamed configuration prefetches "the next slab" without
    // caring about contraction or stage boundaries (noc_rollout.cuh: WStream).
    int nseq;
    int seq_off[2 * MAXL + 1], seq_N[2 * MAXL + 1], seq_K[2 * MAXL + 1];
    int ntile_d, ksplit;       // D-wide contractions (sym, W4) in streamed configs: output tiles, and the most K-slices a tile gets
};

// Raw (reference-layout) pointers handed to the pack kernel.
template <typename real>
struct PhiRaw {
    const real* A; const real* c_w; const real* c_b; const real* w;
    const real* K[MAXL]; const real* b[MAXL];
};

struct ProbPack {
    int kind, obstacle, training, nAgents, agentDim, nctrl;
    double alph_Q, alph_W, r, mass, grav;
    double cutW;               // interaction cut-off: 2r (eval) / 2.2r / 3.2r (train) — Cross2D.py:139-155, SwarmTraj.py:140-156
    const void* xtarget;       // dev [d]
};

// Row offsets (in rows of TSP elements) of every shared-memory array of a tile; computed on the host
// (noc_api.cu: plan_smem) so that aliasing decisions live in one place.
struct SmemPlan {
    int U, U2;                 // current hidden vector (u_i, then y / v in the reverse sweep); U2 == U when in place
    int T[MAXL];               // tanh(a_i), i = 0..nTh-2 (the last layer's tanh is folded into y)
    int Zb;                    // z_{i+1} of the reverse sweep (nTh > 2 only)
    int S;                     // s = [x_stage, t]   (D rows)
    int G;                     // grad Phi           (D rows; may alias U)
    int Qs;                    // A'A s kept apart in the terminal pass (aliases T[0])
    int Z0, ZA;                // augmented state at step start / RK accumulation (d+4 rows each)
    int SC;                    // per-sample scalars: L, |Phi_t - H|, Q, W, + terminal costs
    int RED;                   // [3][TPS] partial sums of the problem phase
    int PN;                    // [NWO*WO] partial sums of w . u_last (terminal pass)
    int QX;                    // Quadcopter per-agent scalars [5 * nAgents]: u/mass, f7, f8, f9, u
    int GP;                    // partial sums of the K-split D-wide contractions (streamed configs)
    int rows;                  // total rows
    int wsm_off;               // element offset of the weight blob copy (WSMEM configs) / of the warps' rings (streamed)
    int ring_slab, ring_ns;    // streamed configs: elements per ring slot (GRP rows x WB columns), slots per warp (2..8)
    int z_global;              // Z0/ZA live in a per-CTA global scratch instead of shared memory
};

constexpr int SC_L = 0, SC_HJ = 1, SC_Q = 2, SC_W = 3, SC_G = 4, SC_HJF = 5, SC_HJG = 6, SC_ROWS = 8;

template <typename real>
struct RolloutArgs {
    PhiPack<real> phi;
    ProbPack prob;
    SmemPlan sp;
    const real* x;             // [n, d]   (phi-eval: [n, D]; prob-eval: x and p)
    const real* p_in;          // prob-eval only
    long long n;
    int nt, stepper, mode;
    const double* times;       // dev [nt*5]: t_a, t_mid, t_b, t_ctrl, h'   (noc_stage_times)
    real alph0, alph3, alph4, alph5;
    real t_end;                // tspan[1] rounded to real (OCflow.py:62)
    double* partials;          // [gridDim.x][8] per-CTA cost sums (mean mode)
    real* out_a;               // noMean: [n,8]; phi-eval: phi [n];  prob-eval: lhqw [n,4]
    real* out_b;               // intermediates: zFull; phi-eval: grad [n,D]; prob-eval: gradpH [n,d]
    real* out_c;               // intermediates: ctrlFull; prob-eval: ctrls [n,nctrl]
    real* zscratch;            // per-CTA [2][(d+4)][TSP] augmented-state scratch when sp.z_global
    int zstride;               // elements per CTA in zscratch
    int ntiles;
};

enum { KMODE_ROLLOUT = 0, KMODE_PHI = 1, KMODE_PROB = 2 };

}  // namespace noc
