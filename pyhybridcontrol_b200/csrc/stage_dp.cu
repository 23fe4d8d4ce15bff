// stage_dp.cu -- K3s/K4s: exact mixed-integer solve for MLD models with a SCALAR state, by a value-table
// bound plus an exact forward search.  Replaces the cvxpy -> Gurobi/CPLEX call inside
// ConstraintSolvedController.solve (reference: controllers/controller_base.py:509-512) for the problem class
// the reference's own example poses for every water heater (examples/residential_mg_with_pv_and_dewhs/
// modelling/micro_grid_models.py:27-100: nx = 1, one binary input, slack-softened box constraints, linear cost).
//
// Problem class (per agent; v(k) = [binaries (nb = nu + ndelta); mu (nmu = nc or 0)], k = 0..Nt-1):
//     minimise  sum_k  cu_k' alpha_k + q_k' mu_k
//     s.t.      e_i p_k + f_i' alpha_k - d_i mu_k,i <= rhs_k,i ,   mu >= 0 ,   alpha_k in {0,1}^nb (within lb/ub)
//               p_0 = 0 ,  p_k+1 = a p_k + g' alpha_k            (p = forced response = row k of Gamma_v times v)
// with a = A, g = [B1 B2], e = E + G C, f = [F1 F2] + G [D1 D2], d_i = -Psi_ii (0: hard row), rhs from K2.
// This is exactly  min c'v  s.t.  H_v v <= rhs  of the condensed form (mld_evolution_matrices.py:237-240), whose
// continuous part has the closed form  mu_k,i = max(0, e_i p_k + f_i' alpha_k - rhs_k,i) / d_i.
//
// Algorithm (round 2)
//   * scaled state s_k = p_k / a^k turns the recursion into a pure translation s_k+1 = s_k + beta_k(alpha),
//     beta_k = g' alpha / a^(k+1).  All stages share ONE lattice of cells of width w (absolute cell a holds the
//     states [a w, (a+1) w)), so the "no input" action maps a cell onto a cell exactly.
//   * MOVING WINDOW: stage k tabulates the G absolute cells [off_k, off_k + G), off_k following the violation-free
//     band of the stage (the band drifts with the free response; one global window over a long horizon wastes most of
//     its cells).  Everything below / above the window is one SEMI-INFINITE cell each, whose bound is valid for all
//     of its states (penalty at its favourable edge, minimum over every next-stage cell its image can touch), so a
//     state that leaves the window never falls back to the trivial bound.
//   * kernel 1 (stage_dp_table_kernel, one CTA per agent): backward sweep of a LOWER BOUND of the cost-to-go per
//     cell.  Two cell formats: a constant per cell (default), or a LINE per cell (value at the left and right edge,
//     cells need not agree at their common edge) built from the lower convex hull of the translated pieces -- exact
//     where slack penalties make the cost-to-go steep (full-horizon robust constraint sets), and never below the
//     constant-cell bound.  The two stage buffers are exactly 2 G cells of shared memory (no guard cells: the two
//     semi-infinite bounds of a stage are computed by a helper warp next to the sweep); only the stages the search can
//     read (k = D, 2D, ...: the depth of one search expansion) leave the SM, each through one TMA bulk copy.
//   * search: exact depth-first search over the binary sequence in time order -- states and costs are exact FP64 --
//     pruned by cost so far + table bound >= min(incumbent, T), where the threshold T starts just above the root bound
//     and grows whenever the tree below it is exhausted without a solution (iterative deepening on the bound: a poor
//     first dive can no longer trap the search in a bad subtree).  The unit of work is the depth-D subtree under an
//     open node (32 lanes = 32 action sequences, one table read each).  One warp searches an agent first (a descent is
//     sequential); what that leaves is searched by a team of warps, several open nodes per round.  Up to 296 agents
//     all of it is the tail of kernel 1 (ONE launch per solve); above, kernel 1 is followed by stage_dp_solo_kernel
//     (one warp per agent) and stage_dp_search_kernel (one 16-warp CTA per agent still pending).
// Optional convex stage terms (quadratic / L1 atoms on the state, outputs and slacks) make the problem an MIQP; they
// run through the general loops of both kernels.
#include <limits.h>
#include <string.h>
#include "common.cuh"

namespace hmpc {

constexpr int kTableThreads = 512;     // sweep threads of the table kernel (a thread owns the cells n * 512 + tid)
constexpr int kTableShift = 9;        // log2(kTableThreads)
constexpr int kTableBlock = kTableThreads + 32;   // + one helper warp (semi-infinite cells, off the sweep's critical path)
constexpr int kDpMaxNb = 4;
constexpr int kDpMaxAct = 1 << kDpMaxNb;
constexpr int kDpMaxNc = 8;
constexpr int kDpMaxNt = 128;
constexpr int kDpMaxT = 4;           // extra state terms per stage
constexpr int kSearchWarps = 16;      // warps of the search kernel's CTA (one agent per CTA)
constexpr int kStackCap = 2048;       // open nodes per agent in the wide search kernel
constexpr int kSoloStack = 896;       // ... per warp in the one-warp search kernel, at most (less when shared memory is short)
constexpr int kSoloWarps = 4;         // agents per CTA in the one-warp search kernel
constexpr int kSoloTail = 48;         // expansions beyond one descent that warp 0 spends alone in the fused tail before the CTA joins in
constexpr int kSplit = 8;             // CTAs that share one hard agent's tree in the team-search kernel (by the root's children)
constexpr int kSplitWords = 3 + 6 * kSplit;   // per-agent scratch of the split, in 8-byte words
constexpr int kSoloBudget = 128;      // ... and a warp of the one-warp kernel before it hands the agent to the wide kernel
constexpr int kPending = -1;          // status of an agent that kernel 2 still has to search
constexpr double kEdgeEps = 1e-9;     // cell-boundary guard (fraction of a cell)
constexpr double kLinSlop = 4e-15;    // relative floating-point slop subtracted per stage from a linear cell

enum { FMT_F32 = 0, FMT_F64 = 1, FMT_LIN = 2 };

struct DpArgs {
    hmpc_dims d;
    const double* mats[HMPC_NUM_MATS];
    int64_t stride[HMPC_NUM_MATS];
    const double* rhs;                 // [B, Nt*nc]
    const double* cost; int64_t sc;    // [B|1, Nt*nv]
    const double* lb; const double* ub;  // [Nt*nv] shared by the batch
    const uint8_t* is_bin;             // [Nt*nv]
    hmpc_stage_dp_opts o;
    hmpc_stage_terms t;                // optional convex state / slack terms (T == 0 and qmu == NULL: none)
    int G, nb, nact, nv, T;
    int D, nstore, fmt, fuse;          // search depth per expansion, stored stages (k = D, 2D, ...), cell format
    int solo_cap;                      // open nodes per warp in the one-warp search kernel
    long long buf_bytes;               // shared memory behind the plan (stage buffers / set-up staging / search stack)
    void* table;                       // [B, nstore, G] cells
    double* pblk;                      // [B, plan doubles]: the agent's stage data, handed from kernel 1 to kernel 2
    unsigned long long* split;         // [B, kSplitWords]: shared incumbent key, arrival counter, one result per part
    int* pend;                         // [2 + B]: number of agents the one-warp kernel left to the team kernel, cursor of the
                                       // team kernel's work queue, the list
    double* v; double* obj; int32_t* status; int32_t* stats;
};

__host__ __device__ inline int fmt_bytes(int fmt) { return fmt == FMT_LIN ? 16 : (fmt == FMT_F64 ? 8 : 4); }

// per-agent stage data in shared memory (all doubles; small integers are stored as doubles or packed ints so that
// the block can be handed to the search kernel with one coalesced copy)
struct DpPlan {
    int ak, iak, cu, qs, rhs, tailmin, amask, e, dscale, galpha, falpha, misc, off, loinf, hiinf, nd;   // persistent
    int x_eak, x_foff, x_hq, x_ca, x_shift;
    int t_h, t_ga, t_r, t_wq, t_w1, qq;
    int scr;                                                                      // load-time scratch [9*Nt]
    int mst;                                                                      // staged MLD blocks [11][64]
    int sc_ca, sc_base, sc_slope, sc_fr, sc_i0, sc_span, sc_flags, sc_z, sc_semi, sc_semd, sc_blk;   // per-stage sweep constants
    int total;
};

constexpr int kMstMats = 11, kMstElems = 64;

__host__ __device__ inline DpPlan make_dp_plan(int Nt, int nb, int nc, int T = kDpMaxT) {
    DpPlan p;
    const int nact = 1 << nb;
    int o = 0;
    auto take = [&](int n) { int r = o; o += n; return r; };
    p.misc = take(16);
    p.ak = take(Nt + 2); p.iak = take(Nt + 2); p.cu = take(Nt * nb); p.qs = take(Nt * nc); p.rhs = take(Nt * nc);
    p.tailmin = take(Nt + 1); p.amask = take(Nt);
    p.e = take(nc); p.dscale = take(nc); p.galpha = take(nact); p.falpha = take(nc * nact);
    p.off = take((Nt + 3) / 2); p.loinf = take(Nt + 1); p.hiinf = take(Nt + 1);
    // per-stage constants of the forward search: viol_i = x_eak[k][i] * s + x_foff[k][i][al], penalty weight / 2,
    // action cost, translation of s
    p.x_eak = take(Nt * nc); p.x_foff = take(Nt * nc * nact); p.x_hq = take(Nt * nc); p.x_ca = take(Nt * nact);
    p.x_shift = take(Nt * nact);
    // optional convex terms: tau = t_h p + t_ga[al] + t_r[k];  cost += t_wq tau^2 + t_w1 |tau|;  slack: qq v^2
    p.t_h = take(T); p.t_ga = take(T * nact); p.t_r = take(Nt * T); p.t_wq = take(Nt * T); p.t_w1 = take(Nt * T);
    p.qq = take(Nt * nc);
    p.nd = (o + 1) & ~1;
    o = p.nd;
    p.scr = take(9 * Nt);
    p.mst = take(kMstMats * kMstElems);
    const int nc_ = nc > 0 ? nc : 1;
    p.sc_ca = take(Nt * nact); p.sc_base = take(Nt * nc_ * nact); p.sc_slope = take(Nt * nc_); p.sc_fr = take(Nt * nact);
    p.sc_i0 = take((Nt * nact + 1) / 2); p.sc_span = take((Nt * nact + 1) / 2); p.sc_flags = take((Nt + 1) / 2);
    p.sc_z = take(Nt + 1); p.sc_semi = take(4 * Nt); p.sc_semd = take(4 * Nt); p.sc_blk = take(8 * Nt);
    p.total = (o + 1) & ~1;
    return p;
}

enum { MISC_W = 0, MISC_INVW, MISC_A, MISC_FLAG, MISC_SIMPLE, MISC_TERMS, MISC_LINOK, MISC_INC_OBJ, MISC_INC_P0,
       MISC_INC_P1, MISC_NODES0, MISC_IMPR0, MISC_KINS, MISC_T_SETUP, MISC_T_SWEEP, MISC_T_SEARCH };

struct DpCtx {
    int Nt, nb, nc, nact, nmu, nv, G;
    double *ak, *iak, *cu, *qs, *rhs, *tailmin, *amask, *e, *dscale, *galpha, *falpha, *misc, *scr, *mst;
    double *loinf, *hiinf; int* off;
    double *sc_ca, *sc_base, *sc_slope, *sc_fr, *sc_semd; int *sc_i0, *sc_span, *sc_flags, *sc_z, *sc_semi, *sc_blk;
    double *x_eak, *x_foff, *x_hq, *x_ca, *x_shift;
    double *t_h, *t_ga, *t_r, *t_wq, *t_w1, *qq;
    int T;
    double feas_tol;
};

__device__ inline DpCtx bind_ctx(const DpArgs& A, unsigned char* smem) {
    DpCtx c;
    c.Nt = A.d.Nt; c.nb = A.nb; c.nc = A.d.nc; c.nact = A.nact; c.nmu = A.d.nmu; c.nv = A.nv; c.G = A.G;
    c.T = A.T;
    const DpPlan p = make_dp_plan(c.Nt, c.nb, c.nc, c.T);
    double* sd = reinterpret_cast<double*>(smem);
    c.ak = sd + p.ak; c.iak = sd + p.iak; c.cu = sd + p.cu; c.qs = sd + p.qs; c.rhs = sd + p.rhs;
    c.tailmin = sd + p.tailmin; c.amask = sd + p.amask; c.e = sd + p.e; c.dscale = sd + p.dscale;
    c.galpha = sd + p.galpha; c.falpha = sd + p.falpha; c.misc = sd + p.misc; c.scr = sd + p.scr; c.mst = sd + p.mst;
    c.off = reinterpret_cast<int*>(sd + p.off); c.loinf = sd + p.loinf; c.hiinf = sd + p.hiinf;
    c.sc_ca = sd + p.sc_ca; c.sc_base = sd + p.sc_base; c.sc_slope = sd + p.sc_slope; c.sc_fr = sd + p.sc_fr;
    c.sc_i0 = reinterpret_cast<int*>(sd + p.sc_i0); c.sc_span = reinterpret_cast<int*>(sd + p.sc_span);
    c.sc_flags = reinterpret_cast<int*>(sd + p.sc_flags); c.sc_z = reinterpret_cast<int*>(sd + p.sc_z);
    c.sc_semi = reinterpret_cast<int*>(sd + p.sc_semi); c.sc_semd = sd + p.sc_semd;
    c.sc_blk = reinterpret_cast<int*>(sd + p.sc_blk);
    c.x_eak = sd + p.x_eak; c.x_foff = sd + p.x_foff; c.x_hq = sd + p.x_hq; c.x_ca = sd + p.x_ca; c.x_shift = sd + p.x_shift;
    c.t_h = sd + p.t_h; c.t_ga = sd + p.t_ga; c.t_r = sd + p.t_r; c.t_wq = sd + p.t_wq; c.t_w1 = sd + p.t_w1; c.qq = sd + p.qq;
    c.feas_tol = A.o.feas_tol;
    return c;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ const double* mat_of(const DpArgs& A, int which, int b) {
    return A.mats[which] ? A.mats[which] + (int64_t)b * A.stride[which] : nullptr;
}

// Block-cooperative load of one agent's stage data (table kernel).  On return misc[MISC_FLAG] != 0 marks an
// agent outside the supported class.
__device__ inline void dp_load(const DpArgs& A, int b, DpCtx& c, double* stage) {
    const int Nt = c.Nt, nb = c.nb, nc = c.nc, nv = c.nv, nu = A.d.nu, nmu = c.nmu, nact = c.nact;
    const int tid = threadIdx.x, nthr = blockDim.x;
    {   // all MLD blocks of this agent in one round of loads (missing blocks are zero)
        const int which[kMstMats] = {HMPC_A, HMPC_B1, HMPC_B2, HMPC_C, HMPC_D1, HMPC_D2, HMPC_E, HMPC_F1, HMPC_F2, HMPC_G, HMPC_Psi};
        const int ny_ = A.d.ny, nd_ = A.d.ndelta;
        const int cnt[kMstMats] = {1, nu, nd_, ny_, ny_ * nu, ny_ * nd_, nc, nc * nu, nc * nd_, nc * ny_, nc * nmu};
        for (int i = tid; i < kMstMats * kMstElems; i += nthr) {
            const int mi = i / kMstElems, e = i - mi * kMstElems;
            const double* src = mat_of(A, which[mi], b);
            c.mst[i] = (src && e < cnt[mi]) ? src[e] : 0.0;
        }
    }
    // every per-stage input of the agent (cost row, right-hand side, bounds, integrality) in the same round of loads:
    // one trip to HBM / L2 instead of one per use in the stage loop below
    double* g_cost = stage; double* g_lb = g_cost + Nt * nv; double* g_ub = g_lb + Nt * nv; double* g_rhs = g_ub + Nt * nv;
    double* g_bin = g_rhs + Nt * nc;
    {
        const double* cost_g = A.cost + (int64_t)b * A.sc;
        const double* rhs_g = A.rhs + (int64_t)b * Nt * nc;
        for (int i = tid; i < Nt * nv; i += nthr) {
            g_cost[i] = cost_g[i]; g_lb[i] = A.lb[i]; g_ub[i] = A.ub[i]; g_bin[i] = A.is_bin[i] ? 1.0 : 0.0;
        }
        for (int i = tid; i < Nt * nc; i += nthr) g_rhs[i] = rhs_g[i];
    }
    __syncthreads();
    // a^k (the reference's A_pow_tilde, mld_evolution_matrices.py:266-272) by binary powers, one stage per thread
    for (int k = tid; k <= Nt + 1; k += nthr) {
        double r = 1.0, pw = c.mst[0];
        for (int e = k; e > 0; e >>= 1) { if (e & 1) r *= pw; pw *= pw; }
        c.ak[k] = r; c.iak[k] = 1.0 / r;
    }
    if (tid == 0) {
        int flag = 0;
        const double* Am = c.mst;
        const double a = Am[0];
        c.misc[MISC_A] = a;
        {   // a^Nt must stay within three decades (same product, computed here so that thread 0 need not wait)
            double r = 1.0, pw = a;
            for (int e = Nt; e > 0; e >>= 1) { if (e & 1) r *= pw; pw *= pw; }
            if (!(a > 0.0) || !(r > 1e-3) || !(r < 1e3)) flag = 1;
        }
        const double* B1 = c.mst + 1 * kMstElems; const double* B2 = c.mst + 2 * kMstElems;
        const double* Cm = c.mst + 3 * kMstElems;
        const double* D1 = c.mst + 4 * kMstElems; const double* D2 = c.mst + 5 * kMstElems;
        const double* E = c.mst + 6 * kMstElems;
        const double* F1 = c.mst + 7 * kMstElems; const double* F2 = c.mst + 8 * kMstElems;
        const double* Gm = c.mst + 9 * kMstElems; const double* Psi = c.mst + 10 * kMstElems;
        const int ny = A.d.ny, nd = A.d.ndelta;
        double g[kDpMaxNb], f[kDpMaxNc][kDpMaxNb];
        for (int j = 0; j < nb; ++j) g[j] = j < nu ? B1[j] : B2[j - nu];
        for (int i = 0; i < nc; ++i) {
            double ei = E[i];                                 // E is [nc, nx = 1]
            for (int j = 0; j < nb; ++j) f[i][j] = j < nu ? F1[i * nu + j] : F2[i * nd + (j - nu)];
            for (int r = 0; r < ny; ++r) {                    // y = C x + D1 u + D2 delta (+ terms already in rhs)
                const double gir = Gm[i * ny + r];
                if (gir == 0.0) continue;
                ei += gir * Cm[r];
                for (int j = 0; j < nb; ++j) f[i][j] += gir * (j < nu ? D1[r * nu + j] : D2[r * nd + (j - nu)]);
            }
            c.e[i] = ei;
            double di = 0.0;
            if (nmu) {
                for (int j = 0; j < nmu; ++j) {
                    const double pij = Psi[i * nmu + j];
                    if (j == i) di = -pij; else if (pij != 0.0) flag = 1;   // every row owns at most its own slack
                }
                if (di < 0.0) flag = 1;
            }
            c.dscale[i] = di;
        }
        for (int al = 0; al < nact; ++al) {
            double ga = 0.0;
            for (int j = 0; j < nb; ++j) if (al >> j & 1) ga += g[j];
            c.galpha[al] = ga;
            for (int i = 0; i < nc; ++i) {
                double fa = 0.0;
                for (int j = 0; j < nb; ++j) if (al >> j & 1) fa += f[i][j];
                c.falpha[i * nact + al] = fa;
            }
        }
        for (int t = 0; t < c.T; ++t) {
            c.t_h[t] = A.t.h[(int64_t)b * A.t.h_stride_b + t];
            for (int al = 0; al < nact; ++al) {
                double ga = 0.0;
                if (A.t.ga) for (int j = 0; j < nb; ++j) if (al >> j & 1) ga += A.t.ga[(int64_t)b * A.t.ga_stride_b + t * nb + j];
                c.t_ga[t * nact + al] = ga;
            }
        }
        c.misc[MISC_FLAG] = (double)flag;
        c.misc[MISC_TERMS] = (c.T > 0 || A.t.qmu) ? 1.0 : 0.0;
        c.misc[MISC_INC_OBJ] = INFINITY; c.misc[MISC_INC_P0] = 0.0; c.misc[MISC_INC_P1] = 0.0;
        c.misc[MISC_NODES0] = 0.0; c.misc[MISC_IMPR0] = 0.0; c.misc[MISC_T_SEARCH] = 0.0;
    }
    __syncthreads();
    // ---- per-stage data, one stage per thread
    const double* cost = g_cost;
    const double* rhs = g_rhs;
    double* s_lo = c.scr; double* s_hi = c.scr + Nt; double* s_smin = c.scr + 2 * Nt; double* s_smax = c.scr + 3 * Nt;
    double* s_cmin = c.scr + 4 * Nt; double* s_marg = c.scr + 5 * Nt;
    int bad = 0;
    for (int k = tid; k < Nt; k += nthr) {
        int mask = 0;
        for (int al = 0; al < nact; ++al) {
            bool ok = true;
            for (int j = 0; j < nb; ++j) {
                const double bit = (double)(al >> j & 1);
                if (bit < g_lb[k * nv + j] || bit > g_ub[k * nv + j]) ok = false;
            }
            if (ok) mask |= 1 << al;
        }
        c.amask[k] = (double)mask;
        for (int j = 0; j < nb; ++j) { c.cu[k * nb + j] = cost[k * nv + j]; if (g_bin[k * nv + j] == 0.0) bad = 1; }
        for (int i = 0; i < nc; ++i) {
            c.rhs[k * nc + i] = rhs[k * nc + i];
            double qs = INFINITY;                                // hard row
            if (nmu) {
                const int col = k * nv + nb + i;
                const double q = cost[col], di = c.dscale[i], ubm = g_ub[col], lbm = g_lb[col];
                if (g_bin[col] != 0.0 || lbm > 0.0 || (lbm < 0.0 && di > 0.0)) bad = 1;
                if (di > 0.0 && ubm > 0.0) {
                    if (isfinite(ubm) || q < 0.0) bad = 1;      // bounded or rewarded slack: not this class
                    qs = q / di;
                } else if (q < 0.0 && ubm > 0.0) bad = 1;         // free column with negative cost: unbounded
            }
            c.qs[k * nc + i] = qs;
            // quadratic slack price: Qmu mu^2 = (Qmu / d^2) max(0, violation)^2
            double qq = 0.0;
            if (A.t.qmu && nmu) {
                const double Q = A.t.qmu[(int64_t)b * A.t.qmu_stride_b + k * nc + i], di = c.dscale[i];
                if (Q < 0.0) bad = 1;
                if (Q > 0.0) { if (isinf(qs)) qq = 0.0; else qq = Q / (di * di); }
            }
            c.qq[k * nc + i] = qq;
        }
        for (int t = 0; t < c.T; ++t) {
            c.t_r[k * c.T + t] = A.t.r[((int64_t)b * Nt + k) * c.T + t];
            const double wq = A.t.wq ? A.t.wq[(int64_t)b * A.t.wq_stride_b + k * c.T + t] : 0.0;
            const double w1 = A.t.w1 ? A.t.w1[(int64_t)b * A.t.w1_stride_b + k * c.T + t] : 0.0;
            if (wq < 0.0 || w1 < 0.0) bad = 1;          // concave term: not a convex stage cost
            c.t_wq[k * c.T + t] = wq; c.t_w1[k * c.T + t] = w1;
        }
        // shifts, cheapest action, violation-free band of this stage (all in s = p / a^k)
        const double ik = 1.0 / c.ak[k], ik1 = 1.0 / c.ak[k + 1];
        double smin = 0.0, smax = 0.0, cmin = INFINITY, marg = 0.0;
        bool any = false;
        for (int al = 0; al < nact; ++al) if (mask >> al & 1) {
            const double sh = c.galpha[al] * ik1;
            smin = any ? fmin(smin, sh) : sh; smax = any ? fmax(smax, sh) : sh; any = true;
            marg = fmax(marg, fabs(sh));
            double ca = 0.0;
            for (int j = 0; j < nb; ++j) if (al >> j & 1) ca += cost[k * nv + j];
            cmin = fmin(cmin, ca);
        }
        double lo_k = -INFINITY, hi_k = INFINITY;
        for (int i = 0; i < nc; ++i) {
            const double ei = c.e[i];
            if (ei == 0.0) continue;
            double fmin_a = INFINITY;
            for (int al = 0; al < nact; ++al) if (mask >> al & 1) fmin_a = fmin(fmin_a, c.falpha[i * nact + al]);
            if (!isfinite(fmin_a)) fmin_a = 0.0;
            const double lim = (rhs[k * nc + i] - fmin_a) / ei * ik;
            if (ei > 0.0) hi_k = fmin(hi_k, lim); else lo_k = fmax(lo_k, lim);
        }
        s_lo[k] = lo_k; s_hi[k] = hi_k; s_smin[k] = smin; s_smax[k] = smax; s_cmin[k] = cmin; s_marg[k] = marg;
    }
    if (bad) c.misc[MISC_FLAG] = 1.0;   // benign race: every writer stores the same value
    __syncthreads();
    // ---- grid window and trivial bound, one stage per thread (each thread sums / scans its own prefix: O(Nt) steps of
    // independent loads, three short phases):
    //   tailmin_k  = sum_{i>=k} min(cmin_i, 0): trivial bound on the cost-to-go from ANY state
    //   moving window: stage k covers [band_k - 2 shifts, band_k + 1 shift], the band clamped to what is reachable at
    //   that stage and slew-limited -- the state climbs by at most the largest shift per stage and falls by at most
    //   the most negative one, so after a jump of the band the window follows the states that are catching up, not the
    //   band itself:  env_lo_k = min(lo_k, env_lo_k-1 + smax_k-1) = P_k + min_{j<=k} (lo_j - P_j),  P = prefix sum of
    //   smax (likewise env_hi with the most negative shifts).  Every stage has the same width (the widest of them, at
    //   least 5.4 shifts).
    // One warp does all of it with shuffle scans (lane l owns the CH consecutive stages l*CH ...; kDpMaxNt = 128: CH <= 4):
    // thread-per-stage loops over the prefixes cost 19 k cycles of a 53 k-cycle set-up, and six barriers.
    if (tid < 32) {
        const int lane = tid, CH = (Nt + 31) / 32;
        const unsigned full = 0xffffffffu;
        double v_smin[4], v_smax[4], v_cmin[4], v_lo[4], v_hi[4];
        double margin = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = lane * CH + i;
            const bool in = i < CH && k < Nt;
            v_smin[i] = in ? s_smin[k] : 0.0; v_smax[i] = in ? s_smax[k] : 0.0; v_cmin[i] = in ? s_cmin[k] : 0.0;
            v_lo[i] = in ? s_lo[k] : INFINITY; v_hi[i] = in ? s_hi[k] : -INFINITY;
            margin = fmax(margin, in ? s_marg[k] : 0.0);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) margin = fmax(margin, __shfl_xor_sync(full, margin, o));
        // exclusive prefix sums: rlo / rhi (reachable states), pp / pm (positive / negative shifts only)
        double e_rlo[4], e_rhi[4], e_pp[4], e_pm[4];
        double t_rlo = 0.0, t_rhi = 0.0, t_pp = 0.0, t_pm = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            e_rlo[i] = t_rlo; e_rhi[i] = t_rhi; e_pp[i] = t_pp; e_pm[i] = t_pm;
            t_rlo += v_smin[i]; t_rhi += v_smax[i]; t_pp += v_smax[i] > 0.0 ? v_smax[i] : 0.0; t_pm += v_smin[i] < 0.0 ? v_smin[i] : 0.0;
        }
        double i_rlo = t_rlo, i_rhi = t_rhi, i_pp = t_pp, i_pm = t_pm;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double a1 = __shfl_up_sync(full, i_rlo, o), a2 = __shfl_up_sync(full, i_rhi, o);
            const double a3 = __shfl_up_sync(full, i_pp, o), a4 = __shfl_up_sync(full, i_pm, o);
            if (lane >= o) { i_rlo += a1; i_rhi += a2; i_pp += a3; i_pm += a4; }
        }
        const double o_rlo = i_rlo - t_rlo, o_rhi = i_rhi - t_rhi, o_pp = i_pp - t_pp, o_pm = i_pm - t_pm;
        // suffix sums of min(cmin, 0): the trivial bound on the cost-to-go
        double suf[4], t_tail = 0.0;
#pragma unroll
        for (int i = 3; i >= 0; --i) { t_tail += v_cmin[i] < 0.0 ? v_cmin[i] : 0.0; suf[i] = t_tail; }
        double i_tail = t_tail;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double a1 = __shfl_down_sync(full, i_tail, o); if (lane + o < 32) i_tail += a1; }
        const double o_tail = i_tail - t_tail;
        // band clamped to what is reachable, minus the slew; running min / max over the stages up to k
        double x_lo[4], x_hi[4], pp_[4], pm_[4], rl_[4], rh_[4];
        double r_mn = INFINITY, r_mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = lane * CH + i;
            const bool in = i < CH && k < Nt;
            rl_[i] = o_rlo + e_rlo[i]; rh_[i] = o_rhi + e_rhi[i]; pp_[i] = o_pp + e_pp[i]; pm_[i] = o_pm + e_pm[i];
            if (in) c.tailmin[k] = o_tail + suf[i];
            const double lo_c = fmin(fmax(v_lo[i], rl_[i]), rh_[i]), hi_c = fmax(fmin(v_hi[i], rh_[i]), rl_[i]);
            r_mn = in ? fmin(r_mn, lo_c - pp_[i]) : r_mn; r_mx = in ? fmax(r_mx, hi_c - pm_[i]) : r_mx;
            x_lo[i] = r_mn; x_hi[i] = r_mx;                       // inclusive within the lane
        }
        double i_mn = r_mn, i_mx = r_mx;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double a1 = __shfl_up_sync(full, i_mn, o), a2 = __shfl_up_sync(full, i_mx, o);
            if (lane >= o) { i_mn = fmin(i_mn, a1); i_mx = fmax(i_mx, a2); }
        }
        double p_mn = __shfl_up_sync(full, i_mn, 1), p_mx = __shfl_up_sync(full, i_mx, 1);      // lanes before this one
        if (lane == 0) { p_mn = INFINITY; p_mx = -INFINITY; }
        double a0[4], Wmax = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = lane * CH + i;
            const bool in = i < CH && k < Nt;
            const double env_lo = pp_[i] + fmin(p_mn, x_lo[i]), env_hi = pm_[i] + fmax(p_mx, x_hi[i]);
            a0[i] = fmax(fmin(env_lo, env_hi) - 2.0 * margin, rl_[i] - 0.01 * margin);
            double b0 = fmin(fmax(env_lo, env_hi) + margin, rh_[i] + 0.01 * margin);
            if (!(b0 > a0[i])) b0 = a0[i];
            if (in) Wmax = fmax(Wmax, b0 - a0[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Wmax = fmax(Wmax, __shfl_xor_sync(full, Wmax, o));
        // every stage has the same width: the widest of them, at least 5.4 shifts
        double W = fmax(Wmax, 5.4 * margin * (1.0 + 16.0 / (double)c.G));
        if (!(W > 0.0) || !isfinite(W)) W = 1.0;
        const double w = W / (double)c.G, invw = 1.0 / w;
        int off_last = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = lane * CH + i;
            if (i < CH && k < Nt) {
                const int o_ = (int)fmax(fmin(floor(a0[i] * invw), 1.0e9), -1.0e9);
                c.off[k] = o_;
                if (k == Nt - 1) off_last = o_;
            }
        }
        off_last = __shfl_sync(full, off_last, (Nt - 1) / CH);
        if (lane == 0) {
            c.tailmin[Nt] = 0.0;
            c.misc[MISC_W] = w; c.misc[MISC_INVW] = invw; c.misc[MISC_SIMPLE] = 1.0;
            c.loinf[Nt] = 0.0; c.hiinf[Nt] = 0.0;
            c.off[Nt] = off_last; c.off[Nt + 1] = off_last;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double action_cost(const DpCtx& c, int k, int al) {
    double ca = 0.0;
    for (int j = 0; j < c.nb; ++j) if (al >> j & 1) ca += c.cu[k * c.nb + j];
    return ca;
}

// exact stage cost of action `al` at stage k and forced response p
__device__ __forceinline__ double stage_cost(const DpCtx& c, int k, int al, double p) {
    double st = action_cost(c, k, al);
    for (int i = 0; i < c.nc; ++i) {
        const double viol = fma(c.e[i], p, c.falpha[i * c.nact + al] - c.rhs[k * c.nc + i]);
        const double qs = c.qs[k * c.nc + i];
        if (isinf(qs)) { if (viol > c.feas_tol) st = INFINITY; }
        else {
            const double v = fmax(viol, 0.0);
            st = fma(qs, v, st);
            st = fma(c.qq[k * c.nc + i] * v, v, st);
        }
    }
    for (int t = 0; t < c.T; ++t) {
        const double tau = fma(c.t_h[t], p, c.t_ga[t * c.nact + al] + c.t_r[k * c.T + t]);
        st = fma(c.t_wq[k * c.T + t] * tau, tau, st);
        st = fma(c.t_w1[k * c.T + t], fabs(tau), st);
    }
    return st;
}

// lower bound over p in [plo, phi] of the convex extra terms of stage k under action al
__device__ __forceinline__ double terms_lower_bound(const DpCtx& c, int k, int al, double plo, double phi) {
    double st = 0.0;
    for (int t = 0; t < c.T; ++t) {
        const double cst = c.t_ga[t * c.nact + al] + c.t_r[k * c.T + t];
        const double t1 = fma(c.t_h[t], plo, cst), t2 = fma(c.t_h[t], phi, cst);
        const double m = (t1 <= 0.0 && t2 >= 0.0) || (t2 <= 0.0 && t1 >= 0.0) ? 0.0 : fmin(fabs(t1), fabs(t2));
        st = fma(c.t_wq[k * c.T + t] * m, m, st);
        st = fma(c.t_w1[k * c.T + t], m, st);
    }
    return st;
}

// ---- cell formats.  F32: constant, rounded DOWN (half the shared memory; still a valid bound).  F64: constant
// (default: sequences that tie with the incumbent are pruned at gap 0).  LIN: a line (x = value at the left edge,
// y = value at the right edge).
template <int FMT> struct Cell;
template <> struct Cell<FMT_F32> {
    typedef float T;
    static __device__ __forceinline__ T pack(double v) { return __double2float_rd(v); }
    static __device__ __forceinline__ double lo(T a) { return (double)a; }
};
template <> struct Cell<FMT_F64> {
    typedef double T;
    static __device__ __forceinline__ T pack(double v) { return v; }
    static __device__ __forceinline__ double lo(T a) { return a; }
};
template <> struct Cell<FMT_LIN> {
    typedef double2 T;
    static __device__ __forceinline__ T pack(double v) { return make_double2(v, v); }
    static __device__ __forceinline__ double lo(T a) { return a.x < a.y ? a.x : a.y; }
};
__device__ __forceinline__ float cmin2(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double cmin2(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }      // (no NaN handling: two instructions less than fmin)
__device__ __forceinline__ void put_cell(float* p, double v) { *p = __double2float_rd(v); }   // rounded DOWN: still a bound
__device__ __forceinline__ void put_cell(double* p, double v) { *p = v; }

// TMA bulk copy (cp.async.bulk, shared -> global) of one finished stage of the table: one elected thread issues it,
// the copy engine streams the stage to HBM / L2 while the CTA already sweeps the next stage.
__device__ __forceinline__ void bulk_store_stage(void* gdst, const void* ssrc, unsigned bytes) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the async proxy
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_source_free() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- hot loops of the backward sweep.  A thread owns the cells n * 512 + tid; a warp's 32 cells of block n are
// INTERIOR when every cell they read lies inside the next stage's window (no bound checks) and PENALTY-FREE when none
// of them violates a row at its favourable edge; both sets are ranges of n (per warp, per stage), so the sweep is a
// few straight loops with every per-stage constant in registers.  q max(v, 0) is evaluated as (q/2) (v + |v|) --
// exact, and |v| is a free operand modifier of the FP64 add.
struct Semi { double lo, hi; };      // bounds of the two semi-infinite cells of the next stage

// (a) the DEWH shape: two actions, action 0 = no input (a cell maps onto the cell `d0` away in the next stage's
//     window), action 1 = a translation over two cells; row violations do not depend on the action.
template <int NC, bool PEN, bool CHECKED, typename TT>
__device__ __forceinline__ void sweep_same2(const TT* __restrict__ cur, TT* __restrict__ nxt, int G, int n0, int n1,
                                            int d0, int i1, const double (&slope)[NC], const double (&hq)[NC],
                                            const double (&base)[NC], double c0, double c1, Semi out) {
    if (n0 >= n1) return;
    int cell = n0 * kTableThreads + threadIdx.x;
    const TT* ps = cur + cell + d0;
    const TT* pm = ps + i1;
    TT* po = nxt + cell;
    double cd = (double)cell;
#pragma unroll 4
    for (int n = n0; n < n1; ++n, cell += kTableThreads, cd += (double)kTableThreads, ps += kTableThreads,
                             pm += kTableThreads, po += kTableThreads) {
        double st, m0, m1;
        if (CHECKED) {
            if (cell >= G) break;
            const int js = cell + d0, jm = js + i1;
            st = js < 0 ? out.lo : (js >= G ? out.hi : (double)ps[0]);
            m0 = jm < 0 ? out.lo : (jm >= G ? out.hi : (double)pm[0]);
            m1 = jm + 1 < 0 ? out.lo : (jm + 1 >= G ? out.hi : (double)pm[1]);
        } else {
            st = (double)ps[0]; m0 = (double)pm[0]; m1 = (double)pm[1];
        }
        st += c0;
        const double mv = (m0 < m1 ? m0 : m1) + c1;
        double best = st < mv ? st : mv;
        if (PEN) {
            double pen = 0.0;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const double v = fma(cd, slope[i], base[i]);
                pen = fma(hq[i], v + fabs(v), pen);
            }
            best += pen;
        }
        put_cell(po, best);
    }
}

// (b) any action count, violations may depend on the action; interior blocks only
template <int NC, int NACT, bool SAME, typename TT>
__device__ __forceinline__ void sweep_interior(const TT* __restrict__ cur, TT* __restrict__ nxt, int n0, int n1, int d0,
                                               const double* __restrict__ s_slope, const double* __restrict__ s_q,
                                               const double* __restrict__ s_base, const double* __restrict__ s_ca,
                                               const int* __restrict__ s_i0, const int* __restrict__ s_span) {
    if (n0 >= n1) return;
    double slope[NC], hq[NC], base[NC * NACT], ca[NACT];
    int i0[NACT]; bool two[NACT];
#pragma unroll
    for (int i = 0; i < NC; ++i) { slope[i] = s_slope[i]; hq[i] = 0.5 * s_q[i]; }
#pragma unroll
    for (int al = 0; al < NACT; ++al) {
        ca[al] = s_ca[al]; i0[al] = s_i0[al] + d0; two[al] = (s_span[al] & 1) != 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) base[i * NACT + al] = s_base[i * NACT + al];
    }
    int cell = n0 * kTableThreads + threadIdx.x;
    double cd = (double)cell;
#pragma unroll 2
    for (int n = n0; n < n1; ++n, cell += kTableThreads, cd += (double)kTableThreads) {
        double pen = 0.0;
        if (SAME) {
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const double v = fma(cd, slope[i], base[i * NACT]);
                pen = fma(hq[i], v + fabs(v), pen);
            }
        }
        double best = 0.0;
#pragma unroll
        for (int al = 0; al < NACT; ++al) {
            double st = ca[al];
            if (!SAME) {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const double v = fma(cd, slope[i], base[i * NACT + al]);
                    st = fma(hq[i], v + fabs(v), st);
                }
            }
            const TT* src = cur + (cell + i0[al]);
            TT nx = src[0];
            if (two[al]) nx = cmin2(nx, src[1]);
            st += (double)nx;
            best = (al == 0 || st < best) ? st : best;
        }
        put_cell(nxt + cell, best + pen);
    }
}

// ---- linear cells.  One backward step for a cell with the two actions of the DEWH shape:
//     g(s) = min( stay: line of the cell d0 away,  move: line of cell m on [lo, t) and of cell m+1 on [t, hi) ) + costs
//     out  = tightest supporting line of the lower convex hull of g at the hull's lowest vertex
//            + the stage penalty of every row that the WHOLE cell violates (a cell that straddles a kink gets zero)
// -- never below the constant-cell bound min g.
__device__ __forceinline__ double2 hull_line(double y_lo, double y_t, double y_hi, double th, double ith, double i1th) {
    const double chord = fma(th, y_hi - y_lo, y_lo);
    if (!(y_t < chord)) return make_double2(y_lo, y_hi);         // no kink below the chord of the end points
    if (y_t <= fmin(y_lo, y_hi)) return make_double2(y_t, y_t);
    if (y_lo > y_hi) { const double s2 = (y_hi - y_t) * i1th; return make_double2(y_hi - s2, y_hi); }
    const double s1 = (y_t - y_lo) * ith;
    return make_double2(y_lo, y_lo + s1);
}

struct LinStage {
    int mask, d0, i1, span1; double fr, c0, c1;
};

template <int NC, bool CHECKED>
__device__ __forceinline__ double2 lin_cell(const double2* __restrict__ cur, int cell, int G, double2 lo_out, double2 hi_out,
                                            const LinStage& L, const double* slope, const double* q, const double* lbase,
                                            double cd, double feas_tol) {
    auto at = [&](int i) -> double2 {
        if (CHECKED) { if (i < 0) return lo_out; if (i >= G) return hi_out; }
        return cur[i];
    };
    const double BIG = INFINITY;
    double y_lo = BIG, y_hi = BIG, y_t = BIG;
    const double th = 1.0 - L.fr;
    if (L.mask & 2) {
        const int m = cell + L.d0 + L.i1;
        if (L.span1 == 0) {                       // zero translation: the move is a second "stay"
            const double2 s = at(m);
            y_lo = s.x + L.c1; y_hi = s.y + L.c1;
        } else if (L.span1 == 1) {
            const double2 a = at(m), bq = at(m + 1);
            y_lo = fma(L.fr, a.y - a.x, a.x) + L.c1;
            y_hi = fma(L.fr, bq.y - bq.x, bq.x) + L.c1;
            y_t = fmin(a.y, bq.x) + L.c1;
        } else {                                  // the translation lands on a cell boundary: constant over all it can touch
            const int j0 = (L.span1 & 2) ? m - 1 : m, j1 = (L.span1 & 4) ? m + 2 : m + 1;
            double mv = BIG;
            for (int j = j0; j <= j1; ++j) { const double2 a = at(j); mv = fmin(mv, fmin(a.x, a.y)); }
            y_lo = y_hi = mv + L.c1;
        }
    }
    if (L.mask & 1) {
        const double2 s = at(cell + L.d0);
        y_lo = fmin(y_lo, s.x + L.c0); y_hi = fmin(y_hi, s.y + L.c0);
    }
    double2 o = hull_line(y_lo, y_t, y_hi, th, 1.0 / th, 1.0 / fmax(L.fr, 1e-300));
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const double vL = fma(cd, slope[i], lbase[i]), vR = vL + slope[i];
        if (fmin(vL, vR) >= 0.0) { o.x = fma(q[i], vL, o.x); o.y = fma(q[i], vR, o.y); }
    }
    (void)feas_tol;
    o.x -= kLinSlop * fabs(o.x); o.y -= kLinSlop * fabs(o.y);
    return o;
}


// monotone map double -> unsigned 64 (smaller value <=> smaller key), and back
__device__ __forceinline__ unsigned long long order_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
    if (k == ~0ull) return INFINITY;                                  // untouched slot
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// minimum over the warp of a key / of the key of a value: two 32-bit redux.sync
__device__ __forceinline__ unsigned long long warp_min_key_u(unsigned long long key) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __forceinline__ unsigned long long warp_min_key(double v) { return warp_min_key_u(order_key(v)); }

// NC / NACT > 0: compile-time row and action counts (the DEWH shape is <2, 2>); 0: run-time loops.
// Per-stage sweep constants (computed for all stages at once, one thread per stage, before the sweep):
//   ca[k][al]      action cost
//   base[k][i][al] violation of row i under action al at the favourable edge of local cell 0,
//   slope[k][i]    ... and its increment per cell:  viol(cell) = slope * cell + base
//   i0[k][al]      translation of a cell in whole cells; span bit0: also i0+1, bit1: also i0-1, bit2: also i0+2
//                  (span 0 = identity: the "no input" action maps a cell onto itself exactly);  fr[k][al] its fraction
//   flags[k]       bit0 fast (every action allowed, no hard row, no boundary-case translation, every read inside the
//                  guard cells), bit1 samepen (row violations do not depend on the action: F = 0, G D = 0)
//   z[k]           cells [z0, z1) of a samepen stage violate no row
template <int NC, int NACT, int FMT, int MINB>
__global__ void __launch_bounds__(kTableBlock, MINB) stage_dp_table_kernel(const DpArgs A);

struct Node { double s, cost, bound; unsigned long long p0, p1; int k, pad; };

struct SearchShared {
    double best, T, root_lb, delta, open_lb, gbest;
    unsigned long long bp0, bp1;
    int sp, nodes, improvements, cut_by_T, limit, defer, pass, cap, last, gnodes;
    double cand[kTableBlock / 32];
    unsigned long long cp0[kTableBlock / 32], cp1[kTableBlock / 32];
    int cnt[kTableBlock / 32];
};

template <bool SOLO>
__device__ bool dp_search(const DpArgs& A, const DpCtx& c, int b, Node* stack, int cap, double* ptraj, SearchShared* sh,
                          int W, int budget, bool resume = false, int part = 0, int nparts = 1);

// MINB = 1: the register budget of one CTA per SM (small batches: the step time is one agent's latency);  MINB = 2: half
// the registers (a few spills outside the hot loops) so that two CTAs share an SM and hide each other's latencies -- 24 %
// more agents per second at 10,000 agents, 12 % slower at 100.
template <int NC, int NACT, int FMT, int MINB>
__global__ void __launch_bounds__(kTableBlock, MINB) stage_dp_table_kernel(const DpArgs A) {
    typedef typename Cell<FMT>::T TT;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ SearchShared sh_search;
    const int b = blockIdx.x;
    const int nthr = kTableBlock;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool helper = warp == kTableThreads / 32;
    DpCtx c = bind_ctx(A, smem);
    const DpPlan plan = make_dp_plan(c.Nt, c.nb, c.nc, c.T);
    TT* buf0 = reinterpret_cast<TT*>(smem + (size_t)plan.total * 8);
    const unsigned long long t_start = global_ns();
    if (b == 0 && tid == 0 && !A.fuse) { A.pend[0] = 0; A.pend[1] = 0; }      // list of agents for the team kernel: filled by the one-warp kernel
    dp_load(A, b, c, reinterpret_cast<double*>(buf0));        // (the stage buffers are free until the sweep starts)
    const int G = c.G, Nt = c.Nt;
    const int nc = NC > 0 ? NC : c.nc, nact = NACT > 0 ? NACT : c.nact;
    const double w = c.misc[MISC_W];
    TT* tab = reinterpret_cast<TT*>(A.table) + (int64_t)b * A.nstore * G;
    TT* cur = buf0;                        // stage k+1
    TT* nxt = buf0 + G;                    // stage k (being written)
    for (int cell = tid; cell < G; cell += nthr) cur[cell] = Cell<FMT>::pack(0.0);   // terminal stage
    // ---- per-stage constants, one stage per thread
    for (int k = tid; k < Nt; k += nthr) {
        const double akk = c.ak[k];
        const int mask = (int)c.amask[k];
        const int offk = c.off[k], d0 = offk - c.off[k + 1];
        int fast = mask == (1 << nact) - 1 && c.misc[MISC_TERMS] == 0.0, same = 1, hard = 0;
        for (int i = 0; i < nc; ++i) {
            c.sc_slope[k * nc + i] = c.e[i] * akk * w;
            if (isinf(c.qs[k * nc + i])) { fast = 0; hard = 1; }
        }
        int rd_lo = 0, rd_hi = 0;
        for (int al = 0; al < nact; ++al) {
            c.sc_ca[k * nact + al] = action_cost(c, k, al);
            const double ga = c.galpha[al];
            int i0 = 0, span = 0;
            double fr = 0.0;
            if (ga != 0.0) {
                const double r = ga * c.iak[k + 1] * c.misc[MISC_INVW];   // translation of a cell, in cells
                const double fl = floor(r);
                fr = r - fl;
                i0 = (int)fmax(fmin(fl, 1.0e9), -1.0e9);
                span = 1 | (fr < kEdgeEps ? 2 : 0) | (fr > 1.0 - kEdgeEps ? 4 : 0);
            }
            c.sc_i0[k * nact + al] = i0; c.sc_span[k * nact + al] = span; c.sc_fr[k * nact + al] = fr;
            if (span & 6) fast = 0;
            rd_lo = min(rd_lo, i0); rd_hi = max(rd_hi, i0 + (span & 1));
            // the cell is widened by a hair (kEdgeEps of its width on both sides) so that the bound also holds for
            // states that floating-point rounding assigns to it from just outside; linear cells use the exact left
            // edge (their look-up takes the neighbour's edge value when a state sits on a boundary)
            for (int i = 0; i < nc; ++i) {
                const double ei = c.e[i];
                const double pos = FMT == FMT_LIN ? (double)offk : ((double)offk + (ei >= 0.0 ? -kEdgeEps : 1.0 + kEdgeEps));
                const double bs = fma(ei, akk * (pos * w), c.falpha[i * nact + al] - c.rhs[k * nc + i]);
                c.sc_base[(k * nc + i) * nact + al] = bs;
                if (c.falpha[i * nact + al] != c.falpha[i * nact]) same = 0;
            }
        }
        c.sc_flags[k] = fast | (same << 1) | (hard << 2);
        // cells [z0, z1) of the stage violate no row at their favourable edge (exact test with the sweep's own formula)
        {
            int z0 = 0, z1 = same && !hard ? G : 0;
            auto clean = [&](int cell) {
                for (int i = 0; i < nc; ++i) if (fma((double)cell, c.sc_slope[k * nc + i], c.sc_base[(k * nc + i) * nact]) > 0.0) return false;
                return true;
            };
            for (int i = 0; i < nc && z1 > z0; ++i) {
                const double sl = c.sc_slope[k * nc + i], bs = c.sc_base[(k * nc + i) * nact];
                if (sl > 0.0) { const double t = floor(-bs / sl); z1 = (int)fmin((double)z1, fmax(t + 1.0, 0.0)); }
                else if (sl < 0.0) { const double t = ceil(-bs / sl); z0 = (int)fmax((double)z0, fmin(t, (double)G)); }
                else if (bs > 0.0) z1 = z0;
            }
            int guard = 0;
            while (z0 < z1 && !clean(z0) && guard++ < 64) ++z0;
            while (z1 > z0 && !clean(z1 - 1) && guard++ < 128) --z1;
            if (z1 <= z0 || !clean(z0) || !clean(z1 - 1)) { z0 = 0; z1 = 0; }
            c.sc_z[2 * k] = z0; c.sc_z[2 * k + 1] = z1;
        }
        // the two semi-infinite cells of the stage: which next-stage cells their images can touch (below the window
        // under the identity action: <= jl0, under any other action: <= jl1; above: >= jh0 / >= jh1), the cheapest
        // action cost of either kind, and the stage penalty at the favourable edge of each of them
        {
            int jl0 = INT_MIN, jl1 = INT_MIN, jh0 = INT_MAX, jh1 = INT_MAX;
            double c_id = INFINITY, c_rest = INFINITY;
            for (int al = 0; al < nact; ++al) {
                if (!(mask >> al & 1)) continue;
                const int sp = c.sc_span[k * nact + al], i0 = c.sc_i0[k * nact + al];
                if (al == 0 && sp == 0) { jl0 = d0 - 1; jh0 = d0 + G; c_id = c.sc_ca[k * nact]; continue; }
                jl1 = max(jl1, d0 - 1 + i0 + (sp ? 2 : 0));
                jh1 = min(jh1, d0 + G + i0 - (sp ? 1 : 0));
                c_rest = fmin(c_rest, c.sc_ca[k * nact + al]);
            }
            double plo = 0.0, phi = 0.0;
            for (int i = 0; i < nc; ++i) {
                const double ei = c.e[i], q = c.qs[k * nc + i];
                if (ei == 0.0) continue;
                double fm = INFINITY;
                for (int al = 0; al < nact; ++al) if (mask >> al & 1) fm = fmin(fm, c.falpha[i * nact + al]);
                const double pos = ei < 0.0 ? ((double)offk + kEdgeEps) : ((double)offk + (double)G - kEdgeEps);
                const double viol = fma(ei, akk * (pos * w), fm - c.rhs[k * nc + i]);
                double add = 0.0;
                if (isinf(q)) { if (viol > c.feas_tol) add = INFINITY; }
                else if (viol > 0.0) add = q * viol;
                if (ei < 0.0) plo += add; else phi += add;
            }
            c.sc_semi[8 * k] = jl0; c.sc_semi[8 * k + 1] = jl1; c.sc_semi[8 * k + 2] = jh0; c.sc_semi[8 * k + 3] = jh1;
            c.sc_semi[8 * k + 4] = rd_lo; c.sc_semi[8 * k + 5] = rd_hi;
            c.sc_semd[4 * k] = c_id; c.sc_semd[4 * k + 1] = c_rest; c.sc_semd[4 * k + 2] = plo; c.sc_semd[4 * k + 3] = phi;
        }
        // forward-search constants
        for (int i = 0; i < nc; ++i) {
            c.x_eak[k * nc + i] = c.e[i] * akk;
            c.x_hq[k * nc + i] = 0.5 * c.qs[k * nc + i];
            for (int al = 0; al < nact; ++al) c.x_foff[(k * nc + i) * nact + al] = c.falpha[i * nact + al] - c.rhs[k * nc + i];
        }
        for (int al = 0; al < nact; ++al) {
            c.x_ca[k * nact + al] = c.sc_ca[k * nact + al];
            c.x_shift[k * nact + al] = c.galpha[al] * c.iak[k + 1];
        }
        if (!(mask == (1 << nact) - 1) || c.misc[MISC_TERMS] != 0.0) c.misc[MISC_SIMPLE] = 0.0;   // benign race: same value from every writer
        for (int i = 0; i < nc; ++i) if (isinf(c.qs[k * nc + i])) c.misc[MISC_SIMPLE] = 0.0;
    }
    __syncthreads();
    // per (stage, sweep warp): the block ranges of the hand-tuned paths, one byte each -- blocks [0, na) and
    // [nb, nblocks) read beyond the next stage's window (checked), [na, nb) are interior, of which [za, zb) are
    // penalty-free.  Computed once here so that a warp starts a stage with one load instead of a chain of them.
    for (int idx = tid; idx < Nt * (kTableThreads / 32); idx += nthr) {
        const int k = idx / (kTableThreads / 32), wb = (idx % (kTableThreads / 32)) * 32;
        const int nblk = (G + kTableThreads - 1) / kTableThreads;
        int na = 0, nb_ = 0, za = 0, zb = 0;
        if (k >= 1 && (c.sc_flags[k] & 1) && NC > 0) {
            const int d0 = c.off[k] - c.off[k + 1];
            const int rd_lo = c.sc_semi[8 * k + 4], rd_hi = c.sc_semi[8 * k + 5];
            const int lo_need = -(wb + d0 + rd_lo);                    // n * 512 >= lo_need
            const int hi_room = G - 1 - wb - 31 - d0 - max(rd_hi, -d0);  // n * 512 <= hi_room (and the cells exist)
            na = min(max((lo_need + kTableThreads - 1) >> kTableShift, 0), nblk);
            nb_ = hi_room >= 0 ? min((hi_room >> kTableShift) + 1, nblk) : 0;
            if (nb_ < na) nb_ = na;
            const int z0 = c.sc_z[2 * k], z1 = c.sc_z[2 * k + 1];
            za = max((z0 - wb + kTableThreads - 1) >> kTableShift, 0); zb = ((z1 - 32 - wb) >> kTableShift) + 1;
            za = min(max(za, na), nb_); zb = min(max(zb, za), nb_);
        }
        c.sc_blk[idx] = na | (nb_ << 8) | (za << 16) | (zb << 24);
    }
    __syncthreads();
    // linear cells proper need the DEWH shape on every stage (else the agent's cells carry constant lines)
    bool lin_ok = false;
    if (FMT == FMT_LIN && NACT == 2) {
        lin_ok = c.misc[MISC_TERMS] == 0.0 && c.galpha[0] == 0.0;
        for (int k = 1; k < Nt && lin_ok; ++k) {
            const int fl = c.sc_flags[k], mask = (int)c.amask[k];
            if (!(fl & 2) || (fl & 4) || mask == 0) lin_ok = false;
        }
    }
    if (helper) {
        // executed FP64-pipe instructions of the sweep (for the roofline): per cell 4 (two compares, two adds) without
        // and 5 + 3 NC with penalty arithmetic on the hand-tuned path; the other paths at their mix per action.  The
        // helper warp adds them up on its own (nobody needs the number before the search tail).
        double kins = 0.0;
        for (int k = 1 + lane; k < Nt; k += 32) {
            const int zc = c.sc_z[2 * k + 1] - c.sc_z[2 * k];
            const int fl = c.sc_flags[k];
            if (FMT == FMT_LIN) kins += (double)G * (22.0 + 4.0 * nc);
            else if ((fl & 1) && (fl & 2) && nact == 2) kins += 4.0 * zc + (5.0 + 3.0 * nc) * (G - zc);
            else kins += (double)G * nact * (3.0 + 3.0 * nc);
        }
        kins = warp_sum(kins);
        if (lane == 0) { c.misc[MISC_KINS] = kins; c.misc[MISC_LINOK] = lin_ok ? 1.0 : 0.0; }
    }
    const bool bad_agent = c.misc[MISC_FLAG] != 0.0;
    if (bad_agent) {
        double* vout = A.v + (int64_t)b * Nt * c.nv;
        for (int j = tid; j < Nt * c.nv; j += nthr) vout[j] = nan("");
        if (tid == 0) {
            A.status[b] = HMPC_SOLVE_UNSUPPORTED; A.obj[b] = INFINITY;
            for (int i = 0; i < 8; ++i) A.stats[(int64_t)b * 8 + i] = 0;
        }
        return;
    }
    const unsigned long long t_setup = global_ns();
    // bounds of the two semi-infinite cells: S_k = f(minima of stage k+1 over the cells their images can touch, S_k+1).
    // The HELPER warp computes S_k while the 16 sweep warps work on the cells of stage k (which only need S_k+1, in
    // shared memory since the previous iteration): a chain of dependent shared-memory loads and warp reductions that
    // would otherwise sit on every warp's critical path, 48 times.  One barrier per stage.
    Semi S_help; S_help.lo = 0.0; S_help.hi = 0.0;        // helper warp: S_k+1 (the terminal stage costs nothing)
    const int nblocks = (G + kTableThreads - 1) / kTableThreads;
    for (int k = Nt - 1; k >= 1; --k) {
        const int flags = c.sc_flags[k];
        const int mask = (int)c.amask[k];
        const int offk = c.off[k], d0 = offk - c.off[k + 1];
        const double* s_ca = c.sc_ca + k * nact; const double* s_base = c.sc_base + k * nc * nact;
        const double* s_slope = c.sc_slope + k * nc; const double* s_q = c.qs + k * nc;
        const int* s_i0 = c.sc_i0 + k * nact; const int* s_span = c.sc_span + k * nact;
        if (helper) {
            const int jl0 = c.sc_semi[8 * k], jl1 = c.sc_semi[8 * k + 1], jh0 = c.sc_semi[8 * k + 2], jh1 = c.sc_semi[8 * k + 3];
            const int jl = min(max(jl0, jl1), G - 1), jh = min(jh0, jh1);
            // (the helper warp is the only one on this chain: its loads are batched eight deep so that the scan of a
            // range costs a few shared-memory latencies, not one per 32 cells)
            auto range_min = [&](int a, int bq) -> double {      // min over cells [a, bq] (this lane's share)
                double acc[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] = INFINITY;
                int j = a + lane;
                for (; j + 7 * 32 <= bq; j += 8 * 32) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc[u] = dmin(acc[u], Cell<FMT>::lo(cur[j + u * 32]));
                }
                for (; j <= bq; j += 32) acc[0] = dmin(acc[0], Cell<FMT>::lo(cur[j]));
#pragma unroll
                for (int u = 1; u < 8; ++u) acc[0] = dmin(acc[0], acc[u]);
                return acc[0];
            };
            (void)jl; (void)jh;
            double m0 = jl0 >= 0 ? range_min(0, min(jl0, G - 1)) : INFINITY;
            double m1 = jl1 >= 0 ? range_min(0, min(jl1, G - 1)) : INFINITY;
            double m2 = jh0 < G ? range_min(max(jh0, 0), G - 1) : INFINITY;
            double m3 = jh1 < G ? range_min(max(jh1, 0), G - 1) : INFINITY;
            m0 = jl0 >= 0 ? key_value(warp_min_key(m0)) : INFINITY;
            m1 = jl1 >= 0 ? key_value(warp_min_key(m1)) : INFINITY;
            m2 = jh0 < G ? key_value(warp_min_key(m2)) : INFINITY;
            m3 = jh1 < G ? key_value(warp_min_key(m3)) : INFINITY;
            // images that reach beyond the next window on either side
            const Semi above = S_help;
            if (jl0 != INT_MIN) { m0 = dmin(m0, above.lo); if (jl0 >= G) m0 = dmin(m0, above.hi); m2 = dmin(m2, above.hi); if (jh0 < 0) m2 = dmin(m2, above.lo); }
            if (jl1 != INT_MIN) { m1 = dmin(m1, above.lo); if (jl1 >= G) m1 = dmin(m1, above.hi); m3 = dmin(m3, above.hi); if (jh1 < 0) m3 = dmin(m3, above.lo); }
            const double c_id = c.sc_semd[4 * k], c_rest = c.sc_semd[4 * k + 1];
            S_help.lo = dmin(c_id + m0, c_rest + m1) + c.sc_semd[4 * k + 2];
            S_help.hi = dmin(c_id + m2, c_rest + m3) + c.sc_semd[4 * k + 3];
            if (lane == 0) { c.loinf[k] = S_help.lo; c.hiinf[k] = S_help.hi; }
        }
        Semi S_next; S_next.lo = c.loinf[k + 1]; S_next.hi = c.hiinf[k + 1];
        if (!helper) {
        // ---- (c) the cells of the window.  Blocks [na, nb) of this warp are interior: every cell they read is inside
        // the next stage's window.
        const bool fast = (flags & 1) && NC > 0;
        const unsigned blk = (unsigned)c.sc_blk[k * (kTableThreads / 32) + warp];
        int na = blk & 255u, nb_ = (blk >> 8) & 255u;
        Semi out = S_next;
        if (FMT == FMT_LIN && lin_ok) {
            LinStage L;
            L.mask = mask; L.d0 = d0; L.i1 = s_i0[1]; L.span1 = s_span[1]; L.fr = c.sc_fr[k * nact + 1];
            L.c0 = s_ca[0]; L.c1 = s_ca[1];
            double slope[NC > 0 ? NC : 1], q[NC > 0 ? NC : 1], lbase[NC > 0 ? NC : 1];
#pragma unroll
            for (int i = 0; i < (NC > 0 ? NC : 1); ++i) { slope[i] = s_slope[i]; q[i] = s_q[i]; lbase[i] = s_base[i * nact]; }
            const double2* cur2 = reinterpret_cast<const double2*>(cur);
            double2* nxt2 = reinterpret_cast<double2*>(nxt);
            const double2 lo2 = make_double2(out.lo, out.lo), hi2 = make_double2(out.hi, out.hi);
            if (!(fast && L.span1 == 1)) { na = 0; nb_ = 0; }
            for (int n = 0; n < na; ++n) {
                const int cell = n * kTableThreads + tid;
                if (cell < G) nxt2[cell] = lin_cell<(NC > 0 ? NC : 1), true>(cur2, cell, G, lo2, hi2, L, slope, q, lbase, (double)cell, c.feas_tol);
            }
#pragma unroll 2
            for (int n = na; n < nb_; ++n) {
                const int cell = n * kTableThreads + tid;
                nxt2[cell] = lin_cell<(NC > 0 ? NC : 1), false>(cur2, cell, G, lo2, hi2, L, slope, q, lbase, (double)cell, c.feas_tol);
            }
            for (int n = nb_; n < nblocks; ++n) {
                const int cell = n * kTableThreads + tid;
                if (cell < G) nxt2[cell] = lin_cell<(NC > 0 ? NC : 1), true>(cur2, cell, G, lo2, hi2, L, slope, q, lbase, (double)cell, c.feas_tol);
            }
        } else if (fast && FMT != FMT_LIN && NACT == 2 && (flags & 2) && s_span[0] == 0 && s_span[1] == 1) {
            typedef typename Cell<FMT == FMT_LIN ? FMT_F64 : FMT>::T ST;     // (scalar formats only)
            const ST* curs = reinterpret_cast<const ST*>(cur);
            ST* nxts = reinterpret_cast<ST*>(nxt);
            constexpr int NCc = NC > 0 ? NC : 1;
            double slope[NCc], hq[NCc], base[NCc];
#pragma unroll
            for (int i = 0; i < NCc; ++i) { slope[i] = s_slope[i]; hq[i] = 0.5 * s_q[i]; base[i] = s_base[i * 2]; }
            const double c0 = s_ca[0], c1 = s_ca[1];
            const int i1 = s_i0[1];
            // penalty-free blocks of this warp: all 32 cells inside [z0, z1)
            const int za = (blk >> 16) & 255u, zb = blk >> 24;
            sweep_same2<NCc, true, true, ST>(curs, nxts, G, 0, na, d0, i1, slope, hq, base, c0, c1, out);
            sweep_same2<NCc, true, false, ST>(curs, nxts, G, na, za, d0, i1, slope, hq, base, c0, c1, out);
            sweep_same2<NCc, false, false, ST>(curs, nxts, G, za, zb, d0, i1, slope, hq, base, c0, c1, out);
            sweep_same2<NCc, true, false, ST>(curs, nxts, G, zb, nb_, d0, i1, slope, hq, base, c0, c1, out);
            sweep_same2<NCc, true, true, ST>(curs, nxts, G, nb_, nblocks, d0, i1, slope, hq, base, c0, c1, out);
        } else {
            if (fast && FMT != FMT_LIN) {
                typedef typename Cell<FMT == FMT_LIN ? FMT_F64 : FMT>::T ST;
                const ST* curs = reinterpret_cast<const ST*>(cur);
                ST* nxts = reinterpret_cast<ST*>(nxt);
                if (flags & 2)
                    sweep_interior<(NC > 0 ? NC : 1), (NACT > 0 ? NACT : 1), true, ST>(curs, nxts, na, nb_, d0, s_slope, s_q, s_base, s_ca, s_i0, s_span);
                else
                    sweep_interior<(NC > 0 ? NC : 1), (NACT > 0 ? NACT : 1), false, ST>(curs, nxts, na, nb_, d0, s_slope, s_q, s_base, s_ca, s_i0, s_span);
            } else { na = 0; nb_ = 0; }
            // ---- general loop (the blocks outside [na, nb)): restricted action sets, hard rows, convex terms,
            //      translations that land on a cell boundary, reads beyond the window (constant bound per cell in
            //      every format)
            for (int n = 0; n < nblocks; ++n) {
                if (n == na) { n = nb_; if (n >= nblocks) break; }
                const int cell = n * kTableThreads + tid;
                if (cell >= G) break;
                const double cd = (double)cell;
                double best = INFINITY;
                for (int al = 0; al < nact; ++al) {
                    if (!(mask >> al & 1)) continue;
                    double st = s_ca[al];
                    for (int i = 0; i < nc; ++i) {
                        double viol = fma(cd, s_slope[i], s_base[i * nact + al]);
                        if (FMT == FMT_LIN)        // base is the exact left edge here: move to the favourable, widened edge
                            viol += s_slope[i] * (s_slope[i] >= 0.0 ? -kEdgeEps : 1.0 + kEdgeEps);
                        const double q = s_q[i];
                        if (isinf(q)) { if (viol > c.feas_tol) st = INFINITY; }
                        else {
                            const double v = viol > 0.0 ? viol : 0.0;
                            st = fma(q, v, st);
                            st = fma(c.qq[k * nc + i] * v, v, st);
                        }
                    }
                    if (c.T > 0) {
                        const double plo_ = c.ak[k] * (((double)offk + cd - kEdgeEps) * w);
                        const double phi_ = c.ak[k] * (((double)offk + cd + 1.0 + kEdgeEps) * w);
                        st += terms_lower_bound(c, k, al, plo_, phi_);
                    }
                    const int span = s_span[al];
                    const long long c0 = (long long)cell + d0 + s_i0[al];
                    auto at = [&](long long i) -> double { return i < 0 ? out.lo : (i >= G ? out.hi : Cell<FMT>::lo(cur[i])); };
                    double nx = at(c0);
                    if (span & 1) nx = dmin(nx, at(c0 + 1));
                    if (span & 2) nx = dmin(nx, at(c0 - 1));
                    if (span & 4) nx = dmin(nx, at(c0 + 2));
                    st += nx;
                    best = st < best ? st : best;
                }
                nxt[cell] = Cell<FMT>::pack(best);
            }
        }
        }
        // the buffer the NEXT stage overwrites is the source of the bulk copy issued at most D stages ago: it must have
        // been read completely before anybody passes the barrier
        if (tid == 0) bulk_wait_source_free();
        __syncthreads();
        if (tid == 0 && k % A.D == 0 && k / A.D <= A.nstore)
            bulk_store_stage(tab + (int64_t)(k / A.D - 1) * G, nxt, (unsigned)G * sizeof(TT));   // stage k -> L2 / HBM, asynchronously
        TT* t = cur; cur = nxt; nxt = t;
    }
    if (tid == 0) {
        bulk_wait_all(); asm volatile("fence.proxy.async;" ::: "memory");
        c.misc[MISC_T_SETUP] = (double)(t_setup - t_start); c.misc[MISC_T_SWEEP] = (double)(global_ns() - t_setup);
    }
    __syncthreads();
    const DpPlan plan2 = plan;
    if (A.fuse) {
        // ---- tail: the whole CTA searches this agent right away (stage data still in shared memory, table in L2);
        // the stage buffers become the stack of open nodes
        double* ptraj = reinterpret_cast<double*>(buf0);
        Node* stack = reinterpret_cast<Node*>(ptraj + (kDpMaxNt + 1));
        const int cap = (int)((A.buf_bytes - 8 * (kDpMaxNt + 1)) / sizeof(Node));
        // warp 0 starts alone (a first descent is sequential; most agents of a nominal batch are done within it) ...
        int done = 0;
        if (warp == 0) done = dp_search<true>(A, c, b, stack, cap, ptraj, &sh_search, 1, (c.Nt + A.D - 1) / A.D + kSoloTail);
        done = __syncthreads_or(done);
        // ... and the whole CTA takes over what is left, to the end: a fused launch has no second kernel
        if (!done) done = dp_search<false>(A, c, b, stack, cap, ptraj, &sh_search, kTableBlock / 32, INT_MAX, true);
        if (done) return;
        __syncthreads();
    }
    {   // hand the stage data to the search kernel
        const double* src = reinterpret_cast<const double*>(smem);
        double* dst = A.pblk + (int64_t)b * plan2.nd;
        for (int i = tid; i < plan2.nd; i += nthr) dst[i] = src[i];
        if (tid == 0) A.status[b] = kPending;
    }
}

// ------------------------------------------------------------------------------------------------ search
__device__ __forceinline__ void path_set(unsigned long long& p0, unsigned long long& p1, int k, int nb, int al) {
    const int pos = k * nb;
    const unsigned long long a = (unsigned long long)al;
    if (pos < 64) { p0 |= a << pos; if (pos + nb > 64) p1 |= a >> (64 - pos); }
    else p1 |= a << (pos - 64);
}
__device__ __forceinline__ int path_get(unsigned long long p0, unsigned long long p1, int k, int nb) {
    const int pos = k * nb;
    unsigned long long a;
    if (pos < 64) { a = p0 >> pos; if (pos + nb > 64) a |= p1 << (64 - pos); }
    else a = p1 >> (pos - 64);
    return (int)(a & ((1ull << nb) - 1ull));
}

__device__ __forceinline__ void path_set_bits(unsigned long long& p0, unsigned long long& p1, int pos, int width,
                                              unsigned long long val) {
    if (pos < 64) { p0 |= val << pos; if (pos + width > 64) p1 |= val >> (64 - pos); }
    else p1 |= val << (pos - 64);
}

// lane holding the smallest `val` among the lanes with `flag` (lowest lane on exact ties), -1 if none:
// two 32-bit redux.sync + one ballot instead of a 5-step shuffle butterfly
__device__ __forceinline__ int warp_argmin(bool flag, double val) {
    const unsigned fm = __ballot_sync(0xffffffffu, flag);
    if (fm == 0) return -1;
    const unsigned long long key = flag ? order_key(val) : ~0ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const bool cand = flag && hi == mhi;
    const unsigned mlo = __reduce_min_sync(0xffffffffu, cand ? lo : 0xffffffffu);
    const unsigned wm = __ballot_sync(0xffffffffu, cand && lo == mlo);
    return wm ? __ffs(wm) - 1 : __ffs(fm) - 1;
}

// Table bound of the cost-to-go from the exact state s at a stored stage k (a multiple of D).  The table may have been
// written by this very kernel (fused tail): read through L2 (ld.global.cg), never through the non-coherent path.
struct TabRef { const void* tab; int fmt, G, D; double invw; const int* off; const double* loinf; const double* hiinf; bool linok; };

__device__ __forceinline__ double table_bound(const TabRef& t, int k, double s) {
    const double x = fma(s, t.invw, -(double)t.off[k]);
    const double fl = floor(x);
    if (!(fl >= 0.0)) return t.loinf[k];
    if (!(fl < (double)t.G)) return t.hiinf[k];
    const int64_t idx = (int64_t)(k / t.D - 1) * t.G + (int)fl;
    if (t.fmt == FMT_F64) return __ldcg(reinterpret_cast<const double*>(t.tab) + idx);
    if (t.fmt == FMT_F32) return (double)__ldcg(reinterpret_cast<const float*>(t.tab) + idx);
    const double2* tp = reinterpret_cast<const double2*>(t.tab) + idx;
    const double2 e = __ldcg(tp);
    if (!t.linok) return fmin(e.x, e.y);
    const double fr = x - fl;
    double v = fma(fr, e.y - e.x, e.x);
    // a state on a cell boundary (within rounding) may belong to the neighbour: take the lower of the two edge values
    if (fr < kEdgeEps) v = fmin(v, fl >= 1.0 ? __ldcg(tp - 1).y : t.loinf[k]);
    if (fr > 1.0 - kEdgeEps) v = fmin(v, fl + 1.0 < (double)t.G ? __ldcg(tp + 1).x : t.hiinf[k]);
    return v;
}

// Branch-free evaluation of lane's action sequence (depth D, NB binaries per stage, NC soft rows, every action
// allowed): the D states are a short FMA chain, the table read of the final state is issued before the D
// independent stage costs are computed, so its latency is hidden behind them.
template <int NC, int NB, int D>
__device__ __forceinline__ bool expand_simple(const DpCtx& c, const TabRef& tr, int k0, int lane,
                                              double& s, double& cost, unsigned long long& q0,
                                              unsigned long long& q1, double cut, double& bd, bool& leaf) {
    constexpr int NACT = 1 << NB;
    const int Nt = c.Nt;
    const int De = (Nt - k0) < D ? (Nt - k0) : D;
    leaf = (k0 + De == Nt);
    bool ok = lane < (1 << (NB * De));
    double st[D + 1];
    int al[D];
    st[0] = s;
#pragma unroll
    for (int t = 0; t < D; ++t) {
        const int k = (k0 + t) < Nt ? (k0 + t) : (Nt - 1);
        al[t] = (lane >> (t * NB)) & (NACT - 1);
        st[t + 1] = st[t] + (t < De ? c.x_shift[k * NACT + al[t]] : 0.0);
    }
    double lbf = 0.0;
    if (!leaf && ok) lbf = table_bound(tr, k0 + D, st[D]);
#pragma unroll
    for (int t = 0; t < D; ++t) {
        if (t < De) {
            const int k = k0 + t;
            double lc = c.x_ca[k * NACT + al[t]];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const double v = fma(c.x_eak[k * NC + i], st[t], c.x_foff[(k * NC + i) * NACT + al[t]]);
                lc = fma(c.x_hq[k * NC + i], v + fabs(v), lc);
            }
            cost += lc;
        }
    }
    path_set_bits(q0, q1, k0 * NB, NB * De, (unsigned long long)(lane & ((1 << (NB * De)) - 1)));
    s = st[D];
    bd = leaf ? cost : cost + lbf;
    if (!ok) bd = INFINITY;
    ok = ok && cost + c.tailmin[k0 + De] < cut;      // (the costs still ahead may be negative: tailmin <= 0)
    ok = ok && bd < cut;
    return ok;
}

// Exact search of one agent by a team of W warps: one warp (SOLO: the calling warp, synchronised with __syncwarp; the
// other warps of the CTA are elsewhere) or a whole CTA.  The unit of work is the depth-D subtree below one open node,
// D = the largest depth with nact^D <= 32: lane l of a warp evaluates the action sequence whose base-nact digits are l
// -- exact stage costs, exact states -- and closes it with the table bound of the state it reaches (one table read
// per lane).  The open nodes are a stack in shared memory; every ROUND the team takes the nodes on top of it
// (depth-first order: the best child of the last expansion is on top, its siblings below), expands them side by side,
// and pushes the survivors back in a fixed order (the top node's children end on top again), so the run is
// deterministic.  A first descent is as sequential as the problem, so a nominal agent is searched by one warp; the
// proof phase of the hard agents (robust configurations) runs W wide.  Nodes beyond the top one are taken only while
// the stack keeps `reserve` entries free -- what a plain depth-first descent from any open node can need -- so that
// working ahead never causes an overflow the sequential search would not have had.
//   T      : threshold of the pass.  The first pass sets it a hair above the best bound of the root's children; a pass
//            that exhausts the tree below T without a solution is repeated with four times the distance;
//   dive   : when the stack could not hold the siblings of a first descent, a greedy dive (warp 0, nothing pushed)
//            produces an incumbent first; the same dive is the best effort of a search that hit its limits without one.
// Returns false when `budget` expansions were not enough: the incumbent goes to misc[] for whoever continues.
// `resume`: continue the search another team left unfinished in this very shared memory (stack, threshold, pass and
// incumbent as they are in *sh) instead of starting over from the root with its incumbent.
// `part` of `nparts`: several CTAs share one agent's tree.  The root's children are ranked by their bounds and part p
// searches the subtrees of ranks p, p + nparts, ... (every part computes the same ranking); the parts prune with each
// other's incumbents through one 64-bit key in global memory (atomicMin), and the last part to finish merges the
// results and writes the solution.  The optimum does not depend on the timing; the expansion counts, and which of
// several equally good plans is returned, may.
template <bool SOLO>
__device__ bool dp_search(const DpArgs& A, const DpCtx& c, int b, Node* stack, int cap, double* ptraj, SearchShared* sh,
                          int W, int budget, bool resume, int part, int nparts) {
    const int Nt = c.Nt, nb = c.nb, nc = c.nc, nv = c.nv, nact = c.nact;
    const int tid = SOLO ? (int)(threadIdx.x & 31) : (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = SOLO ? 32 : (int)blockDim.x;
    auto bar = [] { if (SOLO) __syncwarp(); else __syncthreads(); };
    double* vout = A.v + (int64_t)b * Nt * nv;
    int32_t* st_out = A.stats + (int64_t)b * 8;
    const unsigned long long t_search = global_ns();
    TabRef tr;
    tr.fmt = A.fmt; tr.G = c.G; tr.D = A.D; tr.invw = c.misc[MISC_INVW]; tr.off = c.off; tr.loinf = c.loinf; tr.hiinf = c.hiinf;
    tr.linok = c.misc[MISC_LINOK] != 0.0;
    tr.tab = reinterpret_cast<const unsigned char*>(A.table) + (int64_t)b * A.nstore * c.G * fmt_bytes(A.fmt);
    const int D = A.D;
    const int nodes_in = resume ? sh->nodes : (int)c.misc[MISC_NODES0];

    const bool simple21 = c.misc[MISC_SIMPLE] != 0.0 && nb == 1 && nc == 2 && D == 5;
    // evaluate lane's action sequence below node (k0, s, cost, path); returns ok, and (k1, s, cost, path, bd)
    auto expand = [&](int k0, double& s, double& cost, unsigned long long& q0, unsigned long long& q1, double cut,
                      double& bd, bool& leaf) -> bool {
        if (simple21) return expand_simple<2, 1, 5>(c, tr, k0, lane, s, cost, q0, q1, cut, bd, leaf);
        const int De = (Nt - k0) < D ? (Nt - k0) : D;
        bool ok = lane < (1 << (nb * De));
        for (int t = 0; t < De; ++t) {
            if (!ok) break;
            const int k = k0 + t, al = (lane >> (t * nb)) & (nact - 1);
            if (!(((int)c.amask[k] >> al) & 1)) { ok = false; cost = INFINITY; break; }
            cost += stage_cost(c, k, al, c.ak[k] * s);
            s = fma(c.galpha[al], c.iak[k + 1], s);
            path_set(q0, q1, k, nb, al);
            if (!(cost + c.tailmin[k + 1] < cut)) ok = false;   // partial cost + the most negative costs still ahead
        }
        leaf = (k0 + De == Nt);
        bd = lane < (1 << (nb * De)) ? cost : INFINITY;
        if (ok && !leaf) {
            bd = cost + table_bound(tr, k0 + De, s);
            ok = bd < cut;
        }
        return ok;
    };
    // warp 0: follow the best lane of every subtree down to the horizon, nothing pushed
    auto greedy_dive = [&]() {
        int k0 = 0, nodes = 0; double s0 = 0.0, cost0 = 0.0; unsigned long long r0 = 0, r1 = 0;
        while (k0 < Nt) {
            double s = s0, cost = cost0, bd; unsigned long long q0 = r0, q1 = r1; bool leaf;
            const bool ok = expand(k0, s, cost, q0, q1, INFINITY, bd, leaf);
            ++nodes;
            const int win = warp_argmin(ok, bd);
            if (win < 0) break;                               // dead end (hard rows): the search proper takes over
            s0 = __shfl_sync(0xffffffffu, s, win); cost0 = __shfl_sync(0xffffffffu, cost, win);
            r0 = __shfl_sync(0xffffffffu, q0, win); r1 = __shfl_sync(0xffffffffu, q1, win);
            k0 += (Nt - k0) < D ? (Nt - k0) : D;
        }
        if (lane == 0) {
            sh->nodes += nodes;
            if (k0 >= Nt && cost0 < sh->best) { sh->best = cost0; sh->bp0 = r0; sh->bp1 = r1; sh->improvements += 1; }
        }
    };

    if (tid == 0 && resume) { sh->defer = 0; sh->open_lb = INFINITY; }
    if (tid == 0 && !resume) {
        sh->best = c.misc[MISC_INC_OBJ];
        sh->bp0 = (unsigned long long)__double_as_longlong(c.misc[MISC_INC_P0]);
        sh->bp1 = (unsigned long long)__double_as_longlong(c.misc[MISC_INC_P1]);
        sh->nodes = nodes_in; sh->improvements = (int)c.misc[MISC_IMPR0];
        sh->T = INFINITY; sh->root_lb = INFINITY; sh->delta = 0.0; sh->open_lb = INFINITY;
        sh->limit = 0; sh->defer = 0; sh->cap = cap; sh->gbest = INFINITY; sh->last = 1; sh->gnodes = nodes_in;
    }
    bar();
    unsigned long long* gsplit = nparts > 1 ? A.split + (int64_t)b * kSplitWords : nullptr;
    int flushed = nodes_in, rounds = 0;          // (thread 0: this part's expansions already added to the agent's count)
    const int fan = (1 << (nb * D)) - 1;
    const int reserve = ((Nt + D - 1) / D) * fan;                 // growth of a sequential descent from any open node
    if (!resume && reserve + 64 > cap && !isfinite(sh->best)) {
        if (warp == 0) greedy_dive();
        bar();
    }
    // ---- exact search, pass by pass
    const int pass0 = resume ? sh->pass : 0;
    bar();
    for (int pass = pass0; pass < 200; ++pass) {
        if (tid == 0 && !(resume && pass == pass0)) {
            Node r; r.s = 0.0; r.cost = 0.0; r.bound = pass == 0 ? -INFINITY : sh->root_lb; r.p0 = r.p1 = 0; r.k = 0; r.pad = 0;
            stack[0] = r;
            sh->sp = 1; sh->cut_by_T = 0; sh->pass = pass;
        }
        bar();
        while (true) {
            int sp = sh->sp;
            const int nodes = sh->nodes;
            const double best = fmin(sh->best, sh->gbest), T = sh->T;       // (gbest: the other parts' incumbent)
            if (sp == 0) break;
            // the node budget is the agent's: parts add their expansions to one counter (every 16th round)
            if ((nparts > 1 ? sh->gnodes : nodes) >= A.o.max_nodes) { if (tid == 0) sh->limit = 1; break; }
            if (nodes - nodes_in >= budget) { if (tid == 0) sh->defer = 1; break; }
            const double tol = isfinite(best) ? fmax(1e-11 * fmax(1.0, fabs(best)), A.o.mip_rel_gap * fabs(best)) : 0.0;
            const double bcut = best - tol;
            const double cut = fmin(bcut, T);
            // discard the run of dominated nodes on top of the stack (every warp finds the same new top; nothing is written)
            while (sp > 0) {
                const unsigned lm = __ballot_sync(0xffffffffu, lane < sp && stack[sp - 1 - lane].bound < cut);
                if (lm) { sp -= __ffs(lm) - 1; break; }
                sp -= sp < 32 ? sp : 32;
            }
            if (sp == 0) { if (tid == 0) sh->sp = 0; break; }
            // how many nodes this round: the top one, and more while their children leave the reserve untouched
            const int ahead = (cap - reserve - sp) / (fan + 1);
            const int take = SOLO ? 1 : max(1, min(min(W, sp), ahead));
            const bool mine = warp < take;
            Node nd;
            bool live = false;
            if (mine) { nd = stack[sp - 1 - warp]; live = nd.bound < cut; }
            bar();                                 // everybody holds its node: the stack may be overwritten from here
            double s = 0.0, cost = 0.0, bd = INFINITY; unsigned long long q0 = 0, q1 = 0; bool leaf = false, ok = false;
            double cand = INFINITY; int npush = 0, win = -1; unsigned pm = 0;
            if (live) {
                s = nd.s; cost = nd.cost; q0 = nd.p0; q1 = nd.p1;
                ok = expand(nd.k, s, cost, q0, q1, cut, bd, leaf);
                if (nd.k == 0 && pass == 0) {
                    // the root's children fix the first threshold: a hair above the best of their bounds
                    const double rl = warp_min(bd);
                    double dl = fmax(1e-3 * fmax(1.0, fabs(rl)), 1e-9);
                    // a search that starts with an incumbent (handed over by the one-warp leg) knows the scale of the
                    // gap: a quarter of it per pass, so the tree is descended twice at most instead of once per factor 4
                    if (isfinite(best) && best > rl) dl = fmax(dl, 0.25 * (best - rl));
                    if (lane == 0) { sh->root_lb = rl; sh->delta = dl; sh->T = rl + dl; }
                    ok = ok && bd < rl + dl;
                }
                if (__any_sync(0xffffffffu, !ok && bd < bcut) && lane == 0) atomicOr(&sh->cut_by_T, 1);
                if (nparts > 1 && nd.k == 0) {
                    // rank of this child among the root's children by bound (ties by lane): the parts deal them out
                    int rank = 0;
                    for (int l = 0; l < 32; ++l) {
                        const double o = __shfl_sync(0xffffffffu, bd, l);
                        rank += (o < bd || (o == bd && l < lane)) ? 1 : 0;
                    }
                    if (rank % nparts != part) ok = false;
                }
                if (leaf) {
                    win = warp_argmin(ok, cost);
                    if (win >= 0) cand = __shfl_sync(0xffffffffu, cost, win);
                } else {
                    pm = __ballot_sync(0xffffffffu, ok);
                    npush = __popc(pm);
                    win = npush > 1 ? warp_argmin(ok, bd) : (npush ? __ffs(pm) - 1 : -1);
                }
            }
            if (SOLO) {
                // one warp: the outcome of the expansion is already in every lane's registers -- no exchange through
                // shared memory, one warp barrier per round
                const int base = sp - 1;
                if (live) {
                    if (base + npush > cap) {           // the children do not fit: stop with what is known
                        if (lane == 0) { sh->limit = 1; sh->sp = base; sh->open_lb = fmin(sh->open_lb, sh->root_lb); }
                        __syncwarp();
                        break;
                    }
                    if (!leaf && ok) {
                        int pos = __popc(pm & ((1u << lane) - 1u));           // rank among the survivors
                        if (lane == win) pos = npush - 1;                     // (the best one goes on top)
                        else if (lane > win) pos -= 1;
                        Node ch; ch.s = s; ch.cost = cost; ch.bound = bd; ch.k = nd.k + D; ch.pad = 0; ch.p0 = q0; ch.p1 = q1;
                        stack[base + pos] = ch;
                    }
                    if (leaf && win >= 0 && cand < best && lane == win) {
                        sh->best = cand; sh->bp0 = q0; sh->bp1 = q1; sh->improvements += 1;
                    }
                    if (lane == 0) sh->nodes = nodes + 1;
                }
                if (lane == 0) sh->sp = base + npush;
                __syncwarp();
                continue;
            }
            if (mine && lane == 0) {
                sh->cand[warp] = cand; sh->cnt[warp] = npush;
                if (live) atomicAdd(&sh->nodes, 1);
            }
            if (live && leaf && win >= 0) {          // the winning lane publishes its path
                if (lane == win) { sh->cp0[warp] = q0; sh->cp1[warp] = q1; }
            }
            bar();
            // ---- merge: new incumbent (lowest warp wins ties), push offsets (warp 0's children on top)
            double nbest = best; int bw = -1, total = 0, above = 0;
            for (int w2 = 0; w2 < take; ++w2) {
                const double cv = sh->cand[w2];
                if (cv < nbest) { nbest = cv; bw = w2; }
                const int cn = sh->cnt[w2];
                total += cn;
                if (w2 < warp) above += cn;
            }
            const int base = sp - take;
            if (base + total > cap) {               // the children do not fit: stop with what is known (never silently drop)
                if (tid == 0) { sh->limit = 1; sh->sp = base; sh->open_lb = fmin(sh->open_lb, sh->root_lb); }
                bar();
                break;
            }
            if (live && !leaf && npush > 0) {
                // this warp's block sits below the blocks of the warps with a smaller index (those are nearer the top)
                const int blk = base + (total - above - npush);
                if (ok) {
                    int pos = __popc(pm & ((1u << lane) - 1u));           // rank among the survivors
                    if (lane == win) pos = npush - 1;
                    else if (lane > win) pos -= 1;
                    Node ch; ch.s = s; ch.cost = cost; ch.bound = bd; ch.k = nd.k + D; ch.pad = 0; ch.p0 = q0; ch.p1 = q1;
                    stack[blk + pos] = ch;
                }
            }
            if (tid == 0) {
                sh->sp = base + total;
                if (bw >= 0) { sh->best = nbest; sh->bp0 = sh->cp0[bw]; sh->bp1 = sh->cp1[bw]; sh->improvements += 1; }
                if (nparts > 1) {                      // (the other parts' incumbent: looked up every fourth round)
                    if (bw >= 0) atomicMin(gsplit, order_key(nbest));
                    if (bw >= 0 || (rounds & 3) == 0) sh->gbest = key_value(__ldcg(gsplit));
                    if ((++rounds & 15) == 0) {
                        const int mine_now = sh->nodes, add = mine_now - flushed;
                        flushed = mine_now;
                        sh->gnodes = nodes_in + (int)atomicAdd(gsplit + 2, (unsigned long long)add) + add;
                    }
                }
            }
            bar();
        }
        bar();
        if (sh->limit || sh->defer) {
            // what is still open bounds the optimum from below (certified gap of an unfinished search)
            if (warp == 0) {
                double m = INFINITY;
                for (int i = lane; i < sh->sp; i += 32) m = fmin(m, stack[i].bound);
                m = warp_min(m);
                if (lane == 0) { sh->open_lb = fmin(sh->open_lb, m); if (sh->cut_by_T) sh->open_lb = fmin(sh->open_lb, sh->T); }
            }
            bar();
            break;
        }
        const double pbest = fmin(sh->best, sh->gbest);
        const bool finished = !sh->cut_by_T || (isfinite(pbest) && pbest <= sh->T);
        bar();
        if (finished) break;                                     // nothing was held back by the threshold: done
        if (tid == 0) {
            sh->delta *= 4.0;
            // with an incumbent the scale of the gap is known: at least a quarter of it per pass
            if (isfinite(pbest) && pbest > sh->root_lb) sh->delta = fmax(sh->delta, 0.25 * (pbest - sh->root_lb));
            double Tn = sh->root_lb + sh->delta;
            sh->T = isfinite(Tn) ? Tn : INFINITY;
        }
        bar();
    }
    if (sh->defer) {
        if (tid == 0) {
            double* misc = c.misc;
            misc[MISC_INC_OBJ] = sh->best;
            misc[MISC_INC_P0] = __longlong_as_double((long long)sh->bp0); misc[MISC_INC_P1] = __longlong_as_double((long long)sh->bp1);
            misc[MISC_NODES0] = (double)sh->nodes; misc[MISC_IMPR0] = (double)sh->improvements;
            misc[MISC_T_SEARCH] += (double)(global_ns() - t_search);
        }
        bar();
        return false;
    }
    if (nparts > 1) {
        // this part's result; the last part to arrive merges them (lowest part wins ties) and writes the solution
        unsigned long long* R = gsplit + 3 + 6 * part;
        if (tid == 0) {
            R[0] = (unsigned long long)__double_as_longlong(sh->best);
            R[1] = (unsigned long long)__double_as_longlong(sh->limit ? sh->open_lb : INFINITY);
            R[2] = sh->bp0; R[3] = sh->bp1;
            R[4] = ((unsigned long long)(unsigned)sh->limit << 32) | (unsigned)(sh->nodes - nodes_in);
            R[5] = (unsigned long long)(unsigned)(sh->improvements - (int)c.misc[MISC_IMPR0]);
            __threadfence();
            sh->last = atomicAdd(reinterpret_cast<unsigned int*>(gsplit + 1), 1u) == (unsigned)(nparts - 1);
        }
        bar();
        if (!sh->last) return true;
        if (tid == 0) {
            __threadfence();
            double bb = INFINITY, olb = INFINITY; unsigned long long b0 = 0, b1 = 0; int lim = 0, nn = nodes_in;
            int im = (int)c.misc[MISC_IMPR0];
            for (int q = 0; q < nparts; ++q) {
                const unsigned long long* Q = gsplit + 3 + 6 * q;
                const double qb = __longlong_as_double((long long)__ldcg(Q));
                if (qb < bb) { bb = qb; b0 = __ldcg(Q + 2); b1 = __ldcg(Q + 3); }
                olb = fmin(olb, __longlong_as_double((long long)__ldcg(Q + 1)));
                const unsigned long long w4 = __ldcg(Q + 4);
                lim |= (int)(w4 >> 32); nn += (int)(unsigned)w4; im += (int)(unsigned)__ldcg(Q + 5);
            }
            sh->best = bb; sh->bp0 = b0; sh->bp1 = b1; sh->limit = lim; sh->open_lb = olb; sh->nodes = nn; sh->improvements = im;
        }
        bar();
    }
    if (sh->limit && !isfinite(sh->best)) {          // limits hit before the first descent finished: best effort
        if (warp == 0) greedy_dive();
        bar();
    }
    const double best = sh->best;
    const unsigned long long bp0 = sh->bp0, bp1 = sh->bp1;
    const bool limit = sh->limit != 0;
    // ---- write the solution: binaries from the path, mu in closed form along the exact trajectory
    const bool have = isfinite(best);
    if (warp == 0 && have) {
        // exact trajectory: the scaled state is a prefix sum, s_k = sum_{j<k} galpha[al_j] / a^(j+1), and p_k = a^k s_k
        const int ch = (Nt + 31) / 32;                       // consecutive steps per lane (kDpMaxNt = 128: at most 4)
        double loc[4], sum = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = lane * ch + i;
            loc[i] = sum;
            if (i < ch && k < Nt) sum = fma(c.galpha[path_get(bp0, bp1, k, nb)], c.iak[k + 1], sum);
        }
        double incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const double excl = incl - sum;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = lane * ch + i;
            if (i < ch && k < Nt) ptraj[k] = c.ak[k] * (excl + loc[i]);
        }
    }
    bar();
    double rdi[kDpMaxNc];
    if (c.nmu)
        for (int i = 0; i < nc; ++i) rdi[i] = c.dscale[i] > 0.0 ? __drcp_rn(c.dscale[i]) : 0.0;
    for (int k = tid; k < Nt; k += nthreads) {
        double* vk = vout + (int64_t)k * nv;
        if (!have) { for (int i = 0; i < nv; ++i) vk[i] = nan(""); continue; }
        const int ak_ = path_get(bp0, bp1, k, nb);
        for (int i = 0; i < nb; ++i) vk[i] = (double)(ak_ >> i & 1);
        if (c.nmu) {
            const double p = ptraj[k];
            for (int i = 0; i < nc; ++i) {
                const double viol = fma(c.e[i], p, c.falpha[i * nact + ak_] - c.rhs[k * nc + i]);
                vk[nb + i] = !isinf(c.qs[k * nc + i]) ? fmax(viol, 0.0) * rdi[i] : 0.0;
            }
        }
    }
    if (tid == 0) {
        A.obj[b] = have ? best : INFINITY;
        A.status[b] = limit ? HMPC_SOLVE_NODE_LIMIT : (have ? HMPC_SOLVE_OPTIMAL : HMPC_SOLVE_INFEASIBLE);
        // certified relative gap of an unfinished search, in units of 1e-9 (0 when proven)
        int gap9 = 0;
        if (limit) {
            const double lbv = fmin(sh->open_lb, best);
            const double g = have ? (best - lbv) / fmax(fabs(best), 1e-300) : INFINITY;
            gap9 = (int)fmin(fmax(g, 0.0) * 1e9, 2.0e9);
        }
        const int nodes = sh->nodes;
        // executed FP64-pipe instructions (sweep + search), in thousands (stats[7])
        const double kins = c.misc[MISC_KINS] + (double)nodes * 32.0 * D * (2.0 + 3.0 * nc);
        // device time of this agent's phases in units of 0.1 us: [1] load + set-up, [2] sweep, [4] search (all legs)
        st_out[0] = nodes; st_out[1] = (int32_t)fmin(c.misc[MISC_T_SETUP] * 0.01, 2.0e9);
        st_out[2] = (int32_t)fmin(c.misc[MISC_T_SWEEP] * 0.01, 2.0e9); st_out[3] = c.G;
        st_out[4] = (int32_t)fmin((c.misc[MISC_T_SEARCH] + (double)(global_ns() - t_search)) * 0.01, 2.0e9);
        st_out[5] = sh->improvements;
        st_out[6] = gap9;
        st_out[7] = (int32_t)fmin(kins / 1000.0, 2.0e9);
    }
    bar();
    return true;
}

// One-warp search kernel (batches too large for the fused tail): kSoloWarps agents per CTA, each warp with its own copy
// of the stage data and its own stack.  An agent that needs more than kSoloBudget expansions stays pending, incumbent
// saved, for the wide kernel.
__global__ void __launch_bounds__(kSoloWarps * 32) stage_dp_solo_kernel(const DpArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ SearchShared sh[kSoloWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kSoloWarps + warp;
    if (b >= A.d.B || A.status[b] != kPending) return;
    const DpPlan plan = make_dp_plan(A.d.Nt, A.nb, A.d.nc, A.T);
    const size_t per_warp = (size_t)plan.nd * 8 + sizeof(Node) * A.solo_cap + 8 * (size_t)(kDpMaxNt + 1);
    unsigned char* mine = smem + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    DpCtx c = bind_ctx(A, mine);
    Node* stack = reinterpret_cast<Node*>(mine + (size_t)plan.nd * 8);
    double* ptraj = reinterpret_cast<double*>(stack + A.solo_cap);
    double* pb = A.pblk + (int64_t)b * plan.nd;
    double* dst = reinterpret_cast<double*>(mine);
    for (int i = lane; i < plan.nd; i += 32) dst[i] = pb[i];
    __syncwarp();
    if (lane == 0) { unsigned long long* g = A.split + (int64_t)b * kSplitWords; g[0] = ~0ull; g[1] = 0ull; g[2] = 0ull; }
    if (!dp_search<true>(A, c, b, stack, A.solo_cap, ptraj, &sh[warp], 1, kSoloBudget)) {
        const int m0 = (int)(c.misc - dst);
        if (lane < 16) pb[m0 + lane] = c.misc[lane];              // the incumbent and the counters travel on
        if (lane == 0) A.pend[2 + atomicAdd(A.pend, 1)] = b;      // (list order does not matter: no result depends on it)
    }
}

// Team-search kernel: a fixed grid of 16-warp CTAs works through the list of (pending agent, part) items the one-warp
// kernel left -- kSplit parts per agent, each a share of the root's subtrees (see dp_search).  Nothing waits for
// anything: the parts of an agent may run side by side or one after the other.
__global__ void __launch_bounds__(kSearchWarps * 32) stage_dp_search_kernel(const DpArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ SearchShared sh;
    const int items = __ldcg(A.pend) * kSplit;
    const DpPlan plan = make_dp_plan(A.d.Nt, A.nb, A.d.nc, A.T);
    DpCtx c = bind_ctx(A, smem);
    Node* stack = reinterpret_cast<Node*>(smem + (size_t)plan.nd * 8);
    double* ptraj = reinterpret_cast<double*>(stack + kStackCap);
    __shared__ int s_item;
    while (true) {
        // work queue: a CTA takes the next item when it is free (a static deal would queue items behind a hard one)
        if (threadIdx.x == 0) s_item = atomicAdd(A.pend + 1, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= items) break;
        const int b = __ldcg(A.pend + 2 + item / kSplit), part = item % kSplit;
        {
            const double* src = A.pblk + (int64_t)b * plan.nd;
            double* dst = reinterpret_cast<double*>(smem);
            for (int i = threadIdx.x; i < plan.nd; i += blockDim.x) dst[i] = src[i];
        }
        __syncthreads();
        dp_search<false>(A, c, b, stack, kStackCap, ptraj, &sh, kSearchWarps, INT_MAX, false, part, kSplit);
        __syncthreads();
    }
}

static size_t table_bytes(int B, int nstore, int G, int fmt) { return (size_t)B * nstore * G * fmt_bytes(fmt); }
static int search_depth(int nb) { return nb == 1 ? 5 : (nb == 2 ? 2 : 1); }      // nact^D <= 32

}  // namespace hmpc

extern "C" void hmpc_stage_dp_default_opts(hmpc_stage_dp_opts* o) {
    if (!o) return;
    o->mip_rel_gap = 0.0; o->feas_tol = 1e-9; o->cells = 4096; o->max_nodes = 4000000; o->table_fp64 = 1;
    o->bound = HMPC_DP_BOUND_CONSTANT; o->fuse_search = -1; o->reserved = 0;
}

extern "C" int hmpc_stage_dp_supported(const hmpc_dims* d) {
    using namespace hmpc;
    if (!d) return 0;
    const int nb = d->nu + d->ndelta;
    if (d->nx != 1 || d->nz != 0 || nb < 1 || nb > kDpMaxNb || d->nc > kDpMaxNc || d->ny > 8 || d->Nt < 1 || d->Nt > kDpMaxNt) return 0;
    if (d->nmu != 0 && d->nmu != d->nc) return 0;
    if (nb * d->Nt > 128) return 0;
    return 1;
}

static int dp_format(const hmpc_stage_dp_opts& o) {
    using namespace hmpc;
    return o.bound == HMPC_DP_BOUND_LINEAR ? FMT_LIN : (o.table_fp64 ? FMT_F64 : FMT_F32);
}

extern "C" int hmpc_stage_dp_workspace_bytes(const hmpc_dims* d, const hmpc_stage_dp_opts* opts, size_t* bytes) {
    using namespace hmpc;
    if (!d || !bytes || d->B < 0) return HMPC_ERR_ARG;
    hmpc_stage_dp_opts o;
    if (opts) o = *opts; else hmpc_stage_dp_default_opts(&o);
    if (o.cells < 64) return HMPC_ERR_ARG;
    const int nb = d->nu + d->ndelta;
    const DpPlan plan = make_dp_plan(d->Nt, nb, d->nc);
    const int D = search_depth(nb), nstore = (d->Nt - 1) / D;
    // sized for the widest format the options can end up with (an FP64 table that does not fit shared memory falls
    // back to FP32, which is smaller)
    *bytes = ((table_bytes(d->B, nstore > 0 ? nstore : 1, o.cells, dp_format(o)) + 255) & ~(size_t)255) +
             (((size_t)d->B * plan.nd * sizeof(double) + 255) & ~(size_t)255) + (size_t)d->B * kSplitWords * 8 +
             (size_t)(d->B + 4) * 4 + 512;
    return HMPC_OK;
}

extern "C" int hmpc_stage_dp_solve_f64(const hmpc_dims* dims, const double* const mats[HMPC_NUM_MATS],
                                       const int64_t mat_stride_b[HMPC_NUM_MATS], const double* rhs,
                                       const double* cost_v, int64_t cost_v_stride_b, const double* lb_v,
                                       const double* ub_v, const uint8_t* is_bin_v, const hmpc_stage_terms* terms,
                                       const hmpc_stage_dp_opts* opts, void* workspace, size_t workspace_bytes,
                                       double* v, double* obj, int32_t* status, int32_t* stats, void* stream) {
    using namespace hmpc;
    if (!dims || !mats || !mat_stride_b || !cost_v || !lb_v || !ub_v || !is_bin_v || !v || !obj || !status || !stats)
        return HMPC_ERR_ARG;
    if (!hmpc_stage_dp_supported(dims)) return HMPC_ERR_ARG;
    if (dims->nc > 0 && !rhs) return HMPC_ERR_ARG;
    if (dims->B == 0) return HMPC_OK;
    DpArgs a;
    a.d = *dims;
    for (int i = 0; i < HMPC_NUM_MATS; ++i) { a.mats[i] = mats[i]; a.stride[i] = mat_stride_b[i]; }
    if (opts) a.o = *opts; else hmpc_stage_dp_default_opts(&a.o);
    if (a.o.cells < 64 || a.o.cells % 64 != 0) return HMPC_ERR_ARG;   // bulk copies move whole 16-byte units
    a.G = a.o.cells; a.nb = dims->nu + dims->ndelta; a.nact = 1 << a.nb; a.nv = a.nb + dims->nmu;
    memset(&a.t, 0, sizeof(a.t));
    if (terms) {
        if (terms->T < 0 || terms->T > kDpMaxT) return HMPC_ERR_ARG;
        if (terms->T > 0 && (!terms->h || !terms->r || (!terms->wq && !terms->w1))) return HMPC_ERR_ARG;
        a.t = *terms;
    }
    a.T = a.t.T;
    a.rhs = rhs; a.cost = cost_v; a.sc = cost_v_stride_b; a.lb = lb_v; a.ub = ub_v; a.is_bin = is_bin_v;
    a.D = search_depth(a.nb); a.nstore = (dims->Nt - 1) / a.D;
    size_t need = 0;
    hmpc_stage_dp_workspace_bytes(dims, &a.o, &need);
    if (!workspace || workspace_bytes < need) return HMPC_ERR_WORKSPACE;
    const DpPlan plan = make_dp_plan(dims->Nt, a.nb, dims->nc, a.T);
    int dev = 0, smem_optin = 0;
    HMPC_CUDA_TRY(cudaGetDevice(&dev));
    HMPC_CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t tail_bytes = sizeof(Node) * 1024 + 8 * (kDpMaxNt + 1);            // least stack of the fused tail
    auto smem_table = [&](int fmt) {
        size_t bufs = 2 * (size_t)a.G * fmt_bytes(fmt);
        if (bufs < tail_bytes) bufs = tail_bytes;
        const size_t stage_bytes = (size_t)(4 * a.nv + dims->nc) * dims->Nt * 8;      // set-up staging of the inputs
        if (bufs < stage_bytes) bufs = stage_bytes;
        return (size_t)plan.total * 8 + bufs + 1024;    // + the kernel's static shared memory
    };
    // a format that does not fit the two stage buffers into shared memory falls back to the next smaller one
    a.fmt = dp_format(a.o);
    if (a.fmt == FMT_LIN && smem_table(FMT_LIN) > (size_t)smem_optin) return HMPC_ERR_ARG;     // (the caller picks the cell count)
    if (a.fmt == FMT_F64 && smem_table(FMT_F64) > (size_t)smem_optin) a.fmt = FMT_F32;
    a.table = workspace;
    a.pblk = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + ((table_bytes(dims->B, a.nstore > 0 ? a.nstore : 1, a.G, dp_format(a.o)) + 255) & ~(size_t)255));
    a.split = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(a.pblk) +
                                                    (((size_t)dims->B * plan.nd * sizeof(double) + 255) & ~(size_t)255));
    a.pend = reinterpret_cast<int*>(a.split + (size_t)dims->B * kSplitWords);
    a.v = v; a.obj = obj; a.status = status; a.stats = stats;
    // the tail search keeps an agent's CTA (and its shared memory) for the length of one warp's search: worth it while
    // the batch is at most two CTAs per SM, else kernel 2 searches many agents per SM concurrently
    a.fuse = a.o.fuse_search < 0 ? (dims->B <= 2 * kNumSM ? 1 : 0) : (a.o.fuse_search != 0);
    const size_t smem1 = smem_table(a.fmt) - 1024;
    a.buf_bytes = (long long)(smem1 - (size_t)plan.total * 8);
    const size_t smem2 = (size_t)plan.nd * 8 + sizeof(Node) * kStackCap + 8 * (kDpMaxNt + 1);
    // one-warp kernel: a stack that holds a whole sequential descent when shared memory has the room (two CTAs per SM)
    const size_t solo_fixed = (size_t)plan.nd * 8 + 8 * (size_t)(kDpMaxNt + 1) + 16;
    const long long solo_room = ((long long)(114 * 1024) / kSoloWarps - (long long)solo_fixed) / (long long)sizeof(Node);
    const int levels = (dims->Nt + a.D - 1) / a.D;
    a.solo_cap = levels * ((1 << (a.nb * a.D)) - 1) + 72;
    if (a.solo_cap > kSoloStack) a.solo_cap = kSoloStack;
    if ((long long)a.solo_cap > solo_room) a.solo_cap = (int)solo_room;
    if (a.solo_cap < 128) a.solo_cap = 128;
    const size_t per_warp = ((solo_fixed - 16 + sizeof(Node) * a.solo_cap) + 15) & ~(size_t)15;
    const size_t smem_solo = per_warp * kSoloWarps;
    if (smem1 + 1024 > (size_t)smem_optin || smem2 + 1024 > (size_t)smem_optin || smem_solo + 1024 > (size_t)smem_optin)
        return HMPC_ERR_ARG;
    const bool dewh_shape = dims->nc == 2 && a.nact == 2;
    void (*table_kernel)(const DpArgs) = nullptr;
    // two CTAs per SM when the batch has that many and their shared memory fits twice
    const bool two_per_sm = dims->B > 2 * kNumSM && 2 * (smem1 + 2048) <= (size_t)228 * 1024;
#define HMPC_DP_PICK(NC_, NA_, F_) (two_per_sm ? stage_dp_table_kernel<NC_, NA_, F_, 2> : stage_dp_table_kernel<NC_, NA_, F_, 1>)
    if (a.fmt == FMT_LIN) table_kernel = dewh_shape ? HMPC_DP_PICK(2, 2, FMT_LIN) : HMPC_DP_PICK(0, 0, FMT_LIN);
    else if (a.fmt == FMT_F64) table_kernel = dewh_shape ? HMPC_DP_PICK(2, 2, FMT_F64) : HMPC_DP_PICK(0, 0, FMT_F64);
    else table_kernel = dewh_shape ? HMPC_DP_PICK(2, 2, FMT_F32) : HMPC_DP_PICK(0, 0, FMT_F32);
#undef HMPC_DP_PICK
    HMPC_CUDA_TRY(cudaFuncSetAttribute(table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    HMPC_CUDA_TRY(cudaFuncSetAttribute(stage_dp_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    cudaStream_t s = (cudaStream_t)stream;
    // one fat CTA per SM: measured, two 256-thread CTAs per SM sweep 40 % fewer agents per second than one 512-thread CTA
    table_kernel<<<dims->B, kTableBlock, smem1, s>>>(a);
    HMPC_LAUNCH_CHECK("stage_dp_table_kernel");
    if (!a.fuse) {
        HMPC_CUDA_TRY(cudaFuncSetAttribute(stage_dp_solo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solo));
        stage_dp_solo_kernel<<<(dims->B + kSoloWarps - 1) / kSoloWarps, kSoloWarps * 32, smem_solo, s>>>(a);
        HMPC_LAUNCH_CHECK("stage_dp_solo_kernel");
        const int wide_ctas = (int)std::min<long long>((long long)dims->B * kSplit, kNumSM);   // one team per SM: a team alone on its SM runs rounds 1.5x faster
        stage_dp_search_kernel<<<wide_ctas, kSearchWarps * 32, smem2, s>>>(a);
        HMPC_LAUNCH_CHECK("stage_dp_search_kernel");
    }
    return HMPC_OK;
}

extern "C" int hmpc_stage_dp_max_cells(const hmpc_dims* dims, const hmpc_stage_dp_opts* opts, int32_t* cells) {
    using namespace hmpc;
    if (!dims || !cells || !hmpc_stage_dp_supported(dims)) return HMPC_ERR_ARG;
    hmpc_stage_dp_opts o;
    if (opts) o = *opts; else hmpc_stage_dp_default_opts(&o);
    int dev = 0, smem_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
        (void)cudaGetLastError();
        smem_optin = 232448;                      // sm_100: 227 KB per CTA (no device here: a host-side sizing query)
    }
    const DpPlan plan = make_dp_plan(dims->Nt, dims->nu + dims->ndelta, dims->nc);
    const int fmt = dp_format(o);
    int best = 0;
    for (int G = 256; G <= 32768; G += 256) {
        const size_t need = (size_t)plan.total * 8 + 2 * (size_t)G * fmt_bytes(fmt) + 1024;
        if (need <= (size_t)smem_optin) best = G;
    }
    *cells = best;
    return best > 0 ? HMPC_OK : HMPC_ERR_ARG;
}
