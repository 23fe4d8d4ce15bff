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
// Algorithm
//   * scaled state s_k = p_k / a^k turns the recursion into a pure translation s_k+1 = s_k + beta_k(alpha),
//     beta_k = g' alpha / a^(k+1): the "no input" action maps every cell of a uniform s-grid onto itself.
//   * kernel 1 (stage_dp_table_kernel, one CTA per agent): backward sweep over the stages of a LOWER BOUND
//     LB_k[cell] of the cost-to-go that is valid for every state in the cell (stage penalties are bounded from
//     below over the cell, a translated cell overlaps two cells of the next stage and takes their min).  The
//     stage in flight lives in shared memory between guard cells (no bound checks in the hot loop); every stage is
//     streamed to HBM by a TMA bulk copy -- as FP64 (default: sequences that tie with the incumbent are then pruned
//     at gap 0) or as FP32 rounded DOWN (half the workspace; still a valid bound).  States outside the grid
//     window get the trivial bound (sum of negative costs), so the window only affects speed, never correctness.
//   * kernel 2 (stage_dp_search_kernel, one warp per agent): exact search over the binary sequence in time
//     order; states and costs are exact FP64, a node is pruned when cost so far + LB >= incumbent.  The unit of
//     work is the depth-5 subtree under an open node (32 lanes = 32 action sequences, one table read each).  A
//     greedy dive gives the incumbent (it is optimal on > 99 % of the DEWH instances); the depth-first pass that
//     follows is the optimality proof.
// Optional convex stage terms (quadratic / L1 atoms on the state, outputs and slacks) make the problem an MIQP; they
// run through the general loops of both kernels.
// Work: Nt * G cell updates (~45 instructions each) + ~15 subtree expansions per agent -- against ~10^3..10^5
// dense simplex pivots of the general branch-and-cut kernel (milp_bnc.cu) on the same problems.
#include <string.h>
#include "common.cuh"

namespace hmpc {

constexpr int kDpThreads = 256;
constexpr int kDpMaxNb = 4;
constexpr int kDpMaxAct = 1 << kDpMaxNb;
constexpr int kDpMaxNc = 8;
constexpr int kDpMaxNt = 128;
constexpr int kDpMaxT = 4;           // extra state terms per stage
constexpr int kSearchWarps = 4;
constexpr int kStackCap = 512;        // open nodes per agent
constexpr double kEdgeEps = 1e-9;     // cell-boundary guard (fraction of a cell)
__host__ __device__ constexpr int G_PAD(int G) { return G / 4; }   // guard cells on each side of a stage buffer

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
    float* table;                      // [B, Nt, G]   (stage 0 unused)
    double* pblk;                      // [B, plan doubles]: the agent's stage data, handed from kernel 1 to kernel 2
    double* v; double* obj; int32_t* status; int32_t* stats;
};

// per-agent stage data in shared memory (all doubles; amask is stored as doubles too so that the block can be
// handed to the search kernel with one coalesced copy)
struct DpPlan {
    int ak, iak, cu, qs, rhs, tailmin, amask, e, dscale, galpha, falpha, misc, nd;   // persistent part
    int x_eak, x_foff, x_hq, x_ca, x_shift;
    int t_h, t_ga, t_r, t_wq, t_w1, qq;
    int scr;                                                                      // load-time scratch [6*Nt]
    int mst;                                                                      // staged MLD blocks [11][64]
    int sc_ca, sc_base, sc_slope, sc_i0, sc_span, sc_flags;                       // per-stage sweep constants
    int total;
};

constexpr int kMstMats = 11, kMstElems = 64;

__host__ __device__ inline DpPlan make_dp_plan(int Nt, int nb, int nc, int T = kDpMaxT) {
    DpPlan p;
    const int nact = 1 << nb;
    int o = 0;
    auto take = [&](int n) { int r = o; o += n; return r; };
    p.misc = take(8);
    p.ak = take(Nt + 2); p.iak = take(Nt + 2); p.cu = take(Nt * nb); p.qs = take(Nt * nc); p.rhs = take(Nt * nc);
    p.tailmin = take(Nt + 1); p.amask = take(Nt);
    p.e = take(nc); p.dscale = take(nc); p.galpha = take(nact); p.falpha = take(nc * nact);
    // per-stage constants of the forward search: viol_i = x_eak[k][i] * s + x_foff[k][i][al], penalty weight / 2,
    // action cost, translation of s
    p.x_eak = take(Nt * nc); p.x_foff = take(Nt * nc * nact); p.x_hq = take(Nt * nc); p.x_ca = take(Nt * nact);
    p.x_shift = take(Nt * nact);
    // optional convex terms: tau = t_h p + t_ga[al] + t_r[k];  cost += t_wq tau^2 + t_w1 |tau|;  slack: qq v^2
    p.t_h = take(T); p.t_ga = take(T * nact); p.t_r = take(Nt * T); p.t_wq = take(Nt * T); p.t_w1 = take(Nt * T);
    p.qq = take(Nt * nc);
    p.nd = (o + 1) & ~1;
    o = p.nd;
    p.scr = take(6 * Nt);
    p.mst = take(kMstMats * kMstElems);
    const int nc_ = nc > 0 ? nc : 1;
    p.sc_ca = take(Nt * nact); p.sc_base = take(Nt * nc_ * nact); p.sc_slope = take(Nt * nc_);
    p.sc_i0 = take((Nt * nact + 1) / 2); p.sc_span = take((Nt * nact + 1) / 2); p.sc_flags = take((Nt + 1) / 2);
    p.total = (o + 1) & ~1;
    return p;
}

enum { MISC_S0 = 0, MISC_W, MISC_A, MISC_FLAG, MISC_INVW, MISC_SIMPLE, MISC_TERMS };

struct DpCtx {
    int Nt, nb, nc, nact, nmu, nv, G;
    double *ak, *iak, *cu, *qs, *rhs, *tailmin, *amask, *e, *dscale, *galpha, *falpha, *misc, *scr, *mst;
    double *sc_ca, *sc_base, *sc_slope; int *sc_i0, *sc_span, *sc_flags;
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
    c.sc_ca = sd + p.sc_ca; c.sc_base = sd + p.sc_base; c.sc_slope = sd + p.sc_slope;
    c.sc_i0 = reinterpret_cast<int*>(sd + p.sc_i0); c.sc_span = reinterpret_cast<int*>(sd + p.sc_span);
    c.sc_flags = reinterpret_cast<int*>(sd + p.sc_flags);
    c.x_eak = sd + p.x_eak; c.x_foff = sd + p.x_foff; c.x_hq = sd + p.x_hq; c.x_ca = sd + p.x_ca; c.x_shift = sd + p.x_shift;
    c.t_h = sd + p.t_h; c.t_ga = sd + p.t_ga; c.t_r = sd + p.t_r; c.t_wq = sd + p.t_wq; c.t_w1 = sd + p.t_w1; c.qq = sd + p.qq;
    c.feas_tol = A.o.feas_tol;
    return c;
}

__device__ __forceinline__ const double* mat_of(const DpArgs& A, int which, int b) {
    return A.mats[which] ? A.mats[which] + (int64_t)b * A.stride[which] : nullptr;
}

// Block-cooperative load of one agent's stage data (table kernel).  On return misc[MISC_FLAG] != 0 marks an
// agent outside the supported class.
__device__ inline void dp_load(const DpArgs& A, int b, DpCtx& c) {
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
    __syncthreads();
    if (tid == 0) {
        int flag = 0;
        const double* Am = c.mst;
        const double a = Am[0];
        c.misc[MISC_A] = a;
        // a^k by running products, like the reference's A_pow_tilde (mld_evolution_matrices.py:266-272)
        double ap = 1.0;
        for (int k = 0; k <= Nt + 1; ++k) { c.ak[k] = ap; ap *= a; }
        if (!(a > 0.0) || !(c.ak[Nt] > 1e-3) || !(c.ak[Nt] < 1e3)) flag = 1;
        const double* B1 = c.mst + 1 * kMstElems; const double* B2 = c.mst + 2 * kMstElems;
        const double* Cm = c.mst + 3 * kMstElems;
        const double* D1 = c.mst + 4 * kMstElems; const double* D2 = c.mst + 5 * kMstElems;
        const double* E = c.mst + 6 * kMstElems;
        const double* F1 = c.mst + 7 * kMstElems; const double* F2 = c.mst + 8 * kMstElems;
        const double* Gm = c.mst + 9 * kMstElems; const double* Psi = c.mst + 10 * kMstElems;
        const int ny = A.d.ny, nd = A.d.ndelta;
        double g[kDpMaxNb], f[kDpMaxNc][kDpMaxNb];
        for (int j = 0; j < nb; ++j) g[j] = j < nu ? (B1 ? B1[j] : 0.0) : (B2 ? B2[j - nu] : 0.0);
        for (int i = 0; i < nc; ++i) {
            double ei = E ? E[i] : 0.0;                       // E is [nc, nx = 1]
            for (int j = 0; j < nb; ++j) f[i][j] = j < nu ? (F1 ? F1[i * nu + j] : 0.0) : (F2 ? F2[i * nd + (j - nu)] : 0.0);
            if (Gm) for (int r = 0; r < ny; ++r) {            // y = C x + D1 u + D2 delta (+ terms already in rhs)
                const double gir = Gm[i * ny + r];
                if (gir == 0.0) continue;
                ei += gir * (Cm ? Cm[r] : 0.0);
                for (int j = 0; j < nb; ++j) f[i][j] += gir * (j < nu ? (D1 ? D1[r * nu + j] : 0.0) : (D2 ? D2[r * nd + (j - nu)] : 0.0));
            }
            c.e[i] = ei;
            double di = 0.0;
            if (nmu) {
                for (int j = 0; j < nmu; ++j) {
                    const double pij = Psi ? Psi[i * nmu + j] : 0.0;
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
    }
    __syncthreads();
    // ---- per-stage data, one stage per thread
    const double* cost = A.cost + (int64_t)b * A.sc;
    const double* rhs = A.rhs + (int64_t)b * Nt * nc;
    double* s_lo = c.scr; double* s_hi = c.scr + Nt; double* s_smin = c.scr + 2 * Nt; double* s_smax = c.scr + 3 * Nt;
    double* s_cmin = c.scr + 4 * Nt; double* s_marg = c.scr + 5 * Nt;
    int bad = 0;
    for (int k = tid; k <= Nt + 1; k += nthr) c.iak[k] = 1.0 / c.ak[k];
    for (int k = tid; k < Nt; k += nthr) {
        int mask = 0;
        for (int al = 0; al < nact; ++al) {
            bool ok = true;
            for (int j = 0; j < nb; ++j) {
                const double bit = (double)(al >> j & 1);
                if (bit < A.lb[k * nv + j] || bit > A.ub[k * nv + j]) ok = false;
            }
            if (ok) mask |= 1 << al;
        }
        c.amask[k] = (double)mask;
        for (int j = 0; j < nb; ++j) { c.cu[k * nb + j] = cost[k * nv + j]; if (!A.is_bin[k * nv + j]) bad = 1; }
        for (int i = 0; i < nc; ++i) {
            c.rhs[k * nc + i] = rhs[k * nc + i];
            double qs = INFINITY;                                // hard row
            if (nmu) {
                const int col = k * nv + nb + i;
                const double q = cost[col], di = c.dscale[i], ubm = A.ub[col], lbm = A.lb[col];
                if (A.is_bin[col] || lbm > 0.0 || (lbm < 0.0 && di > 0.0)) bad = 1;
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
    if (tid == 0) {
        // trivial bound on the cost-to-go from ANY state: the negative action costs that are still ahead
        c.tailmin[Nt] = 0.0;
        for (int k = Nt - 1; k >= 0; --k) c.tailmin[k] = c.tailmin[k + 1] + fmin(s_cmin[k], 0.0);
        // grid window: hull over the stages of the violation-free band (clamped to what is reachable at that
        // stage), one max shift of margin, intersected with what is reachable at all
        double rlo = 0.0, rhi = 0.0, blo = INFINITY, bhi = -INFINITY, margin = 0.0;
        for (int k = 0; k < Nt; ++k) {
            const double lo_c = fmin(fmax(s_lo[k], rlo), rhi), hi_c = fmax(fmin(s_hi[k], rhi), rlo);
            blo = fmin(blo, fmin(lo_c, hi_c)); bhi = fmax(bhi, fmax(lo_c, hi_c));
            margin = fmax(margin, s_marg[k]);
            rlo += s_smin[k]; rhi += s_smax[k];
        }
        double S0 = fmax(rlo, blo - margin), S1 = fmin(rhi, bhi + margin);
        if (!(S1 > S0)) { S0 = rlo; S1 = rhi; }
        // keep every translation within the guard cells of the stage buffers (G / 4 on each side), so that the whole
        // sweep runs through the hot loop: a window narrower than four shifts is widened around its centre (the
        // extra cells are unreachable or dead; they cost nothing but their share of the sweep)
        {
            const double minw = 4.0 * margin * (1.0 + 16.0 / (double)c.G);
            if (S1 - S0 < minw) { const double mid = 0.5 * (S0 + S1); S0 = mid - 0.5 * minw; S1 = mid + 0.5 * minw; }
        }
        double w = (S1 - S0) / (double)c.G;
        if (!(w > 0.0) || !isfinite(w)) w = 1.0;
        c.misc[MISC_S0] = S0; c.misc[MISC_W] = w; c.misc[MISC_INVW] = 1.0 / w; c.misc[MISC_SIMPLE] = 1.0;
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

// Table storage: FP32 rounded DOWN (default: half the HBM stream and shared memory) or FP64 (exact: sequences that
// tie with the incumbent are then pruned instead of explored, see DESIGN.md section 4.2 "known limits").
template <typename TT> __device__ __forceinline__ TT to_table(double v);
template <> __device__ __forceinline__ float to_table<float>(double v) { return __double2float_rd(v); }
template <> __device__ __forceinline__ double to_table<double>(double v) { return v; }
template <typename TT> __device__ __forceinline__ TT table_min(TT a, TT b);
template <> __device__ __forceinline__ float table_min<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double table_min<double>(double a, double b) { return a < b ? a : b; }

// TMA bulk copy (cp.async.bulk, shared -> global) of one finished stage of the table: one elected thread issues it,
// the copy engine streams the 4 G bytes to HBM while the CTA already sweeps the next stage.
__device__ __forceinline__ void bulk_store_stage(void* gdst, const void* ssrc, unsigned bytes) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the async proxy
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_source_free() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Hot loop of the backward sweep over the cells [lo, hi) whose translated neighbours are all inside the table:
// no bound checks, every per-stage constant in registers.  q max(v, 0) is evaluated as (q/2) (v + |v|) -- exact,
// and |v| is a free operand modifier of the FP64 add.
template <int NC, int NACT, bool SAME, typename TT>
__device__ __forceinline__ void sweep_interior(const TT* __restrict__ cur, TT* __restrict__ nxt,
                                               int lo, int hi, int nthr,
                                               const double* __restrict__ s_slope, const double* __restrict__ s_q,
                                               const double* __restrict__ s_base, const double* __restrict__ s_ca,
                                               const int* __restrict__ s_i0, const int* __restrict__ s_span) {
    double slope[NC], hq[NC], base[NC * NACT], ca[NACT];
    int i0[NACT]; bool two[NACT];
#pragma unroll
    for (int i = 0; i < NC; ++i) { slope[i] = s_slope[i]; hq[i] = 0.5 * s_q[i]; }
#pragma unroll
    for (int al = 0; al < NACT; ++al) {
        ca[al] = s_ca[al]; i0[al] = s_i0[al]; two[al] = (s_span[al] & 1) != 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) base[i * NACT + al] = s_base[i * NACT + al];
    }
    int cell = lo + threadIdx.x;
    double cd = (double)cell;
    const double dstep = (double)nthr;
#pragma unroll 4
    for (; cell < hi; cell += nthr, cd += dstep) {
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
            if (two[al]) nx = table_min<TT>(nx, src[1]);
            st += (double)nx;
            best = (al == 0 || st < best) ? st : best;
        }
        nxt[cell] = to_table<TT>(best + pen);   // FP32: rounded DOWN, so the stored table stays a lower bound
    }
}

// NC / NACT > 0: compile-time row and action counts (the DEWH shape is <2, 2>); 0: run-time loops.
// Per-stage sweep constants (computed for all stages at once, one thread per stage, before the sweep):
//   ca[k][al]      action cost
//   base[k][i][al] violation of row i under action al at the favourable edge of cell 0,
//   slope[k][i]    ... and its increment per cell:  viol(cell) = slope * cell + base
//   i0[k][al]      translation of a cell in whole cells; span bit0: also i0+1, bit1: also i0-1, bit2: also i0+2
//                  (span 0 = identity: the "no input" action maps a cell onto itself exactly)
//   flags[k]       bit0 fast (every action allowed, no hard row, no boundary-case translation),
//                  bit1 samepen (row violations do not depend on the action: F = 0, G D = 0)
template <int NC, int NACT, typename TT>
__global__ void __launch_bounds__(512) stage_dp_table_kernel(const DpArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    const int nthr = blockDim.x;
    DpCtx c = bind_ctx(A, smem);
    const DpPlan plan = make_dp_plan(c.Nt, c.nb, c.nc, c.T);
    // two stage buffers (k+1 / k), each with `pad` guard cells on both sides that hold the out-of-window bound, so
    // that the hot loop needs no bound checks
    const int pad = G_PAD(A.G);
    TT* buf0 = reinterpret_cast<TT*>(smem + (size_t)plan.total * 8);
    dp_load(A, b, c);
    const int G = c.G, Nt = c.Nt;
    const int nc = NC > 0 ? NC : c.nc, nact = NACT > 0 ? NACT : c.nact;
    const double S0 = c.misc[MISC_S0], w = c.misc[MISC_W];
    TT* tab = reinterpret_cast<TT*>(A.table) + (int64_t)b * Nt * G;
    TT* cur = buf0 + pad;                  // stage k+1
    TT* nxt = buf0 + (G + 2 * pad) + pad;  // stage k (being written)
    for (int cell = threadIdx.x - pad; cell < G + pad; cell += nthr) cur[cell] = (TT)0;      // LB of the terminal stage
    __shared__ int s_maxshift;
    if (threadIdx.x == 0) s_maxshift = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < Nt; k += nthr) {
        const double akk = c.ak[k];
        const int mask = (int)c.amask[k];
        int fast = mask == (1 << nact) - 1 && c.misc[MISC_TERMS] == 0.0, same = 1;
        for (int i = 0; i < nc; ++i) {
            c.sc_slope[k * nc + i] = c.e[i] * akk * w;
            if (isinf(c.qs[k * nc + i])) fast = 0;
        }
        for (int al = 0; al < nact; ++al) {
            c.sc_ca[k * nact + al] = action_cost(c, k, al);
            const double ga = c.galpha[al];
            int i0 = 0, span = 0;
            if (ga != 0.0) {
                const double r = ga * c.iak[k + 1] * c.misc[MISC_INVW];   // translation of a cell, in cells
                const double fl = floor(r), fr = r - fl;
                i0 = (int)fmax(fmin(fl, 1.0e9), -1.0e9);
                span = 1 | (fr < kEdgeEps ? 2 : 0) | (fr > 1.0 - kEdgeEps ? 4 : 0);
            }
            c.sc_i0[k * nact + al] = i0; c.sc_span[k * nact + al] = span;
            if (span & 6) fast = 0;
            atomicMax(&s_maxshift, (i0 < 0 ? -i0 : i0) + 2);
            // the cell is widened by a hair (kEdgeEps of its width on both sides) so that the bound also holds for
            // states that floating-point rounding assigns to it from just outside
            for (int i = 0; i < nc; ++i) {
                const double ei = c.e[i];
                const double edge = akk * (ei >= 0.0 ? fma(-kEdgeEps, w, S0) : fma(1.0 + kEdgeEps, w, S0));
                const double bs = fma(ei, edge, c.falpha[i * nact + al] - c.rhs[k * nc + i]);
                c.sc_base[(k * nc + i) * nact + al] = bs;
                if (bs != c.sc_base[(k * nc + i) * nact]) same = 0;
            }
        }
        c.sc_flags[k] = fast | (same << 1);
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
    {   // hand the stage data to the search kernel
        const double* src = reinterpret_cast<const double*>(smem);
        double* dst = A.pblk + (int64_t)b * plan.nd;
        for (int i = threadIdx.x; i < plan.nd; i += nthr) dst[i] = src[i];
    }
    if (c.misc[MISC_FLAG] != 0.0) return;
    for (int k = Nt - 1; k >= 1; --k) {
        const TT out_next = to_table<TT>(c.tailmin[k + 1]);
        TT* tabk = tab + (int64_t)k * G;
        const int flags = c.sc_flags[k];
        const double* s_ca = c.sc_ca + k * nact; const double* s_base = c.sc_base + k * nc * nact;
        const double* s_slope = c.sc_slope + k * nc; const double* s_q = c.qs + k * nc;
        const int* s_i0 = c.sc_i0 + k * nact; const int* s_span = c.sc_span + k * nact;
        // every translated neighbour lands inside the guard cells: the whole stage goes through the hot loop
        const bool fast = (flags & 1) && NC > 0 && s_maxshift <= pad;
        const int lo = 0, hi = G;
        if (fast) {
            if (flags & 2) sweep_interior<(NC > 0 ? NC : 1), (NACT > 0 ? NACT : 1), true, TT>(cur, nxt, lo, hi, nthr, s_slope, s_q, s_base, s_ca, s_i0, s_span);
            else sweep_interior<(NC > 0 ? NC : 1), (NACT > 0 ? NACT : 1), false, TT>(cur, nxt, lo, hi, nthr, s_slope, s_q, s_base, s_ca, s_i0, s_span);
        }
        {
            // ---- general loop: restricted action sets, hard rows, translations that land on a cell boundary or
            //      beyond the guard cells
            const int mask = (int)c.amask[k];
            const int nedge = fast ? lo + (G - hi) : G;
            for (int e = threadIdx.x; e < nedge; e += nthr) {
                const int cell = fast ? (e < lo ? e : hi + (e - lo)) : e;
                const double cd = (double)cell;
                double best = INFINITY;
                for (int al = 0; al < nact; ++al) {
                    if (!(mask >> al & 1)) continue;
                    double st = s_ca[al];
                    for (int i = 0; i < nc; ++i) {
                        const double viol = fma(cd, s_slope[i], s_base[i * nact + al]);
                        const double q = s_q[i];
                        if (isinf(q)) { if (viol > c.feas_tol) st = INFINITY; }
                        else {
                            const double v = viol > 0.0 ? viol : 0.0;
                            st = fma(q, v, st);
                            st = fma(c.qq[k * nc + i] * v, v, st);
                        }
                    }
                    if (c.T > 0) {
                        const double plo = c.ak[k] * fma(cd - kEdgeEps, w, S0);
                        const double phi = c.ak[k] * fma(cd + 1.0 + kEdgeEps, w, S0);
                        st += terms_lower_bound(c, k, al, plo, phi);
                    }
                    const int span = s_span[al];
                    const long long c0 = (long long)cell + s_i0[al];
                    auto at = [&](long long i) -> TT { return (i < 0 || i >= G) ? out_next : cur[i]; };
                    TT nx = at(c0);
                    if (span & 1) nx = table_min<TT>(nx, at(c0 + 1));
                    if (span & 2) nx = table_min<TT>(nx, at(c0 - 1));
                    if (span & 4) nx = table_min<TT>(nx, at(c0 + 2));
                    st += (double)nx;
                    best = st < best ? st : best;
                }
                nxt[cell] = to_table<TT>(best);
            }
        }
        // guard cells of the stage just written: the bound of states outside the window at stage k
        {
            const TT out_k = to_table<TT>(c.tailmin[k]);
            for (int i = threadIdx.x; i < pad; i += nthr) { nxt[-1 - i] = out_k; nxt[G + i] = out_k; }
        }
        // the buffer the NEXT stage overwrites is the source of the bulk copy issued one stage ago: it must have
        // been read completely before anybody passes the barrier
        if (threadIdx.x == 0) bulk_wait_source_free();
        __syncthreads();
        if (threadIdx.x == 0) bulk_store_stage(tabk, nxt, (unsigned)G * sizeof(TT));   // stage k -> HBM, asynchronously
        TT* t = cur; cur = nxt; nxt = t;
    }
    if (threadIdx.x == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------ kernel 2
struct Node { double s, cost, bound; unsigned long long p0, p1; int k, pad; };

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

// monotone map double -> unsigned 64 (smaller value <=> smaller key)
__device__ __forceinline__ unsigned long long order_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// lane holding the smallest `val` among the lanes with `flag` (lowest lane on exact ties), -1 if none:
// two 32-bit redux.sync + one ballot instead of a 5-step shuffle butterfly
__device__ __forceinline__ int warp_argmin(bool flag, double val, int lane) {
    const unsigned fm = __ballot_sync(0xffffffffu, flag);
    if (fm == 0) return -1;
    const unsigned long long key = flag ? order_key(val) : ~0ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const bool cand = flag && hi == mhi;
    const unsigned mlo = __reduce_min_sync(0xffffffffu, cand ? lo : 0xffffffffu);
    const unsigned wm = __ballot_sync(0xffffffffu, cand && lo == mlo);
    (void)lane;
    return wm ? __ffs(wm) - 1 : __ffs(fm) - 1;
}

// Branch-free evaluation of lane's action sequence (depth D, NB binaries per stage, NC soft rows, every action
// allowed): the D states are a short FMA chain, the table read of the final state is issued before the D
// independent stage costs are computed, so its latency is hidden behind them.
__device__ __forceinline__ double table_read(const void* tab, bool fp64, int64_t idx) {
    return fp64 ? __ldg(reinterpret_cast<const double*>(tab) + idx) : (double)__ldg(reinterpret_cast<const float*>(tab) + idx);
}

template <int NC, int NB, int D>
__device__ __forceinline__ bool expand_simple(const DpCtx& c, const void* __restrict__ tab, bool fp64, int k0, int lane, double S0,
                                              double invw, double& s, double& cost, unsigned long long& q0,
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
    bool inwin = false;
    if (!leaf) {
        const double fl = floor((st[D] - S0) * invw);
        inwin = fl >= 0.0 && fl < (double)c.G;
        if (ok && inwin) lbf = table_read(tab, fp64, (int64_t)(k0 + D) * c.G + (int)fl);
    }
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
    ok = ok && cost + c.tailmin[k0 + De] < cut;      // (the costs still ahead may be negative: tailmin <= 0)
    bd = cost;
    if (!leaf) {
        bd = cost + (inwin ? lbf : c.tailmin[k0 + D]);
        ok = ok && bd < cut;
    }
    return ok;
}

// One warp per agent.  The unit of work is the depth-D subtree below one open node, D = the largest depth with
// nact^D <= 32: lane l evaluates the action sequence whose base-nact digits are l -- exact stage costs, exact
// states -- and closes it with the table bound of the state it reaches (one table read per lane per iteration).
//   search : depth-first from the root; a sequence survives when cost + bound < incumbent, the best survivor goes
//            on top of the stack; whole runs of dominated nodes are discarded in one step.  The first descent has no
//            incumbent yet, so it is a greedy dive that also leaves every sibling (with its exact bound) on the stack;
//   dive   : on long horizons those siblings would not fit, so a separate greedy dive (best lane of every subtree,
//            nothing pushed) produces the incumbent first and the search then starts from the root with it.
__global__ void __launch_bounds__(kSearchWarps * 32) stage_dp_search_kernel(const DpArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kSearchWarps + warp;
    if (b >= A.d.B) return;
    const DpPlan plan = make_dp_plan(A.d.Nt, A.nb, A.d.nc, A.T);
    const size_t per_warp = (size_t)plan.nd * 8 + sizeof(Node) * kStackCap + 8 * (kDpMaxNt + 1);
    unsigned char* base = smem + per_warp * warp;
    DpCtx c = bind_ctx(A, base);
    Node* stack = reinterpret_cast<Node*>(base + (size_t)plan.nd * 8);
    double* ptraj = reinterpret_cast<double*>(stack + kStackCap);
    {
        const double* src = A.pblk + (int64_t)b * plan.nd;
        double* dst = reinterpret_cast<double*>(base);
        for (int i = lane; i < plan.nd; i += 32) dst[i] = src[i];
    }
    __syncwarp();
    const int Nt = c.Nt, nb = c.nb, nc = c.nc, nv = c.nv, nact = c.nact;
    double* vout = A.v + (int64_t)b * Nt * nv;
    int32_t* st_out = A.stats + (int64_t)b * 8;
    if (c.misc[MISC_FLAG] != 0.0) {
        for (int j = lane; j < Nt * nv; j += 32) vout[j] = nan("");
        if (lane == 0) { A.status[b] = HMPC_SOLVE_UNSUPPORTED; A.obj[b] = INFINITY; for (int i = 0; i < 8; ++i) st_out[i] = 0; }
        return;
    }
    const double S0 = c.misc[MISC_S0], invw = c.misc[MISC_INVW], a = c.misc[MISC_A];
    const bool fp64 = A.o.table_fp64 != 0;
    const void* tab = fp64 ? (const void*)(reinterpret_cast<const double*>(A.table) + (int64_t)b * Nt * c.G)
                           : (const void*)(A.table + (int64_t)b * Nt * c.G);
    const int D = nb == 1 ? 5 : (nb == 2 ? 2 : 1);      // nact^D <= 32

    double best = INFINITY;
    unsigned long long bp0 = 0, bp1 = 0;
    int nodes = 0, improvements = 0, max_sp = 0;
    bool limit = false;

    const bool simple21 = c.misc[MISC_SIMPLE] != 0.0 && nb == 1 && nc == 2;
    // evaluate lane's action sequence below node (k0, s, cost, path); returns ok, and (k1, s, cost, path, bd)
    auto expand = [&](int k0, double& s, double& cost, unsigned long long& q0, unsigned long long& q1, double cut,
                      double& bd, bool& leaf) -> bool {
        if (simple21) return expand_simple<2, 1, 5>(c, tab, fp64, k0, lane, S0, invw, s, cost, q0, q1, cut, bd, leaf);
        const int De = (Nt - k0) < D ? (Nt - k0) : D;
        bool ok = lane < (1 << (nb * De));
        for (int t = 0; t < De; ++t) {
            if (!ok) break;
            const int k = k0 + t, al = (lane >> (t * nb)) & (nact - 1);
            if (!(((int)c.amask[k] >> al) & 1)) { ok = false; break; }
            cost += stage_cost(c, k, al, c.ak[k] * s);
            s = fma(c.galpha[al], c.iak[k + 1], s);
            path_set(q0, q1, k, nb, al);
            if (!(cost + c.tailmin[k + 1] < cut)) ok = false;   // partial cost + the most negative costs still ahead
        }
        leaf = (k0 + De == Nt);
        bd = cost;
        if (ok && !leaf) {
            const int k1 = k0 + De;
            const double fl = floor((s - S0) * invw);
            const double lbv = (fl >= 0.0 && fl < (double)c.G) ? table_read(tab, fp64, (int64_t)k1 * c.G + (int)fl)
                                                               : c.tailmin[k1];
            bd = cost + lbv;
            ok = bd < cut;
        }
        return ok;
    };
    auto argmin_lane = [&](bool flag, double val) -> int { return warp_argmin(flag, val, lane); };

    // ---- phase A: greedy dive.  Only when the stack could not hold every sibling of a first dive that runs inside
    // phase B itself (long horizons): otherwise phase B starts without an incumbent, its first descent IS the dive,
    // the siblings it pushes carry their exact bounds and are discarded 32 at a time once the incumbent exists --
    // cheaper than walking the optimal path a second time.
    const bool two_phase = ((Nt + D - 1) / D) * ((1 << (nb * D)) - 1) + 64 > kStackCap;
    auto greedy_dive = [&]() {
        int k0 = 0; double s0 = 0.0, cost0 = 0.0; unsigned long long r0 = 0, r1 = 0;
        while (k0 < Nt) {
            double s = s0, cost = cost0, bd; unsigned long long q0 = r0, q1 = r1; bool leaf;
            const bool ok = expand(k0, s, cost, q0, q1, INFINITY, bd, leaf);
            ++nodes;
            const int win = argmin_lane(ok, bd);
            if (win < 0) break;                                   // dead end (hard rows): phase B searches properly
            s0 = __shfl_sync(0xffffffffu, s, win); cost0 = __shfl_sync(0xffffffffu, cost, win);
            r0 = __shfl_sync(0xffffffffu, q0, win); r1 = __shfl_sync(0xffffffffu, q1, win);
            k0 += (Nt - k0) < D ? (Nt - k0) : D;
            if (k0 >= Nt && cost0 < best) { best = cost0; bp0 = r0; bp1 = r1; ++improvements; }
        }
    };
    if (two_phase) greedy_dive();
    // ---- phase B: exact search
    int sp = 1;
    if (lane == 0) { Node r; r.s = 0.0; r.cost = 0.0; r.bound = -INFINITY; r.p0 = r.p1 = 0; r.k = 0; r.pad = 0; stack[0] = r; }
    __syncwarp();
    while (sp > 0) {
        if (nodes >= A.o.max_nodes) { limit = true; break; }
        const double tol = isfinite(best) ? fmax(1e-11 * fmax(1.0, fabs(best)), A.o.mip_rel_gap * fabs(best)) : 0.0;
        const double cut = best - tol;
        // discard the run of dominated nodes on top of the stack, pop the first live one
        const bool live = lane < sp && stack[sp - 1 - lane].bound < cut;
        const unsigned lm = __ballot_sync(0xffffffffu, live);
        if (lm == 0) { sp -= sp < 32 ? sp : 32; continue; }
        const int first = __ffs(lm) - 1;
        const Node nd = stack[sp - 1 - first];
        sp -= first + 1;
        __syncwarp();
        double s = nd.s, cost = nd.cost, bd; unsigned long long q0 = nd.p0, q1 = nd.p1; bool leaf;
        const bool ok = expand(nd.k, s, cost, q0, q1, cut, bd, leaf);
        ++nodes;
        if (leaf) {
            const int win = argmin_lane(ok, cost);
            if (win >= 0) {
                best = __shfl_sync(0xffffffffu, cost, win);
                bp0 = __shfl_sync(0xffffffffu, q0, win); bp1 = __shfl_sync(0xffffffffu, q1, win);
                ++improvements;
            }
            continue;
        }
        const unsigned pm = __ballot_sync(0xffffffffu, ok);
        const int npush = __popc(pm);
        if (npush == 0) continue;
        if (sp + npush > kStackCap) { limit = true; break; }
        // best survivor on top, the others below it in lane order
        const int win = npush > 1 ? argmin_lane(ok, bd) : __ffs(pm) - 1;
        if (ok) {
            int pos = __popc(pm & ((1u << lane) - 1u));           // rank among the survivors
            if (lane == win) pos = npush - 1;
            else if (lane > win) pos -= 1;
            Node ch; ch.s = s; ch.cost = cost; ch.bound = bd; ch.k = nd.k + D; ch.pad = 0; ch.p0 = q0; ch.p1 = q1;
            stack[sp + pos] = ch;
        }
        sp += npush;
        max_sp = sp > max_sp ? sp : max_sp;
        __syncwarp();
    }
    if (limit && !isfinite(best)) greedy_dive();     // budget gone before the first descent finished: best effort
    // ---- write the solution: binaries from the path, mu in closed form along the exact trajectory
    const bool have = isfinite(best);
    if (lane == 0 && have) {
        double p = 0.0;
        for (int k = 0; k < Nt; ++k) { ptraj[k] = p; p = fma(a, p, c.galpha[path_get(bp0, bp1, k, nb)]); }
    }
    __syncwarp();
    for (int k = lane; k < Nt; k += 32) {
        double* vk = vout + (int64_t)k * nv;
        if (!have) { for (int i = 0; i < nv; ++i) vk[i] = nan(""); continue; }
        const int ak_ = path_get(bp0, bp1, k, nb);
        for (int i = 0; i < nb; ++i) vk[i] = (double)(ak_ >> i & 1);
        if (c.nmu) {
            const double p = ptraj[k];
            for (int i = 0; i < nc; ++i) {
                const double viol = fma(c.e[i], p, c.falpha[i * nact + ak_] - c.rhs[k * nc + i]);
                const double di = c.dscale[i];
                vk[nb + i] = (di > 0.0 && !isinf(c.qs[k * nc + i])) ? fmax(viol, 0.0) / di : 0.0;
            }
        }
    }
    if (lane == 0) {
        A.obj[b] = have ? best : INFINITY;
        A.status[b] = limit ? HMPC_SOLVE_NODE_LIMIT : (have ? HMPC_SOLVE_OPTIMAL : HMPC_SOLVE_INFEASIBLE);
        st_out[0] = nodes; st_out[1] = 0; st_out[2] = 0; st_out[3] = c.G; st_out[4] = max_sp; st_out[5] = improvements;
        st_out[6] = 0;
        st_out[7] = (int32_t)fmin(((double)(Nt - 1) * c.G * nact * (4.0 + 3.0 * nc) + (double)nodes * 32.0 * D * (4.0 + 3.0 * nc)) / 1024.0, 2.0e9);
    }
}

static size_t table_bytes(int B, int Nt, int G, bool fp64) { return (size_t)B * Nt * G * (fp64 ? sizeof(double) : sizeof(float)); }

}  // namespace hmpc

extern "C" void hmpc_stage_dp_default_opts(hmpc_stage_dp_opts* o) {
    if (!o) return;
    o->mip_rel_gap = 0.0; o->feas_tol = 1e-9; o->cells = 8192; o->max_nodes = 4000000; o->table_fp64 = 1; o->reserved = 0;
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

extern "C" int hmpc_stage_dp_workspace_bytes(const hmpc_dims* d, const hmpc_stage_dp_opts* opts, size_t* bytes) {
    using namespace hmpc;
    if (!d || !bytes || d->B < 0) return HMPC_ERR_ARG;
    hmpc_stage_dp_opts o;
    if (opts) o = *opts; else hmpc_stage_dp_default_opts(&o);
    if (o.cells < 64) return HMPC_ERR_ARG;
    const DpPlan plan = make_dp_plan(d->Nt, d->nu + d->ndelta, d->nc);
    *bytes = ((table_bytes(d->B, d->Nt, o.cells, o.table_fp64 != 0) + 255) & ~(size_t)255) + (size_t)d->B * plan.nd * sizeof(double) + 256;
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
    size_t need = 0;
    hmpc_stage_dp_workspace_bytes(dims, &a.o, &need);
    if (!workspace || workspace_bytes < need) return HMPC_ERR_WORKSPACE;
    a.table = reinterpret_cast<float*>(workspace);
    a.pblk = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + ((table_bytes(dims->B, dims->Nt, a.G, a.o.table_fp64 != 0) + 255) & ~(size_t)255));
    a.v = v; a.obj = obj; a.status = status; a.stats = stats;
    const DpPlan plan = make_dp_plan(dims->Nt, a.nb, dims->nc, a.T);
    int dev = 0, smem_optin = 0;
    HMPC_CUDA_TRY(cudaGetDevice(&dev));
    HMPC_CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    auto smem_table = [&](bool f64) { return (size_t)plan.total * 8 + 2 * (size_t)(a.G + 2 * G_PAD(a.G)) * (f64 ? sizeof(double) : sizeof(float)); };
    // an FP64 table that does not fit the two stage buffers into shared memory (cells > ~9000) falls back to FP32
    if (a.o.table_fp64 && smem_table(true) > (size_t)smem_optin) a.o.table_fp64 = 0;
    const bool fp64 = a.o.table_fp64 != 0;
    const size_t smem1 = smem_table(fp64);
    const size_t smem2 = ((size_t)plan.nd * 8 + sizeof(Node) * kStackCap + 8 * (kDpMaxNt + 1)) * kSearchWarps;
    if (smem1 > (size_t)smem_optin || smem2 > (size_t)smem_optin) return HMPC_ERR_ARG;
    const bool dewh_shape = dims->nc == 2 && a.nact == 2;
    auto table_kernel = fp64 ? (dewh_shape ? stage_dp_table_kernel<2, 2, double> : stage_dp_table_kernel<0, 0, double>)
                             : (dewh_shape ? stage_dp_table_kernel<2, 2, float> : stage_dp_table_kernel<0, 0, float>);
    HMPC_CUDA_TRY(cudaFuncSetAttribute(table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    HMPC_CUDA_TRY(cudaFuncSetAttribute(stage_dp_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    cudaStream_t s = (cudaStream_t)stream;
    // one fat CTA per SM: measured, two 256-thread CTAs per SM sweep 40 % fewer agents per second than one 512-thread CTA
    const int table_threads = 512;
    table_kernel<<<dims->B, table_threads, smem1, s>>>(a);
    HMPC_LAUNCH_CHECK("stage_dp_table_kernel");
    stage_dp_search_kernel<<<ceil_div(dims->B, kSearchWarps), kSearchWarps * 32, smem2, s>>>(a);
    HMPC_LAUNCH_CHECK("stage_dp_search_kernel");
    return HMPC_OK;
}
