// milp_bnc.cu -- K3/K4: batched mixed-integer LP solve, one CTA per problem, everything resident in shared
// memory.  Replaces the cvxpy -> Gurobi/CPLEX call inside ConstraintSolvedController.solve
// (reference: controllers/controller_base.py:509-512) for the linear-cost problems the reference example
// poses (hard part 2).
//
//   minimise c'x   s.t.   H x <= rhs,   lo <= x <= hi,   x[j] in {0,1} for is_bin[j]
//
// Algorithm (tools/bnc_proto.py is its numpy twin):
//   * bounded DUAL simplex on a dense tableau T = B^-1 N that holds only the ACTIVE rows.  The tableau starts
//     empty; violated rows of H are brought in on demand, expressed in the current basis (row generation), so
//     the working set is ~30-60 rows x n columns instead of m x (n+m) and lives in shared memory.
//   * every node of the search re-uses the SAME tableau: a node is only a set of variable bounds, and because
//     every binary is boxed the current basis stays dual feasible under any re-assignment of bounds.  Moving
//     to another node = re-seat the non-basic variables on the bound their reduced cost asks for, one mat-vec
//     for the basic values, then a few dual pivots.  No per-node basis storage, no refactorisation.
//   * complemented mixed-integer-rounding (c-MIR) cuts from single rows of H, separated at every node; they
//     are globally valid so they simply become further tableau rows.  Inactive rows/cuts are purged when the
//     tableau is full.
//   * depth-first branch and bound on the most fractional binary (nearest child first), incumbent cutoff
//     inside the dual simplex, integral candidates are polished (binaries fixed, LP re-solved) and verified
//     against ALL original rows before they become incumbents.
//
// Work per pivot is R x n FMAs (R active rows) out of shared memory; the kernel is latency / FP64-FMA /
// shared-memory bound, not HBM bound -- H is read from L2/HBM only when rows are scanned or brought in.
#include "common.cuh"

namespace hmpc {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxAdd = 16;     // rows brought in per scan
constexpr int kMaxCutsRound = 16;

struct MilpArgs {
    int B, n, m;
    const double* c; int64_t sc;
    const double* H; int64_t sH;
    const double* rhs;
    const double* lb; const double* ub; int64_t sbnd;
    const uint8_t* is_bin;
    hmpc_milp_opts o;
    int rmax, ldT, nbin;
    double* v; double* obj; int32_t* status; int32_t* stats;
};

// ---------------------------------------------------------------- shared-memory plan
struct Plan {
    // offsets in doubles
    int T, bbar, xB, colq, gB, d, xN, lo, hi, glo, ghi, cc, x, rowr, gbuf, rhs, rownorm, sepEff, sepDelta, sepF0,
        rowbuf, redd, stackBound, bestx_unused;
    int nd;  // number of doubles
    // offsets in ints (after the doubles)
    int basis, nb, fracList, binList, pathVar, pathVal, stackVar, stackVal, stackDepth, redi, sel, keepIdx, sint;
    int ni;
    // bytes
    int isbin, poolState;
    int nbytes_tail;
    size_t total;
};

__host__ __device__ inline Plan make_plan(int n, int m, int rmax, int ldT, int nbin) {
    Plan p;
    int o = 0;
    auto take = [&](int k) { int r = o; o += k; return r; };
    p.T = take(rmax * ldT);
    p.bbar = take(rmax); p.xB = take(rmax); p.colq = take(rmax); p.gB = take(rmax);
    p.d = take(n); p.xN = take(n); p.lo = take(n); p.hi = take(n); p.glo = take(n); p.ghi = take(n);
    p.cc = take(n); p.x = take(n); p.rowr = take(n); p.gbuf = take(n);
    p.rhs = take(m); p.rownorm = take(m); p.sepEff = take(m); p.sepDelta = take(m); p.sepF0 = take(m);
    p.rowbuf = take(kWarps * n);
    p.redd = take(64);
    p.stackBound = take(nbin + 2);
    p.bestx_unused = o;
    p.nd = o;
    int q = 0;
    auto takei = [&](int k) { int r = q; q += k; return r; };
    p.basis = takei(rmax); p.nb = takei(n); p.fracList = takei(n); p.binList = takei(nbin + 1);
    p.pathVar = takei(nbin + 2); p.pathVal = takei(nbin + 2);
    p.stackVar = takei(nbin + 2); p.stackVal = takei(nbin + 2); p.stackDepth = takei(nbin + 2);
    p.redi = takei(64); p.sel = takei(32); p.keepIdx = takei(rmax); p.sint = takei(32);
    p.ni = q;
    p.isbin = 0; p.poolState = n;
    p.nbytes_tail = n + m;
    p.total = (size_t)p.nd * 8 + (size_t)p.ni * 4 + (size_t)p.nbytes_tail;
    p.total = (p.total + 15) & ~(size_t)15;
    return p;
}

enum { LP_OPT = 0, LP_INF = 1, LP_CUT = 2, LP_LIM = 3 };
// slots of the shared int scratch `sint`
enum { SI_R = 0, SI_ROW, SI_COL, SI_FLAG, SI_NSEL, SI_NFRAC, SI_BRANCH, SI_STATUS, SI_SP, SI_DEPTH, SI_PIVOTS,
       SI_NODES, SI_CUTS, SI_ROWS_ADDED, SI_MAX_ROWS, SI_LP, SI_PURGES, SI_CUTID };

struct Ctx {
    int n, m, rmax, ldT, nbin;
    double *T, *bbar, *xB, *colq, *gB, *d, *xN, *lo, *hi, *glo, *ghi, *cc, *x, *rowr, *gbuf, *rhs, *rownorm, *sepEff,
        *sepDelta, *sepF0, *rowbuf, *redd, *stackBound;
    int *basis, *nb, *fracList, *binList, *pathVar, *pathVal, *stackVar, *stackVal, *stackDepth, *redi, *sel,
        *keepIdx, *sint;
    uint8_t *isbin, *poolState;
    const double* H;   // this problem's rows in global memory [m][n]
    double ptol, itol, big;
    double z;          // running LP objective (uniform across threads)
    int R;             // active rows (uniform across threads)
    double fmas;       // algorithmic FMA count (uniform), reported in stats[7] in units of 1024
};

// ---------------------------------------------------------------- block reductions (uniform result)
struct ArgVal { double v; int i; };

__device__ __forceinline__ ArgVal warp_argmax(ArgVal a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, a.v, o);
        int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
        if (ov > a.v || (ov == a.v && oi < a.i)) { a.v = ov; a.i = oi; }
    }
    return a;
}

// every thread gets the block-wide (max value, smallest index among ties); two barriers
__device__ inline ArgVal block_argmax(Ctx& c, double v, int i) {
    ArgVal a{v, i};
    a = warp_argmax(a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { c.redd[warp] = a.v; c.redi[warp] = a.i; }
    __syncthreads();
    ArgVal r{c.redd[0], c.redi[0]};
#pragma unroll
    for (int w = 1; w < kWarps; ++w) {
        double ov = c.redd[w]; int oi = c.redi[w];
        if (ov > r.v || (ov == r.v && oi < r.i)) { r.v = ov; r.i = oi; }
    }
    return r;
}

__device__ inline double block_sum(Ctx& c, double v) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) c.redd[32 + warp] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) r += c.redd[32 + w];
    return r;
}

// ---------------------------------------------------------------- bounds of a variable id
__device__ __forceinline__ void var_bounds(const Ctx& c, int var, double& l, double& h) {
    if (var >= c.n) { l = 0.0; h = INFINITY; }
    else { l = c.lo[var]; h = c.hi[var]; }
}

// x (structural values) from the basic / non-basic split; caller syncs afterwards
__device__ inline void compute_x(Ctx& c) {
    for (int j = threadIdx.x; j < c.n; j += kThreads) { int v = c.nb[j]; if (v < c.n) c.x[v] = c.xN[j]; }
    for (int r = threadIdx.x; r < c.R; r += kThreads) { int v = c.basis[r]; if (v < c.n) c.x[v] = c.xB[r]; }
}

// seat the non-basics on the bound their reduced cost asks for, recompute basics and the objective
__device__ inline void node_setup(Ctx& c) {
    __syncthreads();
    for (int j = threadIdx.x; j < c.n; j += kThreads) {
        double l, h; var_bounds(c, c.nb[j], l, h);
        double xv = (c.d[j] >= 0.0) ? l : h;
        if (!isfinite(xv)) xv = isfinite(l) ? l : h;
        c.xN[j] = xv;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < c.R; r += kWarps) {
        const double* row = c.T + r * c.ldT;
        double acc = 0.0;
        for (int j = lane; j < c.n; j += 32) acc += row[j] * c.xN[j];
        acc = warp_sum(acc);
        if (lane == 0) c.xB[r] = c.bbar[r] - acc;
    }
    __syncthreads();
    compute_x(c);
    __syncthreads();
    double part = 0.0;
    for (int j = threadIdx.x; j < c.n; j += kThreads) part += c.cc[j] * c.x[j];
    c.z = block_sum(c, part);
    c.fmas += (double)c.R * c.n;
}

// ---------------------------------------------------------------- bounded dual simplex on the active rows
__device__ inline int dual_simplex(Ctx& c, double cutoff, int& pivots_left) {
    const int n = c.n, ldT = c.ldT;
    while (true) {
        const int R = c.R;
        if (R == 0) return LP_OPT;
        // A. leaving row: largest bound violation among the basics
        double bv = -INFINITY; int bi = 0x7fffffff;
        for (int r = threadIdx.x; r < R; r += kThreads) {
            double l, h; var_bounds(c, c.basis[r], l, h);
            const double xb = c.xB[r];
            const double v = fmax(l - xb, xb - h);
            if (v > bv) { bv = v; bi = r; }
        }
        ArgVal lv = block_argmax(c, bv, bi);
        if (!(lv.v > c.ptol)) return LP_OPT;
        if (pivots_left <= 0) return LP_LIM;
        const int r = lv.i;
        double lr, hr; var_bounds(c, c.basis[r], lr, hr);
        const double xBr = c.xB[r];
        const bool below = (lr - xBr) > (xBr - hr);
        const double target = below ? lr : hr;
        const double* rowp = c.T + r * ldT;
        // B. dual ratio test (two passes: min ratio, then the largest pivot among the near-ties)
        double myratio = INFINITY, myalpha = 0.0; int myj = 0x7fffffff;
        for (int j = threadIdx.x; j < n; j += kThreads) {
            double l, h; var_bounds(c, c.nb[j], l, h);
            if (!(h > l)) continue;
            const double a = rowp[j];
            const double xn = c.xN[j];
            const bool atl = xn <= l, atu = xn >= h;
            bool cand;
            if (below) cand = (a < -1e-9 && atl) || (a > 1e-9 && atu);
            else       cand = (a > 1e-9 && atl) || (a < -1e-9 && atu);
            if (!cand) continue;
            const double ratio = fabs(c.d[j]) / fabs(a);
            if (ratio < myratio || (ratio == myratio && fabs(a) > myalpha)) { myratio = ratio; myalpha = fabs(a); myj = j; }
        }
        ArgVal rm = block_argmax(c, -myratio, myj);
        if (rm.i == 0x7fffffff || !isfinite(rm.v)) return LP_INF;
        const double rmin = -rm.v;
        double ta = -1.0; int tj = 0x7fffffff;
        for (int j = threadIdx.x; j < n; j += kThreads) {
            double l, h; var_bounds(c, c.nb[j], l, h);
            if (!(h > l)) continue;
            const double a = rowp[j];
            const double xn = c.xN[j];
            const bool atl = xn <= l, atu = xn >= h;
            bool cand;
            if (below) cand = (a < -1e-9 && atl) || (a > 1e-9 && atu);
            else       cand = (a > 1e-9 && atl) || (a < -1e-9 && atu);
            if (!cand) continue;
            const double ratio = fabs(c.d[j]) / fabs(a);
            if (ratio <= rmin + 1e-12 && fabs(a) > ta) { ta = fabs(a); tj = j; }
        }
        ArgVal pq = block_argmax(c, ta, tj);
        const int q = pq.i;
        // C. snapshot of the pivot row / column (all reads before any write)
        const double piv = rowp[q];
        const double t = (xBr - target) / piv;
        const double dq = c.d[q];
        const double bq = c.bbar[r] / piv;
        const double xq_new = c.xN[q] + t;
        const double ipiv = 1.0 / piv;
        for (int i = threadIdx.x; i < R; i += kThreads) c.colq[i] = c.T[i * ldT + q];
        for (int j = threadIdx.x; j < n; j += kThreads) c.rowr[j] = rowp[j] * ipiv;
        __syncthreads();
        // D. rank-1 update; warp <-> rows, lanes <-> columns
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (int i = warp; i < R; i += kWarps) {
                double* Ti = c.T + i * ldT;
                if (i == r) {
                    for (int j = lane; j < n; j += 32) Ti[j] = (j == q) ? ipiv : c.rowr[j];
                } else {
                    const double ci = c.colq[i];
                    if (ci != 0.0) {
                        for (int j = lane; j < n; j += 32) {
                            if (j == q) Ti[j] = -ci * ipiv;
                            else Ti[j] = fma(-ci, c.rowr[j], Ti[j]);
                        }
                    }
                }
            }
            for (int i = threadIdx.x; i < R; i += kThreads) {
                if (i == r) { c.xB[i] = xq_new; c.bbar[i] = bq; }
                else { const double ci = c.colq[i]; c.xB[i] -= ci * t; c.bbar[i] -= ci * bq; }
            }
            for (int j = threadIdx.x; j < n; j += kThreads) {
                if (j == q) { c.d[j] = -dq * ipiv; c.xN[j] = target; }
                else c.d[j] = fma(-dq, c.rowr[j], c.d[j]);
            }
            if (threadIdx.x == 0) {
                const int leaving = c.basis[r];
                c.basis[r] = c.nb[q];
                c.nb[q] = leaving;
                c.sint[SI_PIVOTS] += 1;
            }
        }
        c.z += dq * t;
        c.fmas += (double)R * n;
        --pivots_left;
        __syncthreads();
        if (c.z >= cutoff) return LP_CUT;
    }
}

// ---------------------------------------------------------------- row generation
// append  g'x <= g0  (g over the n structurals, readable by all threads) as tableau row R with slack id
__device__ inline void add_row(Ctx& c, const double* g, double g0, int slack_id) {
    const int n = c.n, ldT = c.ldT, R = c.R;
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += kThreads) { int v = c.basis[r]; c.gB[r] = (v < n) ? g[v] : 0.0; }
    __syncthreads();
    double* Tn = c.T + R * ldT;
    double part = 0.0;
    for (int j = threadIdx.x; j < n; j += kThreads) {
        const int v = c.nb[j];
        double acc = (v < n) ? g[v] : 0.0;
        for (int r = 0; r < R; ++r) { const double gb = c.gB[r]; if (gb != 0.0) acc = fma(-gb, c.T[r * ldT + j], acc); }
        Tn[j] = acc;
        part = fma(acc, c.xN[j], part);
    }
    double pb = 0.0;
    for (int r = threadIdx.x; r < R; r += kThreads) pb = fma(c.gB[r], c.bbar[r], pb);
    const double tx = block_sum(c, part);
    const double tb = block_sum(c, pb);
    if (threadIdx.x == 0) {
        c.bbar[R] = g0 - tb;
        c.xB[R] = (g0 - tb) - tx;
        c.basis[R] = slack_id;
        c.sint[SI_ROWS_ADDED] += 1;
        if (R + 1 > c.sint[SI_MAX_ROWS]) c.sint[SI_MAX_ROWS] = R + 1;
    }
    c.R = R + 1;
    c.fmas += (double)R * n;
    __syncthreads();
}

// drop rows whose basic variable is a strictly positive slack (the constraint is inactive right now)
__device__ inline void purge(Ctx& c) {
    const int n = c.n, ldT = c.ldT, R = c.R;
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = 0;
        for (int r = 0; r < R; ++r) {
            const int v = c.basis[r];
            const bool drop = (v >= n) && (c.xB[r] > 1e-7);
            if (drop) { if (v < n + c.m) c.poolState[v - n] = 0; c.keepIdx[r] = -1; }
            else c.keepIdx[r] = k++;
        }
        c.sint[SI_FLAG] = k;
        if (k != R) c.sint[SI_PURGES] += 1;
    }
    __syncthreads();
    const int newR = c.sint[SI_FLAG];
    if (newR == R) return;
    for (int r = 0; r < R; ++r) {
        const int k = c.keepIdx[r];
        if (k >= 0 && k != r) {
            for (int j = threadIdx.x; j < n; j += kThreads) c.T[k * ldT + j] = c.T[r * ldT + j];
            if (threadIdx.x == 0) { c.bbar[k] = c.bbar[r]; c.xB[k] = c.xB[r]; c.basis[k] = c.basis[r]; }
            __syncthreads();
        }
    }
    c.R = newR;
    __syncthreads();
}

// bring in the most violated inactive original rows; returns how many were added (uniform)
__device__ inline int scan_rows(Ctx& c) {
    const int n = c.n, m = c.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    compute_x(c);
    __syncthreads();
    for (int i = warp; i < m; i += kWarps) {
        double v = -INFINITY;
        if (!c.poolState[i]) {
            const double* h = c.H + (int64_t)i * n;
            double acc = 0.0;
            for (int j = lane; j < n; j += 32) acc = fma(h[j], c.x[j], acc);
            acc = warp_sum(acc) - c.rhs[i];
            v = (acc > 1e-7) ? acc / c.rownorm[i] : -INFINITY;
        }
        if (lane == 0) c.sepEff[i] = v;
    }
    __syncthreads();
    if (warp == 0) {
        int nsel = 0;
        for (int k = 0; k < kMaxAdd; ++k) {
            ArgVal a{-INFINITY, 0x7fffffff};
            for (int i = lane; i < m; i += 32) { double v = c.sepEff[i]; if (v > a.v) { a.v = v; a.i = i; } }
            a = warp_argmax(a);
            if (!(a.v > -INFINITY)) break;
            if (lane == 0) { c.sel[nsel] = a.i; c.sepEff[a.i] = -INFINITY; }
            ++nsel;
            __syncwarp();
        }
        if (lane == 0) c.sint[SI_NSEL] = nsel;
    }
    __syncthreads();
    const int nsel = c.sint[SI_NSEL];
    c.fmas += (double)m * n;
    if (nsel == 0) return 0;
    if (c.R + nsel > c.rmax) purge(c);
    int added = 0;
    for (int k = 0; k < nsel; ++k) {
        if (c.R >= c.rmax) break;
        const int i = c.sel[k];
        add_row(c, c.H + (int64_t)i * n, c.rhs[i], n + i);
        if (threadIdx.x == 0) c.poolState[i] = 1;
        ++added;
    }
    __syncthreads();
    return added;
}

// dual simplex + row generation until no original row is violated
__device__ inline int solve_lp(Ctx& c, double cutoff, int& pivots_left) {
    if (threadIdx.x == 0) c.sint[SI_LP] += 1;
    while (true) {
        const int st = dual_simplex(c, cutoff, pivots_left);
        if (st != LP_OPT) return st;
        if (scan_rows(c) == 0) return LP_OPT;
    }
}

// ---------------------------------------------------------------- c-MIR separation from single rows of H
// For row  h'x <= r :  continuous columns are shifted to their global lower bound (positive coefficients are
// relaxed away, negative ones form the continuous term s), binaries close to 1 are complemented, the row is
// scaled by 1/delta and rounded:   sum_j F(a_j/delta) xs_j - s / (delta (1 - f0)) <= floor(b/delta),
// F(t) = floor(t) + max(0, frac(t) - f0) / (1 - f0),  f0 = frac(b/delta).
struct RowPrep { double bb, sneg; bool ok; };

__device__ inline RowPrep mir_prepare(const Ctx& c, const double* h, double rhs_i, int lane) {
    // returns bb (rhs after shifting/complementing) and the value of the continuous term at x
    double shift = 0.0, compsum = 0.0, sneg = 0.0;
    int bad = 0;
    for (int j = lane; j < c.n; j += 32) {
        const double hj = h[j];
        if (hj == 0.0) continue;
        if (c.isbin[j]) { if (c.x[j] > 0.5) compsum += hj; }
        else {
            const double gl = c.glo[j];
            if (!(fabs(gl) < 0.5 * c.big)) { bad = 1; continue; }
            shift += hj * gl;
            if (hj < 0.0) sneg += -hj * (c.x[j] - gl);
        }
    }
    shift = warp_sum(shift); compsum = warp_sum(compsum); sneg = warp_sum(sneg);
    bad = __any_sync(0xffffffffu, bad);
    RowPrep p; p.bb = rhs_i - shift - compsum; p.sneg = sneg; p.ok = !bad;
    return p;
}

__device__ inline void separate_rows(Ctx& c) {
    const int n = c.n, m = c.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nfrac = c.sint[SI_NFRAC];
    double* hb = c.rowbuf + warp * n;
    for (int i = warp; i < m; i += kWarps) {
        const double* h = c.H + (int64_t)i * n;
        for (int j = lane; j < n; j += 32) hb[j] = h[j];
        __syncwarp();
        // candidate scalings: |h_ij| of the fractional binaries in this row, plus max |h_ij| over binaries
        double amax = 0.0; int has_frac = 0;
        for (int j = lane; j < n; j += 32) if (c.isbin[j]) amax = fmax(amax, fabs(hb[j]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        for (int k = 0; k < nfrac; ++k) if (fabs(hb[c.fracList[k]]) > 1e-9) { has_frac = 1; break; }
        double bestEff = 0.0, bestDelta = 0.0, bestF0 = 0.0;
        if (has_frac && amax > 1e-9) {
            const RowPrep pr = mir_prepare(c, hb, c.rhs[i], lane);
            if (pr.ok) {
                for (int k = 0; k <= nfrac; ++k) {
                    const double d0 = (k < nfrac) ? fabs(hb[c.fracList[k]]) : amax;
                    if (!(d0 > 1e-9)) continue;
                    for (int s = 0; s < 4; ++s) {
                        const double dl = d0 / (double)(1 << s);
                        const double bt = pr.bb / dl;
                        const double fb = floor(bt);
                        const double f0 = bt - fb;
                        if (f0 < 0.05 || f0 > 0.95) continue;
                        const double inv1f = 1.0 / (1.0 - f0);
                        double lhs = 0.0, nrm = 0.0;
                        for (int j = lane; j < n; j += 32) {
                            if (!c.isbin[j]) continue;
                            const double hj = hb[j];
                            if (hj == 0.0) continue;
                            const bool comp = c.x[j] > 0.5;
                            const double at = (comp ? -hj : hj) / dl;
                            const double fl = floor(at);
                            const double Fa = fl + fmax(0.0, (at - fl) - f0) * inv1f;
                            const double xs = comp ? 1.0 - c.x[j] : c.x[j];
                            lhs = fma(Fa, xs, lhs);
                            nrm = fma(Fa, Fa, nrm);
                        }
                        lhs = warp_sum(lhs); nrm = warp_sum(nrm);
                        const double viol = lhs - pr.sneg * inv1f / dl - fb;
                        const double eff = viol / sqrt(nrm + 1e-12);
                        if (viol > 1e-6 && eff > bestEff) { bestEff = eff; bestDelta = dl; bestF0 = f0; }
                    }
                }
            }
        }
        if (lane == 0) { c.sepEff[i] = bestEff > 0.0 ? bestEff : -INFINITY; c.sepDelta[i] = bestDelta; c.sepF0[i] = bestF0; }
        __syncwarp();
    }
}

// materialise the chosen cut of row i into gbuf (structural space); returns g0 (uniform)
__device__ inline double build_cut(Ctx& c, int i) {
    const int n = c.n;
    const double* h = c.H + (int64_t)i * n;
    const double dl = c.sepDelta[i], f0 = c.sepF0[i];
    const double inv1f = 1.0 / (1.0 - f0);
    // rhs after shifting / complementing (block-wide recomputation, same arithmetic as mir_prepare)
    double shift = 0.0, compsum = 0.0;
    for (int j = threadIdx.x; j < n; j += kThreads) {
        const double hj = h[j];
        if (hj == 0.0) continue;
        if (c.isbin[j]) { if (c.x[j] > 0.5) compsum += hj; }
        else shift += hj * c.glo[j];
    }
    const double tshift = block_sum(c, shift);
    const double tcomp = block_sum(c, compsum);
    const double bb = c.rhs[i] - tshift - tcomp;
    const double fb = floor(bb / dl);
    double g0part = 0.0;
    for (int j = threadIdx.x; j < n; j += kThreads) {
        const double hj = h[j];
        double g = 0.0;
        if (hj != 0.0) {
            if (c.isbin[j]) {
                const bool comp = c.x[j] > 0.5;
                const double at = (comp ? -hj : hj) / dl;
                const double fl = floor(at);
                const double Fa = fl + fmax(0.0, (at - fl) - f0) * inv1f;
                if (comp) { g = -Fa; g0part -= Fa; } else g = Fa;
            } else if (hj < 0.0) {
                const double k = hj * inv1f / dl;
                g = k; g0part += k * c.glo[j];
            }
        }
        c.gbuf[j] = g;
    }
    const double g0 = fb + block_sum(c, g0part);
    __syncthreads();
    return g0;
}

// fractional binaries of the current x -> fracList, branching candidate; x must be current. uniform result.
__device__ inline ArgVal find_fractional(Ctx& c) {
    __syncthreads();
    if (threadIdx.x == 0) c.sint[SI_NFRAC] = 0;
    __syncthreads();
    double bf = -1.0; int bj = 0x7fffffff;
    for (int k = threadIdx.x; k < c.nbin; k += kThreads) {
        const int j = c.binList[k];
        const double xv = c.x[j];
        const double f = fabs(xv - rint(xv));
        if (f > c.itol) { const int pos = atomicAdd(&c.sint[SI_NFRAC], 1); c.fracList[pos] = j; }
        if (f > bf) { bf = f; bj = j; }
    }
    ArgVal a = block_argmax(c, bf, bj);
    __syncthreads();
    // deterministic order of the candidate list (atomics may permute it): sort ascending, tiny list
    if (threadIdx.x == 0) {
        const int nf = c.sint[SI_NFRAC];
        for (int p = 1; p < nf; ++p) { int key = c.fracList[p], q = p - 1; while (q >= 0 && c.fracList[q] > key) { c.fracList[q + 1] = c.fracList[q]; --q; } c.fracList[q + 1] = key; }
    }
    __syncthreads();
    return a;
}

__device__ inline int cut_loop(Ctx& c, int rounds, int per_round, double cutoff, int& pivots_left, int max_cuts) {
    for (int rd = 0; rd < rounds; ++rd) {
        __syncthreads();
        compute_x(c);
        ArgVal fr = find_fractional(c);
        if (!(fr.v > c.itol)) return LP_OPT;
        if (c.sint[SI_CUTS] >= max_cuts) return LP_OPT;
        separate_rows(c);
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (warp == 0) {
            int nsel = 0;
            for (int k = 0; k < per_round; ++k) {
                ArgVal a{-INFINITY, 0x7fffffff};
                for (int i = lane; i < c.m; i += 32) { double v = c.sepEff[i]; if (v > a.v) { a.v = v; a.i = i; } }
                a = warp_argmax(a);
                if (!(a.v > -INFINITY)) break;
                if (lane == 0) { c.sel[nsel] = a.i; c.sepEff[a.i] = -INFINITY; }
                ++nsel;
                __syncwarp();
            }
            if (lane == 0) c.sint[SI_NSEL] = nsel;
        }
        __syncthreads();
        const int nsel = c.sint[SI_NSEL];
        if (nsel == 0) return LP_OPT;
        if (c.R + nsel > c.rmax) purge(c);
        int added = 0;
        for (int k = 0; k < nsel; ++k) {
            if (c.R >= c.rmax) break;
            const double g0 = build_cut(c, c.sel[k]);
            const int id = c.n + c.m + c.sint[SI_CUTID];
            add_row(c, c.gbuf, g0, id);
            if (threadIdx.x == 0) { c.sint[SI_CUTID] += 1; c.sint[SI_CUTS] += 1; }
            ++added;
        }
        __syncthreads();
        if (added == 0) return LP_OPT;
        const int st = solve_lp(c, cutoff, pivots_left);
        if (st != LP_OPT) return st;
    }
    return LP_OPT;
}

// ---------------------------------------------------------------- the kernel: one CTA per problem
__global__ void __launch_bounds__(kThreads, 1) milp_bnc_kernel(const MilpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int n = a.n, m = a.m;
    const Plan p = make_plan(n, m, a.rmax, a.ldT, a.nbin);
    double* sd = reinterpret_cast<double*>(smem_raw);
    int* si = reinterpret_cast<int*>(sd + p.nd);
    uint8_t* sb = reinterpret_cast<uint8_t*>(si + p.ni);
    Ctx c;
    c.n = n; c.m = m; c.rmax = a.rmax; c.ldT = a.ldT; c.nbin = a.nbin;
    c.T = sd + p.T; c.bbar = sd + p.bbar; c.xB = sd + p.xB; c.colq = sd + p.colq; c.gB = sd + p.gB; c.d = sd + p.d;
    c.xN = sd + p.xN; c.lo = sd + p.lo; c.hi = sd + p.hi; c.glo = sd + p.glo; c.ghi = sd + p.ghi; c.cc = sd + p.cc;
    c.x = sd + p.x; c.rowr = sd + p.rowr; c.gbuf = sd + p.gbuf; c.rhs = sd + p.rhs; c.rownorm = sd + p.rownorm;
    c.sepEff = sd + p.sepEff; c.sepDelta = sd + p.sepDelta; c.sepF0 = sd + p.sepF0; c.rowbuf = sd + p.rowbuf;
    c.redd = sd + p.redd; c.stackBound = sd + p.stackBound;
    c.basis = si + p.basis; c.nb = si + p.nb; c.fracList = si + p.fracList; c.binList = si + p.binList;
    c.pathVar = si + p.pathVar; c.pathVal = si + p.pathVal; c.stackVar = si + p.stackVar; c.stackVal = si + p.stackVal;
    c.stackDepth = si + p.stackDepth; c.redi = si + p.redi; c.sel = si + p.sel; c.keepIdx = si + p.keepIdx;
    c.sint = si + p.sint;
    c.isbin = sb + p.isbin; c.poolState = sb + p.poolState;
    c.H = a.H + (int64_t)b * a.sH;
    c.ptol = a.o.feas_tol; c.itol = a.o.int_tol; c.big = a.o.big_bound;
    c.R = 0; c.z = 0.0; c.fmas = 0.0;

    // ---- load problem data
    const double* cg = a.c + (int64_t)b * a.sc;
    const double* lbg = a.lb + (int64_t)b * a.sbnd;
    const double* ubg = a.ub + (int64_t)b * a.sbnd;
    for (int j = threadIdx.x; j < n; j += kThreads) {
        const double cj = cg[j];
        double l = lbg[j], h = ubg[j];
        const uint8_t ib = a.is_bin[j];
        if (ib) { l = fmax(l, 0.0); h = fmin(h, 1.0); }
        if (!isfinite(l)) l = -c.big;                       // artificial box keeps the slack basis dual feasible
        if (!isfinite(h) && cj < 0.0) h = c.big;
        c.cc[j] = cj; c.d[j] = cj; c.glo[j] = l; c.ghi[j] = h; c.lo[j] = l; c.hi[j] = h;
        c.nb[j] = j; c.isbin[j] = ib; c.xN[j] = 0.0; c.x[j] = 0.0;
    }
    for (int i = threadIdx.x; i < m; i += kThreads) { c.rhs[i] = a.rhs[(int64_t)b * m + i]; c.poolState[i] = 0; }
    if (threadIdx.x < 32) c.sint[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) { int k = 0; for (int j = 0; j < n; ++j) if (c.isbin[j]) c.binList[k++] = j; c.sint[SI_FLAG] = k; }
    __syncthreads();
    c.nbin = c.sint[SI_FLAG];
    {   // row norms (max |h_ij|, >= 1) for the violation ranking
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int i = warp; i < m; i += kWarps) {
            const double* h = c.H + (int64_t)i * n;
            double mx = 0.0;
            for (int j = lane; j < n; j += 32) mx = fmax(mx, fabs(h[j]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) c.rownorm[i] = fmax(mx, 1.0);
        }
    }
    __syncthreads();

    // ---- depth-first branch and cut
    double best = INFINITY;
    int have_inc = 0;
    int pivots_left = a.o.max_pivots;
    int nodes = 0;
    int final_status = HMPC_SOLVE_OPTIMAL;
    double* vout = a.v + (int64_t)b * n;
    if (threadIdx.x == 0) { c.stackVar[0] = -1; c.stackVal[0] = 0; c.stackDepth[0] = 0; c.stackBound[0] = -INFINITY; c.sint[SI_SP] = 1; }
    __syncthreads();
    int cur_depth = 0;   // number of fixings currently applied (path length)
    while (true) {
        __syncthreads();
        const int sp = c.sint[SI_SP];
        if (sp == 0) break;
        if (nodes >= a.o.max_nodes) { final_status = HMPC_SOLVE_NODE_LIMIT; break; }
        if (pivots_left <= 0) { final_status = HMPC_SOLVE_ITER_LIMIT; break; }
        const int top = sp - 1;
        const int bvar = c.stackVar[top], bval = c.stackVal[top], depth = c.stackDepth[top];
        const double bound = c.stackBound[top];
        __syncthreads();
        if (threadIdx.x == 0) c.sint[SI_SP] = top;
        const double gaptol = isfinite(best) ? fmax(1e-9 * fmax(1.0, fabs(best)), a.o.mip_rel_gap * fabs(best)) : 0.0;
        if (bound >= best - gaptol) continue;
        ++nodes;
        // rebuild this node's bounds: undo fixings deeper than depth-1, then apply the new one
        if (threadIdx.x == 0) {
            for (int k = depth > 0 ? depth - 1 : 0; k < cur_depth; ++k) { const int j = c.pathVar[k]; c.lo[j] = c.glo[j]; c.hi[j] = c.ghi[j]; }
            if (depth > 0) { c.pathVar[depth - 1] = bvar; c.pathVal[depth - 1] = bval; c.lo[bvar] = (double)bval; c.hi[bvar] = (double)bval; }
        }
        cur_depth = depth;
        node_setup(c);
        int st = solve_lp(c, best - gaptol, pivots_left);
        if (st == LP_OPT)
            st = cut_loop(c, depth == 0 ? a.o.cut_rounds_root : a.o.cut_rounds_node, a.o.cuts_per_round, best - gaptol,
                          pivots_left, a.o.max_cuts);
        if (st == LP_LIM) { final_status = HMPC_SOLVE_ITER_LIMIT; break; }
        if (st != LP_OPT) continue;
        // exact objective of the node's vertex
        __syncthreads();
        compute_x(c);
        __syncthreads();
        double part = 0.0;
        for (int j = threadIdx.x; j < n; j += kThreads) part += c.cc[j] * c.x[j];
        const double obj = block_sum(c, part);
        c.z = obj;
        if (obj >= best - gaptol) continue;
        ArgVal fr = find_fractional(c);
        if (!(fr.v > c.itol)) {
            // integral: polish (fix every binary at its rounded value, re-solve the continuous part)
            __syncthreads();
            for (int k = threadIdx.x; k < c.nbin; k += kThreads) { const int j = c.binList[k]; const double rv = rint(c.x[j]); c.lo[j] = rv; c.hi[j] = rv; }
            node_setup(c);
            const int st2 = solve_lp(c, INFINITY, pivots_left);
            __syncthreads();
            // restore this node's bounds (path fixings) for the bookkeeping of later pops
            for (int k = threadIdx.x; k < c.nbin; k += kThreads) { const int j = c.binList[k]; c.lo[j] = c.glo[j]; c.hi[j] = c.ghi[j]; }
            __syncthreads();
            if (threadIdx.x == 0) for (int k = 0; k < cur_depth; ++k) { const int j = c.pathVar[k]; c.lo[j] = (double)c.pathVal[k]; c.hi[j] = (double)c.pathVal[k]; }
            __syncthreads();
            if (st2 == LP_LIM) { final_status = HMPC_SOLVE_ITER_LIMIT; break; }
            if (st2 == LP_OPT) {
                compute_x(c);
                __syncthreads();
                double p2 = 0.0;
                for (int j = threadIdx.x; j < n; j += kThreads) p2 += c.cc[j] * c.x[j];
                const double o2 = block_sum(c, p2);
                if (o2 < best) {
                    best = o2; have_inc = 1;
                    for (int j = threadIdx.x; j < n; j += kThreads) vout[j] = c.isbin[j] ? rint(c.x[j]) : c.x[j];
                }
            }
            continue;
        }
        // branch: nearest-integer child is explored first (pushed last)
        if (threadIdx.x == 0) {
            const int j = fr.i;
            const int first = c.x[j] >= 0.5 ? 1 : 0;
            int s = c.sint[SI_SP];
            c.stackVar[s] = j; c.stackVal[s] = 1 - first; c.stackDepth[s] = depth + 1; c.stackBound[s] = obj; ++s;
            c.stackVar[s] = j; c.stackVal[s] = first; c.stackDepth[s] = depth + 1; c.stackBound[s] = obj; ++s;
            c.sint[SI_SP] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int stt = final_status;
        if (!have_inc) {
            stt = (final_status == HMPC_SOLVE_OPTIMAL) ? HMPC_SOLVE_INFEASIBLE : final_status;
        }
        a.status[b] = stt;
        a.obj[b] = have_inc ? best : INFINITY;
        int32_t* s = a.stats + (int64_t)b * 8;
        s[0] = nodes; s[1] = c.sint[SI_PIVOTS]; s[2] = c.sint[SI_CUTS]; s[3] = c.sint[SI_ROWS_ADDED];
        s[4] = c.sint[SI_MAX_ROWS]; s[5] = c.sint[SI_LP]; s[6] = c.sint[SI_PURGES]; s[7] = (int32_t)fmin(c.fmas / 1024.0, 2.0e9);
    }
    if (!have_inc) for (int j = threadIdx.x; j < n; j += kThreads) vout[j] = nan("");
}

static int choose_rmax(int n, int m, int nbin, const hmpc_milp_opts& o, size_t smem_limit, int* ldT_out) {
    const int ldT = n | 1;  // odd leading dimension: conflict-free column reads
    int rmax = o.max_rows > 0 ? o.max_rows : 128;
    if (rmax > m + 96) rmax = m + 96;
    if (rmax < 8) rmax = 8;
    while (rmax > 8 && make_plan(n, m, rmax, ldT, nbin).total > smem_limit) rmax -= 8;
    *ldT_out = ldT;
    return rmax;
}

}  // namespace hmpc

extern "C" void hmpc_milp_default_opts(hmpc_milp_opts* o) {
    if (!o) return;
    o->mip_rel_gap = 0.0; o->int_tol = 1e-6; o->feas_tol = 1e-9; o->big_bound = 1e7;
    o->max_nodes = 200000; o->max_pivots = 2000000; o->max_cuts = 512; o->max_rows = 0;
    o->cut_rounds_root = 30; o->cut_rounds_node = 2; o->cuts_per_round = 8; o->force_general = 0;
}

extern "C" int hmpc_milp_workspace_bytes(int32_t B, int32_t n, int32_t m, const hmpc_milp_opts* opts, size_t* bytes) {
    if (!bytes || B < 0 || n < 0 || m < 0) return HMPC_ERR_ARG;
    (void)opts;
    *bytes = 256;  // everything lives in shared memory; a token scratch keeps the ABI stable
    return HMPC_OK;
}

extern "C" int hmpc_milp_solve_f64(int32_t B, int32_t n, int32_t m, const double* c, int64_t stride_c_b,
                                   const double* H, int64_t stride_H_b, const double* rhs, const double* lb,
                                   const double* ub, int64_t stride_bnd_b, const uint8_t* is_bin,
                                   const hmpc_milp_opts* opts, void* workspace, size_t workspace_bytes, double* v,
                                   double* obj, int32_t* status, int32_t* stats, void* stream) {
    using namespace hmpc;
    (void)workspace; (void)workspace_bytes;
    if (B < 0 || n <= 0 || m < 0 || !c || !lb || !ub || !is_bin || !v || !obj || !status || !stats) return HMPC_ERR_ARG;
    if (m > 0 && (!H || !rhs)) return HMPC_ERR_ARG;
    if (B == 0) return HMPC_OK;
    hmpc_milp_opts o;
    if (opts) o = *opts; else hmpc_milp_default_opts(&o);
    if (o.cuts_per_round > kMaxCutsRound) o.cuts_per_round = kMaxCutsRound;
    if (o.cuts_per_round < 1) o.cuts_per_round = 1;
    const int nbin = n;  // capacity of the per-binary tables; the kernel counts the actual binaries itself
    int dev = 0, smem_optin = 0;
    HMPC_CUDA_TRY(cudaGetDevice(&dev));
    HMPC_CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    MilpArgs a;
    a.B = B; a.n = n; a.m = m; a.c = c; a.sc = stride_c_b; a.H = H; a.sH = stride_H_b; a.rhs = rhs; a.lb = lb; a.ub = ub;
    a.sbnd = stride_bnd_b; a.is_bin = is_bin; a.o = o; a.nbin = nbin;
    a.rmax = choose_rmax(n, m, nbin, o, (size_t)smem_optin, &a.ldT);
    a.v = v; a.obj = obj; a.status = status; a.stats = stats;
    const Plan p = make_plan(n, m, a.rmax, a.ldT, nbin);
    if (p.total > (size_t)smem_optin) return HMPC_ERR_ARG;
    HMPC_CUDA_TRY(cudaFuncSetAttribute(milp_bnc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.total));
    milp_bnc_kernel<<<B, kThreads, p.total, (cudaStream_t)stream>>>(a);
    HMPC_LAUNCH_CHECK("milp_bnc_kernel");
    return HMPC_OK;
}
