// miqp_admm.cu -- K3q/K4q: mixed-integer QUADRATIC programs of any MLD, by branch and bound over an ADMM
// (operator-splitting) QP relaxation.  Replaces the cvxpy -> Gurobi/CPLEX call inside
// ConstraintSolvedController.solve (reference: controllers/controller_base.py:509-512) when the cost carries
// Quadratic / L22 atoms with dense weights (controllers/components/objective_atoms.py:320-343, 185-206), L1 / Linf
// atoms (:338-363; as epigraph columns and rows, added by the host) or non-linear rate atoms (:297-305) on an MLD
// outside the scalar-state class of stage_dp.cu:
//
//     minimise 0.5 v'P v + c'v   s.t.   H v <= rhs,  lb <= v <= ub,  v_j in {0,1} where is_bin_j       (P >= 0)
//
// One CTA per problem.
//   * relaxation: ADMM on  min 0.5 x'Px + q'x,  z = [H; I] x,  z_H <= rhs,  lb <= z_I <= ub  (the OSQP splitting, with
//     the variable bounds as rows of the constraint operator, so that a node of the search only changes the projection,
//     never the linear system).  K = P + sigma I + rho (H'H + I) is factorised ONCE per problem (Cholesky, then the
//     explicit inverse, one column per thread), so an iteration is three dense matrix-vector products with coalesced
//     reads (K^-1 symmetric; H kept in both layouts) and three barriers.  rho is adapted at the root only (residual
//     balancing, refactorising).  Rows of H and the cost are scaled to unit size first.
//   * search: depth-first over the binaries, most fractional first, nearer rounding first; a child starts from its
//     parent's iterates (x, z, y stay in shared memory), is pruned when its relaxation value reaches the incumbent, or
//     when the iterates produce an infeasibility certificate (a direction dy with [H; I]'dy = 0 and a negative support
//     value).  An integral relaxation is POLISHED (every binary fixed, one more warm-started solve) before it becomes
//     the incumbent, so the reported point is the minimiser of the continuous part for its binaries.
// Accuracy is that of the relaxation tolerance (default 1e-9 scaled): objectives agree with an exact solver to ~1e-7
// relative, which is what the parity tests ask of it (1e-6).  This is the general, not the fast path: the DEWH fleet
// of the reference example never comes here.
#include <string.h>
#include "common.cuh"

namespace hmpc {

constexpr int kQpThreads = 256;
constexpr int kQpMaxBin = 512;
constexpr int kQpMaskWords = kQpMaxBin / 32;

struct QpArgs {
    int B, n, m;
    const double* P; int64_t sP;        // [B|1, n, n] or NULL
    const double* c; int64_t sc;        // [B|1, n]
    const double* H; int64_t sH;        // [B|1, m, n]
    const double* rhs;                  // [B, m]
    const double* lb; const double* ub; // [n]
    const uint8_t* is_bin;              // [n]
    hmpc_miqp_opts o;
    double* ws; int64_t ws_stride;      // per problem, doubles
    double* v; double* obj; int32_t* status; int32_t* stats;
};

struct QpLayout {      // offsets (doubles) into a problem's global workspace
    int64_t K, Kinv, Hs, HT, nodes, total;
};
__host__ __device__ inline QpLayout qp_layout(int n, int m, int nbin) {
    QpLayout L; int64_t o = 0;
    L.K = o; o += (int64_t)n * n;
    L.Kinv = o; o += (int64_t)n * n;
    L.Hs = o; o += (int64_t)m * n;
    L.HT = o; o += (int64_t)m * n;
    L.nodes = o; o += (int64_t)(2 * nbin + 4) * (kQpMaskWords + 1);      // (fixed mask, value mask) words + bound
    L.total = (o + 3) & ~(int64_t)3;
    return L;
}

__device__ __forceinline__ double block_reduce_max(double v, double* red) {
    // max over the CTA (absolute values are the callers' business); red: kQpThreads / 32 doubles of shared memory
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < kQpThreads / 32; ++i) r = fmax(r, red[i]);
    return r;
}
__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < kQpThreads / 32; ++i) r += red[i];
    return r;
}

__global__ void __launch_bounds__(kQpThreads) miqp_admm_kernel(const QpArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[kQpThreads / 32];
    __shared__ int s_flag, s_branch;
    const int b = blockIdx.x, tid = threadIdx.x, nthr = kQpThreads;
    const int n = A.n, m = A.m;
    const double* Pm = A.P ? A.P + (int64_t)b * A.sP : nullptr;
    const double* cv = A.c + (int64_t)b * A.sc;
    const double* Hm = A.H + (int64_t)b * A.sH;
    const double* rhs = A.rhs + (int64_t)b * m;
    // ---- shared memory: vectors of the iteration
    double* sm = reinterpret_cast<double*>(smem_raw);
    double* x = sm;            double* xt = x + n;        double* zI = xt + n;       double* yI = zI + n;
    double* q = yI + n;        double* rv = q + n;        double* lbn = rv + n;      double* ubn = lbn + n;
    double* bestx = ubn + n;   double* yIp = bestx + n;   double* Px = yIp + n;
    double* zH = Px + n;       double* yH = zH + m;       double* rs = yH + m;       double* er = rs + m;
    double* tm = er + m;       double* yHp = tm + m;
    int* binidx = reinterpret_cast<int*>(yHp + m);          // [nbin]
    // ---- binaries
    __shared__ int s_nbin;
    if (tid == 0) {
        int nb = 0;
        for (int j = 0; j < n; ++j) if (A.is_bin[j]) { if (nb < kQpMaxBin) binidx[nb] = j; ++nb; }
        s_nbin = nb;
    }
    __syncthreads();
    const int nbin = s_nbin;
    int32_t* st_out = A.stats + (int64_t)b * 8;
    if (nbin > kQpMaxBin) {
        for (int j = tid; j < n; j += nthr) A.v[(int64_t)b * n + j] = nan("");
        if (tid == 0) { A.status[b] = HMPC_SOLVE_UNSUPPORTED; A.obj[b] = INFINITY; for (int i = 0; i < 8; ++i) st_out[i] = 0; }
        return;
    }
    const QpLayout L = qp_layout(n, m, nbin);
    double* W = A.ws + (int64_t)b * A.ws_stride;
    double* K = W + L.K; double* Kinv = W + L.Kinv; double* Hs = W + L.Hs; double* HT = W + L.HT;
    unsigned* nodes = reinterpret_cast<unsigned*>(W + L.nodes);
    const int node_words = 2 * (kQpMaskWords + 1);           // unsigned words per node: fixed[16], value[16], bound (2)
    // ---- scaling: rows of H to unit infinity norm, the cost to unit size
    for (int i = tid; i < m; i += nthr) {
        double mx = 0.0;
        for (int j = 0; j < n; ++j) mx = fmax(mx, fabs(Hm[(int64_t)i * n + j]));
        const double e = mx > 1e-12 ? 1.0 / mx : 1.0;
        er[i] = e; rs[i] = rhs[i] * e;
    }
    double cmx = 0.0;
    for (int j = tid; j < n; j += nthr) {
        cmx = fmax(cmx, fabs(cv[j]));
        if (Pm) cmx = fmax(cmx, fabs(Pm[(int64_t)j * n + j]));
    }
    cmx = block_reduce_max(cmx, red);
    const double cs = cmx > 1e-12 ? 1.0 / cmx : 1.0;
    __syncthreads();
    for (int64_t e = tid; e < (int64_t)m * n; e += nthr) {
        const int i = (int)(e / n), j = (int)(e - (int64_t)i * n);
        const double hv = Hm[e] * er[i];
        Hs[e] = hv; HT[(int64_t)j * m + i] = hv;
    }
    for (int j = tid; j < n; j += nthr) {
        q[j] = cv[j] * cs; lbn[j] = A.lb[j]; ubn[j] = A.ub[j];
        x[j] = fmin(fmax(0.0, lbn[j]), ubn[j]); zI[j] = x[j]; yI[j] = 0.0; bestx[j] = nan("");
    }
    for (int i = tid; i < m; i += nthr) { zH[i] = fmin(0.0, rs[i]); yH[i] = 0.0; }
    __syncthreads();

    const double sigma = 1e-6, alpha = 1.6;
    double rho = A.o.rho > 0.0 ? A.o.rho : 0.1;
    // ---- factorisation: K = cs P + sigma I + rho (Hs'Hs + I), Cholesky in place, inverse by columns
    auto factorise = [&](double rho_) {
        for (int64_t e = tid; e < (int64_t)n * n; e += nthr) {
            const int a = (int)(e / n), c2 = (int)(e - (int64_t)a * n);
            if (c2 > a) continue;                       // lower triangle
            double s = 0.0;
            for (int i = 0; i < m; ++i) s = fma(HT[(int64_t)a * m + i], HT[(int64_t)c2 * m + i], s);
            s *= rho_;
            if (Pm) s += cs * 0.5 * (Pm[(int64_t)a * n + c2] + Pm[(int64_t)c2 * n + a]);
            if (a == c2) s += sigma + rho_;
            K[e] = s;
        }
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            if (tid == 0) K[(int64_t)k * n + k] = sqrt(fmax(K[(int64_t)k * n + k], 1e-300));
            __syncthreads();
            const double d = K[(int64_t)k * n + k];
            for (int i = k + 1 + tid; i < n; i += nthr) K[(int64_t)i * n + k] /= d;
            __syncthreads();
            // trailing update: row i gets  K[i][j] -= K[i][k] K[j][k]  (k < j <= i), rows over threads
            for (int i = k + 1 + tid; i < n; i += nthr) {
                const double lik = K[(int64_t)i * n + k];
                for (int j = k + 1; j <= i; ++j) K[(int64_t)i * n + j] = fma(-lik, K[(int64_t)j * n + k], K[(int64_t)i * n + j]);
            }
            __syncthreads();
        }
        // K^-1 column by column: L y = e_c, L' w = y  (column c of a thread lives in Kinv[. * n + c]: coalesced)
        for (int c2 = tid; c2 < n; c2 += nthr) {
            for (int i = 0; i < n; ++i) {
                double s = i == c2 ? 1.0 : 0.0;
                for (int j = (c2 < i ? c2 : i); j < i; ++j) s = fma(-K[(int64_t)i * n + j], Kinv[(int64_t)j * n + c2], s);
                Kinv[(int64_t)i * n + c2] = i < c2 ? 0.0 : s / K[(int64_t)i * n + i];
            }
            for (int i = n - 1; i >= 0; --i) {
                double s = Kinv[(int64_t)i * n + c2];
                for (int j = i + 1; j < n; ++j) s = fma(-K[(int64_t)j * n + i], Kinv[(int64_t)j * n + c2], s);
                Kinv[(int64_t)i * n + c2] = s / K[(int64_t)i * n + i];
            }
        }
        __syncthreads();
    };

    // ---- one ADMM solve from the current iterates with the current node bounds (lbn / ubn).
    // returns 0 converged, 1 infeasible, 2 iteration limit; *objv = scaled objective at x
    long long iters_total = 0;
    auto admm = [&](int max_iter, double eps, bool adapt, double* objv) -> int {
        int result = 2;
        for (int it = 1; it <= max_iter; ++it) {
            // rv = sigma x - q + Hs'(rho zH - yH) + (rho zI - yI)
            for (int i = tid; i < m; i += nthr) tm[i] = fma(rho, zH[i], -yH[i]);
            __syncthreads();
            for (int j = tid; j < n; j += nthr) {
                double s = fma(sigma, x[j], -q[j]) + fma(rho, zI[j], -yI[j]);
                for (int i = 0; i < m; ++i) s = fma(Hs[(int64_t)i * n + j], tm[i], s);
                rv[j] = s;
            }
            __syncthreads();
            for (int j = tid; j < n; j += nthr) {
                double s = 0.0;
                for (int i = 0; i < n; ++i) s = fma(Kinv[(int64_t)i * n + j], rv[i], s);
                xt[j] = s;
            }
            __syncthreads();
            const bool check = (it % 20 == 0) || it == max_iter;
            if (check) { for (int i = tid; i < m; i += nthr) yHp[i] = yH[i]; for (int j = tid; j < n; j += nthr) yIp[j] = yI[j]; }
            for (int i = tid; i < m; i += nthr) {
                double zt = 0.0;
                for (int j = 0; j < n; ++j) zt = fma(HT[(int64_t)j * m + i], xt[j], zt);
                const double mix = fma(alpha, zt, (1.0 - alpha) * zH[i]);
                const double zn = fmin(mix + yH[i] / rho, rs[i]);
                yH[i] += rho * (mix - zn);
                zH[i] = zn;
            }
            for (int j = tid; j < n; j += nthr) {
                const double mix = fma(alpha, xt[j], (1.0 - alpha) * zI[j]);
                const double zn = fmin(fmax(mix + yI[j] / rho, lbn[j]), ubn[j]);
                yI[j] += rho * (mix - zn);
                zI[j] = zn;
                x[j] = fma(alpha, xt[j], (1.0 - alpha) * x[j]);
            }
            __syncthreads();
            if (!check) continue;
            // ---- residuals (scaled space)
            double rp = 0.0, nz = 0.0, nAx = 0.0;
            for (int i = tid; i < m; i += nthr) {
                double ax = 0.0;
                for (int j = 0; j < n; ++j) ax = fma(HT[(int64_t)j * m + i], x[j], ax);
                rp = fmax(rp, fabs(ax - zH[i])); nz = fmax(nz, fabs(zH[i])); nAx = fmax(nAx, fabs(ax));
            }
            for (int j = tid; j < n; j += nthr) { rp = fmax(rp, fabs(x[j] - zI[j])); nz = fmax(nz, fabs(zI[j])); nAx = fmax(nAx, fabs(x[j])); }
            double rd = 0.0, nPx = 0.0, nAty = 0.0, nq = 0.0, obj = 0.0;
            for (int j = tid; j < n; j += nthr) {
                double px = 0.0;
                if (Pm) for (int i = 0; i < n; ++i) px = fma(0.5 * (Pm[(int64_t)i * n + j] + Pm[(int64_t)j * n + i]), x[i], px);
                px *= cs;
                double aty = yI[j];
                for (int i = 0; i < m; ++i) aty = fma(Hs[(int64_t)i * n + j], yH[i], aty);
                rd = fmax(rd, fabs(px + q[j] + aty)); nPx = fmax(nPx, fabs(px)); nAty = fmax(nAty, fabs(aty)); nq = fmax(nq, fabs(q[j]));
                obj += x[j] * (0.5 * px + q[j]);
                Px[j] = px;
            }
            rp = block_reduce_max(rp, red); nz = block_reduce_max(nz, red); nAx = block_reduce_max(nAx, red);
            rd = block_reduce_max(rd, red); nPx = block_reduce_max(nPx, red); nAty = block_reduce_max(nAty, red);
            nq = block_reduce_max(nq, red);
            obj = block_reduce_sum(obj, red);
            *objv = obj;
            const double ep = eps + eps * fmax(nAx, nz), ed = eps + eps * fmax(fmax(nPx, nAty), nq);
            if (rp <= ep && rd <= ed) { result = 0; iters_total += it; break; }
            // ---- infeasibility certificate from dy = y - y(20 iterations ago)
            double ndy = 0.0, sup = 0.0, bad = 0.0;
            for (int i = tid; i < m; i += nthr) {
                const double dy = yH[i] - yHp[i];
                ndy = fmax(ndy, fabs(dy));
                if (dy > 0.0) sup += rs[i] * dy; else bad = fmax(bad, -dy);      // rows have no lower side
            }
            for (int j = tid; j < n; j += nthr) {
                const double dy = yI[j] - yIp[j];
                ndy = fmax(ndy, fabs(dy));
                if (dy > 0.0) { if (isfinite(ubn[j])) sup += ubn[j] * dy; else bad = fmax(bad, dy); }
                else if (dy < 0.0) { if (isfinite(lbn[j])) sup += lbn[j] * dy; else bad = fmax(bad, -dy); }
            }
            ndy = block_reduce_max(ndy, red); bad = block_reduce_max(bad, red);
            sup = block_reduce_sum(sup, red);
            double natdy = 0.0;
            for (int j = tid; j < n; j += nthr) {
                double s = yI[j] - yIp[j];
                for (int i = 0; i < m; ++i) s = fma(Hs[(int64_t)i * n + j], yH[i] - yHp[i], s);
                natdy = fmax(natdy, fabs(s));
            }
            natdy = block_reduce_max(natdy, red);
            const double einf = 1e-7;
            if (ndy > 1e-12 && natdy <= einf * ndy && bad <= einf * ndy && sup < -einf * ndy) { result = 1; iters_total += it; break; }
            // ---- residual balancing (root only): rho <- rho sqrt(rp_rel / rd_rel), refactorise
            if (adapt && it % 100 == 0) {
                const double rpn = rp / fmax(fmax(nAx, nz), 1e-12), rdn = rd / fmax(fmax(fmax(nPx, nAty), nq), 1e-12);
                const double ratio = sqrt(rpn / fmax(rdn, 1e-30));
                if (ratio > 5.0 || ratio < 0.2) {
                    const double rho_new = fmin(fmax(rho * ratio, 1e-6), 1e6);
                    __syncthreads();
                    rho = rho_new;
                    factorise(rho);
                }
            }
            if (it == max_iter) iters_total += it;
        }
        __syncthreads();
        return result;
    };

    factorise(rho);
    // ---- branch and bound
    const double eps = A.o.eps > 0.0 ? A.o.eps : 1e-9;
    const double int_tol = A.o.int_tol > 0.0 ? A.o.int_tol : 1e-6;
    const int max_iter = A.o.max_iter > 0 ? A.o.max_iter : 50000;
    double best = INFINITY;
    int nnodes = 0, status = HMPC_SOLVE_INFEASIBLE, improvements = 0;
    bool limit = false;
    int sp = 0;
    auto node_ptr = [&](int i) { return nodes + (int64_t)i * node_words; };
    if (tid == 0) { unsigned* p = node_ptr(0); for (int w2 = 0; w2 < node_words; ++w2) p[w2] = 0; double lbv = -INFINITY; memcpy(p + 2 * kQpMaskWords, &lbv, 8); }
    sp = 1;
    __syncthreads();
    bool first = true;
    while (sp > 0) {
        if (nnodes >= A.o.max_nodes) { limit = true; break; }
        --sp;
        const unsigned* nd = node_ptr(sp);
        double nbound; memcpy(&nbound, nd + 2 * kQpMaskWords, 8);
        const double tol = isfinite(best) ? fmax(1e-9 * fmax(1.0, fabs(best)), A.o.mip_rel_gap * fabs(best)) : 0.0;
        if (nbound >= best - tol) { __syncthreads(); continue; }
        // node bounds
        for (int t = tid; t < nbin; t += nthr) {
            const int j = binidx[t];
            const bool fx = (nd[t >> 5] >> (t & 31)) & 1u, vl = (nd[kQpMaskWords + (t >> 5)] >> (t & 31)) & 1u;
            lbn[j] = fx ? (vl ? 1.0 : 0.0) : fmax(A.lb[j], 0.0);
            ubn[j] = fx ? (vl ? 1.0 : 0.0) : fmin(A.ub[j], 1.0);
        }
        // copy the node's masks (the slot may be overwritten by its children)
        __shared__ unsigned cur_fix[kQpMaskWords], cur_val[kQpMaskWords];
        __syncthreads();
        if (tid < kQpMaskWords) { cur_fix[tid] = nd[tid]; cur_val[tid] = nd[kQpMaskWords + tid]; }
        __syncthreads();
        double objv = 0.0;
        int r = admm(max_iter, eps, first, &objv);
        first = false;
        ++nnodes;
        if (r == 1) continue;                               // infeasible node
        if (r == 2) status = HMPC_SOLVE_ITER_LIMIT;         // (kept going with what the iterates say)
        if (objv >= best - tol) continue;
        // most fractional binary
        double frac = -1.0; int which = -1;
        for (int t = tid; t < nbin; t += nthr) {
            if ((cur_fix[t >> 5] >> (t & 31)) & 1u) continue;
            const double xv = zI[binidx[t]], f = fabs(xv - rint(xv));
            if (f > frac) { frac = f; which = t; }
        }
        {   // arg-max over the CTA
            const double mx = block_reduce_max(frac, red);
            if (tid == 0) s_branch = -1;
            __syncthreads();
            if (which >= 0 && frac == mx && mx > int_tol) atomicMax(&s_branch, which);
            __syncthreads();
        }
        const int br = s_branch;
        if (br < 0) {
            // integral relaxation: fix every binary at its rounding and polish
            for (int t = tid; t < nbin; t += nthr) { const int j = binidx[t]; const double v2 = rint(fmin(fmax(zI[j], 0.0), 1.0)); lbn[j] = v2; ubn[j] = v2; }
            __syncthreads();
            double objp = 0.0;
            const int r2 = admm(max_iter, eps, false, &objp);
            if (r2 == 1) continue;        // (cannot happen for an integral feasible relaxation, up to tolerances)
            if (objp < best - tol) {
                best = objp; ++improvements;
                for (int j = tid; j < n; j += nthr) bestx[j] = A.is_bin[j] ? rint(zI[j]) : zI[j];
                __syncthreads();
            }
            continue;
        }
        // two children; the nearer rounding is explored first (pushed last)
        const int jv = binidx[br];
        const bool up_first = zI[jv] >= 0.5;
        if (sp + 2 > 2 * nbin + 4) { limit = true; break; }
        if (tid < 2) {
            const bool val = tid == 0 ? !up_first : up_first;
            unsigned* p = node_ptr(sp + tid);
            for (int w2 = 0; w2 < kQpMaskWords; ++w2) { p[w2] = cur_fix[w2]; p[kQpMaskWords + w2] = cur_val[w2]; }
            p[br >> 5] |= 1u << (br & 31);
            if (val) p[kQpMaskWords + (br >> 5)] |= 1u << (br & 31);
            memcpy(p + 2 * kQpMaskWords, &objv, 8);
        }
        sp += 2;
        __syncthreads();
    }
    // ---- result
    const bool have = isfinite(best);
    double* vout = A.v + (int64_t)b * n;
    for (int j = tid; j < n; j += nthr) vout[j] = have ? bestx[j] : nan("");
    if (tid == 0) {
        A.obj[b] = have ? best / cs : INFINITY;
        A.status[b] = limit ? HMPC_SOLVE_NODE_LIMIT : (have ? (status == HMPC_SOLVE_ITER_LIMIT ? HMPC_SOLVE_ITER_LIMIT : HMPC_SOLVE_OPTIMAL) : HMPC_SOLVE_INFEASIBLE);
        st_out[0] = nnodes; st_out[1] = (int32_t)fmin((double)iters_total, 2.0e9); st_out[2] = improvements; st_out[3] = nbin;
        st_out[4] = 0; st_out[5] = 0; st_out[6] = 0;
        st_out[7] = (int32_t)fmin((double)iters_total * (2.0 * m * n + (double)n * n) / 1000.0, 2.0e9);
    }
}

static size_t qp_smem_bytes(int n, int m) { return (size_t)(11 * n + 6 * m) * 8 + (size_t)kQpMaxBin * 4 + 64; }

}  // namespace hmpc

extern "C" void hmpc_miqp_default_opts(hmpc_miqp_opts* o) {
    if (!o) return;
    o->mip_rel_gap = 0.0; o->int_tol = 1e-6; o->eps = 1e-9; o->rho = 0.1; o->max_nodes = 100000; o->max_iter = 50000;
}

extern "C" int hmpc_miqp_workspace_bytes(int32_t B, int32_t n, int32_t m, size_t* bytes) {
    using namespace hmpc;
    if (!bytes || B < 0 || n < 1 || m < 0) return HMPC_ERR_ARG;
    const QpLayout L = qp_layout(n, m, n < kQpMaxBin ? n : kQpMaxBin);
    *bytes = (size_t)B * (size_t)L.total * sizeof(double) + 256;
    return HMPC_OK;
}

extern "C" int hmpc_miqp_solve_f64(int32_t B, int32_t n, int32_t m, const double* P, int64_t stride_P_b,
                                   const double* c, int64_t stride_c_b, const double* H, int64_t stride_H_b,
                                   const double* rhs, const double* lb, const double* ub, const uint8_t* is_bin,
                                   const hmpc_miqp_opts* opts, void* workspace, size_t workspace_bytes,
                                   double* v, double* obj, int32_t* status, int32_t* stats, void* stream) {
    using namespace hmpc;
    if (B < 0 || n < 1 || m < 0 || !c || (m > 0 && (!H || !rhs)) || !lb || !ub || !is_bin || !v || !obj || !status || !stats)
        return HMPC_ERR_ARG;
    if (B == 0) return HMPC_OK;
    QpArgs a;
    a.B = B; a.n = n; a.m = m; a.P = P; a.sP = stride_P_b; a.c = c; a.sc = stride_c_b; a.H = H; a.sH = stride_H_b;
    a.rhs = rhs; a.lb = lb; a.ub = ub; a.is_bin = is_bin;
    if (opts) a.o = *opts; else hmpc_miqp_default_opts(&a.o);
    size_t need = 0;
    hmpc_miqp_workspace_bytes(B, n, m, &need);
    if (!workspace || workspace_bytes < need) return HMPC_ERR_WORKSPACE;
    const QpLayout L = qp_layout(n, m, n < kQpMaxBin ? n : kQpMaxBin);
    a.ws = reinterpret_cast<double*>(workspace); a.ws_stride = L.total;
    a.v = v; a.obj = obj; a.status = status; a.stats = stats;
    int dev = 0, smem_optin = 0;
    HMPC_CUDA_TRY(cudaGetDevice(&dev));
    HMPC_CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t smem = qp_smem_bytes(n, m);
    if (smem + 1024 > (size_t)smem_optin) return HMPC_ERR_ARG;
    HMPC_CUDA_TRY(cudaFuncSetAttribute(miqp_admm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    miqp_admm_kernel<<<B, kQpThreads, smem, (cudaStream_t)stream>>>(a);
    HMPC_LAUNCH_CHECK("miqp_admm_kernel");
    return HMPC_OK;
}
