// coupling.cu -- price coordination of the fleet for the reference's CENTRALISED micro-grid problem.
//
// The reference hands every device's constraints and objectives to the grid controller and solves one MILP
// (examples/.../micro_grid_agents.py:691-735): the energy price sits on the grid import z_k = max(0, y_k),
// y_k = sum_i P_i u_i,k + (PV + residential demand), not on the devices (micro_grid_control_simulation.py:228-229;
// grid MLD micro_grid_models.py:145-168).  Relaxing the coupling  a_k = sum_i P_i u_i,k  with a multiplier lambda_k
// separates the problem: every agent solves its own exact MILP with lambda_k P_i as the price of u_i,k (the batched
// K3/K4 solve), and the grid side is a one-dimensional piecewise-linear problem per step,
//     min over a in [lo_k, hi_k] of   price_k max(0, a + other_k) - lambda_k a .
// The dual value is a LOWER bound of the centralised optimum, the agents' plans evaluated at the true price are an
// UPPER bound (a feasible centralised plan), and a projected subgradient step with Polyak's step length moves
// lambda.  Kernels (all tiny, one iteration = sums -> dual step -> keep best, no host round trip):
//   coupling_sums_kernel       P_agg[k] (K6 by columns), sum of the agents' objectives, number of failed agents
//   coupling_dual_step_kernel  bounds, subgradient, step (one CTA)
//   coupling_price_cost_kernel writes lambda_k P_i into the agents' cost vectors
//   coupling_keep_best_kernel  copies the plans when the upper bound improved
#include "common.cuh"

namespace hmpc {

// state layout (doubles): see hmpc.h
enum { ST_LB = 0, ST_UB = 1, ST_DUAL = 2, ST_PRIMAL = 3, ST_IMPROVED = 4, ST_ITERS = 5, ST_GNORM2 = 6, ST_BAD = 7 };

// block k < Nt: P_agg[k]; block Nt: sum of obj; block Nt+1: agents whose status is not 0.  Fixed-order tree: the
// result does not depend on scheduling.
__global__ void __launch_bounds__(256) coupling_sums_kernel(int B, int Nt, const double* __restrict__ u, int64_t sb,
                                                            int sk, const double* __restrict__ P_nom,
                                                            const double* __restrict__ obj,
                                                            const int32_t* __restrict__ status,
                                                            double* __restrict__ sums) {
    __shared__ double red[256];
    const int k = blockIdx.x;
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        if (k < Nt) acc += P_nom[b] * u[(int64_t)b * sb + (int64_t)k * sk];
        else if (k == Nt) acc += obj[b];
        else acc += (status && status[b] != 0) ? 1.0 : 0.0;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[k] = red[0];
}

__device__ inline double block_sum_128(double v, double* red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const double out = red[0];
    __syncthreads();
    return out;
}

// Grid side of the dual at one step: value of  min over a in [lo, hi] of  pr max(0, a + r) - lam a  and the component
// of the projected subgradient.  The minimiser is the kink a = -r for 0 < lam < pr, but a whole interval when lam sits
// on a bound of its box ([lo, kink] at lam = 0, [kink, hi] at lam = pr): the member closest to the agents' aggregate
// is taken (minimum-norm subgradient), and a component that would push lam out of [0, pr] is dropped (projection).
// Without this the free-energy hours, where lam = 0 and the agents use a fraction of the surplus, dominate |g|^2
// and Polyak's step crawls.
__device__ inline void grid_side(double a, double r, double pr, double lo, double hi, double lam, double& best,
                                 double& g) {
    const double kink = fmin(fmax(-r, lo), hi);
    const double f_lo = pr * fmax(0.0, lo + r) - lam * lo, f_hi = pr * fmax(0.0, hi + r) - lam * hi,
                 f_k = pr * fmax(0.0, kink + r) - lam * kink;
    best = fmin(f_k, fmin(f_lo, f_hi));
    double arg;
    if (lam <= 0.0) arg = fmin(fmax(a, lo), kink);
    else if (lam >= pr) arg = fmin(fmax(a, kink), hi);
    else arg = kink;
    g = a - arg;
    if ((lam <= 0.0 && g < 0.0) || (lam >= pr && g > 0.0)) g = 0.0;
}

__global__ void __launch_bounds__(128) coupling_dual_step_kernel(int Nt, const double* __restrict__ sums,
                                                                 const double* __restrict__ p_other,
                                                                 const double* __restrict__ price,
                                                                 const double* __restrict__ a_lo,
                                                                 const double* __restrict__ a_hi, double theta,
                                                                 const double* __restrict__ lambda,
                                                                 double* __restrict__ lambda_next,
                                                                 double* __restrict__ state) {
    __shared__ double red[128];
    const double sum_obj = sums[Nt], bad = sums[Nt + 1];
    double lam_agg = 0.0, imp = 0.0, dual_agg = 0.0, g2 = 0.0, infeasible = 0.0;
    // every thread owns the steps k = tid, tid + 128, ... (Nt <= 128 in practice: one step per thread)
    for (int k = threadIdx.x; k < Nt; k += blockDim.x) {
        const double lam = lambda[k], a = sums[k], r = p_other[k], pr = price[k], lo = a_lo[k], hi = a_hi[k];
        lam_agg += lam * a;
        imp += pr * fmax(0.0, a + r);
        if (a < lo - 1e-9 * fmax(1.0, fabs(lo)) || a > hi + 1e-9 * fmax(1.0, fabs(hi))) infeasible += 1.0;
        double best, g;
        grid_side(a, r, pr, lo, hi, lam, best, g);
        dual_agg += best;
        g2 += g * g;
    }
    lam_agg = block_sum_128(lam_agg, red);
    imp = block_sum_128(imp, red);
    dual_agg = block_sum_128(dual_agg, red);
    g2 = block_sum_128(g2, red);
    infeasible = block_sum_128(infeasible, red);
    __shared__ double s_alpha;
    if (threadIdx.x == 0) {
        double alpha = 0.0;
        state[ST_IMPROVED] = 0.0;
        state[ST_ITERS] += 1.0;
        if (bad > 0.0 || !(sum_obj == sum_obj)) {
            state[ST_BAD] += 1.0;                                  // some agent failed: no bound from this iterate
        } else {
            const double dual = sum_obj + dual_agg;                // valid for any lambda
            const double primal = infeasible > 0.0 ? HUGE_VAL : imp + (sum_obj - lam_agg);
            state[ST_DUAL] = dual;
            state[ST_PRIMAL] = primal;
            state[ST_GNORM2] = g2;
            if (dual > state[ST_LB]) state[ST_LB] = dual;
            if (primal < state[ST_UB]) { state[ST_UB] = primal; state[ST_IMPROVED] = 1.0; }
            const double ub = state[ST_UB];
            if (g2 > 0.0 && ub < HUGE_VAL) alpha = theta * fmax(0.0, ub - dual) / g2;
        }
        s_alpha = alpha;
    }
    __syncthreads();
    const double alpha = s_alpha;
    for (int k = threadIdx.x; k < Nt; k += blockDim.x) {
        if (alpha > 0.0) {
            const double pr = price[k], lam = lambda[k];
            double best, g;
            grid_side(sums[k], p_other[k], pr, a_lo[k], a_hi[k], lam, best, g);
            lambda_next[k] = fmin(fmax(lam + alpha * g, 0.0), pr);
        } else {
            lambda_next[k] = lambda[k];
        }
    }
}

__global__ void __launch_bounds__(256) coupling_price_cost_kernel(int B, int Nt, int nv, int col,
                                                                  const double* __restrict__ lambda,
                                                                  const double* __restrict__ P_nom,
                                                                  double* __restrict__ cost_v, int64_t cost_stride) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * Nt) return;
    const int b = (int)(t / Nt), k = (int)(t - (int64_t)b * Nt);
    cost_v[(int64_t)b * cost_stride + (int64_t)k * nv + col] = lambda[k] * P_nom[b];
}

__global__ void __launch_bounds__(256) coupling_keep_best_kernel(int B, int Nt, const double* __restrict__ u,
                                                                 int64_t sb, int sk, const double* __restrict__ lambda,
                                                                 const double* __restrict__ state,
                                                                 double* __restrict__ u_best,
                                                                 double* __restrict__ lambda_best) {
    if (state[ST_IMPROVED] == 0.0) return;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (int64_t)B * Nt) {
        const int b = (int)(t / Nt), k = (int)(t - (int64_t)b * Nt);
        u_best[t] = u[(int64_t)b * sb + (int64_t)k * sk];
    }
    if (lambda_best && t < Nt) lambda_best[t] = lambda[t];
}

// ---- best-response descent on the true centralised cost (a potential game: an agent's own cost change under its
// marginal price equals the change of the total).  Marginal price of agent b at step k, the others fixed:
//   c_bk = price_k [ max(0, A_k - P_b u_bk + P_b + other_k) - max(0, A_k - P_b u_bk + other_k) ]   in [0, price_k P_b]
enum { BR_TOTAL = 0, BR_ACCEPTED = 1, BR_N_ACCEPT = 2, BR_N_REJECT = 3 };

__global__ void __launch_bounds__(256) coupling_response_cost_kernel(int B, int Nt, int nv, int col,
                                                                     const double* __restrict__ agg,
                                                                     const double* __restrict__ v_cur,
                                                                     const double* __restrict__ P_nom,
                                                                     const double* __restrict__ p_other,
                                                                     const double* __restrict__ price,
                                                                     double* __restrict__ cost_v, int64_t stride) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * Nt) return;
    const int b = (int)(t / Nt), k = (int)(t - (int64_t)b * Nt);
    const double P = P_nom[b];
    const double others = agg[k] - P * v_cur[(int64_t)b * stride + (int64_t)k * nv + col] + p_other[k];
    cost_v[(int64_t)b * stride + (int64_t)k * nv + col] = price[k] * (fmax(0.0, others + P) - fmax(0.0, others));
}

// rows [lo, hi) of the fleet take the new plans; the old rows are kept for a possible restore.  One warp per agent.
__global__ void __launch_bounds__(256) coupling_merge_kernel(int lo, int hi, int Nt, int nv, int col,
                                                             const double* __restrict__ v_new,
                                                             const double* __restrict__ obj_new,
                                                             const int32_t* __restrict__ status_new,
                                                             const double* __restrict__ cost_v, int64_t stride,
                                                             double* __restrict__ v_cur, double* __restrict__ pen_cur,
                                                             double* __restrict__ v_bak, double* __restrict__ pen_bak) {
    const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    const int b = lo + warp;
    if (b >= hi) return;
    const int nvt = Nt * nv;
    const double* vn = v_new + (int64_t)warp * nvt;
    double* vc = v_cur + (int64_t)b * stride;
    double* vb = v_bak + (int64_t)warp * nvt;
    double energy = 0.0;
    for (int j = lane; j < nvt; j += 32) {
        const double x = vn[j];
        vb[j] = vc[j];
        vc[j] = x;
        if (j % nv == col) energy += cost_v[(int64_t)b * stride + j] * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) energy += __shfl_xor_sync(0xffffffffu, energy, o);
    if (lane == 0) {
        pen_bak[warp] = pen_cur[b];
        pen_cur[b] = (status_new && status_new[warp] != 0) ? HUGE_VAL : obj_new[warp] - energy;
    }
}

__global__ void __launch_bounds__(128) coupling_accept_kernel(int Nt, const double* __restrict__ sums_cand,
                                                              const double* __restrict__ p_other,
                                                              const double* __restrict__ price,
                                                              const double* __restrict__ a_lo,
                                                              const double* __restrict__ a_hi,
                                                              double* __restrict__ sums_cur,
                                                              double* __restrict__ br) {
    __shared__ double red[128];
    __shared__ int s_ok;
    double imp = 0.0, infeasible = 0.0;
    for (int k = threadIdx.x; k < Nt; k += blockDim.x) {
        const double a = sums_cand[k], lo = a_lo[k], hi = a_hi[k];
        imp += price[k] * fmax(0.0, a + p_other[k]);
        if (a < lo - 1e-9 * fmax(1.0, fabs(lo)) || a > hi + 1e-9 * fmax(1.0, fabs(hi))) infeasible += 1.0;
    }
    imp = block_sum_128(imp, red);
    infeasible = block_sum_128(infeasible, red);
    if (threadIdx.x == 0) {
        const double total = infeasible > 0.0 ? HUGE_VAL : imp + sums_cand[Nt];      // sums[Nt] = sum of penalties
        const double cur = br[BR_TOTAL];
        // the starting plan is accepted against +inf (inf - inf would be NaN in the tolerance)
        const int ok = (total == total) && total < HUGE_VAL &&
                       (cur == HUGE_VAL || total < cur - 1e-12 * fmax(1.0, fabs(cur)));
        s_ok = ok;
        br[BR_ACCEPTED] = ok ? 1.0 : 0.0;
        if (ok) { br[BR_TOTAL] = total; br[BR_N_ACCEPT] += 1.0; } else { br[BR_N_REJECT] += 1.0; }
    }
    __syncthreads();
    if (s_ok)
        for (int k = threadIdx.x; k < Nt + 2; k += blockDim.x) sums_cur[k] = sums_cand[k];
}

__global__ void __launch_bounds__(256) coupling_restore_kernel(int lo, int hi, int nvt, const double* __restrict__ br,
                                                               const double* __restrict__ v_bak,
                                                               const double* __restrict__ pen_bak,
                                                               double* __restrict__ v_cur, int64_t stride,
                                                               double* __restrict__ pen_cur) {
    if (br[BR_ACCEPTED] != 0.0) return;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)(hi - lo) * nvt;
    if (t >= n) return;
    const int w = (int)(t / nvt), j = (int)(t - (int64_t)w * nvt);
    v_cur[(int64_t)(lo + w) * stride + j] = v_bak[t];
    if (j == 0) pen_cur[lo + w] = pen_bak[w];
}

}  // namespace hmpc

extern "C" int hmpc_coupling_price_cost_f64(int32_t B, int32_t Nt, int32_t nv, int32_t col, const double* lambda,
                                            const double* P_nom, double* cost_v, int64_t cost_stride_b, void* stream) {
    using namespace hmpc;
    if (B < 0 || Nt <= 0 || nv <= 0 || col < 0 || col >= nv || !lambda || !P_nom || !cost_v ||
        cost_stride_b < (int64_t)Nt * nv)
        return HMPC_ERR_ARG;
    if (B == 0) return HMPC_OK;
    const int64_t n = (int64_t)B * Nt;
    coupling_price_cost_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, Nt, nv, col, lambda,
                                                                                               P_nom, cost_v,
                                                                                               cost_stride_b);
    HMPC_LAUNCH_CHECK("coupling_price_cost_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_coupling_sums_f64(int32_t B, int32_t Nt, const double* u, int64_t u_stride_b, int32_t u_stride_k,
                                      const double* P_nom, const double* obj, const int32_t* status, double* sums,
                                      void* stream) {
    using namespace hmpc;
    if (B < 0 || Nt <= 0 || !sums || (B > 0 && (!u || !P_nom || !obj))) return HMPC_ERR_ARG;
    coupling_sums_kernel<<<Nt + 2, 256, 0, (cudaStream_t)stream>>>(B, Nt, u, u_stride_b, u_stride_k, P_nom, obj, status,
                                                                   sums);
    HMPC_LAUNCH_CHECK("coupling_sums_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_coupling_dual_step_f64(int32_t Nt, const double* sums, const double* p_other, const double* price,
                                           const double* a_lo, const double* a_hi, double theta,
                                           const double* lambda, double* lambda_next, double* state, void* stream) {
    using namespace hmpc;
    if (Nt <= 0 || !sums || !p_other || !price || !a_lo || !a_hi || !lambda || !lambda_next || lambda == lambda_next ||
        !state || !(theta > 0.0))
        return HMPC_ERR_ARG;
    coupling_dual_step_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(Nt, sums, p_other, price, a_lo, a_hi, theta, lambda,
                                                                   lambda_next, state);
    HMPC_LAUNCH_CHECK("coupling_dual_step_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_coupling_keep_best_f64(int32_t B, int32_t Nt, const double* u, int64_t u_stride_b,
                                           int32_t u_stride_k, const double* lambda, const double* state,
                                           double* u_best, double* lambda_best, void* stream) {
    using namespace hmpc;
    if (B < 0 || Nt <= 0 || !state || !lambda || (B > 0 && (!u || !u_best))) return HMPC_ERR_ARG;
    const int64_t n = (int64_t)B * Nt > Nt ? (int64_t)B * Nt : Nt;
    coupling_keep_best_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, Nt, u, u_stride_b,
                                                                                              u_stride_k, lambda, state,
                                                                                              u_best, lambda_best);
    HMPC_LAUNCH_CHECK("coupling_keep_best_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_coupling_response_cost_f64(int32_t B, int32_t Nt, int32_t nv, int32_t col, const double* agg,
                                               const double* v_cur, const double* P_nom, const double* p_other,
                                               const double* price, double* cost_v, int64_t stride_b, void* stream) {
    using namespace hmpc;
    if (B < 0 || Nt <= 0 || nv <= 0 || col < 0 || col >= nv || !agg || !v_cur || !P_nom || !p_other || !price ||
        !cost_v || stride_b < (int64_t)Nt * nv)
        return HMPC_ERR_ARG;
    if (B == 0) return HMPC_OK;
    const int64_t n = (int64_t)B * Nt;
    coupling_response_cost_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        B, Nt, nv, col, agg, v_cur, P_nom, p_other, price, cost_v, stride_b);
    HMPC_LAUNCH_CHECK("coupling_response_cost_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_coupling_merge_f64(int32_t lo, int32_t hi, int32_t Nt, int32_t nv, int32_t col,
                                       const double* v_new, const double* obj_new, const int32_t* status_new,
                                       const double* cost_v, int64_t stride_b, double* v_cur, double* pen_cur,
                                       double* v_bak, double* pen_bak, void* stream) {
    using namespace hmpc;
    if (lo < 0 || hi < lo || Nt <= 0 || nv <= 0 || col < 0 || col >= nv || stride_b < (int64_t)Nt * nv) return HMPC_ERR_ARG;
    if (hi == lo) return HMPC_OK;
    if (!v_new || !obj_new || !cost_v || !v_cur || !pen_cur || !v_bak || !pen_bak) return HMPC_ERR_ARG;
    const int64_t threads = (int64_t)(hi - lo) * 32;
    coupling_merge_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        lo, hi, Nt, nv, col, v_new, obj_new, status_new, cost_v, stride_b, v_cur, pen_cur, v_bak, pen_bak);
    HMPC_LAUNCH_CHECK("coupling_merge_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_coupling_accept_f64(int32_t Nt, const double* sums_cand, const double* p_other, const double* price,
                                        const double* a_lo, const double* a_hi, double* sums_cur, double* br_state,
                                        void* stream) {
    using namespace hmpc;
    if (Nt <= 0 || !sums_cand || !p_other || !price || !a_lo || !a_hi || !sums_cur || !br_state) return HMPC_ERR_ARG;
    coupling_accept_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(Nt, sums_cand, p_other, price, a_lo, a_hi, sums_cur,
                                                                br_state);
    HMPC_LAUNCH_CHECK("coupling_accept_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_coupling_restore_f64(int32_t lo, int32_t hi, int32_t nvt, const double* br_state,
                                         const double* v_bak, const double* pen_bak, double* v_cur, int64_t stride_b,
                                         double* pen_cur, void* stream) {
    using namespace hmpc;
    if (lo < 0 || hi < lo || nvt <= 0 || stride_b < nvt || !br_state) return HMPC_ERR_ARG;
    if (hi == lo) return HMPC_OK;
    if (!v_bak || !pen_bak || !v_cur || !pen_cur) return HMPC_ERR_ARG;
    const int64_t n = (int64_t)(hi - lo) * nvt;
    coupling_restore_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lo, hi, nvt, br_state, v_bak,
                                                                                            pen_bak, v_cur, stride_b,
                                                                                            pen_cur);
    HMPC_LAUNCH_CHECK("coupling_restore_kernel");
    return HMPC_OK;
}
