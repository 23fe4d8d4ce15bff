// assemble.cu -- K2 and friends: per-step problem data from the condensed matrices.
//   hmpc_constraint_rhs_f64 : rhs = H_x x0 + H_omega w + H_5, or the scenario-robust row-min form
//                             (reference: controllers/controller_base.py:440-452)
//   hmpc_predict_f64        : x~ / y~ affine predictions (controllers/components/variables.py:245-286)
//   hmpc_linear_cost_f64    : Linear cost atoms pulled back to v-space (objective_atoms.py:308-318)
// All three are HBM-read-bound mat-vec products: one warp per output row, lanes stride the row so that
// every load instruction covers 256 contiguous bytes.
#include "common.cuh"

namespace hmpc {

// rhs[b, r] = H_x[b, r, :] x0[b] + red_s( H_w[b, r, :] W[b, :, s] ) + H_5[b, r]
__global__ void __launch_bounds__(256) constraint_rhs_kernel(int B, int rows, int rows_full, int nx, int nwt,
                                                             const double* __restrict__ H_x,
                                                             const double* __restrict__ H_w,
                                                             const double* __restrict__ H_5,
                                                             const double* __restrict__ x0,
                                                             const double* __restrict__ w, int S,
                                                             double* __restrict__ rhs) {
    extern __shared__ double sw[];  // [nwt * max(S,1)] this agent's disturbance forecast / scenarios
    const int b = blockIdx.x;
    const int Sc = S > 0 ? S : 1;
    for (int e = threadIdx.x; e < nwt * Sc; e += blockDim.x) sw[e] = w[(int64_t)b * nwt * Sc + e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int r = blockIdx.y * nwarp + warp; r < rows; r += gridDim.y * nwarp) {
        const double* hw = H_w + ((int64_t)b * rows_full + r) * nwt;
        double red;
        if (S <= 0) {
            double acc = 0.0;
            for (int c = lane; c < nwt; c += 32) acc += hw[c] * sw[c];
            red = warp_sum(acc);
        } else {
            red = INFINITY;
            for (int s0 = 0; s0 < S; s0 += 4) {   // 4 scenarios per pass over the row
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                for (int c = lane; c < nwt; c += 32) {
                    const double h = hw[c];
                    const double* ws = sw + c * S + s0;
                    a0 += h * ws[0];
                    if (s0 + 1 < S) a1 += h * ws[1];
                    if (s0 + 2 < S) a2 += h * ws[2];
                    if (s0 + 3 < S) a3 += h * ws[3];
                }
                a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
                red = fmin(red, a0);
                if (s0 + 1 < S) red = fmin(red, a1);
                if (s0 + 2 < S) red = fmin(red, a2);
                if (s0 + 3 < S) red = fmin(red, a3);
            }
        }
        if (lane == 0) {
            double hx = 0.0;
            for (int c = 0; c < nx; ++c) hx += H_x[((int64_t)b * rows_full + r) * nx + c] * x0[(int64_t)b * nx + c];
            rhs[(int64_t)b * rows + r] = hx + red + H_5[(int64_t)b * rows_full + r];
        }
    }
}

// out[b, r] = M_x[b,r,:] x0[b] + M_v[b,r,:] v[b] + M_w[b,r,:] w[b] + M_5[b,r]
__global__ void __launch_bounds__(256) predict_kernel(int B, int R, int nx, int nvt, int nwt,
                                                      const double* __restrict__ M_x, const double* __restrict__ M_v,
                                                      const double* __restrict__ M_w, const double* __restrict__ M_5,
                                                      const double* __restrict__ x0, const double* __restrict__ v,
                                                      const double* __restrict__ w, double* __restrict__ out) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int r = blockIdx.y * nwarp + warp; r < R; r += gridDim.y * nwarp) {
        double acc = 0.0;
        if (v) {
            const double* mv = M_v + ((int64_t)b * R + r) * nvt;
            const double* vb = v + (int64_t)b * nvt;
            for (int c = lane; c < nvt; c += 32) acc += mv[c] * vb[c];
        }
        if (w) {
            const double* mw = M_w + ((int64_t)b * R + r) * nwt;
            const double* wb = w + (int64_t)b * nwt;
            for (int c = lane; c < nwt; c += 32) acc += mw[c] * wb[c];
        }
        for (int c = lane; c < nx; c += 32) acc += M_x[((int64_t)b * R + r) * nx + c] * x0[(int64_t)b * nx + c];
        acc = warp_sum(acc);
        if (lane == 0) out[(int64_t)b * R + r] = acc + (M_5 ? M_5[(int64_t)b * R + r] : 0.0);
    }
}

// c[b, j] = w_v[b, j] + sum_r Gamma_v[b, r, j] w_x[b, r] + sum_r L_v[b, r, j] w_y[b, r]
// (column sums: thread <-> column j keeps the loads of each row coalesced)
__global__ void __launch_bounds__(256) linear_cost_kernel(int B, int nvt, int nxt, int nyt,
                                                          const double* __restrict__ w_v, int64_t w_v_stride,
                                                          const double* __restrict__ w_x,
                                                          const double* __restrict__ Gamma_v,
                                                          const double* __restrict__ xc,
                                                          const double* __restrict__ w_y,
                                                          const double* __restrict__ L_v,
                                                          const double* __restrict__ yc, double* __restrict__ c,
                                                          double* __restrict__ c0) {
    const int b = blockIdx.x;
    for (int j = threadIdx.x; j < nvt; j += blockDim.x) {
        double acc = w_v ? w_v[(int64_t)b * w_v_stride + j] : 0.0;
        if (w_x)
            for (int r = 0; r < nxt; ++r) acc += Gamma_v[((int64_t)b * nxt + r) * nvt + j] * w_x[(int64_t)b * nxt + r];
        if (w_y)
            for (int r = 0; r < nyt; ++r) acc += L_v[((int64_t)b * nyt + r) * nvt + j] * w_y[(int64_t)b * nyt + r];
        c[(int64_t)b * nvt + j] = acc;
    }
    if (c0 && threadIdx.x < 32) {
        double acc = 0.0;
        if (w_x) for (int r = threadIdx.x; r < nxt; r += 32) acc += w_x[(int64_t)b * nxt + r] * xc[(int64_t)b * nxt + r];
        if (w_y) for (int r = threadIdx.x; r < nyt; r += 32) acc += w_y[(int64_t)b * nyt + r] * yc[(int64_t)b * nyt + r];
        acc = warp_sum(acc);
        if (threadIdx.x == 0) c0[b] = acc;
    }
}

}  // namespace hmpc

extern "C" int hmpc_constraint_rhs_f64(const hmpc_dims* dims, int32_t rows, const double* H_x,
                                       const double* H_omega, const double* H_5, const double* x0,
                                       const double* w, int32_t S, double* rhs, void* stream) {
    using namespace hmpc;
    if (!dims || !H_5 || !rhs) return HMPC_ERR_ARG;
    const hmpc_dims d = *dims;
    const int rows_full = d.nc * d.Nt, nwt = d.nomega * d.Nt;
    if (rows < 0 || rows > rows_full || S < 0) return HMPC_ERR_ARG;
    if ((d.nx > 0 && (!H_x || !x0)) || (nwt > 0 && (!H_omega || !w))) return HMPC_ERR_ARG;
    if (d.B == 0 || rows == 0) return HMPC_OK;
    const size_t smem = sizeof(double) * (size_t)nwt * (S > 0 ? S : 1);
    if (smem > 200 * 1024) return HMPC_ERR_ARG;
    HMPC_CUDA_TRY(cudaFuncSetAttribute(constraint_rhs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gy = 1;
    while (d.B * gy < 8 * kNumSM && gy * 8 < rows) gy *= 2;   // small batches: about one row per warp
    constraint_rhs_kernel<<<dim3(d.B, gy), 256, smem, (cudaStream_t)stream>>>(d.B, rows, rows_full, d.nx, nwt, H_x,
                                                                             H_omega, H_5, x0, w, S, rhs);
    HMPC_LAUNCH_CHECK("constraint_rhs_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_predict_f64(int32_t B, int32_t R, int32_t nx, int32_t nvt, int32_t nwt, const double* M_x,
                                const double* M_v, const double* M_w, const double* M_5, const double* x0,
                                const double* v, const double* w, double* out, void* stream) {
    using namespace hmpc;
    if (B < 0 || R < 0 || nx < 0 || nvt < 0 || nwt < 0 || !out) return HMPC_ERR_ARG;
    if ((nx > 0 && (!M_x || !x0)) || (v && !M_v) || (w && !M_w)) return HMPC_ERR_ARG;
    if (B == 0 || R == 0) return HMPC_OK;
    int gy = 1;
    while (B * gy < 2 * kNumSM && gy * 8 < R) gy *= 2;
    predict_kernel<<<dim3(B, gy), 256, 0, (cudaStream_t)stream>>>(B, R, nx, nvt, nwt, M_x, M_v, nwt ? M_w : nullptr,
                                                                  M_5, x0, v, nwt ? w : nullptr, out);
    HMPC_LAUNCH_CHECK("predict_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_linear_cost_f64(int32_t B, int32_t nvt, int32_t nxt, int32_t nyt, const double* w_v,
                                    int64_t w_v_stride_b, const double* w_x, const double* Gamma_v, const double* xc,
                                    const double* w_y, const double* L_v, const double* yc, double* c, double* c0,
                                    void* stream) {
    using namespace hmpc;
    if (B < 0 || nvt < 0 || !c) return HMPC_ERR_ARG;
    if ((w_x && (!Gamma_v || (c0 && !xc))) || (w_y && (!L_v || (c0 && !yc)))) return HMPC_ERR_ARG;
    if (B == 0) return HMPC_OK;
    linear_cost_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(B, nvt, nxt, nyt, w_v, w_v_stride_b, w_x, Gamma_v, xc, w_y,
                                                            L_v, yc, c, c0);
    HMPC_LAUNCH_CHECK("linear_cost_kernel");
    return HMPC_OK;
}
