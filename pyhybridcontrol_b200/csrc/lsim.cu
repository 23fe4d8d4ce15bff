// lsim.cu -- K5: batched one-step MLD simulation and the DEWH model closed forms.
//   hmpc_lsim_step_f64         <- MldModel.lsim_k          (reference: models/mld_model.py:647-699)
//   hmpc_dewh_sim_step_f64     <- DewhAgentMpc.sim_step_k  (examples/.../micro_grid_agents.py:389-408) with the
//                                 const_heat=False model   (examples/.../micro_grid_models.py:45-57)
//   hmpc_dewh_control_model_f64<- const_heat=True model    (micro_grid_models.py:37-44, 52-57)
// ~100 B per agent-step: HBM-bound streaming kernels, one thread per agent for the scalar DEWH forms and one
// thread per (agent, output row) for the generic MLD.
#include "common.cuh"

namespace hmpc {

struct LsimArgs {
    hmpc_dims d;
    const double* mats[HMPC_NUM_MATS];
    int64_t stride[HMPC_NUM_MATS];
    const double *x, *u, *delta, *z, *w;
    double cons_tol;
    double *x1, *y;
    uint8_t* cons;
};

__device__ __forceinline__ double row_dot(const double* M, int64_t stride, int b, int r, int n, const double* v) {
    if (!M || n == 0) return 0.0;
    const double* row = M + (int64_t)b * stride + (int64_t)r * n;
    const double* vb = v + (int64_t)b * n;
    double acc = 0.0;
    for (int c = 0; c < n; ++c) acc += row[c] * vb[c];
    return acc;
}

// pass 0: x1 and y ; pass 1: cons (needs y)
__global__ void __launch_bounds__(256) lsim_xy_kernel(const LsimArgs a) {
    const hmpc_dims d = a.d;
    const int per = d.nx + d.ny;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)d.B * per) return;
    const int b = (int)(t / per), r0 = (int)(t - (int64_t)b * per);
    if (r0 < d.nx) {
        const int r = r0;
        // same association as the reference: A x + B1 u + B2 delta + B3 z + B4 w + b5
        double v = row_dot(a.mats[HMPC_A], a.stride[HMPC_A], b, r, d.nx, a.x);
        v += row_dot(a.mats[HMPC_B1], a.stride[HMPC_B1], b, r, d.nu, a.u);
        v += row_dot(a.mats[HMPC_B2], a.stride[HMPC_B2], b, r, d.ndelta, a.delta);
        v += row_dot(a.mats[HMPC_B3], a.stride[HMPC_B3], b, r, d.nz, a.z);
        v += row_dot(a.mats[HMPC_B4], a.stride[HMPC_B4], b, r, d.nomega, a.w);
        v += a.mats[HMPC_b5] ? a.mats[HMPC_b5][(int64_t)b * a.stride[HMPC_b5] + r] : 0.0;
        a.x1[(int64_t)b * d.nx + r] = v;
    } else {
        const int r = r0 - d.nx;
        double v = row_dot(a.mats[HMPC_C], a.stride[HMPC_C], b, r, d.nx, a.x);
        v += row_dot(a.mats[HMPC_D1], a.stride[HMPC_D1], b, r, d.nu, a.u);
        v += row_dot(a.mats[HMPC_D2], a.stride[HMPC_D2], b, r, d.ndelta, a.delta);
        v += row_dot(a.mats[HMPC_D3], a.stride[HMPC_D3], b, r, d.nz, a.z);
        v += row_dot(a.mats[HMPC_D4], a.stride[HMPC_D4], b, r, d.nomega, a.w);
        v += a.mats[HMPC_d5] ? a.mats[HMPC_d5][(int64_t)b * a.stride[HMPC_d5] + r] : 0.0;
        a.y[(int64_t)b * d.ny + r] = v;
    }
}

__global__ void __launch_bounds__(256) lsim_cons_kernel(const LsimArgs a) {
    const hmpc_dims d = a.d;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)d.B * d.nc) return;
    const int b = (int)(t / d.nc), r = (int)(t - (int64_t)b * d.nc);
    // E x + F1 u + F2 delta + F3 z + F4 w + G y + Psi (mu*0) - f5 <= tol     (mu is ignored, :694)
    double v = row_dot(a.mats[HMPC_E], a.stride[HMPC_E], b, r, d.nx, a.x);
    v += row_dot(a.mats[HMPC_F1], a.stride[HMPC_F1], b, r, d.nu, a.u);
    v += row_dot(a.mats[HMPC_F2], a.stride[HMPC_F2], b, r, d.ndelta, a.delta);
    v += row_dot(a.mats[HMPC_F3], a.stride[HMPC_F3], b, r, d.nz, a.z);
    v += row_dot(a.mats[HMPC_F4], a.stride[HMPC_F4], b, r, d.nomega, a.w);
    v += row_dot(a.mats[HMPC_G], a.stride[HMPC_G], b, r, d.ny, a.y);
    v -= a.mats[HMPC_f5] ? a.mats[HMPC_f5][(int64_t)b * a.stride[HMPC_f5] + r] : 0.0;
    a.cons[(int64_t)b * d.nc + r] = (v <= a.cons_tol) ? 1 : 0;
}

// params [B,12] = {C_w, A_h, U_h, m_h, T_w, T_inf, P_h_Nom, T_h_min, T_h_max, T_h_Nom, ts, reserved}
__device__ __forceinline__ void dewh_model(const double* p, bool const_heat, double T_h, double D_h, double& A,
                                           double& B1, double& B4, double& b5) {
    const double C_w = p[0], A_h = p[1], U_h = p[2], m_h = p[3], T_w = p[4], T_inf = p[5], P = p[6], T_nom = p[9],
                 ts = p[10];
    const double p1 = U_h * A_h, p2 = m_h * C_w;
    double a_c, b4_c;
    if (const_heat) {
        a_c = -p1 / p2;
        b4_c = C_w * (T_w - T_nom) / p2;
    } else {
        const double r = (T_nom - T_w) / (T_h - T_w);
        a_c = -((D_h * C_w * r) + p1) / p2;
        b4_c = C_w * T_w * r / p2;
    }
    A = exp(a_c * ts);
    const double em = (A - 1.0) / a_c;
    B1 = em * P / p2;
    B4 = em * b4_c;
    b5 = em * p1 * T_inf / p2;
}

__global__ void __launch_bounds__(256) dewh_sim_step_kernel(int B, const double* __restrict__ params,
                                                            const double* __restrict__ T, const double* __restrict__ u,
                                                            const double* __restrict__ D_h, double* __restrict__ T1,
                                                            double* __restrict__ model, uint8_t* __restrict__ cons) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* p = params + (int64_t)b * 12;
    double x = T[b];
    if (x <= p[4]) x = p[4] + 0.1;            // clamp T_h <= T_w (micro_grid_agents.py:398-399)
    double A, B1, B4, b5;
    dewh_model(p, false, x, D_h[b], A, B1, B4, b5);
    T1[b] = A * x + B1 * u[b] + B4 * D_h[b] + b5;
    if (model) { model[4 * b + 0] = A; model[4 * b + 1] = B1; model[4 * b + 2] = B4; model[4 * b + 3] = b5; }
    if (cons) { cons[2 * b + 0] = (x - p[8] <= 1e-6); cons[2 * b + 1] = (-x + p[7] <= 1e-6); }
}

__global__ void __launch_bounds__(256) dewh_control_model_kernel(int B, const double* __restrict__ params,
                                                                 double* __restrict__ model) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double A, B1, B4, b5;
    dewh_model(params + (int64_t)b * 12, true, 0.0, 0.0, A, B1, B4, b5);
    model[4 * b + 0] = A; model[4 * b + 1] = B1; model[4 * b + 2] = B4; model[4 * b + 3] = b5;
}

// Hysteresis rule of the example's non-predictive controller (theromstat_control.py:50-62): on at or below
// T_max - band_on, off at or above T_max - band_off, in between the previous input is kept (exactly 1 keeps "on").
__global__ void __launch_bounds__(256) dewh_thermostat_kernel(int B, const double* __restrict__ params,
                                                              const double* __restrict__ band, int64_t band_stride_b,
                                                              const double* __restrict__ T,
                                                              const double* __restrict__ u_prev, double* __restrict__ u) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double T_max = params[(int64_t)b * 12 + 8];
    const double on = band[b * band_stride_b + 0], off = band[b * band_stride_b + 1];
    const double x = T[b];
    double out;
    if (x <= T_max - on) out = 1.0;
    else if (x >= T_max - off) out = 0.0;
    else out = (u_prev[b] == 1.0) ? 1.0 : 0.0;
    u[b] = out;
}

}  // namespace hmpc

extern "C" int hmpc_lsim_step_f64(const hmpc_dims* dims, const double* const mats[HMPC_NUM_MATS],
                                  const int64_t mat_stride_b[HMPC_NUM_MATS], const double* x, const double* u,
                                  const double* delta, const double* z, const double* w, double cons_tol, double* x1,
                                  double* y, uint8_t* cons, void* stream) {
    using namespace hmpc;
    if (!dims || !mats || !mat_stride_b) return HMPC_ERR_ARG;
    const hmpc_dims d = *dims;
    if ((d.nx && (!x || !x1)) || (d.nu && !u) || (d.ndelta && !delta) || (d.nz && !z) || (d.nomega && !w) ||
        (d.ny && !y) || (d.nc && !cons))
        return HMPC_ERR_ARG;
    if (d.B == 0) return HMPC_OK;
    LsimArgs a;
    a.d = d;
    for (int i = 0; i < HMPC_NUM_MATS; ++i) { a.mats[i] = mats[i]; a.stride[i] = mat_stride_b[i]; }
    a.x = x; a.u = u; a.delta = delta; a.z = z; a.w = w; a.cons_tol = cons_tol; a.x1 = x1; a.y = y; a.cons = cons;
    const int64_t n1 = (int64_t)d.B * (d.nx + d.ny), n2 = (int64_t)d.B * d.nc;
    if (n1) {
        lsim_xy_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
        HMPC_LAUNCH_CHECK("lsim_xy_kernel");
    }
    if (n2) {
        lsim_cons_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
        HMPC_LAUNCH_CHECK("lsim_cons_kernel");
    }
    return HMPC_OK;
}

extern "C" int hmpc_dewh_sim_step_f64(int32_t B, const double* params, const double* T, const double* u,
                                      const double* D_h, double* T1, double* model, uint8_t* cons, void* stream) {
    using namespace hmpc;
    if (B < 0 || !params || !T || !u || !D_h || !T1) return HMPC_ERR_ARG;
    if (B == 0) return HMPC_OK;
    dewh_sim_step_kernel<<<ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(B, params, T, u, D_h, T1, model, cons);
    HMPC_LAUNCH_CHECK("dewh_sim_step_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_dewh_control_model_f64(int32_t B, const double* params, double* model, void* stream) {
    using namespace hmpc;
    if (B < 0 || !params || !model) return HMPC_ERR_ARG;
    if (B == 0) return HMPC_OK;
    dewh_control_model_kernel<<<ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(B, params, model);
    HMPC_LAUNCH_CHECK("dewh_control_model_kernel");
    return HMPC_OK;
}

extern "C" int hmpc_dewh_thermostat_f64(int32_t B, const double* params, const double* band, int64_t band_stride_b,
                                        const double* T, const double* u_prev, double* u, void* stream) {
    using namespace hmpc;
    if (B < 0 || !params || !band || !T || !u_prev || !u || (band_stride_b != 0 && band_stride_b != 2))
        return HMPC_ERR_ARG;
    if (B == 0) return HMPC_OK;
    dewh_thermostat_kernel<<<ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(B, params, band, band_stride_b, T,
                                                                                u_prev, u);
    HMPC_LAUNCH_CHECK("dewh_thermostat_kernel");
    return HMPC_OK;
}
