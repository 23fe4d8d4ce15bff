// host_api.cu -- host-buffer front door of the C ABI: one whole MPC control step for a batch of agents.
// This is what a foreign-language binding (ctypes / cgo / JNI) of the reference's
// MpcController.build() + solve() would call (controllers/mpc_controller.py:76-101,
// controllers/controller_base.py:491-540); see INTEGRATION.md.
#include <new>
#include <string.h>
#include "common.cuh"

struct hmpc_step_plan {
    hmpc_dims d;
    hmpc_milp_opts opts;
    int nv, nvt, nwt, mrows;
    cudaStream_t stream;
    cudaEvent_t ev[4];
    // device
    double* mats_dev;            // packed copies of the 20 system matrices
    int64_t mat_off[HMPC_NUM_MATS];
    int64_t mat_elems[HMPC_NUM_MATS];
    double *H_x, *H_v, *H_w, *H_5, *x0, *w, *rhs, *cost, *lb, *ub, *v, *obj;
    int32_t *status, *stats;
    uint8_t* is_bin;
    // pinned staging
    double* pin_in;  size_t pin_in_elems;
    double* pin_out; size_t pin_out_elems;
    int32_t* pin_iout;
    uint8_t* pin_bin;
    bool condensed;
    // exact stage-DP path (stage_dp.cu) for scalar-state MLDs; the branch-and-cut kernel is the general path
    bool dp_dims_ok;
    hmpc_stage_dp_opts dp_opts;
    void* dp_ws; size_t dp_ws_bytes;
    const double* dev_mats[HMPC_NUM_MATS];
    int64_t dev_stride[HMPC_NUM_MATS];
    int32_t last_solver;   // 0 = branch and cut, 1 = stage DP
};

namespace {
int mat_rows(const hmpc_dims& d, int i) { return i < HMPC_C ? d.nx : (i < HMPC_E ? d.ny : d.nc); }
int mat_cols(const hmpc_dims& d, int i) {
    switch (i) {
        case HMPC_A: case HMPC_C: case HMPC_E: return d.nx;
        case HMPC_B1: case HMPC_D1: case HMPC_F1: return d.nu;
        case HMPC_B2: case HMPC_D2: case HMPC_F2: return d.ndelta;
        case HMPC_B3: case HMPC_D3: case HMPC_F3: return d.nz;
        case HMPC_B4: case HMPC_D4: case HMPC_F4: return d.nomega;
        case HMPC_G: return d.ny;
        case HMPC_Psi: return d.nmu;
        default: return 1;  // b5, d5, f5
    }
}
}  // namespace

extern "C" int hmpc_step_plan_create(const hmpc_dims* dims, const hmpc_milp_opts* opts, hmpc_step_plan** out) {
    using namespace hmpc;
    if (!dims || !out || dims->B <= 0 || dims->Nt <= 0) return HMPC_ERR_ARG;
    hmpc_step_plan* p = new (std::nothrow) hmpc_step_plan();
    if (!p) return HMPC_ERR_ARG;
    memset(p, 0, sizeof(*p));
    p->d = *dims;
    if (opts) p->opts = *opts; else hmpc_milp_default_opts(&p->opts);
    const hmpc_dims d = *dims;
    p->nv = d.nu + d.ndelta + d.nz + d.nmu;
    p->nvt = p->nv * d.Nt; p->nwt = d.nomega * d.Nt; p->mrows = d.nc * d.Nt;
    const int64_t B = d.B;
    HMPC_CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) HMPC_CUDA_TRY(cudaEventCreate(&p->ev[i]));
    int64_t tot = 0;
    for (int i = 0; i < HMPC_NUM_MATS; ++i) {
        p->mat_elems[i] = (int64_t)mat_rows(d, i) * mat_cols(d, i);
        p->mat_off[i] = tot;
        tot += p->mat_elems[i] * B;
    }
    auto dalloc = [&](double** ptr, int64_t n) { return cudaMalloc((void**)ptr, sizeof(double) * (size_t)(n > 0 ? n : 1)); };
    HMPC_CUDA_TRY(dalloc(&p->mats_dev, tot));
    HMPC_CUDA_TRY(dalloc(&p->H_x, B * p->mrows * d.nx));
    HMPC_CUDA_TRY(dalloc(&p->H_v, B * p->mrows * p->nvt));
    HMPC_CUDA_TRY(dalloc(&p->H_w, B * p->mrows * p->nwt));
    HMPC_CUDA_TRY(dalloc(&p->H_5, B * p->mrows));
    HMPC_CUDA_TRY(dalloc(&p->x0, B * d.nx));
    HMPC_CUDA_TRY(dalloc(&p->w, B * p->nwt));
    HMPC_CUDA_TRY(dalloc(&p->rhs, B * p->mrows));
    HMPC_CUDA_TRY(dalloc(&p->cost, B * p->nvt));
    HMPC_CUDA_TRY(dalloc(&p->lb, p->nvt));
    HMPC_CUDA_TRY(dalloc(&p->ub, p->nvt));
    HMPC_CUDA_TRY(dalloc(&p->v, B * p->nvt));
    HMPC_CUDA_TRY(dalloc(&p->obj, B));
    HMPC_CUDA_TRY(cudaMalloc((void**)&p->status, sizeof(int32_t) * B));
    HMPC_CUDA_TRY(cudaMalloc((void**)&p->stats, sizeof(int32_t) * B * 8));
    HMPC_CUDA_TRY(cudaMalloc((void**)&p->is_bin, (size_t)(p->nvt > 0 ? p->nvt : 1)));
    p->pin_in_elems = (size_t)(tot + B * d.nx + B * p->nwt + B * p->nvt + 2 * p->nvt + 8);
    p->pin_out_elems = (size_t)(B * p->nvt + B + 8);
    HMPC_CUDA_TRY(cudaMallocHost((void**)&p->pin_in, sizeof(double) * p->pin_in_elems));
    HMPC_CUDA_TRY(cudaMallocHost((void**)&p->pin_out, sizeof(double) * p->pin_out_elems));
    HMPC_CUDA_TRY(cudaMallocHost((void**)&p->pin_iout, sizeof(int32_t) * (size_t)B * 9));
    HMPC_CUDA_TRY(cudaMallocHost((void**)&p->pin_bin, (size_t)(p->nvt > 0 ? p->nvt : 1)));
    hmpc_stage_dp_default_opts(&p->dp_opts);
    p->dp_dims_ok = hmpc_stage_dp_supported(dims) != 0;
    if (p->dp_dims_ok) {
        if (hmpc_stage_dp_workspace_bytes(dims, &p->dp_opts, &p->dp_ws_bytes) != HMPC_OK) return HMPC_ERR_ARG;
        HMPC_CUDA_TRY(cudaMalloc(&p->dp_ws, p->dp_ws_bytes));
    }
    *out = p;
    return HMPC_OK;
}

extern "C" int hmpc_step_plan_destroy(hmpc_step_plan* p) {
    if (!p) return HMPC_OK;
    cudaStreamSynchronize(p->stream);
    double* dbl[] = {p->mats_dev, p->H_x, p->H_v, p->H_w, p->H_5, p->x0, p->w, p->rhs, p->cost, p->lb, p->ub, p->v, p->obj};
    for (double* q : dbl) cudaFree(q);
    cudaFree(p->status); cudaFree(p->stats); cudaFree(p->is_bin); cudaFree(p->dp_ws);
    cudaFreeHost(p->pin_in); cudaFreeHost(p->pin_out); cudaFreeHost(p->pin_iout); cudaFreeHost(p->pin_bin);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(p->ev[i]);
    cudaStreamDestroy(p->stream);
    delete p;
    return HMPC_OK;
}

extern "C" int hmpc_step_plan_last_solver(const hmpc_step_plan* p) { return p ? p->last_solver : -1; }

extern "C" int hmpc_mpc_step_host_bytes(const hmpc_step_plan* p, int32_t recondense, int64_t* h2d, int64_t* d2h) {
    if (!p) return HMPC_ERR_ARG;
    const int64_t B = p->d.B;
    int64_t in = 8 * (B * p->d.nx + B * p->nwt + B * p->nvt + 2 * (int64_t)p->nvt) + p->nvt;
    if (recondense) for (int i = 0; i < HMPC_NUM_MATS; ++i) in += 8 * p->mat_elems[i] * B;
    if (h2d) *h2d = in;
    if (d2h) *d2h = 8 * (B * p->nvt + B) + 4 * (B + 8 * B);
    return HMPC_OK;
}

extern "C" int hmpc_mpc_step_host_f64(hmpc_step_plan* p, int32_t recondense, const double* const mats[HMPC_NUM_MATS],
                                      const int64_t mat_stride_b[HMPC_NUM_MATS], const double* x0, const double* w,
                                      const double* cost_v, int64_t cost_v_stride_b, const double* lb_v,
                                      const double* ub_v, const uint8_t* is_bin_v, double* v, double* obj,
                                      int32_t* status, int32_t* stats, float* timing_ms) {
    using namespace hmpc;
    if (!p || !cost_v || !lb_v || !ub_v || !is_bin_v || !v || !obj || !status || !stats) return HMPC_ERR_ARG;
    const hmpc_dims d = p->d;
    const int64_t B = d.B;
    if ((d.nx && !x0) || (p->nwt && !w)) return HMPC_ERR_ARG;
    if (!recondense && !p->condensed) return HMPC_ERR_ARG;
    if (recondense && (!mats || !mat_stride_b)) return HMPC_ERR_ARG;
    cudaStream_t s = p->stream;
    // ---- stage inputs in pinned memory (so the copies are true async DMA), then host -> device
    double* pin = p->pin_in;
    size_t off = 0;
    HMPC_CUDA_TRY(cudaEventRecord(p->ev[0], s));
    const double** dev_mats = p->dev_mats;
    int64_t* dev_stride = p->dev_stride;
    if (recondense) {
        for (int i = 0; i < HMPC_NUM_MATS; ++i) {
            const int64_t e = p->mat_elems[i];
            if (!mats[i] || e == 0) { dev_mats[i] = nullptr; dev_stride[i] = 0; continue; }
            const int64_t nb = mat_stride_b[i] == 0 ? 1 : B;
            if (mat_stride_b[i] == e || nb == 1) memcpy(pin + off, mats[i], sizeof(double) * (size_t)(e * nb));
            else for (int64_t b = 0; b < B; ++b) memcpy(pin + off + b * e, mats[i] + b * mat_stride_b[i], sizeof(double) * (size_t)e);
            HMPC_CUDA_TRY(cudaMemcpyAsync(p->mats_dev + p->mat_off[i], pin + off, sizeof(double) * (size_t)(e * nb), cudaMemcpyHostToDevice, s));
            dev_mats[i] = p->mats_dev + p->mat_off[i];
            dev_stride[i] = nb == 1 ? 0 : e;
            off += (size_t)(e * nb);
        }
    }
    auto h2d = [&](double* dst, const double* src, int64_t n) -> cudaError_t {
        if (n <= 0) return cudaSuccess;
        memcpy(pin + off, src, sizeof(double) * (size_t)n);
        cudaError_t e = cudaMemcpyAsync(dst, pin + off, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s);
        off += (size_t)n;
        return e;
    };
    HMPC_CUDA_TRY(h2d(p->x0, x0, B * d.nx));
    HMPC_CUDA_TRY(h2d(p->w, w, B * p->nwt));
    const bool bc = cost_v_stride_b == 0;
    if (bc || cost_v_stride_b == p->nvt) HMPC_CUDA_TRY(h2d(p->cost, cost_v, (bc ? 1 : B) * (int64_t)p->nvt));
    else for (int64_t b = 0; b < B; ++b) {
        memcpy(pin + off, cost_v + b * cost_v_stride_b, sizeof(double) * (size_t)p->nvt);
        HMPC_CUDA_TRY(cudaMemcpyAsync(p->cost + b * p->nvt, pin + off, sizeof(double) * (size_t)p->nvt, cudaMemcpyHostToDevice, s));
        off += (size_t)p->nvt;
    }
    HMPC_CUDA_TRY(h2d(p->lb, lb_v, p->nvt));
    HMPC_CUDA_TRY(h2d(p->ub, ub_v, p->nvt));
    memcpy(p->pin_bin, is_bin_v, (size_t)p->nvt);
    HMPC_CUDA_TRY(cudaMemcpyAsync(p->is_bin, p->pin_bin, (size_t)p->nvt, cudaMemcpyHostToDevice, s));
    HMPC_CUDA_TRY(cudaEventRecord(p->ev[1], s));
    // ---- kernels
    int rc;
    if (recondense) {
        double* outs[HMPC_NUM_EVO] = {nullptr};
        outs[HMPC_H_X] = d.nx ? p->H_x : nullptr; outs[HMPC_H_V] = p->H_v; outs[HMPC_H_OMEGA] = p->nwt ? p->H_w : nullptr;
        outs[HMPC_H_5] = p->H_5;
        rc = hmpc_condense_f64(&d, dev_mats, dev_stride, outs, s);
        if (rc != HMPC_OK) return rc;
        p->condensed = true;
    }
    rc = hmpc_constraint_rhs_f64(&d, p->mrows, p->H_x, p->H_w, p->H_5, p->x0, p->w, 0, p->rhs, s);
    if (rc != HMPC_OK) return rc;
    auto solve_bnc = [&]() -> int {
        p->last_solver = 0;
        return hmpc_milp_solve_f64(d.B, p->nvt, p->mrows, p->cost, bc ? 0 : p->nvt, p->H_v, (int64_t)p->mrows * p->nvt,
                                   p->rhs, p->lb, p->ub, 0, p->is_bin, &p->opts, nullptr, 0, p->v, p->obj, p->status,
                                   p->stats, s);
    };
    auto fetch = [&]() -> int {
        HMPC_CUDA_TRY(cudaEventRecord(p->ev[2], s));
        HMPC_CUDA_TRY(cudaMemcpyAsync(p->pin_out, p->v, sizeof(double) * (size_t)(B * p->nvt), cudaMemcpyDeviceToHost, s));
        HMPC_CUDA_TRY(cudaMemcpyAsync(p->pin_out + B * p->nvt, p->obj, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, s));
        HMPC_CUDA_TRY(cudaMemcpyAsync(p->pin_iout, p->status, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, s));
        HMPC_CUDA_TRY(cudaMemcpyAsync(p->pin_iout + B, p->stats, sizeof(int32_t) * (size_t)B * 8, cudaMemcpyDeviceToHost, s));
        HMPC_CUDA_TRY(cudaEventRecord(p->ev[3], s));
        HMPC_CUDA_TRY(cudaStreamSynchronize(s));
        return HMPC_OK;
    };
    if (p->dp_dims_ok && p->opts.reserved == 0) {
        // scalar-state class: exact stage-DP kernels straight from the MLD blocks (H_v is not read)
        p->last_solver = 1;
        rc = hmpc_stage_dp_solve_f64(&d, dev_mats, dev_stride, p->rhs, p->cost, bc ? 0 : p->nvt, p->lb, p->ub, p->is_bin,
                                     &p->dp_opts, p->dp_ws, p->dp_ws_bytes, p->v, p->obj, p->status, p->stats, s);
    } else {
        rc = solve_bnc();
    }
    if (rc != HMPC_OK) return rc;
    // ---- device -> host
    rc = fetch();
    if (rc != HMPC_OK) return rc;
    if (p->last_solver == 1) {
        bool unsupported = false;
        for (int64_t b = 0; b < B; ++b) if (p->pin_iout[b] == HMPC_SOLVE_UNSUPPORTED) unsupported = true;
        if (unsupported) {   // some agent's matrices are outside the class: the general kernel takes the batch
            rc = solve_bnc();
            if (rc != HMPC_OK) return rc;
            rc = fetch();
            if (rc != HMPC_OK) return rc;
        }
    }
    memcpy(v, p->pin_out, sizeof(double) * (size_t)(B * p->nvt));
    memcpy(obj, p->pin_out + B * p->nvt, sizeof(double) * (size_t)B);
    memcpy(status, p->pin_iout, sizeof(int32_t) * (size_t)B);
    memcpy(stats, p->pin_iout + B, sizeof(int32_t) * (size_t)B * 8);
    if (timing_ms) {
        cudaEventElapsedTime(&timing_ms[0], p->ev[0], p->ev[1]);
        cudaEventElapsedTime(&timing_ms[1], p->ev[1], p->ev[2]);
        cudaEventElapsedTime(&timing_ms[2], p->ev[2], p->ev[3]);
        cudaEventElapsedTime(&timing_ms[3], p->ev[0], p->ev[3]);
    }
    return HMPC_OK;
}
