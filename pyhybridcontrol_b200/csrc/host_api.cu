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
    // device: ONE input arena [x0 | w | cost | lb | ub | is_bin | mats...] and ONE output arena
    // [v | obj | status | stats], mirrored by pinned host arenas, so that a step is one H2D and one D2H copy
    double* in_dev;  double* pin_in;  size_t in_elems, in_head_elems;   // head = everything but the matrices
    double* out_dev; double* pin_out; size_t out_elems;
    double* mats_dev;            // = in_dev + in_head_elems: packed copies of the 20 system matrices
    int64_t mat_off[HMPC_NUM_MATS];
    int64_t mat_elems[HMPC_NUM_MATS];
    double *H_x, *H_v, *H_w, *H_5, *x0, *w, *rhs, *cost, *lb, *ub, *v, *obj;
    int32_t *status, *stats;
    uint8_t* is_bin;
    bool condensed, have_H_v;
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
    const int64_t bin_elems = (p->nvt + 7) / 8;   // is_bin bytes, rounded up to whole doubles
    p->in_head_elems = (size_t)(B * d.nx + B * p->nwt + B * p->nvt + 2 * p->nvt + bin_elems);
    p->in_elems = p->in_head_elems + (size_t)tot;
    p->out_elems = (size_t)(B * p->nvt + B + (B * 9 + 1) / 2);
    HMPC_CUDA_TRY(dalloc(&p->in_dev, (int64_t)p->in_elems));
    HMPC_CUDA_TRY(dalloc(&p->out_dev, (int64_t)p->out_elems));
    HMPC_CUDA_TRY(cudaMallocHost((void**)&p->pin_in, sizeof(double) * p->in_elems));
    HMPC_CUDA_TRY(cudaMallocHost((void**)&p->pin_out, sizeof(double) * p->out_elems));
    {
        double* q = p->in_dev;
        p->x0 = q; q += B * d.nx; p->w = q; q += B * p->nwt; p->cost = q; q += B * p->nvt;
        p->lb = q; q += p->nvt; p->ub = q; q += p->nvt; p->is_bin = reinterpret_cast<uint8_t*>(q); q += bin_elems;
        p->mats_dev = q;
        double* o = p->out_dev;
        p->v = o; o += B * p->nvt; p->obj = o; o += B;
        p->status = reinterpret_cast<int32_t*>(o); p->stats = p->status + B;
    }
    HMPC_CUDA_TRY(dalloc(&p->H_x, B * p->mrows * d.nx));
    HMPC_CUDA_TRY(dalloc(&p->H_v, B * p->mrows * p->nvt));
    HMPC_CUDA_TRY(dalloc(&p->H_w, B * p->mrows * p->nwt));
    HMPC_CUDA_TRY(dalloc(&p->H_5, B * p->mrows));
    HMPC_CUDA_TRY(dalloc(&p->rhs, B * p->mrows));
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
    double* dbl[] = {p->in_dev, p->out_dev, p->H_x, p->H_v, p->H_w, p->H_5, p->rhs};
    for (double* q : dbl) cudaFree(q);
    cudaFree(p->dp_ws);
    cudaFreeHost(p->pin_in); cudaFreeHost(p->pin_out);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(p->ev[i]);
    cudaStreamDestroy(p->stream);
    delete p;
    return HMPC_OK;
}

extern "C" int hmpc_step_plan_last_solver(const hmpc_step_plan* p) { return p ? p->last_solver : -1; }

extern "C" int hmpc_mpc_step_host_bytes(const hmpc_step_plan* p, int32_t recondense, int64_t* h2d, int64_t* d2h) {
    if (!p) return HMPC_ERR_ARG;
    const int64_t B = p->d.B;
    (void)B;
    if (h2d) *h2d = 8 * (int64_t)(recondense ? p->in_elems : p->in_head_elems);
    if (d2h) *d2h = 8 * (int64_t)p->out_elems;
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
    // ---- stage every input in the pinned arena (same layout as the device arena), then ONE host -> device copy
    HMPC_CUDA_TRY(cudaEventRecord(p->ev[0], s));
    const double** dev_mats = p->dev_mats;
    int64_t* dev_stride = p->dev_stride;
    double* pin = p->pin_in;
    const bool bc = cost_v_stride_b == 0;
    {
        double* q = pin;
        if (d.nx) memcpy(q, x0, sizeof(double) * (size_t)(B * d.nx));
        q += B * d.nx;
        if (p->nwt) memcpy(q, w, sizeof(double) * (size_t)(B * p->nwt));
        q += B * p->nwt;
        if (bc || cost_v_stride_b == p->nvt) memcpy(q, cost_v, sizeof(double) * (size_t)((bc ? 1 : B) * (int64_t)p->nvt));
        else for (int64_t b = 0; b < B; ++b) memcpy(q + b * p->nvt, cost_v + b * cost_v_stride_b, sizeof(double) * (size_t)p->nvt);
        q += B * p->nvt;
        memcpy(q, lb_v, sizeof(double) * (size_t)p->nvt); q += p->nvt;
        memcpy(q, ub_v, sizeof(double) * (size_t)p->nvt); q += p->nvt;
        memcpy(q, is_bin_v, (size_t)p->nvt);
    }
    size_t copy_elems = p->in_head_elems;
    if (recondense) {
        double* pm = pin + p->in_head_elems;
        for (int i = 0; i < HMPC_NUM_MATS; ++i) {
            const int64_t e = p->mat_elems[i];
            if (!mats[i] || e == 0) { dev_mats[i] = nullptr; dev_stride[i] = 0; continue; }
            const int64_t nb = mat_stride_b[i] == 0 ? 1 : B;
            double* dst = pm + p->mat_off[i];
            if (mat_stride_b[i] == e || nb == 1) memcpy(dst, mats[i], sizeof(double) * (size_t)(e * nb));
            else for (int64_t b = 0; b < B; ++b) memcpy(dst + b * e, mats[i] + b * mat_stride_b[i], sizeof(double) * (size_t)e);
            dev_mats[i] = p->mats_dev + p->mat_off[i];
            dev_stride[i] = nb == 1 ? 0 : e;
        }
        copy_elems = p->in_elems;
    }
    HMPC_CUDA_TRY(cudaMemcpyAsync(p->in_dev, pin, sizeof(double) * copy_elems, cudaMemcpyHostToDevice, s));
    HMPC_CUDA_TRY(cudaEventRecord(p->ev[1], s));
    // ---- kernels
    const bool want_dp = p->dp_dims_ok && p->opts.force_general == 0;
    int rc;
    auto condense = [&](bool with_H_v) -> int {
        double* outs[HMPC_NUM_EVO] = {nullptr};
        outs[HMPC_H_X] = d.nx ? p->H_x : nullptr; outs[HMPC_H_V] = with_H_v ? p->H_v : nullptr;
        outs[HMPC_H_OMEGA] = p->nwt ? p->H_w : nullptr; outs[HMPC_H_5] = p->H_5;
        const int r = hmpc_condense_f64(&d, dev_mats, dev_stride, outs, s);
        if (r == HMPC_OK) { p->condensed = true; p->have_H_v = with_H_v; }
        return r;
    };
    if (recondense) {
        // the stage-DP kernels read the MLD blocks directly; the dense H_v (3/4 of K1's bytes) is only materialised
        // for the branch-and-cut kernel
        rc = condense(!want_dp);
        if (rc != HMPC_OK) return rc;
    }
    rc = hmpc_constraint_rhs_f64(&d, p->mrows, p->H_x, p->H_w, p->H_5, p->x0, p->w, 0, p->rhs, s);
    if (rc != HMPC_OK) return rc;
    auto solve_bnc = [&]() -> int {
        p->last_solver = 0;
        if (!p->have_H_v) { const int r = condense(true); if (r != HMPC_OK) return r; }
        return hmpc_milp_solve_f64(d.B, p->nvt, p->mrows, p->cost, bc ? 0 : p->nvt, p->H_v, (int64_t)p->mrows * p->nvt,
                                   p->rhs, p->lb, p->ub, 0, p->is_bin, &p->opts, nullptr, 0, p->v, p->obj, p->status,
                                   p->stats, s);
    };
    const int32_t* pin_status = reinterpret_cast<const int32_t*>(p->pin_out + B * p->nvt + B);
    auto fetch = [&]() -> int {
        HMPC_CUDA_TRY(cudaEventRecord(p->ev[2], s));
        HMPC_CUDA_TRY(cudaMemcpyAsync(p->pin_out, p->out_dev, sizeof(double) * p->out_elems, cudaMemcpyDeviceToHost, s));
        HMPC_CUDA_TRY(cudaEventRecord(p->ev[3], s));
        HMPC_CUDA_TRY(cudaStreamSynchronize(s));
        return HMPC_OK;
    };
    if (want_dp) {
        // scalar-state class: exact stage-DP kernels straight from the MLD blocks
        p->last_solver = 1;
        rc = hmpc_stage_dp_solve_f64(&d, dev_mats, dev_stride, p->rhs, p->cost, bc ? 0 : p->nvt, p->lb, p->ub, p->is_bin,
                                     nullptr, &p->dp_opts, p->dp_ws, p->dp_ws_bytes, p->v, p->obj, p->status, p->stats, s);
    } else {
        rc = solve_bnc();
    }
    if (rc != HMPC_OK) return rc;
    // ---- ONE device -> host copy
    rc = fetch();
    if (rc != HMPC_OK) return rc;
    if (p->last_solver == 1) {
        bool unsupported = false;
        for (int64_t b = 0; b < B; ++b) if (pin_status[b] == HMPC_SOLVE_UNSUPPORTED) unsupported = true;
        if (unsupported) {   // some agent's matrices are outside the class: the general kernel takes the batch
            rc = solve_bnc();
            if (rc != HMPC_OK) return rc;
            rc = fetch();
            if (rc != HMPC_OK) return rc;
        }
    }
    memcpy(v, p->pin_out, sizeof(double) * (size_t)(B * p->nvt));
    memcpy(obj, p->pin_out + B * p->nvt, sizeof(double) * (size_t)B);
    memcpy(status, pin_status, sizeof(int32_t) * (size_t)B);
    memcpy(stats, pin_status + B, sizeof(int32_t) * (size_t)B * 8);
    if (timing_ms) {
        cudaEventElapsedTime(&timing_ms[0], p->ev[0], p->ev[1]);
        cudaEventElapsedTime(&timing_ms[1], p->ev[1], p->ev[2]);
        cudaEventElapsedTime(&timing_ms[2], p->ev[2], p->ev[3]);
        cudaEventElapsedTime(&timing_ms[3], p->ev[0], p->ev[3]);
    }
    return HMPC_OK;
}
