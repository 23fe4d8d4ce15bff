// param_eval.cu -- batched parameter -> system-matrix evaluation: the symbolic / callable model front-end.
//   hmpc_param_eval_f64  <- CallableMatrix.__call__(param_struct=...) for every non-constant matrix of an MldModel
//                           (reference: utils/matrix_utils.py:339-343 sympy.lambdify of the matrix, :441-470 the
//                           call; MldModel.to_numeric models/mld_model.py:791-793; MldSystemModel.get_mld_numeric
//                           :1128-1149), for B parameter sets in one launch.
// The reference turns each symbolic matrix into a Python function (lambdify) and calls it per agent per step.  Here
// the host compiles the expressions of ALL matrices of a model once into one straight-line register program
// (pyhybridcontrol_b200/utils/matrix_utils.py) and this kernel interprets it, one thread per agent:
//   * a CTA owns a tile of `blockDim.x` consecutive agents.  Their parameter rows are one contiguous piece of
//     params[B, P]; it is loaded with coalesced reads and transposed into shared memory (odd pitch, conflict-free).
//   * the virtual-machine registers live in shared memory as regs[r][thread] (a column per thread: no bank
//     conflicts, no local-memory spills); every thread executes the same instruction, read as one broadcast.
//   * OUT instructions fill an output tile in shared memory; at the end each matrix' piece of the tile --
//     contiguous in out because every matrix is stored [B, size] -- goes to HBM with coalesced writes.
// HBM-bound streaming kernel: algorithmic bytes per agent = 8 (P + sum of matrix sizes); the program itself is
// read once per CTA.  ADD and MUL are separate instructions, so nothing is contracted into an FMA and + - * /
// round exactly as the reference's numpy evaluation does; exp / log / pow / trig are CUDA's FP64 functions (<= 2 ulp).
#include "common.cuh"

namespace hmpc {

constexpr int kMaxOutMats = HMPC_NUM_MATS;

struct ParamEvalArgs {
    int B, P, R, n_ins, n_mats, n_out;
    const int4* prog;
    const double* params;
    double* out;
    int mat_sz[kMaxOutMats];
    int mat_off[kMaxOutMats];   // slot of the matrix' first entry = elements per agent before it
};

__device__ __forceinline__ int expr_arity(int op) {      // -1 = unknown opcode
    if (op == HMPC_EXPR_CONST || op == HMPC_EXPR_PARAM) return 0;
    if (op == HMPC_EXPR_OUT) return 1;
    if (op >= HMPC_EXPR_MOV && op <= HMPC_EXPR_POWI) return 1;
    if (op >= HMPC_EXPR_ADD && op <= HMPC_EXPR_ATAN2) return 2;
    return -1;
}

__device__ __forceinline__ double expr_powi(double x, int n) {
    // exponentiation by squaring, |n| small; negative n = reciprocal of the positive power
    unsigned m = (unsigned)(n < 0 ? -n : n);
    double r = 1.0, p = x;
    while (m) {
        if (m & 1u) r *= p;
        m >>= 1;
        if (m) p *= p;
    }
    return n < 0 ? 1.0 / r : r;
}

__global__ void __launch_bounds__(256) param_eval_kernel(const ParamEvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tw = blockDim.x, t = threadIdx.x;
    const int pitch = tw + 1;                                  // odd: transposed tiles are conflict-free
    int4* sprog = reinterpret_cast<int4*>(smem_raw);
    double* sparam = reinterpret_cast<double*>(smem_raw + (size_t)a.n_ins * sizeof(int4));
    double* sregs = sparam + (size_t)a.P * pitch;
    double* sout = sregs + (size_t)a.R * tw;
    __shared__ int s_bad;
    if (t == 0) s_bad = 0;
    __syncthreads();

    const int64_t b0 = (int64_t)blockIdx.x * tw;
    const int nb = (int)min((int64_t)tw, (int64_t)a.B - b0);   // agents of this tile

    // program -> shared memory, every instruction checked once (a bad program yields NaN, never a wild access)
    for (int i = t; i < a.n_ins; i += tw) {
        const int4 ins = a.prog[i];
        const int ar = expr_arity(ins.x);
        bool ok = ar >= 0;
        if (ins.x == HMPC_EXPR_OUT) ok = ok && ins.y >= 0 && ins.y < a.n_out && ins.z >= 0 && ins.z < a.R;
        else {
            ok = ok && ins.y >= 0 && ins.y < a.R;
            if (ins.x == HMPC_EXPR_PARAM) ok = ok && ins.z >= 0 && ins.z < a.P;
            if (ar >= 1) ok = ok && ins.z >= 0 && ins.z < a.R;
            if (ar == 2) ok = ok && ins.w >= 0 && ins.w < a.R;
        }
        if (!ok) atomicOr(&s_bad, 1);
        sprog[i] = ins;
    }
    // parameter rows of the tile: contiguous nb*P doubles, transposed to sparam[p][thread]
    {
        const double* src = a.params + b0 * a.P;
        const int n = nb * a.P;
        for (int i = t; i < n; i += tw) {
            const int tt = i / a.P, p = i - tt * a.P;
            sparam[p * pitch + tt] = src[i];
        }
        if (t >= nb)                                           // idle lanes of the last tile compute on 1.0
            for (int p = 0; p < a.P; ++p) sparam[p * pitch + t] = 1.0;
    }
    for (int o = 0; o < a.n_out; ++o) sout[o * pitch + t] = __longlong_as_double(0x7ff8000000000000LL);
    __syncthreads();

    if (!s_bad) {
        double* r = sregs + t;
        for (int i = 0; i < a.n_ins; ++i) {
            const int4 ins = sprog[i];
            const int op = ins.x;
            if (op == HMPC_EXPR_OUT) {
                sout[ins.y * pitch + t] = r[ins.z * tw];
                continue;
            }
            double v;
            if (op == HMPC_EXPR_CONST) v = __hiloint2double(ins.w, ins.z);
            else if (op == HMPC_EXPR_PARAM) v = sparam[ins.z * pitch + t];
            else {
                const double x = r[ins.z * tw];
                if (op >= HMPC_EXPR_ADD) {
                    const double y = r[ins.w * tw];
                    switch (op) {
                        case HMPC_EXPR_ADD: v = __dadd_rn(x, y); break;
                        case HMPC_EXPR_SUB: v = __dsub_rn(x, y); break;
                        case HMPC_EXPR_MUL: v = __dmul_rn(x, y); break;
                        case HMPC_EXPR_DIV: v = __ddiv_rn(x, y); break;
                        case HMPC_EXPR_POW: v = pow(x, y); break;
                        case HMPC_EXPR_MIN: v = (x != x || y != y) ? (x + y) : fmin(x, y); break;   // numpy: nan wins
                        case HMPC_EXPR_MAX: v = (x != x || y != y) ? (x + y) : fmax(x, y); break;
                        default: v = atan2(x, y); break;                                         // HMPC_EXPR_ATAN2
                    }
                } else {
                    switch (op) {
                        case HMPC_EXPR_MOV: v = x; break;
                        case HMPC_EXPR_NEG: v = -x; break;
                        case HMPC_EXPR_ABS: v = fabs(x); break;
                        case HMPC_EXPR_SIGN: v = (x != x) ? x : (double)((x > 0.0) - (x < 0.0)); break;
                        case HMPC_EXPR_SQRT: v = sqrt(x); break;
                        case HMPC_EXPR_EXP: v = exp(x); break;
                        case HMPC_EXPR_LOG: v = log(x); break;
                        case HMPC_EXPR_SIN: v = sin(x); break;
                        case HMPC_EXPR_COS: v = cos(x); break;
                        case HMPC_EXPR_TAN: v = tan(x); break;
                        case HMPC_EXPR_ASIN: v = asin(x); break;
                        case HMPC_EXPR_ACOS: v = acos(x); break;
                        case HMPC_EXPR_ATAN: v = atan(x); break;
                        case HMPC_EXPR_SINH: v = sinh(x); break;
                        case HMPC_EXPR_COSH: v = cosh(x); break;
                        case HMPC_EXPR_TANH: v = tanh(x); break;
                        case HMPC_EXPR_FLOOR: v = floor(x); break;
                        case HMPC_EXPR_CEIL: v = ceil(x); break;
                        default: v = expr_powi(x, ins.w); break;                                 // HMPC_EXPR_POWI
                    }
                }
            }
            r[ins.y * tw] = v;
        }
    }
    __syncthreads();

    // output tile -> HBM: matrix m of the tile's agents is the contiguous piece out[B*off_m + b0*sz_m ...)
    for (int m = 0; m < a.n_mats; ++m) {
        const int sz = a.mat_sz[m], off = a.mat_off[m];
        if (sz == 0) continue;
        double* dst = a.out + (int64_t)a.B * off + b0 * sz;
        const int n = nb * sz;
        for (int i = t; i < n; i += tw) {
            const int tt = i / sz, e = i - tt * sz;
            dst[i] = sout[(off + e) * pitch + tt];
        }
    }
}

// ---- EXPERIMENTAL second version (opt-in: hmpc_param_eval_v2_f64; not yet run on a B200) --------------------------
// From the ncu capture of the kernel above (profiles/r1_f4_param_eval_ncu_full_summary.csv): issue-bound, ~56 SASS
// instructions per interpreted one, 13 of the DEWH model's 45 instructions are PARAM copies, occupancy limited by the
// separate parameter tile.  Here (1) the parameters ARE registers 0 .. P-1 of the register file (loaded straight from
// the tile, read-only by convention; programs carry no PARAM instruction), and (2) every thread evaluates TWO agents
// (columns t and t + blockDim.x of a 2 x blockDim.x tile), so one instruction fetch + decode serves two evaluations.
__device__ __forceinline__ double expr_unary(int op, double x, int iw) {
    switch (op) {
        case HMPC_EXPR_MOV: return x;
        case HMPC_EXPR_NEG: return -x;
        case HMPC_EXPR_ABS: return fabs(x);
        case HMPC_EXPR_SIGN: return (x != x) ? x : (double)((x > 0.0) - (x < 0.0));
        case HMPC_EXPR_SQRT: return sqrt(x);
        case HMPC_EXPR_EXP: return exp(x);
        case HMPC_EXPR_LOG: return log(x);
        case HMPC_EXPR_SIN: return sin(x);
        case HMPC_EXPR_COS: return cos(x);
        case HMPC_EXPR_TAN: return tan(x);
        case HMPC_EXPR_ASIN: return asin(x);
        case HMPC_EXPR_ACOS: return acos(x);
        case HMPC_EXPR_ATAN: return atan(x);
        case HMPC_EXPR_SINH: return sinh(x);
        case HMPC_EXPR_COSH: return cosh(x);
        case HMPC_EXPR_TANH: return tanh(x);
        case HMPC_EXPR_FLOOR: return floor(x);
        case HMPC_EXPR_CEIL: return ceil(x);
        default: return expr_powi(x, iw);                                            // HMPC_EXPR_POWI
    }
}

__device__ __forceinline__ double expr_binary(int op, double x, double y) {
    switch (op) {
        case HMPC_EXPR_ADD: return __dadd_rn(x, y);
        case HMPC_EXPR_SUB: return __dsub_rn(x, y);
        case HMPC_EXPR_MUL: return __dmul_rn(x, y);
        case HMPC_EXPR_DIV: return __ddiv_rn(x, y);
        case HMPC_EXPR_POW: return pow(x, y);
        case HMPC_EXPR_MIN: return (x != x || y != y) ? (x + y) : fmin(x, y);
        case HMPC_EXPR_MAX: return (x != x || y != y) ? (x + y) : fmax(x, y);
        default: return atan2(x, y);                                                 // HMPC_EXPR_ATAN2
    }
}

__global__ void __launch_bounds__(128) param_eval2_kernel(const ParamEvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nt = blockDim.x, t = threadIdx.x;
    const int tile = 2 * nt;                                   // agents per CTA
    const int pitch = tile + 1;                                // odd: conflict-free columns and transposing stores
    int4* sprog = reinterpret_cast<int4*>(smem_raw);
    double* sregs = reinterpret_cast<double*>(smem_raw + (size_t)a.n_ins * sizeof(int4));   // [R][pitch], rows 0..P-1 = params
    double* sout = sregs + (size_t)a.R * pitch;                                             // [n_out][pitch]
    __shared__ int s_bad;
    if (t == 0) s_bad = 0;
    __syncthreads();
    const int64_t b0 = (int64_t)blockIdx.x * tile;
    const int nb = (int)min((int64_t)tile, (int64_t)a.B - b0);

    for (int i = t; i < a.n_ins; i += nt) {
        const int4 ins = a.prog[i];
        const int ar = expr_arity(ins.x);
        bool ok = ar >= 0 && ins.x != HMPC_EXPR_PARAM;
        if (ins.x == HMPC_EXPR_OUT) ok = ok && ins.y >= 0 && ins.y < a.n_out && ins.z >= 0 && ins.z < a.R;
        else {
            ok = ok && ins.y >= a.P && ins.y < a.R;            // parameters are never written
            if (ar >= 1) ok = ok && ins.z >= 0 && ins.z < a.R;
            if (ar == 2) ok = ok && ins.w >= 0 && ins.w < a.R;
        }
        if (!ok) atomicOr(&s_bad, 1);
        sprog[i] = ins;
    }
    {
        const double* src = a.params + b0 * a.P;
        const int n = nb * a.P;
        for (int i = t; i < n; i += nt) {
            const int tt = i / a.P, p = i - tt * a.P;
            sregs[p * pitch + tt] = src[i];
        }
        for (int c = t; c < tile; c += nt)
            if (c >= nb)
                for (int p = 0; p < a.P; ++p) sregs[p * pitch + c] = 1.0;
    }
    for (int o = 0; o < a.n_out; ++o) {
        sout[o * pitch + t] = __longlong_as_double(0x7ff8000000000000LL);
        sout[o * pitch + t + nt] = __longlong_as_double(0x7ff8000000000000LL);
    }
    __syncthreads();

    if (!s_bad) {
        double* r0 = sregs + t;
        double* r1 = sregs + t + nt;
        for (int i = 0; i < a.n_ins; ++i) {
            const int4 ins = sprog[i];
            const int op = ins.x;
            if (op == HMPC_EXPR_OUT) {
                sout[ins.y * pitch + t] = r0[ins.z * pitch];
                sout[ins.y * pitch + t + nt] = r1[ins.z * pitch];
                continue;
            }
            double v0, v1;
            if (op == HMPC_EXPR_CONST) {
                v0 = v1 = __hiloint2double(ins.w, ins.z);
            } else if (op >= HMPC_EXPR_ADD) {
                const double x0 = r0[ins.z * pitch], y0 = r0[ins.w * pitch];
                const double x1 = r1[ins.z * pitch], y1 = r1[ins.w * pitch];
                v0 = expr_binary(op, x0, y0);
                v1 = expr_binary(op, x1, y1);
            } else {
                const double x0 = r0[ins.z * pitch], x1 = r1[ins.z * pitch];
                v0 = expr_unary(op, x0, ins.w);
                v1 = expr_unary(op, x1, ins.w);
            }
            r0[ins.y * pitch] = v0;
            r1[ins.y * pitch] = v1;
        }
    }
    __syncthreads();

    for (int m = 0; m < a.n_mats; ++m) {
        const int sz = a.mat_sz[m], off = a.mat_off[m];
        if (sz == 0) continue;
        double* dst = a.out + (int64_t)a.B * off + b0 * sz;
        const int n = nb * sz;
        for (int i = t; i < n; i += nt) {
            const int tt = i / sz, e = i - tt * sz;
            dst[i] = sout[(off + e) * pitch + tt];
        }
    }
}

static size_t param_eval2_smem(int R, int n_ins, int n_out, int nt) {
    return (size_t)n_ins * sizeof(int4) + sizeof(double) * (size_t)(R + n_out) * (2 * nt + 1);
}

static size_t param_eval_smem(int P, int R, int n_ins, int n_out, int tw) {
    return (size_t)n_ins * sizeof(int4) + sizeof(double) * ((size_t)P * (tw + 1) + (size_t)R * tw +
                                                            (size_t)n_out * (tw + 1));
}

}  // namespace hmpc

extern "C" int hmpc_param_eval_f64(int32_t B, int32_t n_params, int32_t n_regs, int32_t n_ins,
                                   const hmpc_expr_ins* program, int32_t n_mats, const int32_t* mat_sizes,
                                   const double* params, double* out, void* stream) {
    using namespace hmpc;
    static_assert(sizeof(hmpc_expr_ins) == sizeof(int4), "instruction = 4 x int32");
    if (B < 0 || n_params < 0 || n_regs < 1 || n_ins < 1 || !program || n_mats < 1 || n_mats > kMaxOutMats ||
        !mat_sizes || (n_params && !params))
        return HMPC_ERR_ARG;
    ParamEvalArgs a;
    a.B = B; a.P = n_params; a.R = n_regs; a.n_ins = n_ins; a.n_mats = n_mats;
    a.prog = reinterpret_cast<const int4*>(program); a.params = params; a.out = out;
    int n_out = 0;
    for (int m = 0; m < kMaxOutMats; ++m) {
        const int sz = m < n_mats ? mat_sizes[m] : 0;
        if (sz < 0) return HMPC_ERR_ARG;
        a.mat_sz[m] = sz; a.mat_off[m] = n_out;
        n_out += sz;
    }
    a.n_out = n_out;
    if (B == 0 || n_out == 0) return HMPC_OK;
    if (!out || ((uintptr_t)program & 15u) != 0) return HMPC_ERR_ARG;   // the program is read as int4
    // tile width: enough CTAs to cover the 148 SMs when the batch allows, wide tiles for long batches
    int tw = 256;
    while (tw > 32 && (int64_t)B < (int64_t)kNumSM * tw) tw >>= 1;
    const size_t limit = 227 * 1024 - 64;
    while (tw > 32 && param_eval_smem(n_params, n_regs, n_ins, n_out, tw) > limit) tw >>= 1;
    const size_t smem = param_eval_smem(n_params, n_regs, n_ins, n_out, tw);
    if (smem > limit) return HMPC_ERR_ARG;                          // program too large: split it per matrix
    if (smem > 48 * 1024)
        HMPC_CUDA_TRY(cudaFuncSetAttribute(param_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ctas = ((int64_t)B + tw - 1) / tw;
    if (ctas > 0x7fffffffLL) return HMPC_ERR_ARG;
    param_eval_kernel<<<(unsigned)ctas, tw, smem, (cudaStream_t)stream>>>(a);
    HMPC_LAUNCH_CHECK("param_eval_kernel");
    return HMPC_OK;
}

extern "C" int64_t hmpc_param_eval_bytes_per_agent(int32_t n_params, int32_t n_mats, const int32_t* mat_sizes) {
    if (n_params < 0 || n_mats < 0 || (n_mats && !mat_sizes)) return -1;
    int64_t n = n_params;
    for (int m = 0; m < n_mats; ++m) n += mat_sizes[m];
    return 8 * n;
}

// EXPERIMENTAL (see param_eval2_kernel): same contract as hmpc_param_eval_f64 except that registers 0 .. n_params-1 hold
// the parameters on entry (n_regs >= n_params + 1 counts them) and PARAM instructions are not allowed.
extern "C" int hmpc_param_eval_v2_f64(int32_t B, int32_t n_params, int32_t n_regs, int32_t n_ins,
                                      const hmpc_expr_ins* program, int32_t n_mats, const int32_t* mat_sizes,
                                      const double* params, double* out, void* stream) {
    using namespace hmpc;
    if (B < 0 || n_params < 0 || n_regs < n_params + 1 || n_ins < 1 || !program || n_mats < 1 ||
        n_mats > kMaxOutMats || !mat_sizes || (n_params && !params))
        return HMPC_ERR_ARG;
    ParamEvalArgs a;
    a.B = B; a.P = n_params; a.R = n_regs; a.n_ins = n_ins; a.n_mats = n_mats;
    a.prog = reinterpret_cast<const int4*>(program); a.params = params; a.out = out;
    int n_out = 0;
    for (int m = 0; m < kMaxOutMats; ++m) {
        const int sz = m < n_mats ? mat_sizes[m] : 0;
        if (sz < 0) return HMPC_ERR_ARG;
        a.mat_sz[m] = sz; a.mat_off[m] = n_out;
        n_out += sz;
    }
    a.n_out = n_out;
    if (B == 0 || n_out == 0) return HMPC_OK;
    if (!out || ((uintptr_t)program & 15u) != 0) return HMPC_ERR_ARG;
    int nt = 128;                                               // threads; 2 agents each
    while (nt > 32 && (int64_t)B < (int64_t)kNumSM * 2 * nt) nt >>= 1;
    const size_t limit = 227 * 1024 - 64;
    while (nt > 32 && param_eval2_smem(n_regs, n_ins, n_out, nt) > limit) nt >>= 1;
    const size_t smem = param_eval2_smem(n_regs, n_ins, n_out, nt);
    if (smem > limit) return HMPC_ERR_ARG;
    if (smem > 48 * 1024)
        HMPC_CUDA_TRY(cudaFuncSetAttribute(param_eval2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ctas = ((int64_t)B + 2 * nt - 1) / (2 * nt);
    if (ctas > 0x7fffffffLL) return HMPC_ERR_ARG;
    param_eval2_kernel<<<(unsigned)ctas, nt, smem, (cudaStream_t)stream>>>(a);
    HMPC_LAUNCH_CHECK("param_eval2_kernel");
    return HMPC_OK;
}
