/* hmpc.h -- C ABI of the B200-native hybrid-MPC hot path (libhmpc.so, sm_100a).
 *
 * Every entry point replaces one step of michchr/pyhybridcontrol's per-step hybrid-MPC solve; the
 * reference interface each one stands in for is cited as file:line (relative to the reference root).
 * The reference is pure Python, so "what its FFI would bind" is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - *_f64 functions take DEVICE pointers and a cudaStream_t passed as void* (NULL = default stream);
 *     they enqueue work and return without synchronising.  *_host_f64 functions take HOST pointers, do the
 *     host<->device copies themselves on an internal stream and return after the result is on the host.
 *   - all matrices are FP64, row-major, batch-major: agent b's block starts at base + b*stride_b elements;
 *     stride_b == 0 broadcasts one block to the whole batch.
 *   - the library never allocates or frees caller buffers; scratch comes in through (workspace, bytes),
 *     sized by the matching *_workspace_bytes query.
 *   - return value: HMPC_OK or a negative hmpc_status; per-problem solver outcomes are in status arrays.
 *   - thread-compatible: no global mutable state; concurrent calls must use distinct workspaces/streams.
 */
#ifndef HMPC_H_
#define HMPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    HMPC_OK = 0,
    HMPC_ERR_ARG = -1,        /* bad dimension / null pointer / unsupported size           */
    HMPC_ERR_CUDA = -2,       /* a CUDA runtime call failed (see hmpc_last_cuda_error)      */
    HMPC_ERR_WORKSPACE = -3,  /* workspace too small                                        */
    HMPC_ERR_NO_DEVICE = -4   /* no sm_100 device visible                                   */
} hmpc_status;

/* per-problem outcome of hmpc_milp_solve_f64 (status[b]) */
typedef enum {
    HMPC_SOLVE_OPTIMAL = 0,     /* proven optimal (within mip_rel_gap)                      */
    HMPC_SOLVE_INFEASIBLE = 1,
    HMPC_SOLVE_NODE_LIMIT = 2,  /* best incumbent returned, not proven                      */
    HMPC_SOLVE_ITER_LIMIT = 3,
    HMPC_SOLVE_NUMERIC = 4,     /* lost dual feasibility / artificial bound active          */
    HMPC_SOLVE_UNSUPPORTED = 5
} hmpc_solve_status;

/* MLD dimensions (reference: models/mld_model.py:149-168) + horizon (controllers/controller_base.py:159-160) */
typedef struct {
    int32_t B;        /* agents in the batch                     */
    int32_t Nt;       /* N_tilde, number of stacked steps        */
    int32_t nx, nu, ndelta, nz, nmu, nomega, ny, nc;
} hmpc_dims;

/* index of each system matrix in the mats[] / mat_stride_b[] arrays of hmpc_condense_f64 */
enum { HMPC_A = 0, HMPC_B1, HMPC_B2, HMPC_B3, HMPC_B4, HMPC_b5,
       HMPC_C, HMPC_D1, HMPC_D2, HMPC_D3, HMPC_D4, HMPC_d5,
       HMPC_E, HMPC_F1, HMPC_F2, HMPC_F3, HMPC_F4, HMPC_f5, HMPC_G, HMPC_Psi, HMPC_NUM_MATS };

/* index of each condensed matrix in out[] */
enum { HMPC_PHI_X = 0, HMPC_GAMMA_V, HMPC_GAMMA_OMEGA, HMPC_GAMMA_5,
       HMPC_L_X, HMPC_L_V, HMPC_L_OMEGA, HMPC_L_5,
       HMPC_H_X, HMPC_H_V, HMPC_H_OMEGA, HMPC_H_5, HMPC_NUM_EVO };

/* ---- library ------------------------------------------------------------------------------------ */
int hmpc_version(void);
const char* hmpc_last_cuda_error(void);           /* message of the last CUDA failure in this thread */
int hmpc_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin_bytes);

/* ---- K1 condense: replaces MldEvoMatrices.gen_mld_evo_matrices
 *      (controllers/components/mld_evolution_matrices.py:108-134; formulas :237-240, :253-332, :467-527).
 *  mats[i]   : [B or 1, rows_i, cols_i]; NULL = all-zero block (C must be given explicitly).
 *  out[i]    : [B, rows*Nt, cols*Nt] dense row-major, NULL = skip.  Shapes: Phi_x (nx*Nt, nx), Gamma_v
 *              (nx*Nt, nv*Nt), Gamma_omega (nx*Nt, nomega*Nt), Gamma_5 (nx*Nt, 1), L_* with ny rows/step,
 *              H_* with nc rows/step; nv = nu+ndelta+nz+nmu, v(k) = [u; delta; z; mu].
 *  The *_N_p variants of the reference are row-prefix views of these (:246-250).                     */
int hmpc_condense_f64(const hmpc_dims* dims, const double* const mats[HMPC_NUM_MATS],
                      const int64_t mat_stride_b[HMPC_NUM_MATS], double* const out[HMPC_NUM_EVO], void* stream);
/* algorithmic bytes one agent's 12 matrices occupy (roofline numerator of K1) */
int64_t hmpc_condense_bytes_per_agent(const hmpc_dims* dims);

/* ---- K2 constraint right-hand side: replaces ConstraintSolvedController.gen_evo_constraints
 *      (controllers/controller_base.py:440-452):  rhs = H_x x0 + H_omega w + H_5        (S == 0)
 *                                                rhs = H_x x0 + rowmin_s(H_omega W) + H_5 (S  > 0)
 *  rows <= nc*Nt selects the reduced-horizon prefix (mld_evolution_matrices.py:89-105).
 *  x0 [B,nx]; w [B, nomega*Nt] or scenarios W [B, nomega*Nt, S]; rhs [B, rows].                       */
int hmpc_constraint_rhs_f64(const hmpc_dims* dims, int32_t rows, const double* H_x, const double* H_omega,
                            const double* H_5, const double* x0, const double* w, int32_t S,
                            double* rhs, void* stream);

/* ---- affine prediction: x~ = Phi x0 + Gamma_v v + Gamma_w w + Gamma_5  (or y~ with the L matrices);
 *      replaces EvoVariables.gen_state_output_vars (controllers/components/variables.py:245-286).
 *  M_x [B,R,nx], M_v [B,R,nvt], M_w [B,R,nwt], M_5 [B,R]; v [B,nvt] may be NULL (constant part only).  */
int hmpc_predict_f64(int32_t B, int32_t R, int32_t nx, int32_t nvt, int32_t nwt, const double* M_x,
                     const double* M_v, const double* M_w, const double* M_5, const double* x0,
                     const double* v, const double* w, double* out, void* stream);

/* ---- linear cost in v-space from Linear atoms on v and on the affine predictions
 *      (controllers/components/objective_atoms.py:308-318, 523-532):
 *        c   = w_v + Gamma_v' w_x + L_v' w_y                       [B, nvt]
 *        c0  = w_x'(x~ at v=0) + w_y'(y~ at v=0)                   [B]
 *  any of w_x / w_y may be NULL.  xc / yc are the constant parts from hmpc_predict_f64(v = NULL).       */
int hmpc_linear_cost_f64(int32_t B, int32_t nvt, int32_t nxt, int32_t nyt, const double* w_v, int64_t w_v_stride_b,
                         const double* w_x, const double* Gamma_v, const double* xc,
                         const double* w_y, const double* L_v, const double* yc,
                         double* c, double* c0, void* stream);

/* ---- K3/K4 mixed-integer solve: replaces cvx.Problem.solve -> Gurobi/CPLEX inside
 *      ConstraintSolvedController.solve (controllers/controller_base.py:509-512).
 *      minimise c'v (+c0)  s.t.  H v <= rhs,  lb <= v <= ub,  v[j] in {0,1} where is_bin[j] != 0.
 *  One CTA per problem: row-generating bounded dual simplex on a shared-memory tableau, c-MIR cuts,
 *  depth-first branch and bound.                                               */
typedef struct {
    double  mip_rel_gap;     /* 0 = prove optimality (reference ran MIPGap=1e-2)                        */
    double  int_tol;         /* integrality tolerance, default 1e-6                                     */
    double  feas_tol;        /* primal feasibility tolerance, default 1e-9                              */
    double  big_bound;       /* artificial box for free columns, default 1e7                            */
    int32_t max_nodes;       /* per problem, default 200000                                             */
    int32_t max_pivots;      /* per problem, default 2000000                                            */
    int32_t max_cuts;        /* cut-pool capacity per problem, default 512                              */
    int32_t max_rows;        /* active tableau rows (0 = as many as shared memory allows)               */
    int32_t cut_rounds_root; /* default 30                                                              */
    int32_t cut_rounds_node; /* default 2                                                               */
    int32_t cuts_per_round;  /* default 8                                                               */
    int32_t force_general;   /* host front door only: 1 = always use this kernel, never the stage-DP path   */
} hmpc_milp_opts;

void hmpc_milp_default_opts(hmpc_milp_opts* opts);
int  hmpc_milp_workspace_bytes(int32_t B, int32_t n, int32_t m, const hmpc_milp_opts* opts, size_t* bytes);
/*  c [B,n] (stride_c_b), H [B,m,n] (stride_H_b), rhs [B,m], lb/ub [B,n] (stride_bnd_b; +-inf allowed),
 *  is_bin [n] bytes shared by the batch.  Outputs: v [B,n], obj [B] (= c'v, caller adds c0),
 *  status [B] (hmpc_solve_status), stats [B,8] = {nodes, pivots, cuts, rows_added, max_rows, lp_solves,
 *  purges, kilo-FMAs (1024 algorithmic FP64 FMAs)}.                                                                                 */
int  hmpc_milp_solve_f64(int32_t B, int32_t n, int32_t m,
                         const double* c, int64_t stride_c_b, const double* H, int64_t stride_H_b,
                         const double* rhs, const double* lb, const double* ub, int64_t stride_bnd_b,
                         const uint8_t* is_bin, const hmpc_milp_opts* opts,
                         void* workspace, size_t workspace_bytes,
                         double* v, double* obj, int32_t* status, int32_t* stats, void* stream);

/* ---- K3q/K4q mixed-integer QUADRATIC solve for any MLD: the same call of the reference as hmpc_milp_solve_f64
 *      (controllers/controller_base.py:509-512) when the cost carries Quadratic / L22 atoms with dense weights
 *      (controllers/components/objective_atoms.py:320-343, 185-206), L1 / Linf atoms (:338-363; the caller adds their
 *      epigraph columns and rows) or non-linear rate atoms (:297-305):
 *        minimise 0.5 v'P v + c'v  s.t.  H v <= rhs,  lb <= v <= ub,  v[j] in {0,1} where is_bin[j] != 0,   P >= 0.
 *      One CTA per problem: depth-first branch and bound over an ADMM (OSQP-style splitting) relaxation whose linear
 *      system is factorised once per problem.  Accuracy is that of `eps` (objectives to
 *      ~1e-7 relative at the default).
 *  P [B|1,n,n] (stride_P_b; NULL: an MILP), c [B|1,n], H [B|1,m,n], rhs [B,m], lb/ub/is_bin [n] shared by the batch.
 *  stats [B,8] = {nodes, ADMM iterations, incumbent updates, binaries, 0, 0, 0, thousands of FMAs}.            */
typedef struct {
    double  mip_rel_gap;   /* 0 = prove optimality (up to eps)                        */
    double  int_tol;       /* integrality tolerance of a relaxation, default 1e-6      */
    double  eps;           /* absolute = relative ADMM tolerance (scaled problem), default 1e-9 */
    double  rho;           /* initial step parameter, default 0.1 (adapted at the root) */
    int32_t max_nodes;     /* per problem, default 100000                              */
    int32_t max_iter;      /* ADMM iterations per node, default 50000                  */
} hmpc_miqp_opts;
void hmpc_miqp_default_opts(hmpc_miqp_opts* opts);
int  hmpc_miqp_workspace_bytes(int32_t B, int32_t n, int32_t m, size_t* bytes);
int  hmpc_miqp_solve_f64(int32_t B, int32_t n, int32_t m, const double* P, int64_t stride_P_b,
                         const double* c, int64_t stride_c_b, const double* H, int64_t stride_H_b,
                         const double* rhs, const double* lb, const double* ub, const uint8_t* is_bin,
                         const hmpc_miqp_opts* opts, void* workspace, size_t workspace_bytes,
                         double* v, double* obj, int32_t* status, int32_t* stats, void* stream);

/* ---- K3s/K4s exact solve for SCALAR-STATE MLDs (the reference example's water heaters): same problem and
 *      same boundary as hmpc_milp_solve_f64 (controllers/controller_base.py:509-512), for the class
 *        nx == 1, nz == 0, every input/delta binary (nb = nu + ndelta <= 4), Psi = -diag(d) (each row owns at most
 *        its own slack; nmu in {0, nc}), linear cost, Nt <= 128, nb*Nt <= 128
 *      (examples/residential_mg_with_pv_and_dewhs/modelling/micro_grid_models.py:27-100).  Instead of the condensed
 *      H_v it reads the MLD blocks directly (mats as in hmpc_condense_f64: A, B1, B2, C, D1, D2, E, F1, F2, G, Psi)
 *      and solves  min cost_v'v  s.t.  H_v v <= rhs  exactly: a backward value-table lower bound over the scaled
 *      scalar state + an exact depth-first search over the binary sequence.
 *  rhs [B, nc*Nt] (from hmpc_constraint_rhs_f64; several constraint sets over the same rows fold into their
 *  row-wise minimum), cost_v [B|1, nv*Nt], lb_v/ub_v/is_bin_v [nv*Nt] shared by the batch, v(k) = [u; delta; mu].
 *  status[b] = HMPC_SOLVE_UNSUPPORTED marks an agent outside the class (use hmpc_milp_solve_f64 for it).
 *  stats [B,8] = {search nodes, set-up time, sweep time, cells, search time (device time of the agent's CTA / warp, in
 *  units of 0.1 us), incumbent updates, certified relative gap of an unfinished search in units of 1e-9 (0 when
 *  proven), thousands of FP64-pipe instructions executed}.                                                    */
typedef struct {
    double  mip_rel_gap;   /* 0 = prove optimality                                     */
    double  feas_tol;      /* tolerance of hard (slack-free) rows, default 1e-9        */
    int32_t cells;         /* value-table cells per stage, default 4096 (the optimum does not depend on it: fewer
                              cells = a shorter sweep and a few more search expansions)                          */
    int32_t max_nodes;     /* search nodes per agent, default 4,000,000                */
    int32_t table_fp64;    /* 1 (default): FP64 value table -- sequences that TIE with the incumbent (piecewise-constant
                              tariffs) are pruned at mip_rel_gap = 0;  0: FP32 table rounded down -- half the
                              workspace and shared memory, ~5-10 % faster, but ties are explored unless
                              mip_rel_gap >= 4e-6                                                               */
    int32_t bound;         /* HMPC_DP_BOUND_CONSTANT (default): one value per cell;  HMPC_DP_BOUND_LINEAR: a line per
                              cell (16 bytes), exact where slack penalties make the cost-to-go steep (full-horizon
                              robust constraint sets); DEWH shape only, other agents get constant lines; the cell
                              count must fit shared memory (hmpc_stage_dp_max_cells)                             */
    int32_t fuse_search;   /* -1 (default): the table kernel's tail searches the agent itself, to the end, when
                              B <= 296 (ONE launch per solve); else the table kernel is followed by a one-warp
                              search kernel and a team-search kernel for what that leaves;  0 / 1: never / always */
    int32_t reserved;
} hmpc_stage_dp_opts;
enum { HMPC_DP_BOUND_CONSTANT = 0, HMPC_DP_BOUND_LINEAR = 1 };
/* Optional convex cost terms of the stage-DP solve -- the reference's Quadratic / L22 / L1 atoms on the state, the
 * outputs and the slacks (controllers/components/objective_atoms.py:320-363), which make the problem an MIQP:
 *     cost += sum_k sum_t  wq[k,t] tau_k,t^2 + w1[k,t] |tau_k,t|  +  sum_k sum_i qmu[k,i] mu_k,i^2 ,
 *     tau_k,t = h[t] p_k + ga[t]' alpha_k + r[k,t]
 * p_k = forced response of the state (row k of Gamma_v times v); the caller folds h[t] * (free response) and any
 * constant into r.  All weights must be >= 0 (convex), else the agent reports HMPC_SOLVE_UNSUPPORTED.           */
typedef struct {
    int32_t T;                                 /* state terms per stage, 0..4                                 */
    int32_t reserved;
    const double* h;   int64_t h_stride_b;     /* [B|1, T]                                                    */
    const double* ga;  int64_t ga_stride_b;    /* [B|1, T, nb]   NULL = 0                                     */
    const double* r;                           /* [B, Nt, T]                                                  */
    const double* wq;  int64_t wq_stride_b;    /* [B|1, Nt, T]   NULL = 0                                     */
    const double* w1;  int64_t w1_stride_b;    /* [B|1, Nt, T]   NULL = 0                                     */
    const double* qmu; int64_t qmu_stride_b;   /* [B|1, Nt, nc]  NULL = 0; may be given with T == 0           */
} hmpc_stage_terms;
void hmpc_stage_dp_default_opts(hmpc_stage_dp_opts* opts);
int  hmpc_stage_dp_supported(const hmpc_dims* dims);     /* 1 when the dimensions fit the class */
int  hmpc_stage_dp_workspace_bytes(const hmpc_dims* dims, const hmpc_stage_dp_opts* opts, size_t* bytes);
/* largest cell count (multiple of 256) whose stage buffers fit shared memory for these dimensions and this format */
int  hmpc_stage_dp_max_cells(const hmpc_dims* dims, const hmpc_stage_dp_opts* opts, int32_t* cells);
int  hmpc_stage_dp_solve_f64(const hmpc_dims* dims, const double* const mats[HMPC_NUM_MATS],
                             const int64_t mat_stride_b[HMPC_NUM_MATS], const double* rhs,
                             const double* cost_v, int64_t cost_v_stride_b, const double* lb_v, const double* ub_v,
                             const uint8_t* is_bin_v, const hmpc_stage_terms* terms /* NULL = linear cost */,
                             const hmpc_stage_dp_opts* opts, void* workspace, size_t workspace_bytes,
                             double* v, double* obj, int32_t* status, int32_t* stats, void* stream);

/* ---- K5 simulation step: replaces MldModel.lsim_k (models/mld_model.py:647-699): mu ignored in cons
 *      (:694), tolerance cons_tol (default 1e-6, :648).  mats as in hmpc_condense_f64.
 *  x [B,nx], u [B,nu], delta [B,ndelta], z [B,nz], w [B,nomega] -> x1 [B,nx], y [B,ny], cons [B,nc] bytes */
int hmpc_lsim_step_f64(const hmpc_dims* dims, const double* const mats[HMPC_NUM_MATS],
                       const int64_t mat_stride_b[HMPC_NUM_MATS], const double* x, const double* u,
                       const double* delta, const double* z, const double* w, double cons_tol,
                       double* x1, double* y, uint8_t* cons, void* stream);

/* ---- DEWH simulation model re-parametrisation + step: replaces DewhAgentMpc.sim_step_k
 *      (examples/residential_mg_with_pv_and_dewhs/modelling/micro_grid_agents.py:389-408) and the
 *      const_heat=False model (micro_grid_models.py:45-57).  params [B,12] =
 *      {C_w, A_h, U_h, m_h, T_w, T_inf, P_h_Nom, T_h_min, T_h_max, T_h_Nom, ts, reserved}.
 *  T [B] is clamped to T_w+0.1 first; outputs T1 [B], model [B,4] = {A, B1, B4, b5} (may be NULL),
 *  cons [B,2] bytes (may be NULL).                                                                     */
int hmpc_dewh_sim_step_f64(int32_t B, const double* params, const double* T, const double* u,
                           const double* D_h, double* T1, double* model, uint8_t* cons, void* stream);
/* control model (const_heat=True): model [B,4] = {A, B1, B4, b5} */
int hmpc_dewh_control_model_f64(int32_t B, const double* params, double* model, void* stream);
/* thermostat rule of the example's non-predictive controller: replaces DewhTheromstatController.solve
 *      (examples/residential_mg_with_pv_and_dewhs/theromstat_control.py:38-62).  band [B,2] (stride 2) or [1,2]
 *      (stride 0) = {T_h_max_sub_T_h_on, T_h_max_sub_T_h_off} (parameters.py:21-22); T [B]; u_prev [B] is the input
 *      applied at k-1 (the reference keeps "on" only if it is exactly 1); u [B] out (0.0 / 1.0).         */
int hmpc_dewh_thermostat_f64(int32_t B, const double* params, const double* band, int64_t band_stride_b,
                             const double* T, const double* u_prev, double* u, void* stream);

/* ---- symbolic / callable model front-end: replaces, for B parameter sets at once, the evaluation of every
 *      non-constant system matrix of a callable MldModel -- CallableMatrix.__call__(param_struct=...)
 *      (utils/matrix_utils.py:441-470) on the sympy.lambdify'd matrix functions (:339-343), as driven by
 *      MldModel.to_numeric (models/mld_model.py:791-793) and MldSystemModel.get_mld_numeric / update_param_struct
 *      (:1072-1081, :1128-1149).  The host compiles the expressions of all matrices into ONE straight-line register
 *      program (pyhybridcontrol_b200/utils/matrix_utils.py); one thread per agent interprets it.
 *  program   : DEVICE pointer, 16-byte aligned, n_ins instructions {op, dst, a, b}:
 *                CONST  dst <- the double whose low / high 32 bits are a / b
 *                PARAM  dst <- params[agent, a]
 *                unary  dst <- f(reg a)            (POWI: reg a to the integer power b)
 *                binary dst <- reg a (op) reg b
 *                OUT    output slot dst <- reg a
 *              + - * / are IEEE round-to-nearest and never contracted into an FMA (same rounding as the reference's
 *              numpy evaluation); exp / log / pow / trig are CUDA's FP64 functions (<= 2 ulp).
 *              A malformed program (register / slot / parameter index out of range, unknown opcode) makes the
 *              affected outputs NaN; it never reads or writes out of bounds.
 *  mat_sizes : HOST array, n_mats <= 20 entries = rows*cols of each evaluated matrix; output slots are numbered
 *              matrix by matrix, row-major inside a matrix.
 *  params    : [B, n_params] row-major.
 *  out       : matrix m occupies out[B*off_m .. B*(off_m + size_m)) as [B, size_m] row-major (off_m = sum of the
 *              sizes before it), i.e. each matrix is a contiguous [B, rows, cols] block, ready for the mats[] /
 *              mat_stride_b[] arguments of the entry points above.  Slots no OUT writes stay NaN.
 *  Limits: n_regs, n_ins and the sizes must fit one CTA's shared memory (16 n_ins + 8*33 (n_params + n_regs + n_out)
 *  <= 227 KB), else HMPC_ERR_ARG -- split the program per matrix.                                         */
typedef struct { int32_t op, dst, a, b; } hmpc_expr_ins;
enum { HMPC_EXPR_CONST = 0, HMPC_EXPR_PARAM = 1, HMPC_EXPR_OUT = 2,
       HMPC_EXPR_MOV = 10, HMPC_EXPR_NEG, HMPC_EXPR_ABS, HMPC_EXPR_SIGN, HMPC_EXPR_SQRT, HMPC_EXPR_EXP,
       HMPC_EXPR_LOG, HMPC_EXPR_SIN, HMPC_EXPR_COS, HMPC_EXPR_TAN, HMPC_EXPR_ASIN, HMPC_EXPR_ACOS, HMPC_EXPR_ATAN,
       HMPC_EXPR_SINH, HMPC_EXPR_COSH, HMPC_EXPR_TANH, HMPC_EXPR_FLOOR, HMPC_EXPR_CEIL, HMPC_EXPR_POWI,
       HMPC_EXPR_ADD = 40, HMPC_EXPR_SUB, HMPC_EXPR_MUL, HMPC_EXPR_DIV, HMPC_EXPR_POW, HMPC_EXPR_MIN,
       HMPC_EXPR_MAX, HMPC_EXPR_ATAN2 };
int hmpc_param_eval_f64(int32_t B, int32_t n_params, int32_t n_regs, int32_t n_ins, const hmpc_expr_ins* program,
                        int32_t n_mats, const int32_t* mat_sizes, const double* params, double* out, void* stream);
/* EXPERIMENTAL variant, opt-in, NOT yet run on a B200 (pyhybridcontrol_b200: HMPC_PARAM_EVAL=v2): same contract, but
 * registers 0 .. n_params-1 hold the parameters on entry (n_regs counts them; they must not be written), PARAM
 * instructions are not allowed, and every thread evaluates two agents (one decode, two evaluations).            */
int hmpc_param_eval_v2_f64(int32_t B, int32_t n_params, int32_t n_regs, int32_t n_ins, const hmpc_expr_ins* program,
                           int32_t n_mats, const int32_t* mat_sizes, const double* params, double* out, void* stream);
/* algorithmic bytes per agent of the calls above: 8 (n_params + sum of sizes) -- HBM-bound by design */
int64_t hmpc_param_eval_bytes_per_agent(int32_t n_params, int32_t n_mats, const int32_t* mat_sizes);

/* ---- K6 aggregate power: replaces GridAgentMpc.get_grid_device_powers_N_tilde + GridModel D4 = ones
 *      (micro_grid_agents.py:625-646, micro_grid_models.py:143):  P_agg[k] = sum_b P_nom[b] * u[b,k].
 *  Deterministic two-pass tree; partial [ceil(B/16), Nt] scratch from the caller.                     */
int hmpc_aggregate_power_f64(int32_t B, int32_t Nt, const double* u, int64_t u_stride_b, int32_t u_stride_k,
                             const double* P_nom, double* partial, double* P_agg, void* stream);
/* K6 across the GPUs of one box WITHOUT a collective library (replaces the per-step NCCL all-reduce of the [Nt] sums):
 * `windows` is a device array of `world` pointers, windows[r] = rank r's exchange window of
 * hmpc_aggregate_window_doubles(Nt, world) doubles, zero-initialised, addressable from this GPU (peer mapping over
 * NVLink, e.g. torch symmetric memory).  publish = the local reduction whose last pass stores this rank's sums into
 * EVERY rank's window and then raises the step's flag there (release, system scope); gather waits (bounded spin on
 * local memory; a peer that never arrives turns the result into NaN and sets the window's error word instead of
 * hanging) for the world's flags of this rank's latest step and adds the contributions in rank order.  Both are plain
 * kernels on `stream`: capturable in the step's CUDA graph.  A ring of 8 steps: the ranks may drift up to lag <= 6 steps apart.                                                                                                        */
int64_t hmpc_aggregate_window_doubles(int32_t Nt, int32_t world);
int hmpc_aggregate_publish_f64(int32_t B, int32_t Nt, const double* u, int64_t u_stride_b, int32_t u_stride_k,
                               const double* P_nom, double* partial, int32_t world, int32_t rank,
                               double* const* windows, double* P_total_prev /* NULL, or [Nt]: the same launch also
                               gathers the step `lag` (1..6) publishes back */, int32_t lag, void* stream);
int hmpc_aggregate_gather_f64(int32_t Nt, int32_t world, int32_t rank, double* window, double* P_total,
                              int64_t spin_limit, int32_t lag /* 0: this rank's latest step; 1..6: that many steps
                              back -- a pipelined loop gathers step s-1 while step s is published, so that no rank
                              ever waits for a slower one inside the control loop */, void* stream);


/* ---- price coordination for the CENTRALISED micro-grid problem: replaces GridAgentMpc.build_grid / solve_grid_mpc
 *      (micro_grid_agents.py:691-735), where one MILP holds every device and the price sits on the grid import
 *      z_k = max(0, y_k), y_k = sum_i P_i u_i,k + other_k (grid MLD, micro_grid_models.py:145-168; q_z,
 *      micro_grid_control_simulation.py:228-229).  Lagrangian relaxation of a_k = sum_i P_i u_i,k: one iteration =
 *      price_cost -> the batched agent solve (hmpc_stage_dp_solve_f64 / hmpc_milp_solve_f64) -> sums -> [all-reduce of
 *      sums across ranks] -> dual_step -> keep_best; nothing returns to the host in between.
 *  price_cost: cost_v[b, k*nv + col] = lambda[k] * P_nom[b].
 *  sums [Nt+2] = {P_agg[0..Nt), sum_b obj[b], number of agents with status != 0} (status may be NULL).
 *  dual_step: state [8] = {best lower bound, best upper bound, this iterate's dual value, this iterate's primal cost,
 *      1.0 if the upper bound improved, iterations, |subgradient|^2, iterations skipped because an agent failed};
 *      initialise to {-inf, +inf, 0, 0, 0, 0, 0, 0}.  a_lo / a_hi [Nt] bound the aggregate (0 and sum_i P_i, tightened
 *      by the grid limits P_g_min - other_k, P_g_max - other_k); a plan outside them has no upper bound.
 *      lambda_next = clip(lambda + theta (UB - dual) / |g|^2 * g, 0, price); must not alias lambda.
 *  keep_best: if the upper bound improved, u_best [B,Nt] <- u and lambda_best [Nt] <- lambda (may be NULL).          */
int hmpc_coupling_price_cost_f64(int32_t B, int32_t Nt, int32_t nv, int32_t col, const double* lambda,
                                 const double* P_nom, double* cost_v, int64_t cost_stride_b, void* stream);
int hmpc_coupling_sums_f64(int32_t B, int32_t Nt, const double* u, int64_t u_stride_b, int32_t u_stride_k,
                           const double* P_nom, const double* obj, const int32_t* status, double* sums, void* stream);
int hmpc_coupling_dual_step_f64(int32_t Nt, const double* sums, const double* p_other, const double* price,
                                const double* a_lo, const double* a_hi, double theta, const double* lambda,
                                double* lambda_next, double* state, void* stream);
int hmpc_coupling_keep_best_f64(int32_t B, int32_t Nt, const double* u, int64_t u_stride_b, int32_t u_stride_k,
                                const double* lambda, const double* state, double* u_best, double* lambda_best,
                                void* stream);

/* best-response descent on the true centralised cost, started from the plan of the price coordination: the agents of
 * a block [lo, hi) answer their own marginal price with the others fixed,
 *      c_bk = price_k [ max(0, A_k - P_b u_bk + P_b + other_k) - max(0, A_k - P_b u_bk + other_k) ],
 * and the block's new plans are kept only if the total (import cost + every agent's penalty) went down, so the
 * upper bound never gets worse.  One block step = response_cost -> the agent solve on rows [lo, hi) -> merge -> sums
 * (obj = pen_cur, status NULL) -> accept -> restore.
 *  response_cost: agg [Nt] = the current aggregate (sums_cur); v_cur / cost_v [B, Nt*nv] with row stride stride_b.
 *  merge: v_new [hi-lo, Nt*nv], obj_new / status_new [hi-lo] from the solve; rows [lo, hi) of v_cur / pen_cur are saved
 *      to v_bak / pen_bak and replaced (penalty = objective - energy part; +inf for a failed agent).
 *  accept: br_state [4] = {total cost of the kept plan, 1.0 if the candidate was accepted, accepted, rejected};
 *      initialise to {+inf, 0, 0, 0} and merge the whole fleet once to load the starting plan.
 *  restore: puts rows [lo, hi) back if the candidate was rejected.                                               */
int hmpc_coupling_response_cost_f64(int32_t B, int32_t Nt, int32_t nv, int32_t col, const double* agg,
                                    const double* v_cur, const double* P_nom, const double* p_other,
                                    const double* price, double* cost_v, int64_t stride_b, void* stream);
int hmpc_coupling_merge_f64(int32_t lo, int32_t hi, int32_t Nt, int32_t nv, int32_t col, const double* v_new,
                            const double* obj_new, const int32_t* status_new, const double* cost_v, int64_t stride_b,
                            double* v_cur, double* pen_cur, double* v_bak, double* pen_bak, void* stream);
int hmpc_coupling_accept_f64(int32_t Nt, const double* sums_cand, const double* p_other, const double* price,
                             const double* a_lo, const double* a_hi, double* sums_cur, double* br_state, void* stream);
int hmpc_coupling_restore_f64(int32_t lo, int32_t hi, int32_t nvt, const double* br_state, const double* v_bak,
                              const double* pen_bak, double* v_cur, int64_t stride_b, double* pen_cur, void* stream);

/* ---- host-buffer front door: one whole control step for a batch (what a ctypes/cgo/JNI caller binds).
 *      Replaces, for every agent of the batch, MpcController.build() + solve()
 *      (controllers/mpc_controller.py:76-101, controllers/controller_base.py:491-540) with Linear cost atoms.
 *  All pointers are HOST memory.  The plan owns the device buffers, pinned staging and a stream; the call
 *  copies the inputs host->device, runs K1 (only when `recondense` != 0) -> K2 -> K3/K4, copies the results
 *  device->host and returns after they have landed.  K3/K4 is hmpc_stage_dp_solve_f64 when the MLD is in its
 *  class (falling back to hmpc_milp_solve_f64 if an agent reports HMPC_SOLVE_UNSUPPORTED), else
 *  hmpc_milp_solve_f64.
 *    mats/strides : as hmpc_condense_f64 (ignored when recondense == 0)
 *    x0 [B,nx], w [B,nomega*Nt], cost_v [B or 1, nv*Nt] (stride 0 = broadcast): linear cost on v~
 *    lb_v, ub_v [nv*Nt] (+-inf allowed), is_bin_v [nv*Nt]: shared by the batch
 *    outputs: v [B,nv*Nt], obj [B], status [B], stats [B,8]; timing_ms[4] = {h2d, kernels, d2h, total} (may be NULL)  */
typedef struct hmpc_step_plan hmpc_step_plan;
int hmpc_step_plan_create(const hmpc_dims* dims, const hmpc_milp_opts* opts, hmpc_step_plan** plan);
int hmpc_step_plan_destroy(hmpc_step_plan* plan);
int hmpc_mpc_step_host_f64(hmpc_step_plan* plan, int32_t recondense, const double* const mats[HMPC_NUM_MATS],
                           const int64_t mat_stride_b[HMPC_NUM_MATS], const double* x0, const double* w,
                           const double* cost_v, int64_t cost_v_stride_b, const double* lb_v, const double* ub_v,
                           const uint8_t* is_bin_v, double* v, double* obj, int32_t* status, int32_t* stats,
                           float* timing_ms);
/* which solve kernels the last call used: 0 = hmpc_milp_solve_f64, 1 = hmpc_stage_dp_solve_f64 */
int hmpc_step_plan_last_solver(const hmpc_step_plan* plan);
/* bytes moved per step by the call above: {host->device, device->host} */
int hmpc_mpc_step_host_bytes(const hmpc_step_plan* plan, int32_t recondense, int64_t* h2d, int64_t* d2h);

/* ---- measured FP64 FMA peak of this device (roofline denominator for the solver), TFLOP/s ---------- */
int hmpc_fp64_peak_probe(double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HMPC_H_ */
