"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The MI(Q)P solve the reference delegates to ``cvx.Problem.solve`` -> Gurobi / CPLEX
(controllers/controller_base.py:509-512; models/mld_model.py:750,753).  Neither cvxpy nor those solvers is
installed or pinned by the reference; the offline backend here is **HiGHS 1.12.0 vendored in scipy 1.18.1**:

* ``solve_milp``      -- linear cost: ``scipy.optimize.milp`` with ``mip_rel_gap=0`` (what the reference example
                         poses: all its atoms are Linear, SURVEY.md section 7 hard part 2).
* ``solve_enumerate`` -- exhaustive enumeration over the binaries (<= ~18) with an LP/QP in the continuous
                         variables; independent cross-check of both HiGHS and the GPU solver.
* ``solve_miqp``      -- quadratic cost: depth-first branch-and-bound over HiGHS QP relaxations
                         (private ``scipy.optimize._highspy._core``; HiGHS itself has no MIQP mode).

Parity status: the reference has no tests and its solver (Gurobi / CPLEX through cvxpy) cannot run here, so the
SOLVER stays unpinned; what is pinned is everything around it -- the unmodified reference's ``MpcController.solve`` /
``feedback`` / ``sim_step_k`` run in the build container with HiGHS behind ``oracle/mini_cvxpy.py``, and
``tests/test_oracle_assembly_pinned.py`` holds this module to the objectives, first-step values and the 14-instant
closed loop they produce (``tests/golden/assembly_*.npz``).  HiGHS and enumeration are checked against each other in
tests/test_oracle_solve.py.
"""
import itertools

import numpy as np
from scipy.optimize import Bounds, LinearConstraint, linprog, milp

OPTIMAL, INFEASIBLE, LIMIT = 0, 1, 2


def solve_milp(prob, mip_rel_gap=0.0, time_limit=None, polish=False):
    """-> (status, objective, v).  Requires prob.P is None.

    polish=True re-solves the continuous part with the binaries fixed at HiGHS' decisions and 1e-10 feasibility
    tolerances: HiGHS' MIP answer is only accurate to its 1e-6/1e-7 primal tolerance times the largest cost
    coefficient (the slack penalties here), which is of the order of the 1e-6 parity bar."""
    assert prob.P is None or not np.any(prob.P)
    cons = [LinearConstraint(prob.H, -np.inf, prob.rhs)] if prob.H.shape[0] else []
    opts = {"mip_rel_gap": mip_rel_gap, "presolve": True}
    if time_limit:
        opts["time_limit"] = time_limit
    res = milp(prob.c, constraints=cons, integrality=prob.is_bin.astype(int),
               bounds=Bounds(prob.lb, prob.ub), options=opts)
    if res.status == 0:
        v = np.array(res.x)
        v[prob.is_bin] = np.round(v[prob.is_bin])
        if polish and mip_rel_gap == 0.0:
            obj2, v2 = _continuous_subproblem(prob, v[prob.is_bin])
            if v2 is not None:
                v2[prob.is_bin] = v[prob.is_bin]
                return OPTIMAL, obj2, v2
        return OPTIMAL, float(res.fun + prob.c0), v
    if res.status == 2:
        return INFEASIBLE, np.inf, None
    return LIMIT, (float(res.fun + prob.c0) if res.x is not None else np.inf), (None if res.x is None else np.array(res.x))


def _continuous_subproblem(prob, bin_vals):
    """Fix the binaries, solve the remaining LP/QP exactly."""
    lb, ub = prob.lb.copy(), prob.ub.copy()
    lb[prob.is_bin] = bin_vals
    ub[prob.is_bin] = bin_vals
    if prob.P is None or not np.any(prob.P):
        res = linprog(prob.c, A_ub=prob.H if prob.H.shape[0] else None, b_ub=prob.rhs if prob.H.shape[0] else None,
                      bounds=np.c_[lb, ub], method="highs",
                      options={"primal_feasibility_tolerance": 1e-10, "dual_feasibility_tolerance": 1e-10})
        if res.status != 0:
            return np.inf, None
        return float(res.fun + prob.c0), np.array(res.x)
    st, obj, v = solve_qp(prob, lb, ub)
    return (obj, v) if st == OPTIMAL else (np.inf, None)


def solve_enumerate(prob, max_bin=18):
    """Brute force over all binary assignments -> (status, objective, v, runner_up_objective)."""
    nb = int(prob.is_bin.sum())
    if nb > max_bin:
        raise ValueError("too many binaries for enumeration: %d" % nb)
    best, best_v, second = np.inf, None, np.inf
    for bits in itertools.product((0.0, 1.0), repeat=nb):
        obj, v = _continuous_subproblem(prob, np.array(bits))
        if obj < best:
            second, best, best_v = best, obj, v
        elif obj < second:
            second = obj
    if best_v is None:
        return INFEASIBLE, np.inf, None, np.inf
    return OPTIMAL, best, best_v, second


def solve_qp(prob, lb=None, ub=None):
    """Convex QP relaxation.  HiGHS' QP solver (private scipy binding) first; it gives up (neither optimal nor
    infeasible) on a fraction of a percent of the sub-problems with many fixed columns, so those are retried with the
    fixed columns eliminated and, as a last resort, with scipy's SLSQP."""
    lb = prob.lb if lb is None else lb
    ub = prob.ub if ub is None else ub
    st, obj, v = _solve_qp_highs(prob, lb, ub)
    if st != LIMIT:
        return st, obj, v
    fixed = np.isfinite(lb) & (lb == ub)
    if fixed.any() and not fixed.all():
        from .assemble import Problem
        fr = ~fixed
        xf = lb[fixed]
        red = Problem(int(fr.sum()))
        P = prob.P if prob.P is not None else np.zeros((prob.n, prob.n))
        Ps = 0.5 * (P + P.T)
        red.P = Ps[np.ix_(fr, fr)]
        red.c = prob.c[fr] + Ps[np.ix_(fr, fixed)] @ xf
        red.c0 = prob.c0 + float(prob.c[fixed] @ xf) + 0.5 * float(xf @ Ps[np.ix_(fixed, fixed)] @ xf)
        red.H = prob.H[:, fr]
        red.rhs = prob.rhs - prob.H[:, fixed] @ xf
        red.lb, red.ub = lb[fr], ub[fr]
        st, obj, vr = _solve_qp_highs(red, red.lb, red.ub)
        if st == LIMIT:
            st, obj, vr = _solve_qp_slsqp(red)
        if st == OPTIMAL:
            v = np.zeros(prob.n)
            v[fixed], v[fr] = xf, vr
            return OPTIMAL, prob.objective(v), v
        return st, np.inf, None
    return _solve_qp_slsqp(prob, lb, ub)


def _solve_qp_slsqp(prob, lb=None, ub=None):
    from scipy.optimize import minimize
    lb = prob.lb if lb is None else lb
    ub = prob.ub if ub is None else ub
    P = 0.5 * (prob.P + prob.P.T) if prob.P is not None else np.zeros((prob.n, prob.n))
    x0 = np.clip(np.zeros(prob.n), np.where(np.isfinite(lb), lb, -1e6), np.where(np.isfinite(ub), ub, 1e6))
    cons = [{"type": "ineq", "fun": lambda x: prob.rhs - prob.H @ x, "jac": lambda x: -prob.H}] if prob.H.shape[0] else []
    res = minimize(lambda x: 0.5 * x @ P @ x + prob.c @ x, x0, jac=lambda x: P @ x + prob.c, method="SLSQP",
                   bounds=list(zip(np.where(np.isfinite(lb), lb, None), np.where(np.isfinite(ub), ub, None))),
                   constraints=cons, options={"maxiter": 500, "ftol": 1e-14})
    if res.success and (not prob.H.shape[0] or np.max(prob.H @ res.x - prob.rhs) <= 1e-8):
        return OPTIMAL, prob.objective(res.x), res.x
    # interior-point fallback (slow, robust): scipy trust-constr
    from scipy.optimize import Bounds as _B, LinearConstraint as _LC
    cons = [_LC(prob.H, -np.inf, prob.rhs)] if prob.H.shape[0] else []
    res = minimize(lambda x: 0.5 * x @ P @ x + prob.c @ x, x0, jac=lambda x: P @ x + prob.c, hess=lambda x: P,
                   method="trust-constr", constraints=cons, bounds=_B(lb, ub),
                   options={"gtol": 1e-11, "xtol": 1e-13, "maxiter": 3000})
    if res.status in (1, 2) and (not prob.H.shape[0] or np.max(prob.H @ res.x - prob.rhs) <= 1e-7):
        return OPTIMAL, prob.objective(res.x), res.x
    return LIMIT, np.inf, None


def _solve_qp_highs(prob, lb=None, ub=None):
    """HiGHS' QP solver through the private scipy binding."""
    from scipy.optimize._highspy import _core as hs
    lb = prob.lb if lb is None else lb
    ub = prob.ub if ub is None else ub
    n, m = prob.n, prob.H.shape[0]
    inf = hs.kHighsInf
    h = hs._Highs()
    h.setOptionValue("output_flag", False)
    model = hs.HighsModel()
    lp = model.lp_
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = prob.c.astype(float)
    lp.col_lower_ = np.where(np.isfinite(lb), lb, -inf)
    lp.col_upper_ = np.where(np.isfinite(ub), ub, inf)
    lp.row_lower_ = np.full(m, -inf)
    lp.row_upper_ = prob.rhs.astype(float)
    lp.a_matrix_.format_ = hs.MatrixFormat.kRowwise
    nz = prob.H != 0
    lp.a_matrix_.start_ = np.concatenate([[0], np.cumsum(nz.sum(axis=1))]).astype(np.int32)
    lp.a_matrix_.index_ = np.nonzero(nz)[1].astype(np.int32)
    lp.a_matrix_.value_ = prob.H[nz].astype(float)
    if prob.P is not None and np.any(prob.P):
        P = 0.5 * (prob.P + prob.P.T)
        tri = np.triu(np.ones_like(P, dtype=bool)) & (P != 0)   # upper triangle, column-wise == lower row-wise
        hess = model.hessian_
        hess.dim_ = n
        hess.format_ = hs.HessianFormat.kTriangular
        # HiGHS wants the lower triangle column-wise; by symmetry use P^T's upper triangle row by row
        Pl = np.tril(P)
        nzl = Pl.T != 0  # rows of Pl.T are columns of Pl
        hess.start_ = np.concatenate([[0], np.cumsum(nzl.sum(axis=1))]).astype(np.int32)
        hess.index_ = np.nonzero(nzl)[1].astype(np.int32)
        hess.value_ = Pl.T[nzl].astype(float)
    h.passModel(model)
    h.run()
    ms = h.getModelStatus()
    if ms == hs.HighsModelStatus.kOptimal:
        sol = h.getSolution()
        v = np.array(sol.col_value)
        return OPTIMAL, prob.objective(v), v
    if ms == hs.HighsModelStatus.kInfeasible:
        return INFEASIBLE, np.inf, None
    return LIMIT, np.inf, None


def solve_miqp(prob, tol=1e-9, int_tol=1e-6, max_nodes=200000):
    """Depth-first B&B on HiGHS QP relaxations -> (status, objective, v, nodes)."""
    bin_idx = np.nonzero(prob.is_bin)[0]
    best, best_v, nodes = np.inf, None, 0
    stack = [(prob.lb.copy(), prob.ub.copy())]
    while stack and nodes < max_nodes:
        lb, ub = stack.pop()
        nodes += 1
        st, obj, v = solve_qp(prob, lb, ub)
        if st != OPTIMAL or obj >= best - tol * max(1.0, abs(best)):
            continue
        frac = np.abs(v[bin_idx] - np.round(v[bin_idx]))
        if frac.max(initial=0.0) <= int_tol:
            vv = v.copy()
            vv[bin_idx] = np.round(vv[bin_idx])
            o2, v2 = _continuous_subproblem(prob, vv[bin_idx])
            if o2 < best:
                best, best_v = o2, v2
            continue
        j = bin_idx[int(np.argmax(frac))]
        for val in ((0.0, 1.0) if v[j] >= 0.5 else (1.0, 0.0)):  # pushed last is explored first
            l2, u2 = lb.copy(), ub.copy()
            l2[j] = u2[j] = val
            stack.append((l2, u2))
    if best_v is None:
        return (INFEASIBLE if not stack else LIMIT), np.inf, None, nodes
    return (OPTIMAL if not stack else LIMIT), best, best_v, nodes


def solve(prob, **kw):
    """Dispatch on the cost type -> (status, objective, v)."""
    if prob.P is None or not np.any(prob.P):
        return solve_milp(prob, **kw)
    st, obj, v, _ = solve_miqp(prob)
    return st, obj, v
