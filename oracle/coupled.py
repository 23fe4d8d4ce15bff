"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's CENTRALISED micro-grid problem as one MILP: ``GridAgentMpc.build_grid``
(examples/residential_mg_with_pv_and_dewhs/modelling/micro_grid_agents.py:691-709) hands the grid controller every
device's constraints and objectives, and the grid's omega~ is the stack of the devices' power expressions
(:625-646), so one cvxpy problem holds

    min   sum_k price_k z_k  +  sum_i (device objectives: the DEWH slack penalties q_mu)          (:228-229, :193-198
                                                                            of micro_grid_control_simulation.py)
    s.t.  every device's evolution constraints,
          grid MLD rows per step (micro_grid_models.py:145-168):  F2 delta_k + F3 z_k + G y_k <= f5,
          y_k = sum_i P_h_Nom_i u_i,k + (PV and residential-demand outputs, which are data)

Parity status: PINNED against the unmodified reference's own centralised problem -- GridAgentMpc.build_grid /
solve_grid_mpc / sim_step_k with four heaters, PV and demand, two controllers, three instants, run in the build
container under oracle/ref_shim.load_controllers (cvxpy's modelling layer = oracle/mini_cvxpy.py, HiGHS for Gurobi);
vectors tests/golden/microgrid_loop.npz (tests/golden/make_golden_microgrid.py), checker
tests/test_oracle_microgrid_pinned.py: optimal objectives to 1e-6, the reference's plan re-priced in this model, device
and grid simulation steps.  Solved with HiGHS (oracle.solve.solve_milp).
"""
import numpy as np

from .assemble import Problem
from .lsim import grid_mld


def build_coupled_problem(agent_problems, P_nom, p_other, price, grid_params):
    """agent_problems: per-DEWH oracle Problems whose linear cost holds ONLY the device objective (q_mu; no energy
    price on u); P_nom [N_h]; p_other [Nt] = sum of the uncontrolled devices' outputs per step (PV negative);
    price [Nt].  Decision vector: [v_1 .. v_Nh, delta_0, z_0, delta_1, z_1, ...].
    -> (Problem, list of per-agent column offsets, offset of the grid block)"""
    Nt = len(price)
    n_agents = sum(p.n for p in agent_problems)
    prob = Problem(n_agents + 2 * Nt)
    offs, o, r0 = [], 0, 0
    rows = sum(p.H.shape[0] for p in agent_problems)
    H = np.zeros((rows + 6 * Nt, prob.n))
    rhs = np.zeros(rows + 6 * Nt)
    for p in agent_problems:
        assert p.P is None
        offs.append(o)
        m = p.H.shape[0]
        H[r0:r0 + m, o:o + p.n] = p.H
        rhs[r0:r0 + m] = p.rhs
        prob.c[o:o + p.n] = p.c
        prob.lb[o:o + p.n], prob.ub[o:o + p.n], prob.is_bin[o:o + p.n] = p.lb, p.ub, p.is_bin
        prob.c0 += p.c0
        o += p.n
        r0 += m
    g = {k: np.asarray(v, dtype=float) for k, v in grid_mld(grid_params, 1).items()}
    F2, F3, G, f5 = g["F2"].ravel(), g["F3"].ravel(), g["G"].ravel(), g["f5"].ravel()
    for k in range(Nt):
        jd, jz = n_agents + 2 * k, n_agents + 2 * k + 1
        prob.is_bin[jd] = True
        prob.lb[jd], prob.ub[jd] = 0.0, 1.0
        prob.c[jz] = price[k]
        rr = slice(r0 + 6 * k, r0 + 6 * k + 6)
        H[rr, jd] = F2
        H[rr, jz] = F3
        for i, p in enumerate(agent_problems):
            nv = p.n // Nt
            u_cols = offs[i] + k * nv + np.flatnonzero(p.is_bin[k * nv:(k + 1) * nv])
            assert u_cols.size == 1, "one binary input per DEWH step"
            H[rr, u_cols[0]] = G * P_nom[i]
        rhs[rr] = f5 - G * p_other[k]
    prob.H, prob.rhs = H, rhs
    return prob, offs, n_agents


def coupled_cost(agent_problems, U, P_nom, p_other, price):
    """True centralised cost of given binary plans U [N_h, Nt]: energy import at the price plus every device's
    minimal penalty with its binaries fixed."""
    from .solve import _continuous_subproblem
    pen = 0.0
    for p, u in zip(agent_problems, U):
        obj, _ = _continuous_subproblem(p, np.asarray(u, dtype=float))
        pen += obj
    y = np.asarray(P_nom) @ np.asarray(U) + np.asarray(p_other)
    return float(np.sum(np.asarray(price) * np.maximum(0.0, y)) + pen)
