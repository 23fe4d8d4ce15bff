"""CPU: the oracle's problem assembly (oracle/assemble.py), solve and closed loop against what the UNMODIFIED reference
builds and returns for the same inputs -- MpcController.set_std_obj_atoms / gen_evo_constraints / set_constraints /
build / solve / feedback / sim_step_k, run under oracle/ref_shim.load_controllers with oracle/mini_cvxpy.py standing in
for cvxpy's modelling layer and HiGHS for Gurobi (tests/golden/make_golden_assembly.py).

Checked per fixture: the decision-vector layout (the reference's own stacked v~ as a function of its cvx variables),
every constraint row and right-hand side (standard, scenario row-min, reduced horizon, min / max sets), the soft-
constraint switch, bounds and integrality, the cost vector and constant for affine objectives, the objective VALUE at
random points for every atom type (Linear / Quadratic / L22 / L1 / Linf, matrix and vector weights, rate atoms with the
previous step, N_p / terminal suffixes), the optimal objective, and the first-step values ``feedback`` returns."""
import glob
import os

import numpy as np
import pytest

from oracle import assemble as oa
from oracle import condense as oc
from oracle import lsim as ol
from oracle import mld as omld
from oracle import solve as osv

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[len("assembly_"):-4] for p in glob.glob(os.path.join(GOLDEN, "assembly_*.npz"))
               if "closed_loop" not in p and "update_sequence" not in p)
MAT_NAMES = ("A", "B1", "B2", "B3", "B4", "b5", "C", "D1", "D2", "D3", "D4", "d5",
             "E", "F1", "F2", "F3", "F4", "f5", "G", "Psi")
REF_VAR_NAMES = dict(U_var_N_tilde="u", Delta_var_N_tilde="delta", Z_var_N_tilde="z", Mu_var_N_tilde="mu")


def _load(case):
    z = np.load(os.path.join(GOLDEN, "assembly_%s.npz" % case))
    g = {k: z[k] for k in z.files}
    mats = {k: g["in_" + k] for k in MAT_NAMES if g["in_" + k].size}
    full, dims, vt = omld.complete(mats, nu_l=int(g["nu_l"]))
    Nt, N_p = int(g["Nt"]), int(g["N_p"])
    evo = oc.condense(full, dims, Nt)
    atoms = {str(k): g["atom_%d" % i] for i, k in enumerate(g["atom_keys"])}
    extra = []
    for j in range(int(g["n_extra"])):
        ec = {}
        for key in ("omega_t", "omega_scenarios"):
            if "extra_%d_%s" % (j, key) in g:
                ec[key] = g["extra_%d_%s" % (j, key)]
        if "extra_%d_N_tilde" % j in g:
            ec["N_tilde"] = int(g["extra_%d_N_tilde" % j])
        extra.append(ec)
    k_neg1 = {k[len("k_neg1_"):]: v for k, v in g.items() if k.startswith("k_neg1_")}
    prob = oa.build_problem(evo, dims, vt, Nt, g["x_k"], g["omega_tilde"], atoms=atoms, N_p=N_p,
                            disable_soft_constraints=bool(g["disable_soft"]), extra_constraints=extra,
                            var_k_neg1=k_neg1)
    return g, prob, evo, dims, Nt


def _objective_in_v(prob, v):
    """objective as a function of v~ alone: epigraph columns (L1 atoms) take their minimal feasible value"""
    n_v = prob.n_v
    full = np.zeros(prob.n)
    full[:n_v] = v
    for j in range(n_v, prob.n):
        rows = np.nonzero(prob.H[:, j] < 0)[0]                     # (A v + a0) - t <= 0  and  -(A v + a0) - t <= 0
        full[j] = max(0.0, max(float(prob.H[r, :n_v] @ v - prob.rhs[r]) for r in rows))
    return prob.objective(full)


@pytest.mark.parametrize("case", CASES)
def test_layout_rows_bounds_and_cost(case):
    g, prob, evo, dims, Nt = _load(case)
    n = dims["nv"] * Nt
    P = g["v_of_x"]                                                # v~ = P x_ref
    assert P.shape == (n, n)
    # layout: the reference's variables are U, Delta, Z, Mu stacks, each step-major; v~ interleaves them per step
    idx = oa.var_layout(dims, Nt)
    expect = np.zeros((n, n))
    off = 0
    for name, dim in zip(g["var_names"], g["var_dims"]):
        var = REF_VAR_NAMES[str(name)]
        assert int(dim) == dims["n" + var]
        cols = off + np.arange(int(dim) * Nt)
        expect[idx[var], cols] = 1.0
        off += int(dim) * Nt
    assert off == n and np.array_equal(P, expect)
    # constraint rows, in the reference's order (standard set, then the other sets)
    m = g["G"].shape[0]
    assert prob.H.shape[0] >= m
    G_v = g["G"] @ P.T
    scale = max(1.0, float(np.abs(G_v).max()) if m else 1.0)
    np.testing.assert_allclose(prob.H[:m, :n], G_v, rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(prob.rhs[:m], g["h"], rtol=1e-12, atol=1e-9 * max(1.0, float(np.abs(g["h"]).max()) if m else 1.0))
    if prob.n == n:
        assert prob.H.shape[0] == m                                # no rows the reference does not have
    # bounds, integrality, soft-constraint switch
    x_to_v = P.argmax(axis=0)                                      # position of each reference entry in v~
    assert np.array_equal(prob.is_bin[:n][x_to_v], g["integrality"].astype(bool))
    lb, ub = prob.lb[:n][x_to_v].copy(), prob.ub[:n][x_to_v].copy()
    if bool(g["disable_soft"]):
        mu_x = np.nonzero(np.isin(x_to_v, idx["mu"]))[0]
        A, b = g["A"], g["b"]                                      # reference: mu == 0 as equality rows
        assert A.shape[0] == mu_x.size and np.array_equal(np.sort(A.argmax(axis=1)), mu_x)
        assert np.all(A.sum(axis=1) == 1.0) and np.all(b == 0.0)
        assert np.all(lb[mu_x] == 0.0) and np.all(ub[mu_x] == 0.0)  # oracle: the same thing as bounds
        ub[mu_x] = np.inf
    else:
        assert g["A"].shape[0] == 0
    assert np.array_equal(lb, g["lb"]) and np.array_equal(ub, g["ub"])
    # affine objective: cost vector and constant
    if bool(g["affine"]):
        assert prob.P is None and prob.n == n
        cs = max(1.0, float(np.abs(g["c"]).max()))
        np.testing.assert_allclose(prob.c[x_to_v], g["c"], rtol=0, atol=1e-12 * cs)
        assert abs(prob.c0 - float(g["c0"])) <= 1e-10 * max(1.0, abs(float(g["c0"])))


@pytest.mark.parametrize("case", CASES)
def test_objective_values_at_random_points(case):
    g, prob, evo, dims, Nt = _load(case)
    for x, f in zip(g["points_x"], g["points_f"]):
        got = _objective_in_v(prob, g["v_of_x"] @ x)
        assert abs(got - f) <= 1e-10 * max(1.0, abs(f)), (case, got, f)


@pytest.mark.parametrize("case", [c for c in CASES if "sol_obj" in np.load(os.path.join(GOLDEN, "assembly_%s.npz" % c)).files])
def test_solution_and_feedback(case):
    g, prob, evo, dims, Nt = _load(case)
    assert str(g["sol_status"]) == "optimal"
    status, obj, v = osv.solve_milp(prob, polish=True)
    assert status == osv.OPTIMAL
    ref = float(g["sol_obj"])
    assert abs(obj - ref) <= 1e-6 * max(1.0, abs(ref)), (obj, ref)          # BASELINE.json: objectives within 1e-6
    v_ref = g["v_of_x"] @ g["sol_x"]
    assert abs(_objective_in_v(prob, v_ref) - ref) <= 1e-9 * max(1.0, abs(ref))
    assert np.all(prob.H[:, :v_ref.size] @ v_ref <= prob.rhs + 1e-6 * np.maximum(1.0, np.abs(prob.rhs)))
    # feedback(): first-step slice of every variable, from the reference's own solution
    full, first = oa.split_solution(v_ref, evo, dims, Nt, g["x_k"], g["omega_tilde"])
    for var in ("x", "u", "delta", "z", "omega", "y", "mu", "v"):
        ref_k = g["fb_" + var]
        assert first[var].shape == ref_k.shape, var
        np.testing.assert_allclose(first[var], ref_k, rtol=1e-12, atol=1e-9, err_msg=var)


def test_closed_loop_against_the_reference():
    """14 instants of the reference's own loop (feedback -> sim_step_k on the re-parametrised simulation model):
    same inputs, same objective (1e-6), same first control, same temperature trajectory (1e-9)."""
    z = np.load(os.path.join(GOLDEN, "assembly_dewh_closed_loop.npz"))
    keys = ("C_w", "A_h", "U_h", "m_h", "T_w", "T_inf", "P_h_Nom", "T_h_min", "T_h_max", "T_h_Nom", "ts")
    p = dict(zip(keys, z["params"]))
    N_p, steps = int(z["N_p"]), int(z["steps"])
    Nt = N_p + 1
    full, dims, vt = omld.complete(ol.dewh_mld(p, const_heat=True), nu_l=1)
    evo = oc.condense(full, dims, Nt)
    x = float(z["x"][0])
    for k in range(steps):
        q_u = z["price"][k:k + Nt]
        atoms = dict(q_u=q_u, q_mu=np.array([10.0 * q_u.sum(), 1.0 * q_u.sum()]))
        prob = oa.build_problem(evo, dims, vt, Nt, [x], z["demand"][k:k + Nt], atoms=atoms)
        status, obj, v = osv.solve_milp(prob, polish=True)
        assert status == osv.OPTIMAL
        assert abs(obj - z["obj"][k]) <= 1e-6 * max(1.0, abs(z["obj"][k])), k
        u = float(round(v[0]))
        assert u == round(float(z["u"][k])), (k, u, z["u"][k])
        A_sim = ol.dewh_scalars(p, const_heat=False, T_h=x if x > p["T_w"] else p["T_w"] + 0.1, D_h=z["demand"][k])[0]
        assert abs(A_sim - z["A"][k]) <= 1e-12
        x1, _, cons = ol.dewh_sim_step(p, x, u, z["demand"][k])
        assert abs(float(x1) - z["x"][k + 1]) <= 1e-9 * max(1.0, abs(z["x"][k + 1])), k
        assert np.array_equal(np.asarray(cons).astype(bool).ravel(), z["cons"][k].astype(bool))
        x = float(x1)
