"""Generates tests/golden/assembly_*.npz: the optimisation problem the UNMODIFIED reference assembles for one control
instant -- MpcController.set_std_obj_atoms / gen_evo_constraints / set_constraints / build (controllers/
mpc_controller.py:76-101, controllers/controller_base.py:411-489, controllers/components/variables.py:189-317,
objective_atoms.py) -- and what its solve / feedback / sim_step_k return, run under oracle/ref_shim.load_controllers():
cvxpy's modelling layer is oracle/mini_cvxpy.py (documented cvxpy semantics, problems recovered by probing), the MILP
backend is HiGHS (scipy) instead of the reference's Gurobi.

Run in the build container only:   python tests/golden/make_golden_assembly.py

Every fixture holds
  in_<M>, nu_l, N_p, Nt, x_k, omega_tilde, atom_keys / atom_<i>, k_neg1_<var>, extra_<j>_{omega_t|omega_scenarios|
  N_tilde}, disable_soft                                    -- the inputs
  var_names, var_dims                                       -- the reference's cvx variables, creation order
  v_of_x [n, n]                                             -- reference's stacked v~ as a function of its variables
  c, c0, affine, G, h, A, b, lb, ub, integrality            -- the problem in the reference's variable order
  points_x [P, n], points_f [P]                             -- objective values at random points (any atom type)
  sol_obj, sol_x, sol_status, fb_<var>                      -- solve() / feedback() results (affine objectives)
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.simplefilter("ignore")

from oracle import ref_shim  # noqa: E402

R = ref_shim.load_controllers()
MAT_NAMES = ("A", "B1", "B2", "B3", "B4", "b5", "C", "D1", "D2", "D3", "D4", "d5",
             "E", "F1", "F2", "F3", "F4", "f5", "G", "Psi")
VARS = ("x", "u", "delta", "z", "omega", "y", "mu", "v")


def run_case(name, mld, N_p, x_k, omega_tilde, atoms, extra=(), disable_soft=False, k_neg1=None, seed=0, solve=True):
    rng = np.random.default_rng(seed)
    Nt = N_p + 1
    ctrl = R.MpcController(model=R.MldSystemModel(mld_numeric=mld), N_p=N_p)
    info = mld.mld_info
    data = dict(N_p=np.array(N_p), Nt=np.array(Nt), nu_l=np.array(int(info.nu_l)), disable_soft=np.array(disable_soft),
                x_k=np.asarray(x_k, dtype=float).reshape(-1), omega_tilde=np.asarray(omega_tilde, dtype=float).reshape(-1),
                atom_keys=np.array(list(atoms)), n_extra=np.array(len(extra)))
    for k in MAT_NAMES:
        data["in_" + k] = np.asarray(mld[k], dtype=float)
    for i, (key, val) in enumerate(atoms.items()):
        data["atom_%d" % i] = np.asarray(val, dtype=float)
    ctrl.set_std_obj_atoms(**atoms)
    if info.nx:
        ctrl.x_k = np.asarray(x_k, dtype=float).reshape(-1, 1)
    if info.nomega:
        ctrl.omega_tilde_k = np.asarray(omega_tilde, dtype=float).reshape(-1, 1)
    if k_neg1:
        ctrl.variables_k_neg1 = {k: np.asarray(v, dtype=float).reshape(-1, 1) for k, v in k_neg1.items()}
        for k, v in k_neg1.items():
            data["k_neg1_" + k] = np.asarray(v, dtype=float).reshape(-1)
    others = []
    for j, ec in enumerate(extra):
        kw = {}
        if "omega_t" in ec:
            kw["omega_tilde_k"] = np.asarray(ec["omega_t"], dtype=float).reshape(-1, 1)
            data["extra_%d_omega_t" % j] = np.asarray(ec["omega_t"], dtype=float).reshape(-1)
        if "omega_scenarios" in ec:
            kw["omega_scenarios_k"] = np.asarray(ec["omega_scenarios"], dtype=float)
            data["extra_%d_omega_scenarios" % j] = np.asarray(ec["omega_scenarios"], dtype=float)
        if "N_tilde" in ec:
            kw["N_tilde"] = int(ec["N_tilde"])
            data["extra_%d_N_tilde" % j] = np.array(int(ec["N_tilde"]))
        others.append(ctrl.gen_evo_constraints(**kw))
    if others:
        ctrl.set_constraints(other_constraints=others)
    ctrl.build(disable_soft_constraints=disable_soft)
    prob = ctrl.problem
    cf = prob.canonical_form()
    variables, n = cf["variables"], cf["n"]
    data["var_names"] = np.array([v.name() for v in variables])
    data["var_dims"] = np.array([v.size // Nt for v in variables])
    # the reference's own stacked decision vector v~ (variables.py:233-241) as a function of its cvx variables
    v_expr = ctrl.variables.v.var_N_tilde
    v_of_x = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        prob._set_x(variables, e)
        v_of_x[:, j] = np.asarray(v_expr.value, dtype=float).reshape(-1)
    data["v_of_x"] = v_of_x
    for key in ("c", "G", "h", "A", "b", "lb", "ub", "integrality"):
        data[key] = np.asarray(cf[key])
    data["c0"], data["affine"] = np.array(cf["c0"]), np.array(cf["objective_is_affine"])
    P = 12
    pts = rng.uniform(-1.0, 2.0, size=(P, n))
    pts[:, cf["integrality"]] = rng.integers(0, 2, size=(P, int(cf["integrality"].sum())))
    nonneg = (cf["lb"] == 0.0) & ~cf["integrality"]
    pts[:, nonneg] = np.abs(pts[:, nonneg])                 # inside the variable bounds (slacks >= 0)
    data["points_x"] = pts
    data["points_f"] = np.array([prob.objective_at(p, variables) for p in pts])
    if solve and cf["objective_is_affine"]:
        if k_neg1:
            # solve(k) reloads the previous step from the sim log's entry k-1 (controllers/controller_base.py:500-501);
            # without that entry it takes lsim_k() of nothing, i.e. zeros
            ctrl.sim_log.set_sim_k(k=-1, **{k: np.asarray(v, dtype=float).reshape(-1, 1) for k, v in k_neg1.items()})
        obj = ctrl.solve(k=0)
        data["sol_obj"], data["sol_status"] = np.array(obj), np.array(prob.status)
        data["sol_x"] = np.concatenate([np.asarray(v.value, dtype=float).reshape(-1, order="F") for v in variables])
        fb = ctrl.variables_k
        for var in VARS:
            data["fb_" + var] = np.asarray(fb[var], dtype=float).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "assembly_%s.npz" % name), **data)
    print(name, "n =", n, "rows =", cf["G"].shape[0], "+", cf["A"].shape[0], "affine =", bool(cf["objective_is_affine"]),
          "obj =", data.get("sol_obj"))
    return ctrl


def closed_loop_case(name, N_p, steps, seed):
    """the reference's own closed loop: feedback -> sim_step_k with the const_heat=False model re-evaluated at the
    current temperature (what DewhAgentMpc.sim_step_k does, micro_grid_agents.py:389-408)"""
    rng = np.random.default_rng(seed)
    Nt = N_p + 1
    p = dict(R.params.dewh_param_struct)
    p["T_h_max"] = 65.0
    control = R.models.DewhModel(param_struct=p, const_heat=True)
    sim = R.models.DewhModel(param_struct=p, const_heat=False)
    ctrl = R.MpcController(model=control, N_p=N_p)
    price = rng.uniform(0.5, 3.0, steps + Nt) * 1e-4 * p["P_h_Nom"]
    demand = rng.uniform(0.0, 0.02, steps + Nt) * (rng.random(steps + Nt) < 0.35)
    x = 51.0
    log = dict(x=[x], u=[], obj=[], mu_hat=[], A=[], cons=[])
    for k in range(steps):
        q_u = price[k:k + Nt]
        ctrl.set_std_obj_atoms(q_u=q_u, q_mu=[10.0 * q_u.sum(), 1.0 * q_u.sum()])
        ctrl.x_k = x
        ctrl.omega_tilde_k = demand[k:k + Nt].reshape(-1, 1)
        ctrl.build()
        obj = ctrl.solve(k=k)
        fb = ctrl.variables_k
        T_h = x if x > p["T_w"] else p["T_w"] + 0.1
        mld_sim = sim.get_mld_numeric(param_struct_subset=dict(T_h=T_h, D_h=float(demand[k])))
        step = ctrl.sim_step_k(k=k, x_k=x, u_k=fb.u, omega_k=float(demand[k]), mld_numeric_k=mld_sim)
        x = float(np.asarray(step.x_k1).reshape(-1)[0])
        log["x"].append(x); log["u"].append(float(np.asarray(fb.u).reshape(-1)[0])); log["obj"].append(float(obj))
        log["mu_hat"].append(np.asarray(fb.mu, dtype=float).reshape(-1)); log["A"].append(float(mld_sim.A[0, 0]))
        log["cons"].append(np.asarray(step.cons).reshape(-1))
    data = {k: np.array(v) for k, v in log.items()}
    data.update(N_p=np.array(N_p), steps=np.array(steps), price=price, demand=demand,
                params=np.array([p[k] for k in ("C_w", "A_h", "U_h", "m_h", "T_w", "T_inf", "P_h_Nom", "T_h_min",
                                               "T_h_max", "T_h_Nom", "ts")]))
    np.savez_compressed(os.path.join(HERE, "assembly_%s.npz" % name), **data)
    print(name, "u =", data["u"], "x_end =", data["x"][-1])


def update_sequence_case(dewh, rng):
    """MpcController.update_std_obj_atoms (controllers/mpc_controller.py:47-58, objective_atoms.py:498-521): weights
    merged into existing atoms -- terminal weight next to a horizon weight, N_p weight over it, an all-zero weight
    deleting the atom -- and the cost vector after every call"""
    N_p, Nt = 6, 7
    q_u = rng.uniform(1, 2, Nt)
    steps = [("set", dict(q_u=q_u, q_mu=np.array([5.0, 1.0]))),
             ("update", dict(q_u_f=np.array([9.0]), q_mu=np.array([7.0, 2.0]))),
             ("update", dict(q_u_N_p=np.array([0.5]))),
             ("update", dict(q_mu=np.zeros(2))),
             ("update", dict(q_x=np.array([0.25]), q_u=2.0 * q_u)),
             ("set", dict(q_mu=np.array([3.0, 4.0])))]
    ctrl = R.MpcController(model=R.MldSystemModel(mld_numeric=dewh), N_p=N_p)
    ctrl.x_k = 55.0
    ctrl.omega_tilde_k = rng.uniform(0, 0.01, (Nt, 1))
    data = dict(N_p=np.array(N_p), n_steps=np.array(len(steps)), x_k=np.array([55.0]),
                omega_tilde=np.asarray(ctrl.omega_tilde_k.value if hasattr(ctrl.omega_tilde_k, "value")
                                       else ctrl.omega_tilde_k, dtype=float).reshape(-1))
    for k in MAT_NAMES:
        data["in_" + k] = np.asarray(dewh[k], dtype=float)
    for i, (how, kw) in enumerate(steps):
        getattr(ctrl, "set_std_obj_atoms" if how == "set" else "update_std_obj_atoms")(**kw)
        ctrl.build()
        cf = ctrl.problem.canonical_form()
        c = cf["c"]                                        # reference order: U (Nt) then Mu (2 Nt, step-major)
        c_v = np.zeros(3 * Nt)
        c_v[0::3], c_v[1::3], c_v[2::3] = c[:Nt], c[Nt:].reshape(Nt, 2)[:, 0], c[Nt:].reshape(Nt, 2)[:, 1]
        data["how_%d" % i] = np.array(how)
        data["keys_%d" % i] = np.array(list(kw))
        for j, val in enumerate(kw.values()):
            data["val_%d_%d" % (i, j)] = np.asarray(val, dtype=float)
        data["c_v_%d" % i], data["c0_%d" % i] = c_v, np.array(cf["c0"])
    np.savez_compressed(os.path.join(HERE, "assembly_update_sequence.npz"), **data)
    print("update_sequence", len(steps), "calls")


def random_mld(rng):
    """2 states, a continuous and a binary input, one delta, one z, 2 disturbances, 2 outputs, 5 rows, 3 slacks"""
    nx, nu, nd, nz, nw, ny, nc, nmu = 2, 2, 1, 1, 2, 2, 5, 3
    g = lambda *s: rng.standard_normal(s)   # noqa: E731
    Psi = np.zeros((nc, nmu))
    Psi[[0, 1, 2], [0, 1, 2]] = -1.0
    return R.MldModel(A=0.5 * g(nx, nx), B1=g(nx, nu), B2=g(nx, nd), B3=g(nx, nz), B4=g(nx, nw), b5=g(nx, 1),
                      C=g(ny, nx), D1=g(ny, nu), D2=g(ny, nd), D3=g(ny, nz), D4=g(ny, nw), d5=g(ny, 1),
                      E=g(nc, nx), F1=g(nc, nu), F2=g(nc, nd), F3=g(nc, nz), F4=g(nc, nw), f5=5.0 + np.abs(g(nc, 1)),
                      G=g(nc, ny), Psi=Psi, nu_l=1)


def main():
    rng = np.random.default_rng(77)
    dewh = R.models.DewhModel(const_heat=True).mld_numeric

    def dewh_inputs(N_p, cold=False):
        Nt = N_p + 1
        q_u = rng.uniform(0.5, 3.0, Nt) * 1e-4 * 3000.0
        w = rng.uniform(0.0, 0.02, Nt) * (rng.random(Nt) < 0.35)
        return (48.5 if cold else 56.0), w, dict(q_u=q_u, q_mu=np.array([10.0 * q_u.sum(), 1.0 * q_u.sum()]))

    x, w, atoms = dewh_inputs(8)
    run_case("dewh_N8_linear", dewh, 8, x, w, atoms, seed=1)
    x, w, atoms = dewh_inputs(48, cold=True)
    run_case("dewh_N48_linear", dewh, 48, x, w, atoms, seed=2)
    x, w, atoms = dewh_inputs(8)
    run_case("dewh_N8_hard", dewh, 8, x, w, atoms, disable_soft=True, seed=3)
    x, w, atoms = dewh_inputs(12)
    sc = rng.uniform(0.0, 0.02, (13, 6)) * (rng.random((13, 6)) < 0.35)
    lo, hi = 0.2 * w, w + 0.004
    run_case("dewh_N12_scenarios_minmax", dewh, 12, x, w, atoms, seed=4,
             extra=[dict(omega_scenarios=sc, N_tilde=5), dict(omega_scenarios=sc), dict(omega_t=lo), dict(omega_t=hi)])
    x, w, atoms = dewh_inputs(6)
    atoms.update(Q_x=np.array([[0.3]]), q_L22_y_N_p=np.array([0.2]), Q_x_f=np.array([[2.0]]), q_L1_du=np.array([0.7]),
                 q_L1_x=np.array([0.05]), Q_mu=np.diag([0.4, 0.1]), q_x_N_p=np.array([0.01]), q_Linf_mu=np.array([0.3, 0.2]))
    run_case("dewh_N6_all_atoms", dewh, 6, x, w, atoms, k_neg1=dict(u=[1.0]), seed=5, solve=False)
    x, w, atoms = dewh_inputs(6)
    # the non-linear atoms the stage-DP kernels take (per-step diagonal weights, no rate form, no Linf)
    atoms.update(Q_x=np.array([[0.3]]), q_L22_y_N_p=np.array([0.2]), Q_x_f=np.array([[2.0]]), q_L1_x=np.array([0.05]),
                 Q_mu=np.diag([0.4, 0.1]), q_L1_u=np.array([0.3]), q_L22_u=np.array([0.5]), q_L1_mu=np.array([0.2, 0.6]),
                 Q_L1_y=np.array([[0.15]]), q_y=rng.uniform(-0.01, 0.01, 7))
    run_case("dewh_N6_stage_dp_atoms", dewh, 6, x, w, atoms, seed=10, solve=False)
    x, w, atoms = dewh_inputs(6)
    atoms.update(q_du=np.array([0.2]), q_y=rng.uniform(-0.01, 0.01, 7), q_x_f=np.array([-0.02]), q_v_N_p=np.array([0.0, 0.1, 0.2]))
    run_case("dewh_N6_linear_xy_rate", dewh, 6, x, w, atoms, k_neg1=dict(u=[1.0]), seed=6)

    grid = R.models.GridModel(num_devices=3).mld_numeric
    price = rng.uniform(0.5, 3.0, 9)
    run_case("grid_N8", grid, 8, np.zeros(0), rng.uniform(-4000.0, 6000.0, 27), dict(q_z=price), seed=7)

    mld = random_mld(rng)
    atoms = dict(q_u=np.array([0.3, 1.0]), q_delta=np.array([0.5]), q_z=np.array([0.2]), q_mu=np.array([50.0, 60.0, 70.0]),
                 q_y=np.array([0.1, -0.1]))
    run_case("rand_N5_linear", mld, 5, rng.standard_normal(2), rng.standard_normal(12), atoms, seed=8)
    atoms = dict(Q_x=np.array([[1.0, 0.2], [0.2, 0.5]]), q_L22_u=np.array([0.3, 0.1]), q_L1_y=np.array([0.4, 0.2]),
                 Q_dz=np.array([[0.6]]), q_mu=np.array([50.0, 60.0, 70.0]), Q_L1_x=np.array([[0.3, -0.1], [0.2, 0.4]]))
    run_case("rand_N5_quadratic", mld, 5, rng.standard_normal(2), rng.standard_normal(12), atoms,
             k_neg1=dict(z=[0.7]), seed=9, solve=False)

    closed_loop_case("dewh_closed_loop", N_p=10, steps=14, seed=21)
    update_sequence_case(dewh, rng)


if __name__ == "__main__":
    main()
