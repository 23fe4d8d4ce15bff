"""GPU parity tests of the exact stage-DP solve (csrc/stage_dp.cu, hmpc_stage_dp_solve_f64) on the whole problem
class it claims -- scalar state, 1..3 binary inputs/deltas, output rows through G/C/D, soft AND hard rows, fixed
binaries, negative costs -- against (a) the oracle (numpy condensing + assembly + HiGHS / enumeration) and (b) the
general branch-and-cut kernel on the condensed problem.  Objectives 1e-6 relative, decisions exact (random
continuous costs make the optimum unique)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_scalar_mld(rng, B, nu, ndelta, nc, ny, soft, hard_rows=0):
    """Batch of random MLDs of the stage-DP class (numpy, float64): name -> [B, r, c]."""
    m = {}
    m["A"] = rng.uniform(0.93, 1.03, size=(B, 1, 1))
    m["B1"] = rng.uniform(0.5, 3.0, size=(B, 1, nu)) * rng.choice([-1.0, 1.0], size=(B, 1, nu), p=[0.25, 0.75])
    if ndelta:
        m["B2"] = rng.uniform(-2.0, 2.0, size=(B, 1, ndelta))
    m["B4"] = rng.uniform(-1.0, 1.0, size=(B, 1, 1))
    m["b5"] = rng.uniform(-0.3, 0.3, size=(B, 1, 1))
    m["C"] = rng.uniform(0.5, 1.5, size=(B, ny, 1))
    m["D1"] = rng.uniform(-0.2, 0.2, size=(B, ny, nu))
    m["D4"] = rng.uniform(-0.2, 0.2, size=(B, ny, 1))
    E = rng.uniform(0.5, 1.5, size=(B, nc, 1)) * np.where(np.arange(nc) % 2 == 0, 1.0, -1.0)[None, :, None]
    m["E"] = E
    m["F1"] = rng.uniform(-0.3, 0.3, size=(B, nc, nu))
    if ndelta:
        m["F2"] = rng.uniform(-0.3, 0.3, size=(B, nc, ndelta))
    m["F4"] = rng.uniform(-0.2, 0.2, size=(B, nc, 1))
    m["G"] = rng.uniform(-0.3, 0.3, size=(B, nc, ny)) * (rng.random((B, nc, ny)) < 0.5)
    width = rng.uniform(2.0, 6.0, size=(B, 1, 1))
    m["f5"] = np.where(np.arange(nc)[None, :, None] % 2 == 0, width, 0.5 * width) * rng.uniform(0.8, 1.2, size=(B, nc, 1))
    if soft:
        d = rng.uniform(0.5, 2.0, size=(B, nc))
        d[:, nc - hard_rows:] = 0.0          # rows without a slack: hard constraints
        m["Psi"] = -np.einsum("bi,ij->bij", d, np.eye(nc))
    return m


def solve_both(m, Nt, nu, ndelta, x0, omega, cost, dev, lb=None, ub=None, cells=None):
    import torch
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    out = {}
    for solver in ("stage_dp", "bnc"):
        kw = dict(dp_opts=cabi.stage_dp_default_opts(cells=cells)) if cells else {}
        bm = BatchMpc(m, Nt - 1, Nt, nu_l=nu, device=dev, solver=solver, **kw)
        if lb is not None:
            bm.lb_v, bm.ub_v = lb.copy(), ub.copy()
        bm.build()
        res = bm.solve(x0, omega, cost_v=cost)
        torch.cuda.synchronize()
        out[solver] = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items()}
        out[solver]["bm"] = bm
    return out


def oracle_problem(m, b, Nt, nu, x0, omega, cost):
    from oracle import mld as omld, condense as oc, assemble as oa
    full, d, vt = omld.complete({k: v[b] for k, v in m.items()}, nu_l=nu)
    evo = oc.condense(full, d, Nt)
    prob = oa.build_problem(evo, d, vt, Nt, x0[b], omega[b], atoms=None)
    prob.c = cost[b].copy()
    return prob


CASES = [  # nu, ndelta, nc, ny, soft, hard_rows, Nt
    (1, 0, 2, 1, True, 0, 12),
    (1, 1, 2, 1, True, 0, 10),
    (2, 0, 3, 2, True, 1, 9),
    (1, 0, 2, 1, False, 0, 12),
    (2, 1, 4, 1, True, 2, 6),
]


@pytest.mark.parametrize("case", CASES)
def test_stage_dp_class_vs_oracle_and_bnc(case, cuda_device):
    from oracle import solve as osv
    nu, ndelta, nc, ny, soft, hard_rows, Nt = case
    rng = np.random.default_rng([nu, ndelta, nc, Nt, 77])
    B = 12
    m = random_scalar_mld(rng, B, nu, ndelta, nc, ny, soft, hard_rows)
    nb, nmu = nu + ndelta, (nc if soft else 0)
    nv = nb + nmu
    x0 = rng.uniform(-1.0, 1.0, size=(B, 1))
    omega = rng.uniform(-1.0, 1.0, size=(B, Nt))
    cost = np.zeros((B, Nt, nv))
    cost[:, :, :nb] = rng.uniform(-0.3, 1.0, size=(B, Nt, nb))          # some negative action costs
    cost[:, :, nb:] = rng.uniform(2.0, 20.0, size=(B, 1, nmu))
    cost = cost.reshape(B, -1)
    out = solve_both(m, Nt, nu, ndelta, x0, omega, cost, cuda_device)
    dp, bnc = out["stage_dp"], out["bnc"]
    assert dp["solver"] == "stage_dp" and bnc["solver"] == "bnc"
    isb = out["stage_dp"]["bm"].is_bin_v.astype(bool)
    n_feasible = 0
    for b in range(B):
        prob = oracle_problem(m, b, Nt, nu, x0, omega, cost)
        if prob.is_bin.sum() <= 8 or b == 0 and prob.is_bin.sum() <= 12:
            st, obj, v, second = osv.solve_enumerate(prob)      # independent of HiGHS' branch and bound
        else:
            st, obj, v = osv.solve_milp(prob, polish=True)
            second = np.inf
        if st == osv.INFEASIBLE:
            assert dp["status"][b] == 1, (b, dp["status"][b])
            continue
        n_feasible += 1
        assert dp["status"][b] == 0, (b, dp["status"][b])
        assert abs(dp["obj"][b] - obj) <= 1e-6 * max(1.0, abs(obj)), (b, dp["obj"][b], obj)
        if second - obj > 1e-6:
            assert np.array_equal(np.round(dp["v"][b][isb]), np.round(v[isb])), b
        # the point the kernel returns is feasible for the condensed problem and its objective is c'v
        assert np.all(prob.H @ dp["v"][b] <= prob.rhs + 1e-7)
        assert abs(cost[b] @ dp["v"][b] - dp["obj"][b]) <= 1e-9 * max(1.0, abs(obj))
        if bnc["status"][b] == 0:
            assert abs(bnc["obj"][b] - dp["obj"][b]) <= 1e-6 * max(1.0, abs(obj)), (b, bnc["obj"][b], dp["obj"][b])
    assert n_feasible >= B // 2


def test_stage_dp_fixed_binaries_and_disabled_slack(cuda_device):
    """lb/ub fix some inputs (reference: a Parameter-pinned binary), mu forced to 0 (disable_soft_constraints,
    controllers/controller_base.py:467-472) turns every row hard."""
    from oracle import solve as osv
    rng = np.random.default_rng(5)
    B, Nt, nu, nc = 8, 12, 1, 2
    m = random_scalar_mld(rng, B, nu, 0, nc, 1, True)
    m["f5"] = m["f5"] * 4.0      # wide box so that the hard version stays feasible for most agents
    nv = nu + nc
    x0 = rng.uniform(-0.5, 0.5, size=(B, 1))
    omega = rng.uniform(-0.5, 0.5, size=(B, Nt))
    cost = np.zeros((B, Nt, nv))
    cost[:, :, 0] = rng.uniform(0.1, 1.0, size=(B, Nt))
    cost[:, :, 1:] = 10.0
    cost = cost.reshape(B, -1)
    lb = np.tile([0.0, 0.0, 0.0], Nt)
    ub = np.tile([1.0, np.inf, np.inf], Nt)
    lb[3 * 2] = 1.0                      # u_2 pinned on
    ub[3 * 5] = 0.0                      # u_5 pinned off
    ub[np.arange(Nt) * 3 + 2] = 0.0      # second slack disabled: row 1 is hard
    out = solve_both(m, Nt, nu, 0, x0, omega, cost, cuda_device, lb=lb, ub=ub)
    dp = out["stage_dp"]
    for b in range(B):
        prob = oracle_problem(m, b, Nt, nu, x0, omega, cost)
        prob.lb, prob.ub = lb.copy(), ub.copy()
        st, obj, v = osv.solve_milp(prob, polish=True)
        if st == osv.INFEASIBLE:
            assert dp["status"][b] == 1
            continue
        assert dp["status"][b] == 0 and abs(dp["obj"][b] - obj) <= 1e-6 * max(1.0, abs(obj))
        assert dp["v"][b][6] == 1.0 and dp["v"][b][15] == 0.0
        assert np.all(dp["v"][b][2::3] == 0.0)


@pytest.mark.parametrize("N_p,cells", [(96, 8192), (48, 512), (48, 16384)])
def test_stage_dp_long_horizon_and_cell_counts(N_p, cells, cuda_device):
    """N_p = 96 (BASELINE config 5's longest horizon) and extreme table resolutions: the answer never depends on
    the number of cells (only the search effort does); checked against HiGHS."""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    B = 6
    wl = syn.dewh_batch(B, N_p, seed=21)
    Nt = wl["Nt"]
    bm = BatchMpc(wl["mats"], N_p, nu_l=1, device=cuda_device, solver="stage_dp",
                  dp_opts=cabi.stage_dp_default_opts(cells=cells))
    bm.build()
    cost = np.zeros((B, Nt, 3))
    cost[:, :, 0] = wl["q_u"]
    cost[:, :, 1:] = wl["q_mu"][:, None, :]
    res = bm.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1))
    obj, v, st = res["obj"].cpu().numpy(), res["v"].cpu().numpy(), res["status"].cpu().numpy()
    assert (st == 0).all()
    for b in range(B if N_p <= 48 else 3):
        full, d, vt = omld.complete({k: mm[b] for k, mm in wl["mats"].items()}, nu_l=1)
        prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, wl["x0"][b], wl["omega"][b],
                                atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b]))
        s, o, vr = osv.solve_milp(prob)
        assert abs(obj[b] - o) <= 1e-6 * max(1.0, abs(o))
        assert np.array_equal(np.round(v[b][prob.is_bin]), np.round(vr[prob.is_bin]))


def test_stage_dp_reports_unsupported_agents(cuda_device):
    """An agent whose Psi couples two rows is outside the class: status 5 for it alone, the others are solved; the
    BatchMpc front door refuses solver='stage_dp' for such a batch and 'auto' picks the general kernel."""
    import torch
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    B, N_p = 4, 12
    wl = syn.dewh_batch(B, N_p, seed=3)
    Nt = wl["Nt"]
    mats = {k: v.copy() for k, v in wl["mats"].items()}
    mats["Psi"][2, 0, 1] = -0.5
    with pytest.raises(ValueError):
        BatchMpc(mats, N_p, nu_l=1, device=cuda_device, solver="stage_dp")
    bm = BatchMpc(mats, N_p, nu_l=1, device=cuda_device, solver="auto")
    assert not bm.stage_dp_ok
    bm.build()
    cost = np.zeros((B, Nt, 3))
    cost[:, :, 0] = wl["q_u"]
    cost[:, :, 1:] = wl["q_mu"][:, None, :]
    res = bm.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1))
    assert res["solver"] == "bnc" and (res["status"].cpu().numpy() == 0).all()
    # raw C-ABI call: per-agent flag
    dev = cuda_device
    x0 = torch.as_tensor(wl["x0"]).to(dev)
    w = torch.as_tensor(wl["omega"]).to(dev)
    rhs = cabi.constraint_rhs(bm.dims, bm.evo, x0, w)
    lb, ub, isb = bm._bounds_dev()
    v, obj, st, stats = cabi.stage_dp_solve(bm.dims, bm.mats, rhs, torch.as_tensor(cost.reshape(B, -1)).to(dev), lb, ub, isb)
    st = st.cpu().numpy()
    assert st[2] == 5 and (np.delete(st, 2) == 0).all()
    assert np.allclose(np.delete(obj.cpu().numpy(), 2), np.delete(res["obj"].cpu().numpy(), 2), rtol=1e-9)


def test_stage_dp_folds_extra_constraint_sets(cuda_device):
    """scenario / min-max controllers add constraint sets over the same rows (examples/.../
    micro_grid_control_simulation.py:200-227); stage-DP folds them into a row-wise min of the right-hand sides and
    must agree with the branch-and-cut kernel that stacks them."""
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    B, N_p = 8, 24
    wl = syn.dewh_batch(B, N_p, seed=4)
    Nt = wl["Nt"]
    rng = np.random.default_rng(8)
    scen = wl["omega"][:, :, None] * rng.uniform(0.5, 1.6, size=(B, Nt, 6))
    cost = np.zeros((B, Nt, 3))
    cost[:, :, 0] = wl["q_u"]
    cost[:, :, 1:] = wl["q_mu"][:, None, :]
    extra = [dict(omega_scenarios_k=scen), dict(omega_tilde_k=wl["omega"] * 1.3, N_tilde=9)]
    res = {}
    for solver in ("stage_dp", "bnc"):
        bm = BatchMpc(wl["mats"], N_p, nu_l=1, device=cuda_device, solver=solver)
        bm.build()
        r = bm.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1), extra_constraints=extra)
        res[solver] = (r["obj"].cpu().numpy(), r["v"].cpu().numpy(), r["status"].cpu().numpy())
    assert (res["stage_dp"][2] == 0).all() and (res["bnc"][2] == 0).all()
    np.testing.assert_allclose(res["stage_dp"][0], res["bnc"][0], rtol=1e-6)
    assert np.array_equal(res["stage_dp"][1][:, ::3], np.round(res["bnc"][1][:, ::3]))


def test_closed_loop_fleet_vs_oracle(cuda_device):
    """BASELINE config 4 in miniature: closed-loop simulation (per-step exact solve + re-parametrised DEWH sim step +
    aggregate power) of a small fleet, every step checked against the oracle (HiGHS solve, numpy sim step)."""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv, lsim as ol
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    B, N_p, steps = 5, 16, 10
    Nt = N_p + 1
    params = [syn.dewh_agent_params(100 + b) for b in range(B)]
    T0 = np.array([syn.dewh_initial_state(100 + b) for b in range(B)])
    demand = np.stack([syn.dhw_demand_profile(steps + Nt, seed=100 + b) for b in range(B)])
    price = syn.price_profile(steps + Nt, seed=3)
    fleet = DewhFleet(params, N_p, device=cuda_device)
    log = {k: v.cpu().numpy() for k, v in fleet.closed_loop(T0, demand, price, steps).items()}
    assert (log["status"] == 0).all()
    T = T0.copy()
    for k in range(steps):
        u0 = np.zeros(B)
        p_agg = np.zeros(Nt)
        for b in range(B):
            p = params[b]
            mats = syn.dewh_scalars(p, const_heat=True)
            m = dict(A=[[mats[0]]], B1=[[mats[1]]], B4=[[mats[2]]], b5=[[mats[3]]], E=[[1.0], [-1.0]], F1=[[0.0], [0.0]],
                     Psi=[[-1.0, 0.0], [0.0, -1.0]], f5=[[p["T_h_max"]], [-p["T_h_min"]]])
            full, d, vt = omld.complete({kk: np.array(vv, dtype=float) for kk, vv in m.items()}, nu_l=1)
            q_u = price[k:k + Nt] * p["P_h_Nom"]
            prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, np.array([T[b]]), demand[b, k:k + Nt],
                                    atoms=dict(q_u=q_u, q_mu=[10.0 * q_u.sum(), 1.0 * q_u.sum()]))
            st, obj, v = osv.solve_milp(prob, polish=True)
            assert abs(log["obj"][k, b] - obj) <= 1e-6 * max(1.0, abs(obj)), (k, b)
            u = np.round(v[prob.is_bin])
            u0[b] = u[0]
            p_agg += p["P_h_Nom"] * u
            assert log["u"][k, b] == u0[b], (k, b)
        np.testing.assert_allclose(log["P_agg"][k], p_agg, rtol=1e-12)
        for b in range(B):
            T[b], _, _ = ol.dewh_sim_step(dict(params[b]), T[b], u0[b], demand[b, k])
        np.testing.assert_allclose(log["T"][k + 1], T, rtol=1e-10)


MIQP_CASES = [
    dict(Q_x=0.5, q_x=-2 * 0.5 * 61.0),                                  # (x - 61)^2 tracking, matrix weight
    dict(q_L22_x=0.6, q_x=-2 * 0.36 * 63.0, q_L22_mu=[1.5, 0.7]),         # ||w.x||^2 (w enters squared) + ||w.mu||^2
    dict(q_L1_x=0.3, q_x=-0.5, Q_mu=[[2.0, 0.0], [0.0, 0.5]]),           # |w x| + mu'W mu
    dict(Q_y=0.4, q_y=-2 * 0.4 * 60.0, q_L22_u=0.2, q_L1_u=0.1),          # outputs; binaries (u^2 = |u| = u)
]


@pytest.mark.parametrize("atoms", MIQP_CASES)
def test_stage_dp_miqp_atoms_vs_oracle(atoms, cuda_device):
    """Quadratic / L22 / L1 atoms (reference: controllers/components/objective_atoms.py:320-363) make the problem an
    MIQP; the stage-DP kernels carry them as convex stage terms.  Checked against exhaustive enumeration with HiGHS
    QPs (oracle.solve.solve_enumerate) on problems with 9 binaries; set-point tracking terms (Q_x with a linear q_x)
    compete with the energy cost so that the optimum is not trivial."""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    B, N_p = 4, 8
    wl = syn.dewh_batch(B, N_p, seed=13)
    Nt = wl["Nt"]
    bm = BatchMpc(wl["mats"], N_p, nu_l=1, device=cuda_device, solver="stage_dp")
    bm.build()
    cost = np.zeros((B, Nt, 3))
    cost[:, :, 0] = wl["q_u"]
    cost[:, :, 1:] = wl["q_mu"][:, None, :]
    scale = float(wl["q_u"].mean())          # state / output weights in units of the energy price
    quad, lin = {}, {}
    scaled = {}
    for key, val in atoms.items():
        wt, atom, var, rate, post = oa.parse_atom_key(key)
        w = np.atleast_1d(np.asarray(val, dtype=float)) * (scale if var in ("x", "y") else 1.0)
        scaled[key] = w if np.ndim(val) else float(w[0])
        if atom == "Linear":
            lin[var] = np.tile(w.reshape(1, 1, -1), (B, Nt, 1)).reshape(B, -1)
            continue
        w_eff = (w ** 2 if atom in ("Quadratic", "L22") else np.abs(w)) if wt == "vector" else np.diag(np.atleast_2d(w))
        name = var + ("2" if atom in ("Quadratic", "L22") else "1")
        quad[name] = quad.get(name, 0) + np.tile(w_eff.reshape(1, 1, -1), (1, Nt, 1))
    res = bm.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1), w_x=lin.get("x"), w_y=lin.get("y"), quad=quad)
    obj, v, st = res["obj"].cpu().numpy(), res["v"].cpu().numpy(), res["status"].cpu().numpy()
    assert (st == 0).all()
    nontrivial = 0
    for b in range(B):
        full, d, vt = omld.complete({k: m[b] for k, m in wl["mats"].items()}, nu_l=1)
        at = dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b])
        for key, val in scaled.items():
            at[key] = at[key] + val if key in at else val
        prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, wl["x0"][b], wl["omega"][b], atoms=at)
        s_, o_, v_, second = osv.solve_enumerate(prob)
        assert abs(obj[b] - o_) <= 1e-6 * max(1.0, abs(o_)), (b, obj[b], o_)
        if second - o_ > 1e-6 * max(1.0, abs(o_)):
            assert np.array_equal(np.round(v[b][::3]), np.round(v_[:3 * Nt][::3])), b
        nontrivial += int(np.round(v_[:3 * Nt][::3]).sum() > 0)
    assert nontrivial >= 1


def test_stage_dp_edge_cases(cuda_device):
    """Degenerate shapes and limits: a single agent, horizons of one and two steps, an MLD without constraint rows,
    every binary pinned, a hard row that cannot be met (infeasible), a node budget of one (not proven), three and
    four binaries per step (8 / 16 actions)."""
    import torch
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    rng = np.random.default_rng(42)
    dev = cuda_device

    def run(m, Nt, nu, ndelta, nc, soft, B, lb=None, ub=None, dp_opts=None, x0=None):
        nb, nmu = nu + ndelta, (nc if soft else 0)
        bm = BatchMpc(m, Nt - 1, Nt, nu_l=nu, device=dev, solver="stage_dp", **({"dp_opts": dp_opts} if dp_opts else {}))
        if lb is not None:
            bm.lb_v, bm.ub_v = lb, ub
        bm.build()
        cost = np.zeros((B, Nt, nb + nmu))
        cost[:, :, :nb] = rng.uniform(0.1, 1.0, size=(B, Nt, nb))
        cost[:, :, nb:] = 10.0
        x0 = rng.uniform(-0.5, 0.5, size=(B, 1)) if x0 is None else x0
        om = rng.uniform(-0.5, 0.5, size=(B, Nt))
        res = bm.solve(x0, om, cost_v=cost.reshape(B, -1))
        return bm, cost.reshape(B, -1), {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items()}

    # (1) B = 1, Nt = 1 and Nt = 2: with positive costs and a wide box the optimum is "all off", objective 0
    for Nt in (1, 2):
        m = random_scalar_mld(rng, 1, 1, 0, 2, 1, True)
        m["f5"] = m["f5"] * 50.0
        bm, cost, r = run(m, Nt, 1, 0, 2, True, 1)
        assert r["status"][0] == 0 and abs(r["obj"][0]) < 1e-12 and not r["v"][0].any()
    # (2) every binary pinned to 1: the cost is the sum of the action costs (+ slack), decisions as pinned
    Nt = 5
    m = random_scalar_mld(rng, 3, 1, 0, 2, 1, True)
    m["f5"] = m["f5"] * 50.0
    lb = np.tile([1.0, 0.0, 0.0], Nt); ub = np.tile([1.0, np.inf, np.inf], Nt)
    bm, cost, r = run(m, Nt, 1, 0, 2, True, 3, lb=lb, ub=ub)
    assert (r["status"] == 0).all() and (r["v"][:, ::3] == 1.0).all()
    np.testing.assert_allclose(r["obj"], (cost * r["v"]).sum(axis=1), rtol=1e-12)
    # (3) a hard row that cannot be met -> infeasible, v = nan
    m = random_scalar_mld(rng, 2, 1, 0, 2, 1, False)
    m["f5"] = -np.abs(m["f5"]) * 100.0
    bm, cost, r = run(m, 6, 1, 0, 2, False, 2)
    assert (r["status"] == 1).all() and np.isinf(r["obj"]).all() and np.isnan(r["v"]).all()
    # (4) node budget of one expansion: an incumbent exists (the greedy dive) but optimality is not proven
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    wl = syn.dewh_batch(4, 24, seed=2)
    bm = BatchMpc(wl["mats"], 24, nu_l=1, device=dev, solver="stage_dp", dp_opts=cabi.stage_dp_default_opts(max_nodes=1))
    bm.build()
    c = np.zeros((4, 25, 3)); c[:, :, 0] = wl["q_u"]; c[:, :, 1:] = wl["q_mu"][:, None, :]
    res = bm.solve(wl["x0"], wl["omega"], cost_v=c.reshape(4, -1))
    assert (res["status"].cpu().numpy() == 2).all() and np.isfinite(res["obj"].cpu().numpy()).all()
    # (5) 3 and 4 binaries per step against the branch-and-cut kernel
    for nu, ndelta in ((2, 1), (2, 2)):
        Nt, nc, B = 5, 3, 6
        m = random_scalar_mld(rng, B, nu, ndelta, nc, 1, True)
        nb = nu + ndelta
        cost = np.zeros((B, Nt, nb + nc)); cost[:, :, :nb] = rng.uniform(-0.2, 1.0, size=(B, Nt, nb)); cost[:, :, nb:] = 8.0
        x0 = rng.uniform(-1, 1, size=(B, 1)); om = rng.uniform(-1, 1, size=(B, Nt))
        out = solve_both(m, Nt, nu, ndelta, x0, om, cost.reshape(B, -1), dev)
        assert (out["stage_dp"]["status"] == 0).all() and (out["bnc"]["status"] == 0).all()
        np.testing.assert_allclose(out["stage_dp"]["obj"], out["bnc"]["obj"], rtol=1e-7, atol=1e-9)
    # (6) no constraint rows at all: the solve falls through to the general kernel (nothing to fold)
    m = dict(A=np.full((2, 1, 1), 0.9), B1=np.ones((2, 1, 1)))
    bm = BatchMpc(m, 3, 4, nu_l=1, device=dev)
    bm.build()
    res = bm.solve(np.zeros((2, 1)), None, cost_v=np.tile([-1.0, 2.0, -3.0, 0.5], (2, 1)))
    assert res["solver"] == "bnc" and np.allclose(res["obj"].cpu().numpy(), -4.0)


def test_host_front_door_falls_back_to_general_kernel(cuda_device):
    """hmpc_mpc_step_host_f64 tries the stage-DP kernels first (the dimensions fit); when an agent reports
    HMPC_SOLVE_UNSUPPORTED (here: a Psi that couples two rows) the whole batch is re-solved by the branch-and-cut
    kernel inside the same call, with H_v condensed on demand."""
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    B, N_p = 3, 10
    wl = syn.dewh_batch(B, N_p, seed=8)
    Nt = wl["Nt"]
    mats = {k: v.copy() for k, v in wl["mats"].items()}
    mats["Psi"][1, 0, 1] = -0.25
    mats["C"] = np.ones((1, 1, 1))
    cost = np.zeros((B, Nt, 3))
    cost[:, :, 0] = wl["q_u"]
    cost[:, :, 1:] = wl["q_mu"][:, None, :]
    ref = BatchMpc({k: v for k, v in mats.items() if k != "C"}, N_p, nu_l=1, device=cuda_device, solver="bnc")
    ref.build()
    r = ref.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1))
    plan = cabi.StepPlan(ref.dims)
    v, obj, st, stats, tm = plan.step(mats, wl["x0"], wl["omega"], cost.reshape(B, -1), ref.lb_v, ref.ub_v, ref.is_bin_v,
                                      recondense=True)
    assert plan.last_solver == "bnc" and (st == 0).all()
    np.testing.assert_allclose(obj, r["obj"].cpu().numpy(), rtol=1e-12)
    # and again without re-condensing (H_v is there now)
    v2, obj2, st2, _, _ = plan.step(None, wl["x0"], wl["omega"], cost.reshape(B, -1), ref.lb_v, ref.ub_v, ref.is_bin_v,
                                    recondense=False)
    assert np.array_equal(obj2, obj)
    plan.close()


def test_stage_dp_fp64_table_prunes_exact_ties(cuda_device):
    """Piecewise-constant tariff (the reference's time-of-use price vector, examples/.../tariff_generator.py): many
    heating patterns cost exactly the same.  With the FP64 table a sequence that ties with the incumbent is pruned at
    mip_rel_gap = 0 (bound == incumbent up to 1e-16); the FP32 table is looser by its rounding and explores them.
    Both return the same proven optimum."""
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.parameters import TOU_LEVELS
    B, N_p = 24, 48
    wl = syn.dewh_batch(B, N_p, seed=1)
    Nt = wl["Nt"]
    hour = (np.arange(Nt) % 96) / 4.0
    level = np.full(Nt, TOU_LEVELS["low_off_peak"])
    level[((hour >= 6) & (hour < 7)) | ((hour >= 10) & (hour < 18)) | ((hour >= 20) & (hour < 22))] = TOU_LEVELS["low_stnd"]
    level[((hour >= 7) & (hour < 10)) | ((hour >= 18) & (hour < 20))] = TOU_LEVELS["low_peak"]
    q_u = (level / 3600.0 / 100.0 / 1000.0 * 900.0)[None, :] * np.array([p["P_h_Nom"] for p in wl["params"]])[:, None]
    cost = np.zeros((B, Nt, 3))
    cost[:, :, 0] = q_u
    cost[:, :, 1] = 10 * q_u.sum(1)[:, None]
    cost[:, :, 2] = q_u.sum(1)[:, None]
    out = {}
    for fp64 in (0, 1):
        bm = BatchMpc(wl["mats"], N_p, nu_l=1, device=cuda_device, solver="stage_dp",
                      dp_opts=cabi.stage_dp_default_opts(table_fp64=fp64, max_nodes=2000000))
        bm.build()
        r = bm.solve(wl["x0"], wl["omega"], cost_v=cost.reshape(B, -1))
        out[fp64] = (r["obj"].cpu().numpy(), r["status"].cpu().numpy(), r["stats"].cpu().numpy()[:, 0])
    assert (out[0][1] == 0).all() and (out[1][1] == 0).all()
    np.testing.assert_allclose(out[1][0], out[0][0], rtol=1e-9)
    assert out[1][2].mean() < 40 and out[1][2].mean() * 5 < out[0][2].mean()


def test_stage_dp_fuzz_against_branch_and_cut(cuda_device):
    """Randomised cross-check of the two solve kernels on 1000 problems of the class with NEGATIVE action costs,
    inputs of both signs, A on both sides of 1, hard rows and pinned binaries.  (Found in round 1: pruning on the
    partial cost alone is invalid when the costs still ahead can be negative -- the search now adds their sum.)"""
    from pyhybridcontrol_b200 import cabi
    B = 200
    shapes = [(1, 0, 3, 1, True, 0, 20), (1, 0, 4, 2, False, 0, 12), (2, 1, 4, 1, True, 2, 6), (2, 0, 3, 2, True, 1, 8),
              (1, 1, 2, 1, True, 0, 10)]
    checked = 0
    for seed, (nu, ndelta, nc, ny, soft, hard, Nt) in enumerate(shapes):
        rng = np.random.default_rng(2000 + seed)
        m = random_scalar_mld(rng, B, nu, ndelta, nc, ny, soft, hard)
        if not soft:
            m["f5"] = m["f5"] * 3.0
        nb, nmu = nu + ndelta, (nc if soft else 0)
        x0 = rng.uniform(-1.5, 1.5, size=(B, 1))
        om = rng.uniform(-1.0, 1.0, size=(B, Nt))
        cost = np.zeros((B, Nt, nb + nmu))
        cost[:, :, :nb] = rng.uniform(-0.4, 1.0, size=(B, Nt, nb))
        cost[:, :, nb:] = rng.uniform(1.0, 30.0, size=(B, 1, nmu))
        lb = np.tile(np.r_[np.zeros(nb), np.zeros(nmu)], Nt)
        ub = np.tile(np.r_[np.ones(nb), np.full(nmu, np.inf)], Nt)
        for pidx in rng.integers(0, Nt * (nb + nmu), size=3):
            if pidx % (nb + nmu) < nb:
                lb[pidx] = ub[pidx] = float(rng.integers(0, 2))
        out = solve_both(m, Nt, nu, ndelta, x0, om, cost.reshape(B, -1), cuda_device, lb=lb, ub=ub)
        dp, bnc = out["stage_dp"], out["bnc"]
        both = (dp["status"] == 0) & (bnc["status"] == 0)
        assert not (((dp["status"] == 1) & (bnc["status"] == 0)) | ((dp["status"] == 0) & (bnc["status"] == 1))).any()
        rel = np.abs(dp["obj"][both] - bnc["obj"][both]) / np.maximum(1.0, np.abs(bnc["obj"][both]))
        assert rel.max() <= 1e-6, (seed, float(rel.max()), int(np.argmax(rel)))
        checked += int(both.sum())
    assert checked > 800


def test_stage_dp_team_search_forms_agree(cuda_device):
    """Robust constraint sets (256 DEWHs x 32 demand scenarios over the whole horizon, N_p = 48) need 10^2..10^4
    expansions for some agents, so every leg of the search runs: one warp first, then the team -- resumed in place in
    the fused tail (fuse_search = 1), restarted from the incumbent in the wide kernel (fuse_search = 0).  Both forms
    and both table bounds must prove the same optima bit for bit; a node budget below the tree size must return an
    incumbent that is no better than the optimum, with status 2 and a certified gap that covers the difference."""
    import torch
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    dev = cuda_device
    B, N_p, S = 256, 48, 32
    Nt = N_p + 1
    rng = np.random.default_rng(1048)
    wl = syn.dewh_batch(B, N_p, seed=5)
    scen = wl["omega"][:, :, None] * rng.uniform(0.5, 1.8, size=(B, Nt, S))
    cost = np.zeros((B, Nt, 3)); cost[:, :, 0] = wl["q_u"]; cost[:, :, 1:] = wl["q_mu"][:, None, :]
    x0 = torch.as_tensor(wl["x0"]).to(dev); om = torch.as_tensor(wl["omega"]).to(dev)
    sc = torch.as_tensor(scen).to(dev); cst = torch.as_tensor(cost.reshape(B, -1)).to(dev)

    def run(bound, fuse, max_nodes=4000000):
        bm = BatchMpc(wl["mats"], N_p, nu_l=1, device=dev, dp_bound=bound)
        bm.dp_opts.fuse_search = fuse
        bm.dp_opts.max_nodes = max_nodes
        bm.build(want=("H_x", "H_v", "H_omega", "H_5"))
        res = bm.solve(x0, om, cost_v=cst, scenarios=sc)
        return {k: res[k].cpu().numpy() for k in ("obj", "status", "stats", "v")}

    ref = run("constant", 1)
    assert (ref["status"] == 0).all()
    assert ref["stats"][:, 0].max() > 200            # the team search is really exercised
    for bound, fuse in (("constant", 0), ("linear", 1), ("linear", 0)):
        r = run(bound, fuse)
        assert (r["status"] == 0).all(), (bound, fuse)
        assert np.array_equal(r["obj"], ref["obj"]), (bound, fuse)
    # the returned point reproduces the objective
    np.testing.assert_allclose((cost.reshape(B, -1) * ref["v"]).sum(axis=1), ref["obj"], rtol=1e-9)
    # a node budget below the size of the hard trees
    for fuse in (1, 0):
        r = run("constant", fuse, max_nodes=150)
        lim = r["status"] == 2
        assert lim.any() and ((r["status"] == 0) | lim).all()
        assert np.array_equal(r["obj"][~lim], ref["obj"][~lim])
        assert np.isfinite(r["obj"][lim]).all() and (r["obj"][lim] >= ref["obj"][lim] * (1 - 1e-12)).all()
        gap = r["stats"][lim, 6] * 1e-9
        true_gap = (r["obj"][lim] - ref["obj"][lim]) / np.abs(r["obj"][lim])
        assert (gap + 1e-9 >= true_gap).all(), (fuse, float((true_gap - gap).max()))
