"""GPU: price coordination of the fleet for the reference's CENTRALISED micro-grid problem (SURVEY.md 8(f1);
micro_grid_agents.py:691-735).  The scheme returns a feasible centralised plan (upper bound) and a certified lower
bound; both are checked against HiGHS on the monolithic MILP (oracle.coupled) for small fleets, and the kernels
against a numpy twin."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(N_h, N_p, seed=0):
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    Nt = N_p + 1
    params = [syn.dewh_agent_params(seed * 100 + b) for b in range(N_h)]
    T0 = np.array([syn.dewh_initial_state(seed * 100 + b) for b in range(N_h)])
    dem = np.stack([syn.dhw_demand_profile(Nt, seed=seed * 100 + b) for b in range(N_h)])
    price = syn.price_profile(Nt, seed=seed)
    P = np.array([p["P_h_Nom"] for p in params])
    k = np.arange(Nt)
    pv = -0.6 * P.sum() * np.clip(np.sin((k - 2) / Nt * 2 * np.pi), 0, None)   # a PV surplus hump: free energy
    p_other = pv + 0.1 * P.sum()
    return params, T0, dem, price, P, p_other


def _agent_problem(p, Nt, T, w, price):
    """device objective only (slack penalties): what the reference's grid controller collects from a DEWH"""
    from oracle import mld as omld, condense as oc, assemble as oa
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    mats = syn.dewh_scalars(p, const_heat=True)
    m = dict(A=[[mats[0]]], B1=[[mats[1]]], B4=[[mats[2]]], b5=[[mats[3]]], E=[[1.0], [-1.0]], F1=[[0.0], [0.0]],
             Psi=[[-1.0, 0.0], [0.0, -1.0]], f5=[[p["T_h_max"]], [-p["T_h_min"]]])
    full, d, vt = omld.complete({kk: np.array(vv, dtype=float) for kk, vv in m.items()}, nu_l=1)
    tot = (price * p["P_h_Nom"]).sum()
    return oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, np.array([T]), w, atoms=dict(q_mu=[10.0 * tot, tot]))


@pytest.mark.parametrize("N_h,seed", [(4, 0), (6, 1), (5, 2)])
def test_coordination_brackets_the_centralised_optimum(N_h, seed, cuda_device):
    from oracle import coupled, solve as osv
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.parameters import grid_param_struct
    N_p = 12
    Nt = N_p + 1
    params, T0, dem, price, P, p_other = _case(N_h, N_p, seed)
    fleet = DewhFleet(params, N_p, device=cuda_device)
    fleet.build()
    out = fleet.coupled_step(T0, dem, price, p_other, iters=400, rel_gap=1e-3)
    probs = [_agent_problem(params[b], Nt, T0[b], dem[b], price) for b in range(N_h)]
    prob, _, _ = coupled.build_coupled_problem(probs, P, p_other, price, dict(grid_param_struct, P_g_min=-1e7, P_g_max=1e7))
    st, opt, _ = osv.solve_milp(prob)
    assert st == 0
    tol = 1e-7 * max(1.0, abs(opt))
    assert out["skipped"] == 0
    assert out["lower_bound"] <= opt + tol, (out["lower_bound"], opt)
    assert out["upper_bound"] >= opt - tol, (out["upper_bound"], opt)
    assert out["upper_bound"] <= opt * 1.01 + tol            # the reference's own acceptance gap (MIPGap = 1e-2)
    assert out["lower_bound"] >= opt * 0.90                  # the duality gap of a handful of agents is not small
    U = out["u"].cpu().numpy()
    assert set(np.unique(U)) <= {0.0, 1.0}
    np.testing.assert_allclose(out["upper_bound"], coupled.coupled_cost(probs, U, P, p_other, price), rtol=1e-7)
    # the agents' answer at the best price is that plan
    assert np.array_equal(out["plan"]["u"].cpu().numpy(), U)
    # and coordination pays: the purely decentralised plan (every agent sees the full price) costs more
    dec = fleet.control_step(torch.as_tensor(T0).reshape(-1, 1), dem, fleet.cost_from_prices(price))
    dec_cost = coupled.coupled_cost(probs, dec["u"].cpu().numpy(), P, p_other, price)
    assert out["upper_bound"] <= dec_cost + tol


def test_coordination_without_surplus_is_the_decentralised_plan(cuda_device):
    """no PV surplus -> the import price applies to every watt -> lambda = price is optimal with zero gap"""
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    N_h, N_p = 8, 16
    params, T0, dem, price, P, _ = _case(N_h, N_p, seed=3)
    fleet = DewhFleet(params, N_p, device=cuda_device)
    fleet.build()
    out = fleet.coupled_step(T0, dem, price, np.full(N_p + 1, 500.0), iters=50)
    assert out["iterations"] <= 25 and out["gap"] <= 1e-12
    np.testing.assert_allclose(out["lam"].cpu().numpy(), price, rtol=1e-9)
    dec = fleet.control_step(torch.as_tensor(T0).reshape(-1, 1), dem, fleet.cost_from_prices(price))
    assert np.array_equal(out["u"].cpu().numpy(), dec["u"].cpu().numpy())


def test_coupling_kernels_vs_numpy(cuda_device):
    from pyhybridcontrol_b200 import cabi
    dev = torch.device(cuda_device)
    rng = np.random.default_rng(5)
    B, Nt, nv = 333, 49, 3
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    v = rng.random((B, Nt, nv))
    v[:, :, 0] = rng.random((B, Nt)) > 0.5
    P, obj = 2500 + 1000 * rng.random(B), rng.random(B) * 3
    status = (rng.random(B) > 0.99).astype(np.int32)
    vd = t(v.reshape(B, -1))
    sums = torch.empty(Nt + 2, dtype=torch.float64, device=dev)
    cabi.coupling_sums(vd.view(B, Nt, nv)[:, :, 0], t(P), t(obj), t(status, torch.int32), sums)
    s = sums.cpu().numpy()
    np.testing.assert_allclose(s[:Nt], P @ v[:, :, 0], rtol=1e-13)
    np.testing.assert_allclose(s[Nt], obj.sum(), rtol=1e-13)
    assert s[Nt + 1] == status.sum()
    # price -> cost column
    lam = rng.random(Nt) * 1e-4
    cost = t(np.ones((B, Nt * nv)))
    cabi.coupling_price_cost(t(lam), t(P), cost, nv, 0)
    c = cost.cpu().numpy().reshape(B, Nt, nv)
    np.testing.assert_array_equal(c[:, :, 0], lam[None, :] * P[:, None])
    assert (c[:, :, 1:] == 1.0).all()
    # dual step against the formulas (prototype: tools/coupling_proto.py)
    price = lam * (1 + rng.random(Nt))
    r = (rng.random(Nt) - 0.6) * P.sum()
    lo, hi = np.zeros(Nt), np.full(Nt, P.sum())
    s[Nt + 1] = 0.0
    state = cabi.coupling_state(dev)
    lam_next = torch.empty(Nt, dtype=torch.float64, device=dev)
    cabi.coupling_dual_step(t(s), t(r), t(price), t(lo), t(hi), 1.0, t(lam), lam_next, state)
    agg = s[:Nt]
    cand = np.stack([np.clip(-r, lo, hi), lo, hi])
    vals = price * np.maximum(0, cand + r) - lam * cand
    j = vals.argmin(0)
    a_star = cand[j, np.arange(Nt)]
    dual = s[Nt] + vals.min(0).sum()
    primal = (price * np.maximum(0, agg + r)).sum() + s[Nt] - lam @ agg
    g = agg - a_star
    want = np.clip(lam + (primal - dual) / (g @ g) * g, 0, price)
    st = state.cpu().numpy()
    np.testing.assert_allclose(st[:4], [dual, primal, dual, primal], rtol=1e-12)
    assert st[4] == 1.0 and st[5] == 1.0 and st[7] == 0.0
    np.testing.assert_allclose(st[6], g @ g, rtol=1e-12)
    np.testing.assert_allclose(lam_next.cpu().numpy(), want, rtol=1e-10, atol=1e-18)
    # multipliers on the bounds of their box: minimum-norm member of the subdifferential, projected
    lam2 = lam.copy()
    lam2[::3] = 0.0
    lam2[1::3] = price[1::3]
    state2 = cabi.coupling_state(dev)
    lam_next2 = torch.empty(Nt, dtype=torch.float64, device=dev)
    cabi.coupling_dual_step(t(s), t(r), t(price), t(lo), t(hi), 1.0, t(lam2), lam_next2, state2)
    kink = np.clip(-r, lo, hi)
    arg = np.where(lam2 <= 0, np.clip(agg, lo, kink), np.where(lam2 >= price, np.clip(agg, kink, hi), kink))
    g2v = agg - arg
    g2v[(lam2 <= 0) & (g2v < 0)] = 0.0
    g2v[(lam2 >= price) & (g2v > 0)] = 0.0
    cand2 = np.stack([kink, lo, hi])
    dual2 = s[Nt] + (price * np.maximum(0, cand2 + r) - lam2 * cand2).min(0).sum()
    primal2 = (price * np.maximum(0, agg + r)).sum() + s[Nt] - lam2 @ agg
    st2 = state2.cpu().numpy()
    np.testing.assert_allclose(st2[2:4], [dual2, primal2], rtol=1e-12)
    np.testing.assert_allclose(st2[6], g2v @ g2v, rtol=1e-12)
    np.testing.assert_allclose(lam_next2.cpu().numpy(), np.clip(lam2 + (primal2 - dual2) / (g2v @ g2v) * g2v, 0, price),
                               rtol=1e-10, atol=1e-18)
    assert (g2v == 0).sum() > 0
    # keep_best copies only when the bound improved
    u_best, lam_best = torch.zeros((B, Nt), dtype=torch.float64, device=dev), torch.zeros(Nt, dtype=torch.float64, device=dev)
    cabi.coupling_keep_best(vd.view(B, Nt, nv)[:, :, 0], t(lam), state, u_best, lam_best)
    assert np.array_equal(u_best.cpu().numpy(), v[:, :, 0]) and np.array_equal(lam_best.cpu().numpy(), lam)
    state[4] = 0.0
    u2 = torch.zeros_like(u_best)
    cabi.coupling_keep_best(vd.view(B, Nt, nv)[:, :, 0], t(lam), state, u2, lam_best)
    assert not u2.any()
    # a failed agent: the iterate gives no bound and lambda is carried over
    s[Nt + 1] = 2.0
    state = cabi.coupling_state(dev)
    cabi.coupling_dual_step(t(s), t(r), t(price), t(lo), t(hi), 1.0, t(lam), lam_next, state)
    st = state.cpu().numpy()
    assert st[7] == 1.0 and np.isinf(st[0]) and np.isinf(st[1])
    np.testing.assert_array_equal(lam_next.cpu().numpy(), lam)


def test_coordination_larger_fleet_gap(cuda_device):
    """40 agents, N_p = 24: certified gap below the reference's MIPGap within a few hundred iterations, and a large
    saving over the uncoordinated plan (every agent chasing the same cheap hours ignores the free PV energy)."""
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    N_h, N_p = 40, 24
    params, T0, dem, price, P, p_other = _case(N_h, N_p, seed=0)
    fleet = DewhFleet(params, N_p, device=cuda_device)
    fleet.build()
    out = fleet.coupled_step(T0, dem, price, p_other, iters=400, rel_gap=1e-2)
    assert out["gap"] <= 1e-2 and out["skipped"] == 0, out["gap"]
    assert out["lower_bound"] <= out["upper_bound"] <= out["dual_upper_bound"]
    assert out["response_accepted"] > 0
    # uncoordinated cost on the coupled objective, computed on the device from the same pieces
    dec = fleet.control_step(torch.as_tensor(T0).reshape(-1, 1), dem, fleet.cost_from_prices(price))
    agg = (dec["u"] * fleet.P_nom[:, None]).sum(0).cpu().numpy()
    pen = float(dec["obj"].sum().cpu()) - float(price @ agg)
    dec_cost = float((price * np.maximum(0, agg + p_other)).sum() + pen)
    assert out["upper_bound"] < 0.9 * dec_cost


def test_best_response_kernels_vs_numpy(cuda_device):
    from pyhybridcontrol_b200 import cabi
    dev = torch.device(cuda_device)
    rng = np.random.default_rng(11)
    B, Nt, nv = 37, 25, 3
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    P = 2500 + 1000 * rng.random(B)
    price = rng.random(Nt) * 1e-4
    r = (rng.random(Nt) - 0.7) * P.sum()
    v = rng.random((B, Nt, nv)); v[:, :, 0] = rng.random((B, Nt)) > 0.5
    agg = P @ v[:, :, 0]
    cost = t(np.full((B, Nt * nv), 7.0))
    v_cur = t(v.reshape(B, -1))
    cabi.coupling_response_cost(t(agg), v_cur, t(P), t(r), t(price), cost, nv, 0)
    others = agg[None, :] - P[:, None] * v[:, :, 0] + r[None, :]
    want = price[None, :] * (np.maximum(0, others + P[:, None]) - np.maximum(0, others))
    c = cost.cpu().numpy().reshape(B, Nt, nv)
    np.testing.assert_allclose(c[:, :, 0], want, rtol=1e-12, atol=1e-18)
    assert (c[:, :, 1:] == 7.0).all() and (want >= 0).all() and (want <= price[None, :] * P[:, None] * (1 + 1e-12)).all()
    # merge rows [lo, hi), evaluate, accept / restore
    lo, hi = 10, 22
    pen = rng.random(B)
    pen_cur = t(pen)
    v_new = rng.random((hi - lo, Nt, nv)); v_new[:, :, 0] = rng.random((hi - lo, Nt)) > 0.5
    obj_new = rng.random(hi - lo) + 5
    status = np.zeros(hi - lo, dtype=np.int32)
    v_bak, pen_bak = torch.empty((B, Nt * nv), dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev)
    cabi.coupling_merge(lo, hi, Nt, nv, t(v_new.reshape(hi - lo, -1)), t(obj_new), t(status, torch.int32), cost, v_cur,
                        pen_cur, v_bak, pen_bak)
    v2 = v.copy(); v2[lo:hi] = v_new
    pen2 = pen.copy(); pen2[lo:hi] = obj_new - (want[lo:hi] * v_new[:, :, 0]).sum(1)
    np.testing.assert_array_equal(v_cur.cpu().numpy().reshape(B, Nt, nv), v2)
    np.testing.assert_allclose(pen_cur.cpu().numpy(), pen2, rtol=1e-12)
    sums_cand = torch.empty(Nt + 2, dtype=torch.float64, device=dev)
    cabi.coupling_sums(v_cur.view(B, Nt, nv)[:, :, 0], t(P), pen_cur, None, sums_cand)
    total2 = (price * np.maximum(0, P @ v2[:, :, 0] + r)).sum() + pen2.sum()
    a_lo, a_hi = t(np.zeros(Nt)), t(np.full(Nt, P.sum()))
    for start, accepted in ((total2 * 2, True), (total2 * 0.5, False), (float("inf"), True)):
        brs = t(np.array([start, 0, 0, 0.0]))
        sums_cur = t(np.zeros(Nt + 2))
        vc, pc = v_cur.clone(), pen_cur.clone()
        cabi.coupling_accept(sums_cand, t(r), t(price), a_lo, a_hi, sums_cur, brs)
        cabi.coupling_restore(lo, hi, brs, v_bak, pen_bak, vc, pc)
        st = brs.cpu().numpy()
        if accepted:
            np.testing.assert_allclose(st[0], total2, rtol=1e-12)
            assert st[1] == 1.0 and st[2] == 1.0 and st[3] == 0.0
            np.testing.assert_array_equal(sums_cur.cpu().numpy(), sums_cand.cpu().numpy())
            np.testing.assert_array_equal(vc.cpu().numpy().reshape(B, Nt, nv), v2)
        else:
            assert st[0] == start and st[1] == 0.0 and st[3] == 1.0 and not sums_cur.any()
            np.testing.assert_array_equal(vc.cpu().numpy().reshape(B, Nt, nv), v)
            np.testing.assert_array_equal(pc.cpu().numpy(), pen)
    # a failed agent in the block makes the candidate worthless
    status[3] = 2
    cabi.coupling_merge(lo, hi, Nt, nv, t(v_new.reshape(hi - lo, -1)), t(obj_new), t(status, torch.int32), cost, v_cur,
                        pen_cur, v_bak, pen_bak)
    assert np.isinf(pen_cur.cpu().numpy()[lo + 3])


def test_closed_loop_with_coordination(cuda_device):
    """a few closed-loop steps of the centralised operation: every instant certified, and the import bill of the
    simulated day is lower than the one of the uncoordinated fleet (same tanks, same draws)."""
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    N_h, N_p, steps = 24, 16, 8
    Nt = N_p + 1
    params = [syn.dewh_agent_params(700 + b) for b in range(N_h)]
    T0 = np.array([syn.dewh_initial_state(700 + b) for b in range(N_h)])
    dem = np.stack([syn.dhw_demand_profile(steps + Nt, seed=700 + b) for b in range(N_h)])
    price = syn.price_profile(steps + Nt, seed=7)
    P = np.array([p["P_h_Nom"] for p in params])
    k = np.arange(steps + Nt)
    p_other = -0.7 * P.sum() * np.clip(np.sin(k / 12 * np.pi), 0, None) + 0.05 * P.sum()
    fleet = DewhFleet(params, N_p, device=cuda_device)
    co = {key: v.cpu().numpy() for key, v in
          fleet.closed_loop(T0, dem, price, steps, coupling=dict(p_other=p_other, iters=150)).items()}
    de = {key: v.cpu().numpy() for key, v in fleet.closed_loop(T0, dem, price, steps).items()}
    assert (co["status"] == 0).all() and (co["coupled_gap"] <= 0.05).all()
    bill = lambda lg: float((price[:steps] * np.maximum(0, lg["P_agg"][:, 0] + p_other[:steps])).sum())
    assert bill(co) <= bill(de) * (1 + 1e-12)
    assert np.isfinite(co["T"]).all() and co["mu_hat"].shape == (steps, N_h, 2)


@pytest.mark.parametrize("N_h,N_p,seed", [(3, 8, 0), (4, 10, 1), (5, 12, 2)])
def test_centralised_problem_exact_small_fleets(N_h, N_p, seed, cuda_device):
    """the monolithic micro-grid MILP on the GPU (DewhFleet.coupled_step_exact, branch-and-cut kernel) against HiGHS on
    the reference's formulation with the grid MLD's delta / z rows: the same optimum to 1e-6, a plan whose true cost is
    that optimum, and the price coordination's bounds around it"""
    from oracle import coupled, solve as osv
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.parameters import grid_param_struct
    Nt = N_p + 1
    params, T0, dem, price, P, p_other = _case(N_h, N_p, seed)
    fleet = DewhFleet(params, N_p, device=cuda_device)
    fleet.build()
    ex = fleet.coupled_step_exact(T0, dem, price, p_other)
    assert ex["status"] == 0
    probs = [_agent_problem(params[b], Nt, T0[b], dem[b], price) for b in range(N_h)]
    prob, _, _ = coupled.build_coupled_problem(probs, P, p_other, price, dict(grid_param_struct, P_g_min=-1e7, P_g_max=1e7))
    st, opt, _ = osv.solve_milp(prob, polish=True)
    assert st == 0
    assert abs(ex["obj"] - opt) <= 1e-6 * max(1.0, abs(opt)), (ex["obj"], opt)
    U = np.round(ex["u"].cpu().numpy())
    np.testing.assert_allclose(coupled.coupled_cost(probs, U, P, p_other, price), opt, rtol=1e-6)
    out = fleet.coupled_step(T0, dem, price, p_other, iters=300, rel_gap=1e-3)
    tol = 1e-7 * max(1.0, abs(opt))
    assert out["lower_bound"] <= ex["obj"] + tol and ex["obj"] <= out["upper_bound"] + tol
