"""GPU: the example's controller campaign -- mpc_pb / mpc_ce / mpc_sb_reduced / mpc_sb_full /
mpc_minmax / thermo (examples/.../micro_grid_control_simulation.py:144-152, 200-227; theromstat_control.py:38-62) run
closed loop on a small fleet; every step is checked against the oracle (HiGHS on the stacked constraint sets, numpy
sim step, Python thermostat rule)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fleet_data(B, N_p, steps, seed0=300):
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    Nt = N_p + 1
    params = [syn.dewh_agent_params(seed0 + b) for b in range(B)]
    T0 = np.array([syn.dewh_initial_state(seed0 + b) for b in range(B)])
    forecast = np.stack([syn.dhw_demand_profile(steps + Nt, seed=seed0 + b) for b in range(B)])
    actual = np.stack([syn.dhw_demand_profile(steps + Nt, seed=seed0 + 50 + b) for b in range(B)])
    price = syn.price_profile(steps + Nt, seed=5)
    rng = np.random.default_rng(seed0)
    scen = forecast[:, :, None] * rng.uniform(0.4, 1.8, size=(B, steps + Nt, 5))
    dmin, dmax = 0.5 * forecast.min(axis=0), 1.5 * forecast.max(axis=0)
    return params, T0, forecast, actual, price, scen, (dmin, dmax)


def _oracle_problem(p, Nt, T, w, price_win, extra):
    from oracle import mld as omld, condense as oc, assemble as oa
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    mats = syn.dewh_scalars(p, const_heat=True)
    m = dict(A=[[mats[0]]], B1=[[mats[1]]], B4=[[mats[2]]], b5=[[mats[3]]], E=[[1.0], [-1.0]], F1=[[0.0], [0.0]],
             Psi=[[-1.0, 0.0], [0.0, -1.0]], f5=[[p["T_h_max"]], [-p["T_h_min"]]])
    full, d, vt = omld.complete({kk: np.array(vv, dtype=float) for kk, vv in m.items()}, nu_l=1)
    q_u = price_win * p["P_h_Nom"]
    return oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, np.array([T]), w,
                            atoms=dict(q_u=q_u, q_mu=[10.0 * q_u.sum(), 1.0 * q_u.sum()]), extra_constraints=extra)


@pytest.mark.parametrize("controller", ["mpc_pb", "mpc_ce", "mpc_sb_reduced", "mpc_sb_full", "mpc_minmax"])
def test_campaign_mpc_variant_vs_oracle(controller, cuda_device):
    from oracle import solve as osv, lsim as ol
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    B, N_p, steps, N_sbr = 3, 12, 5, 5
    Nt = N_p + 1
    params, T0, forecast, actual, price, scen, minmax = _fleet_data(B, N_p, steps)
    fleet = DewhFleet(params, N_p, device=cuda_device)
    log = fleet.closed_loop(T0, forecast, price, steps, demand_actual=actual, controller=controller, scenarios=scen,
                            N_sb_reduced=N_sbr, demand_minmax=minmax)
    log = {k: v.cpu().numpy() for k, v in log.items()}
    assert (log["status"] == 0).all()
    T = T0.copy()
    for k in range(steps):
        u0 = np.zeros(B)
        p_agg = np.zeros(Nt)
        for b in range(B):
            w = (actual if controller == "mpc_pb" else forecast)[b, k:k + Nt]
            extra = []
            if controller == "mpc_sb_reduced":
                extra = [dict(omega_scenarios=scen[b, k:k + Nt], N_tilde=N_sbr)]
            elif controller == "mpc_sb_full":
                extra = [dict(omega_scenarios=scen[b, k:k + Nt])]
            elif controller == "mpc_minmax":
                extra = [dict(omega_t=minmax[0][k:k + Nt]), dict(omega_t=minmax[1][k:k + Nt])]
            prob = _oracle_problem(params[b], Nt, T[b], w, price[k:k + Nt], extra)
            st, obj, v = osv.solve_milp(prob, polish=True)
            assert st == 0
            assert abs(log["obj"][k, b] - obj) <= 1e-6 * max(1.0, abs(obj)), (controller, k, b, log["obj"][k, b], obj)
            u = np.round(v[prob.is_bin])
            u0[b] = u[0]
            p_agg += params[b]["P_h_Nom"] * u
            assert log["u"][k, b] == u0[b], (controller, k, b)
        np.testing.assert_allclose(log["P_agg"][k], p_agg, rtol=1e-12)
        for b in range(B):
            T[b], _, _ = ol.dewh_sim_step(dict(params[b]), T[b], u0[b], actual[b, k])
        np.testing.assert_allclose(log["T"][k + 1], T, rtol=1e-10)


def test_thermostat_kernel_vs_oracle(cuda_device):
    """band edges hit exactly, previous input 0 / 1 / NaN / 0.5 (only an exact 1 keeps the element on)."""
    from oracle import lsim as ol
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    rng = np.random.default_rng(0)
    B = 1000
    params = [dict(syn.dewh_agent_params(b), T_h_max_sub_T_h_on=float(rng.uniform(8, 14)),
                   T_h_max_sub_T_h_off=float(rng.uniform(2, 6))) for b in range(B)]
    T = np.array([rng.uniform(p["T_h_max"] - 16, p["T_h_max"]) for p in params])
    for b in range(0, B, 10):                               # exact band edges
        p = params[b]
        T[b] = p["T_h_max"] - (p["T_h_max_sub_T_h_on"] if b % 20 else p["T_h_max_sub_T_h_off"])
    u_prev = rng.choice([0.0, 1.0, np.nan, 0.5], size=B)
    dev = torch.device(cuda_device)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)
    band = np.array([[p["T_h_max_sub_T_h_on"], p["T_h_max_sub_T_h_off"]] for p in params])
    u = cabi.dewh_thermostat(t(cabi.pack_dewh_params(params)), t(band), t(T), t(u_prev)).cpu().numpy()
    want = np.array([ol.dewh_thermostat(params[b], T[b], u_prev[b]) for b in range(B)], dtype=float)
    assert np.array_equal(u, want)
    assert 0 < want.sum() < B
    # one band shared by every agent (stride 0)
    u1 = cabi.dewh_thermostat(t(cabi.pack_dewh_params(params)), t(band[:1]), t(T), t(u_prev)).cpu().numpy()
    want1 = np.array([ol.dewh_thermostat(dict(params[b], T_h_max_sub_T_h_on=band[0, 0], T_h_max_sub_T_h_off=band[0, 1]),
                                         T[b], u_prev[b]) for b in range(B)], dtype=float)
    assert np.array_equal(u1, want1)


def test_campaign_thermostat_closed_loop_vs_oracle(cuda_device):
    from oracle import lsim as ol
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    B, N_p, steps = 6, 12, 40
    params, T0, forecast, actual, price, scen, minmax = _fleet_data(B, N_p, steps, seed0=400)
    T0 = T0 - 6.0                                          # start near the switch-on threshold
    fleet = DewhFleet(params, N_p, device=cuda_device)
    log = {k: v.cpu().numpy() for k, v in
           fleet.closed_loop(T0, forecast, price, steps, demand_actual=actual * 3.0, controller="thermo").items()}
    assert log["P_agg"].shape == (steps, 1) and np.isnan(log["obj"]).all()
    T, u_prev = T0.copy(), np.zeros(B)
    for k in range(steps):
        u = np.array([ol.dewh_thermostat(dict(params[b], T_h_max_sub_T_h_on=12, T_h_max_sub_T_h_off=4), T[b], u_prev[b])
                      for b in range(B)], dtype=float)
        assert np.array_equal(log["u"][k], u), k
        np.testing.assert_allclose(log["P_agg"][k, 0], sum(params[b]["P_h_Nom"] * u[b] for b in range(B)), rtol=1e-12)
        for b in range(B):
            T[b], _, _ = ol.dewh_sim_step(dict(params[b]), T[b], u[b], 3.0 * actual[b, k])
        np.testing.assert_allclose(log["T"][k + 1], T, rtol=1e-10)
        u_prev = u
    assert 0 < log["u"].sum() < B * steps                  # both branches of the rule were exercised


def test_campaign_runs_all_six(cuda_device):
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    B, N_p, steps = 4, 8, 3
    params, T0, forecast, actual, price, scen, minmax = _fleet_data(B, N_p, steps, seed0=500)
    fleet = DewhFleet(params, N_p, device=cuda_device)
    out = fleet.campaign(DewhFleet.CONTROLLERS, T0, forecast, price, steps, demand_actual=actual, scenarios=scen,
                         demand_minmax=minmax)
    assert set(out) == set(DewhFleet.CONTROLLERS)
    for name, log in out.items():
        assert log["T"].shape == (steps + 1, B) and log["u"].shape == (steps, B)
    with pytest.raises(ValueError):
        fleet.closed_loop(T0, forecast, price, steps, controller="mpc_sb_full")
    with pytest.raises(ValueError):
        fleet.closed_loop(T0, forecast, price, steps, controller="nope")


def test_grid_log_and_result_frame(cuda_device):
    """closed-loop logs -> grid-agent arrays (K5 on the grid MLD, batch axis = steps) -> the reference's result frame;
    the grid arrays are checked against the oracle's lsim_k, the frame against the log it was built from."""
    from oracle import lsim as ol, mld as omld
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import results, models
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.parameters import (
        pv_param_struct, res_demand_param_struct, grid_param_struct)
    B, N_p, steps = 5, 8, 6
    params, T0, forecast, actual, price, scen, minmax = _fleet_data(B, N_p, steps, seed0=600)
    ids = [9, 2, 5, 4, 7]
    rng = np.random.default_rng(3)
    pv = dict(omega=rng.random(steps), omega_hat=rng.random(steps), gain=models.pv_gain(dict(pv_param_struct, P_pv_units=B)))
    resd = dict(omega=2 * rng.random(steps), omega_hat=2 * rng.random(steps),
                gain=models.resd_gain(dict(res_demand_param_struct, P_res_units=B)))
    fleet = DewhFleet(params, N_p, device=cuda_device)
    logs = fleet.campaign(("mpc_ce", "thermo"), T0, forecast, price, steps, demand_actual=actual)
    grid = {k: v.cpu().numpy() for k, v in fleet.grid_log(logs["mpc_ce"], price, pv=pv, resd=resd, device_ids=ids).items()}
    lg = {k: v.cpu().numpy() for k, v in logs["mpc_ce"].items()}
    P = np.array([p["P_h_Nom"] for p in params])
    order = np.argsort(ids)
    full, d, _ = omld.complete({k: np.array(v, dtype=float) for k, v in ol.grid_mld(grid_param_struct, B + 2).items()})
    for tag in ("", "_hat"):
        w = np.concatenate([(lg["u"] * P)[:, order], pv["gain"] * pv["omega" + tag][:, None],
                            resd["gain"] * resd["omega" + tag][:, None]], axis=1)
        np.testing.assert_allclose(grid["omega" + tag], w, rtol=1e-15)
        for k in range(steps):
            de, z = ol.grid_aux_closed_form(float(np.ones(B + 2) @ w[k]))
            _, y, cons = ol.lsim_k(full, np.zeros((0, 1)), np.zeros((0, 1)), np.array([[de]]), np.array([[z]]),
                                   np.zeros((0, 1)), w[k].reshape(-1, 1))
            assert grid["delta" + tag][k] == de
            np.testing.assert_allclose(grid["y" + tag][k], float(y.ravel()[0]), rtol=1e-13)
            np.testing.assert_allclose(grid["z" + tag][k], z, rtol=1e-13)
            if not tag:
                assert np.array_equal(grid["cons"][k].astype(bool), np.asarray(cons).ravel())
    assert grid["cons"].all()                                # the closed-form auxiliaries are feasible
    blocks = [("dewh", ids, c, results.dewh_log_blocks({k: v.cpu().numpy() for k, v in logs[c].items()}, params, c))
              for c in ("mpc_ce", "thermo")]
    blocks += [("pv", [1], "mpc_ce", results.source_log_blocks(pv["omega"], pv["omega_hat"], pv["gain"])),
               ("resd", [1], "mpc_ce", results.source_log_blocks(resd["omega"], resd["omega_hat"], resd["gain"])),
               ("grid", [1], "mpc_ce", results.grid_log_blocks(grid))]
    df = results.grid_sim_dataframe(blocks, steps, time_0="2018-12-10")
    assert df.shape[0] == steps and df.columns.names == list(results.LEVELS)
    b = ids.index(5)
    np.testing.assert_array_equal(df[("dewh", 5, "mpc_ce", "x_k1", 0)].to_numpy(), lg["T"][1:, b])
    np.testing.assert_array_equal(df[("dewh", 5, "thermo", "u", 0)].to_numpy(), logs["thermo"]["u"].cpu().numpy()[:, b])
    imp, exp = df[("grid", 1, "mpc_ce", "p_imp", 0)].to_numpy(), df[("grid", 1, "mpc_ce", "p_exp", 0)].to_numpy()
    np.testing.assert_allclose(imp + exp, grid["y"], rtol=1e-13)
    assert (imp >= 0).all() and (exp <= 0).all()
    np.testing.assert_allclose(df[("grid", 1, "mpc_ce", "cost", 0)].to_numpy(), imp * price[:steps], rtol=1e-13)
    assert (lg["solve_ms"] > 0).all()
