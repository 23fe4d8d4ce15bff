"""CPU: the oracle's CENTRALISED micro-grid problem (oracle/coupled.py), device simulation steps and grid bookkeeping
against the UNMODIFIED reference's own example loop on a small micro-grid -- GridAgentMpc.build_grid / solve_grid_mpc /
sim_step_k with four water heaters, a PV plant and a residential demand, the six controllers of the reference's
campaign side by side (perfect forecast, certainty equivalent, scenario-based reduced / full, min-max, thermostat),
three instants (tests/golden/microgrid_loop.npz, tests/golden/make_golden_microgrid.py; cvxpy's modelling
layer = oracle/mini_cvxpy.py, MILP backend = HiGHS).  The forecast / actual / price windows come from the product's
input-side module (examples/.../profiles.py), so this also checks that module inside the loop."""
import os

import numpy as np
import pytest

from oracle import assemble as oa
from oracle import condense as oc
from oracle import coupled as ocp
from oracle import lsim as ol
from oracle import mld as omld
from oracle import solve as osv
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import profiles as P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "microgrid_loop.npz"))
KEYS = ("C_w", "A_h", "U_h", "m_h", "T_w", "T_inf", "P_h_Nom", "T_h_min", "T_h_max", "T_h_Nom", "ts")


MPC = [i for i, c in enumerate(G["controllers"]) if str(c).startswith("mpc")]


def scenario_draws():
    """(controller, k) -> [N_h, Nt, S]: the reference draws every heater's scenario set from numpy's GLOBAL generator,
    inside its loops over instants, controllers and devices (micro_grid_control_simulation.py:184-201); the same seed
    and the same call order through examples/.../profiles.py give the same sets."""
    N_h, Nt, steps, S = int(G["N_h"]), int(G["N_p"]) + 1, int(G["steps"]), int(G["num_scenarios"])
    sc = P.OmegaScenarios(G["scen_days"].flatten(order="F"), 900.0)
    state = np.random.get_state()
    try:
        np.random.seed(int(G["seed"]))
        out = {}
        for k in range(steps):
            for cname in (str(c) for c in G["controllers"]):
                if cname.startswith("mpc_sb"):
                    out[(cname, k)] = sc.fleet_scenarios(k, Nt, S, N_h)
    finally:
        np.random.set_state(state)
    return out, sc


def test_scenario_draws_of_the_loop_are_reproduced():
    draws, _ = scenario_draws()
    assert len(draws) == 2 * int(G["steps"])
    for (cname, k), arr in draws.items():
        assert np.array_equal(arr, G["scen_" + cname][k]), (cname, k)


@pytest.mark.parametrize("ci", MPC)
def test_centralised_loop_against_the_reference(ci):
    cname, det = str(G["controllers"][ci]), bool(G["deterministic"][ci])
    N_h, N_p, steps, lag = int(G["N_h"]), int(G["N_p"]), int(G["steps"]), int(G["lag"])
    Nt = N_p + 1
    params = [dict(zip(KEYS, row)) for row in G["dewh_params"]]
    P_nom = np.array([p["P_h_Nom"] for p in params])
    grid_params = dict(zip(("P_g_min", "P_g_max", "eps"), G["grid_params"]))
    dewh = P.OmegaProfiles(G["dewh_profiles"][:, :, None], 900.0)
    pv = P.OmegaProfiles(G["pv_profile"], 900.0)
    resd = P.OmegaProfiles(G["resd_profile"], 900.0)
    price = P.PriceProfile(G["price"], 900.0)
    assert dewh.lag == lag
    models = []
    for p in params:
        full, dims, vt = omld.complete(ol.dewh_mld(p, const_heat=True), nu_l=1)
        models.append((oc.condense(full, dims, Nt), dims, vt))
    T = G["x0"].astype(float).copy()
    draws, scen = scenario_draws()
    lo_day, hi_day = scen.min_max_day()
    for k in range(steps):
        pk = price.price_tilde_k(k, Nt)
        w = dewh.omega_tilde_k_hat(k, Nt, deterministic=det)
        # the constraint sets a controller variant adds to every heater (micro_grid_control_simulation.py:199-227)
        if cname == "mpc_sb_reduced":
            extra = [[dict(omega_scenarios=draws[(cname, k)][i], N_tilde=int(G["N_sb_reduced"]))] for i in range(N_h)]
        elif cname == "mpc_sb_full":
            extra = [[dict(omega_scenarios=draws[(cname, k)][i])] for i in range(N_h)]
        elif cname == "mpc_minmax":
            lo, hi = P.get_min_max_dhw_scenario(k, Nt, lo_day, hi_day)
            extra = [[dict(omega_t=lo[:, 0]), dict(omega_t=hi[:, 0])]] * N_h
        else:
            extra = [[]] * N_h
        p_other = float(G["pv_gain"]) * pv.omega_tilde_k_hat(k, Nt, deterministic=det)[0] + \
            float(G["resd_gain"]) * resd.omega_tilde_k_hat(k, Nt, deterministic=det)[0]
        max_cost = pk.sum() * 3000.0                       # the script uses the base P_h_Nom for every heater (:196)
        q_mu = np.array([max_cost * G["soft"][0], max_cost * G["soft"][1]])
        np.testing.assert_allclose(T, G[cname + "_T"][k], rtol=1e-12)
        agents = [oa.build_problem(evo, dims, vt, Nt, [T[i]], w[i], atoms=dict(q_mu=q_mu), extra_constraints=extra[i])
                  for i, (evo, dims, vt) in enumerate(models)]
        prob, offs, n_agents = ocp.build_coupled_problem(agents, P_nom, p_other, pk, grid_params)
        status, obj, v = osv.solve_milp(prob, polish=True)
        ref = float(G[cname + "_obj"][k])
        assert status == osv.OPTIMAL and abs(obj - ref) <= 1e-6 * max(1.0, abs(ref)), (k, obj, ref)
        # the reference's own plan costs the same in the oracle's model (so ties cannot hide a modelling difference)
        U_ref = np.round(G[cname + "_u_plan"][k])
        assert abs(ocp.coupled_cost(agents, U_ref, P_nom, p_other, pk) - ref) <= 1e-6 * max(1.0, abs(ref))
        z_ref = G[cname + "_z_plan"][k]
        np.testing.assert_allclose(z_ref, np.maximum(0.0, P_nom @ U_ref + p_other), rtol=1e-9, atol=1e-6)
        # simulation step of every device with the reference's first controls and the actual disturbances
        u0 = np.round(G[cname + "_u"][k])
        assert np.array_equal(u0, U_ref[:, 0])
        w_act = dewh.omega_k_act(k)[:, 0]
        T_next = np.array([ol.dewh_sim_step(params[i], T[i], u0[i], w_act[i])[0] for i in range(N_h)])
        np.testing.assert_allclose(T_next, G[cname + "_T_next"][k], rtol=1e-9)
        powers = np.concatenate([P_nom * u0, [float(G["pv_gain"]) * pv.omega_k_act(k)[0, 0]],
                                 [float(G["resd_gain"]) * resd.omega_k_act(k)[0, 0]]])
        np.testing.assert_allclose(powers, G[cname + "_grid_omega"][k], rtol=1e-12)       # devices ordered (type, id)
        y = powers.sum()
        delta, z = ol.grid_aux_closed_form(y)
        assert abs(y - G[cname + "_grid_y"][k]) <= 1e-9 * max(1.0, abs(y))
        assert float(delta) == G[cname + "_grid_delta"][k] and abs(float(z) - G[cname + "_grid_z"][k]) <= 1e-6
        cost_k = float(z) * price.price_tilde_k(k, 1)[0]
        assert abs(cost_k - G[cname + "_cost"][k]) <= 1e-9 * max(1.0, abs(cost_k)) + 1e-9
        T = T_next


def test_thermostat_loop_against_the_reference():
    """the non-predictive controller of the same loop (DewhTheromstatController on the heaters, NoController elsewhere):
    hysteresis rule with u(-1) = 0, simulation steps, grid bookkeeping"""
    cname = "thermo"
    N_h, steps = int(G["N_h"]), int(G["steps"])
    params = [dict(zip(KEYS, row), T_h_max_sub_T_h_on=12, T_h_max_sub_T_h_off=4) for row in G["dewh_params"]]
    P_nom = np.array([p["P_h_Nom"] for p in params])
    dewh = P.OmegaProfiles(G["dewh_profiles"][:, :, None], 900.0)
    pv, resd = P.OmegaProfiles(G["pv_profile"], 900.0), P.OmegaProfiles(G["resd_profile"], 900.0)
    price = P.PriceProfile(G["price"], 900.0)
    T, u_prev = G["x0"].astype(float).copy(), np.zeros(N_h)
    for k in range(steps):
        u = np.array([ol.dewh_thermostat(params[i], T[i], u_prev[i]) for i in range(N_h)], dtype=float)
        assert np.array_equal(u, G[cname + "_u"][k]), k
        w_act = dewh.omega_k_act(k)[:, 0]
        T_next = np.array([ol.dewh_sim_step(params[i], T[i], u[i], w_act[i])[0] for i in range(N_h)])
        np.testing.assert_allclose(T_next, G[cname + "_T_next"][k], rtol=1e-9)
        powers = np.concatenate([P_nom * u, [float(G["pv_gain"]) * pv.omega_k_act(k)[0, 0]],
                                 [float(G["resd_gain"]) * resd.omega_k_act(k)[0, 0]]])
        np.testing.assert_allclose(powers, G[cname + "_grid_omega"][k], rtol=1e-12)
        delta, z = ol.grid_aux_closed_form(powers.sum())
        assert float(delta) == G[cname + "_grid_delta"][k] and abs(float(z) - G[cname + "_grid_z"][k]) <= 1e-6
        assert abs(float(z) * price.price_tilde_k(k, 1)[0] - G[cname + "_cost"][k]) <= 1e-9
        T, u_prev = T_next, u


def test_reference_frame_shape():
    """the frame of the reference's loop: grid first, then the devices by (type, id), every controller's columns"""
    cols = [c.split("|") for c in G["frame_columns"]]
    assert G["frame_values"].shape == (int(G["steps"]), len(cols))
    order = []
    for c in cols:
        if (c[0], c[1]) not in order:
            order.append((c[0], c[1]))
    assert order == [("grid", "1")] + [("dewh", str(i)) for i in range(1, int(G["N_h"]) + 1)] + [("pv", "1"), ("resd", "1")]
    assert {c[2] for c in cols} == {str(c) for c in G["controllers"]}


def test_result_frame_of_the_real_loop():
    """examples/.../results.py rebuilds the frame the reference's REAL loop produced (two controllers on every device):
    the inputs are the frame's own primary columns (temperatures, controls, disturbances, planned slacks, grid
    quantities); every derived column (y, v, x_hat / y_hat / u_hat / v_hat, p_imp, p_exp, cost, the grid's v), the
    column order and the device order must come out identical.  Solve-time columns are excluded (wall-clock)."""
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import results
    cols = [tuple(c.split("|")) for c in G["frame_columns"]]
    vals = G["frame_values"]
    steps, N_h, lag = int(G["steps"]), int(G["N_h"]), int(G["lag"])
    params = [dict(zip(KEYS, row)) for row in G["dewh_params"]]

    def col(dev, dev_id, cname, var, idx=0):
        return vals[:, cols.index((dev, str(dev_id), cname, var, str(idx)))]

    blocks = []
    for cname in (str(c) for c in G["controllers"]):
        ids = list(range(1, N_h + 1))
        T = np.stack([np.append(col("dewh", i, cname, "x_hat"), col("dewh", i, cname, "x_k1")[-1]) for i in ids], axis=1)
        log = dict(T=T, u=np.stack([col("dewh", i, cname, "u") for i in ids], axis=1),
                   omega=np.stack([col("dewh", i, cname, "omega") for i in ids], axis=1),
                   omega_hat=np.stack([col("dewh", i, cname, "omega_hat") for i in ids], axis=1),
                   mu_hat=np.stack([np.stack([col("dewh", i, cname, "mu_hat", j) for j in (0, 1)], axis=1) for i in ids], axis=1),
                   cons=np.stack([np.stack([col("dewh", i, cname, "cons", j) for j in (0, 1)], axis=1) for i in ids], axis=1))
        for i in ids[:-1]:                                # the loop's own continuity: x_k1 of step k is x_hat of k + 1
            np.testing.assert_allclose(col("dewh", i, cname, "x_k1")[:-1], col("dewh", i, cname, "x_hat")[1:], rtol=1e-12)
        blocks.append(("dewh", ids, cname, results.dewh_log_blocks(log, params, cname)))
        for dev, gain in (("pv", float(G["pv_gain"])), ("resd", float(G["resd_gain"]))):
            blocks.append((dev, [1], cname, results.source_log_blocks(col(dev, 1, cname, "omega"),
                                                                      col(dev, 1, cname, "omega_hat"), gain,
                                                                      is_mpc=cname != "thermo")))
        n_dev = N_h + 2
        grid = dict(y=col("grid", 1, cname, "y"), delta=col("grid", 1, cname, "delta"), z=col("grid", 1, cname, "z"),
                    omega=np.stack([col("grid", 1, cname, "omega", j) for j in range(n_dev)], axis=1),
                    cons=np.stack([col("grid", 1, cname, "cons", j) for j in range(6)], axis=1),
                    y_hat=col("grid", 1, cname, "y_hat"), delta_hat=col("grid", 1, cname, "delta_hat"),
                    z_hat=col("grid", 1, cname, "z_hat"),
                    omega_hat=np.stack([col("grid", 1, cname, "omega_hat", j) for j in range(n_dev)], axis=1),
                    price=G["price"][lag:lag + steps])
        blocks.append(("grid", [1], cname, results.grid_log_blocks(grid, is_mpc=cname != "thermo")))
    df = results.grid_sim_dataframe(blocks, steps)
    got_cols = [tuple(str(x) for x in c) for c in df.columns.tolist()]
    assert got_cols == cols
    keep = np.array([not c[3].startswith("time_") for c in cols])
    np.testing.assert_allclose(df.to_numpy(dtype=float)[:, keep], vals[:, keep], rtol=1e-12, atol=1e-9)
