"""Generates tests/golden/microgrid_loop.npz: the UNMODIFIED reference's whole example loop on a small micro-grid --
GridAgentMpc + DewhAgentMpc x N_h + PvAgentMpc + ResDemandAgentMpc, per instant ``set_device_objective_atoms`` /
``set_std_obj_atoms(q_z=...)`` / ``build_grid`` / ``solve_grid_mpc`` / ``sim_step_k`` exactly as
examples/residential_mg_with_pv_and_dewhs/micro_grid_control_simulation.py:161-236 drives them (the script itself loads
data files that are not in the repository, so the loop is re-typed here around the reference's own classes), for a
certainty-equivalent and a perfect-forecast controller.  cvxpy's modelling layer = oracle/mini_cvxpy.py, MILP backend
= HiGHS (MIPGap 0) instead of Gurobi (MIPGap 1e-2, 20 s).

Run in the build container only:   python tests/golden/make_golden_microgrid.py
"""
import importlib
import itertools
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.simplefilter("ignore")

from oracle import ref_shim  # noqa: E402

R = ref_shim.load_controllers()
ag = R.agents


def main():
    for cls in (ag.DewhAgentMpc, ag.GridAgentMpc, ag.PvAgentMpc, ag.ResDemandAgentMpc):
        cls.delete_all_devices()
    rng = np.random.default_rng(314)
    N_h, N_p, steps = 4, 8, 3
    Nt = N_p + 1
    lag = 96
    n_rows = lag + steps + Nt + 3
    soft_top, soft_bot = 10.0, 1.0
    gp = R.params.grid_param_struct.deepcopy()
    gp.P_g_min, gp.P_g_max = -1e4 * N_h, 1e4 * N_h
    grid = ag.GridAgentMpc(device_id=1, param_struct=gp)
    price = rng.uniform(40.0, 340.0, n_rows) / 3600 / 100 / 1000 * 900.0       # c/kWh -> currency per W per step
    grid.set_price_profile(price_profile=price)
    dewh_params, dewh_profiles, x0 = [], [], []
    for i in range(1, N_h + 1):
        dp = R.params.dewh_param_struct.deepcopy()
        dp.T_h_min, dp.T_h_max = 50.0, 65.0
        dp.P_h_Nom = 3000.0 * (1.0 + 0.1 * rng.uniform(-1, 1))
        dp.m_h = 150.0 * (1.0 + 0.1 * rng.uniform(-1, 1))
        dev = ag.DewhAgentMpc(device_id=i, param_struct=dp)
        prof = rng.uniform(0, 0.02, (n_rows, 1)) * (rng.random((n_rows, 1)) < 0.35)
        dev.set_omega_profile(prof)
        grid.add_device(dev)
        dewh_params.append(dp)
        dewh_profiles.append(prof[:, 0])
        x0.append(float(rng.integers(49, 60)))
    pp = R.params.pv_param_struct.deepcopy()
    pp.P_pv_units = N_h
    pv = ag.PvAgentMpc(device_id=1, param_struct=pp)
    pv_prof = rng.uniform(0, 1, n_rows) * (rng.random(n_rows) < 0.6)
    pv.set_omega_profile(pv_prof.reshape(-1, 1))
    grid.add_device(pv)
    rp = R.params.res_demand_param_struct.deepcopy()
    rp.P_res_units = N_h
    resd = ag.ResDemandAgentMpc(device_id=1, param_struct=rp)
    resd_prof = rng.uniform(0.2, 2.0, n_rows)
    resd.set_omega_profile(resd_prof.reshape(-1, 1))
    grid.add_device(resd)
    dewhs = [d for d in grid.devices if isinstance(d, ag.DewhAgentMpc)]

    thermo_mod = importlib.import_module("examples.residential_mg_with_pv_and_dewhs.theromstat_control")
    no_ctrl = importlib.import_module("controllers.no_controller").NoController
    # name -> is_deterministic, in the order of the reference script (micro_grid_control_simulation.py:144-152, 266)
    controllers = dict(mpc_pb=True, mpc_ce=False, mpc_sb_reduced=False, mpc_sb_full=False, mpc_minmax=False,
                       thermo=False)
    num_scenarios, N_sb_reduced = 5, 4
    scen_days = rng.uniform(0, 0.02, (96, 30)) * (rng.random((96, 30)) < 0.35)
    for dev in grid.devices:
        if isinstance(dev, ag.DewhAgentMpc):
            dev.set_omega_scenarios(omega_scenarios_profile=scen_days.flatten(order="f"))
    min_day, max_day = scen_days.min(axis=1), scen_days.max(axis=1)
    import ast
    path = os.path.join(ref_shim.REFERENCE_ROOT, "examples", "residential_mg_with_pv_and_dewhs",
                        "micro_grid_control_simulation.py")
    fn = [n for n in ast.parse(open(path).read()).body
          if isinstance(n, ast.FunctionDef) and n.name == "get_min_max_dhw_scenario"]
    ns = dict(np=np, steps_per_day=96, atleast_2d_col=importlib.import_module("utils.matrix_utils").atleast_2d_col)
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)
    get_min_max_dhw_scenario = ns["get_min_max_dhw_scenario"]
    scen_log = {c: [] for c in controllers if c.startswith("mpc_sb")}
    seed = 20261018
    for cname in controllers:
        for dev in itertools.chain([grid], grid.devices):
            if cname != "thermo":
                dev.add_controller(cname, R.MpcController, N_p=N_p)
            elif isinstance(dev, ag.DewhAgentMpc):                           # micro_grid_control_simulation.py:169-177
                dev.add_controller(cname, thermo_mod.DewhTheromstatController, N_p=0, N_tilde=1)
            else:
                dev.add_controller(cname, no_ctrl, N_p=0, N_tilde=1)
    for d, x in zip(dewhs, x0):
        d.x_k = x
    keys = ("C_w", "A_h", "U_h", "m_h", "T_w", "T_inf", "P_h_Nom", "T_h_min", "T_h_max", "T_h_Nom", "ts")
    data = dict(N_h=np.array(N_h), N_p=np.array(N_p), steps=np.array(steps), lag=np.array(lag), price=price,
                dewh_params=np.array([[float(p[k]) for k in keys] for p in dewh_params]),
                dewh_profiles=np.array(dewh_profiles), x0=np.array(x0), pv_profile=pv_prof, resd_profile=resd_prof,
                pv_gain=np.array(-float(pp.P_pv_max) * float(pp.P_pv_units)),
                resd_gain=np.array(float(rp.P_res_ave) * float(rp.P_res_units)),
                grid_params=np.array([float(gp.P_g_min), float(gp.P_g_max), float(gp.eps)]),
                soft=np.array([soft_top, soft_bot]), controllers=np.array(list(controllers)),
                deterministic=np.array(list(controllers.values())))
    logs = {c: dict(obj=[], u=[], T=[], T_next=[], mu=[], grid_y=[], grid_z=[], grid_delta=[], cost=[], u_plan=[],
                    z_plan=[], grid_omega=[]) for c in controllers}
    grid.build_grid(k=0, deterministic_or_struct=controllers)
    np.random.seed(seed)                                   # the scenario draws use numpy's global generator
    for k in range(steps):
        prices_tilde = grid.get_price_tilde_k(k=k)
        for cname in controllers:
            if cname == "thermo":
                continue
            for dev in itertools.chain([grid], grid.devices):
                if isinstance(dev, ag.DewhAgentMpc):
                    max_cost = np.sum(prices_tilde[cname]) * R.params.dewh_param_struct.P_h_Nom
                    dev.set_device_objective_atoms(controller_name=cname,
                                                   q_mu=np.hstack([max_cost * soft_top, max_cost * soft_bot]).ravel(order="c"))
                    ctrl = dev.controllers[cname]
                    if cname.startswith("mpc_sb"):         # micro_grid_control_simulation.py:199-213
                        sc = dev.get_omega_tilde_scenario(k, N_tilde=Nt, num_scenarios=num_scenarios)
                        scen_log[cname].append(sc)
                        if cname == "mpc_sb_reduced":
                            ctrl.set_constraints(other_constraints=[
                                ctrl.gen_evo_constraints(N_tilde=N_sb_reduced, omega_scenarios_k=sc)])
                        else:
                            ctrl.set_constraints(other_constraints=[ctrl.gen_evo_constraints(omega_scenarios_k=sc)])
                    elif cname == "mpc_minmax":            # :215-227
                        omega_min, omega_max = get_min_max_dhw_scenario(k=k, N_tilde=Nt, min_dhw_day=min_day,
                                                                        max_dhw_day=max_day)
                        ctrl.set_constraints(other_constraints=[
                            ctrl.gen_evo_constraints(N_tilde=Nt, omega_tilde_k=omega_min),
                            ctrl.gen_evo_constraints(N_tilde=Nt, omega_tilde_k=omega_max)])
                elif isinstance(dev, ag.GridAgentMpc):
                    dev.controllers[cname].set_std_obj_atoms(q_z=prices_tilde[cname])
        grid.build_grid(k=k, deterministic_or_struct=controllers)
        grid.solve_grid_mpc(k=k, verbose=False, TimeLimit=20, MIPGap=0.0)
        plans = {c: dict(u=[np.asarray(d.controllers[c].variables.u.var_N_tilde.value).ravel() for d in dewhs],
                         z=np.asarray(grid.controllers[c].variables.z.var_N_tilde.value).ravel(),
                         obj=float(grid.controllers[c].problem.value)) for c in controllers if c != "thermo"}
        plans["thermo"] = dict(u=[np.full(Nt, np.nan)] * N_h, z=np.full(Nt, np.nan), obj=np.nan)
        T_before = {c: [float(np.asarray(d.controllers[c].x_k.value if hasattr(d.controllers[c].x_k, "value")
                                         else d.controllers[c].x_k).ravel()[0]) for d in dewhs] for c in controllers}
        grid.sim_step_k(k=k)
        for c in controllers:
            lg = logs[c]
            lg["obj"].append(plans[c]["obj"]); lg["u_plan"].append(np.array(plans[c]["u"])); lg["z_plan"].append(plans[c]["z"])
            lg["T"].append(T_before[c])
            lg["u"].append([float(d.sim_logs[c].get(k).u.ravel()[0]) for d in dewhs])
            lg["T_next"].append([float(d.sim_logs[c].get(k).x_k1.ravel()[0]) for d in dewhs])
            lg["mu"].append([d.sim_logs[c].get(k).mu.ravel() for d in dewhs])
            e = grid.sim_logs[c].get(k)
            lg["grid_y"].append(float(e.y.ravel()[0])); lg["grid_z"].append(float(e.z.ravel()[0]))
            lg["grid_delta"].append(float(e.delta.ravel()[0])); lg["cost"].append(float(e.cost.ravel()[0]))
            lg["grid_omega"].append(e.omega.ravel())
        print("k", k, {c: (round(plans[c]["obj"], 6), logs[c]["u"][-1]) for c in controllers})
    for c, lg in logs.items():
        for key, val in lg.items():
            data["%s_%s" % (c, key)] = np.array(val, dtype=float)
    for c, lst in scen_log.items():
        data["scen_" + c] = np.array(lst).reshape(steps, N_h, Nt, num_scenarios)
    data.update(scen_days=scen_days, num_scenarios=np.array(num_scenarios), N_sb_reduced=np.array(N_sb_reduced),
                seed=np.array(seed))
    df = grid.grid_sim_dataframe
    data["frame_columns"] = np.array(["|".join(str(x) for x in col) for col in df.columns])
    data["frame_values"] = np.array(df.values, dtype=float)
    np.savez_compressed(os.path.join(HERE, "microgrid_loop.npz"), **data)
    print("microgrid_loop.npz", df.shape)


if __name__ == "__main__":
    main()
