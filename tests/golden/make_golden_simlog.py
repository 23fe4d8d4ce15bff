"""Generates tests/golden/simlog_campaign.npz: the frame the UNMODIFIED reference builds from a small closed-loop
log -- ``MldSimLog`` (controllers/controller_base.py:58-146) filled the way ``ControllerBase.sim_step_k`` (:229-253)
and ``GridAgentMpc.sim_step_k`` (micro_grid_agents.py:736-756) fill it, entries produced by the reference's own
``MldModel.lsim_k``, then concatenated with the statements of ``sim_dataframe`` (micro_grid_agents.py:142-150) and
``grid_sim_dataframe`` (:522-525) and re-indexed as the campaign script does
(micro_grid_control_simulation.py:246-247).  Run here (needs /root/reference); the fixture travels, the reference
does not.

    python tests/golden/make_golden_simlog.py
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim, lsim as ol  # noqa: E402


def synthetic_log(rng, steps, B):
    """numpy arrays with the keys DewhFleet.closed_loop returns (values are arbitrary but self-consistent)."""
    T = 55.0 + 12.0 * rng.random((steps + 1, B))
    T[1, 0] = 10.0                                  # below T_w: the sim step clamps it (micro_grid_agents.py:398-399)
    T[2, 1] = 70.0                                  # above T_h_max: a violated row and a positive slack
    u = (rng.random((steps, B)) > 0.5).astype(float)
    return dict(T=T, u=u, mu_hat=rng.random((steps, B, 2)) * (rng.random((steps, B, 2)) > 0.6),
                omega=0.01 * rng.random((steps, B)), omega_hat=0.01 * rng.random((steps, B)),
                solve_ms=1.0 + rng.random(steps))


def main():
    warnings.simplefilter("ignore")
    import pandas as pd
    MldModel, _ = ref_shim.load()
    from controllers.controller_base import MldSimLog
    from controllers.components.variables import VariablesStruct_k
    rng = np.random.default_rng(20261018)
    steps, dewh_ids = 4, [7, 3]                     # deliberately not sorted: the grid orders devices by id
    B = len(dewh_ids)
    params = [dict(ol.DEWH_PARAMS, P_h_Nom=3000.0 + 100.0 * b) for b in range(B)]
    logs = {"mpc_pb": synthetic_log(rng, steps, B), "thermo": synthetic_log(rng, steps, B)}
    pv = dict(omega=rng.random(steps), omega_hat=rng.random(steps), gain=-2000.0 * B)
    resd = dict(omega=rng.random(steps) * 2, omega_hat=rng.random(steps) * 2, gain=1200.0 * B)
    price = rng.random(steps) * 1e-4
    times = {c: rng.random((steps, 2)) for c in ("grid", "pv", "resd")}
    col = lambda a: np.asarray(a, dtype=float).reshape(-1, 1)

    def hats(**kw):
        return {n + "_hat": v for n, v in VariablesStruct_k(**kw).items()}

    def concat(device_type, device_id, sim_logs):
        dfs = {}
        for cname, log in sim_logs.items():
            dfs[cname] = pd.concat([log.get_concat_log()], keys=[(device_type, device_id, cname)],
                                   names=["device_type", "device_id", "controller"], axis=1)
        return pd.concat(dfs.values(), axis=1)

    frames = {}
    # ---- DEWHs
    for b, dev_id in enumerate(dewh_ids):
        p = params[b]
        sim_logs = {}
        for cname, lg in logs.items():
            sim_logs[cname] = sl = MldSimLog()
            for k in range(steps):
                x_ctrl = lg["T"][k, b]
                x = p["T_w"] + 0.1 if x_ctrl <= p["T_w"] else x_ctrl
                mld = MldModel(**{n: np.array(v, dtype=float) for n, v in
                                  ol.dewh_mld(p, const_heat=False, T_h=x, D_h=lg["omega"][k, b]).items()}, nu_l=1)
                mu = np.array([max(0.0, x - p["T_h_max"]), max(0.0, p["T_h_min"] - x)])
                ls = mld.lsim_k(x_k=x, u_k=lg["u"][k, b], omega_k=lg["omega"][k, b], mu_k=mu)
                ls["x_k1"] = col(lg["T"][k + 1, b])          # the fleet's own next state (checked elsewhere)
                if cname == "thermo":                        # theromstat_control.py:62 keeps the simulated step
                    hk = mld.lsim_k(x_k=x_ctrl, u_k=lg["u"][k, b], omega_k=lg["omega_hat"][k, b], mu_k=mu)
                    del hk["x_k1"]
                    hk["mu"], hk["v"], hk["cons"] = ls["mu"], ls["v"], ls["cons"]
                    var_hat = {n + "_hat": v for n, v in hk.items()}
                else:
                    mh = lg["mu_hat"][k, b]
                    var_hat = hats(x=col(x_ctrl), u=col(lg["u"][k, b]), delta=np.zeros((0, 1)), z=np.zeros((0, 1)),
                                   omega=col(lg["omega_hat"][k, b]), y=col(x_ctrl), mu=col(mh),
                                   v=col(np.concatenate([[lg["u"][k, b]], mh])))
                ls.update(var_hat)
                sl.set_sim_k(k=k, sim_k=ls)
                sl.update_sim_k(k=k, time_solve_overall=lg["solve_ms"][k] * 1e-3, time_in_solver=lg["solve_ms"][k] * 1e-3)
        frames[("dewh", dev_id)] = concat("dewh", dev_id, sim_logs)
    # ---- PV / residential demand (MPC controller only: the NoController columns need cvxpy to pin)
    for name, src in (("pv", pv), ("resd", resd)):
        mld = MldModel(D4=[[src["gain"]]])
        sl = MldSimLog()
        for k in range(steps):
            ls = mld.lsim_k(omega_k=src["omega"][k], u_k=None, mu_k=None)
            e = np.zeros((0, 1))
            ls.update(hats(x=e, u=e, delta=e, z=e, omega=col(src["omega_hat"][k]), y=col(src["gain"] * src["omega_hat"][k]),
                           mu=e, v=e))
            sl.set_sim_k(k=k, sim_k=ls)
            sl.update_sim_k(k=k, time_solve_overall=times[name][k, 0], time_in_solver=times[name][k, 1])
        frames[(name, 1)] = concat(name, 1, {"mpc_pb": sl})
    # ---- grid: device powers ordered by (type, id) (micro_grid_agents.py:551-556)
    order = np.argsort(dewh_ids)
    lg = logs["mpc_pb"]
    P_nom = np.array([p["P_h_Nom"] for p in params])
    gm = ol.grid_mld(ol.GRID_PARAMS, B + 2)
    grid_mld = MldModel(**{n: np.array(v, dtype=float) for n, v in gm.items()})
    sl = MldSimLog()
    grid_in = dict(omega=[], omega_hat=[])
    for k in range(steps):
        w = np.concatenate([(lg["u"][k] * P_nom)[order], [pv["gain"] * pv["omega"][k]], [resd["gain"] * resd["omega"][k]]])
        wh = np.concatenate([(lg["u"][k] * P_nom)[order], [pv["gain"] * pv["omega_hat"][k]],
                             [resd["gain"] * resd["omega_hat"][k]]])
        y, yh = float(np.ones(B + 2) @ w), float(np.ones(B + 2) @ wh)
        d, z = ol.grid_aux_closed_form(y)
        dh, zh = ol.grid_aux_closed_form(yh)
        ls = grid_mld.lsim_k(omega_k=w, u_k=None, delta_k=d, z_k=z, mu_k=None)
        e = np.zeros((0, 1))
        ls.update(hats(x=e, u=e, delta=col(dh), z=col(zh), omega=col(wh), y=col(yh), mu=e, v=col([dh, zh])))
        sl.set_sim_k(k=k, sim_k=ls)
        sl.update_sim_k(k=k, time_solve_overall=times["grid"][k, 0], time_in_solver=times["grid"][k, 1])
        add = dict(p_imp=ls.z, p_exp=ls.y - ls.z, cost=ls.z * price[k])          # micro_grid_agents.py:750-756
        sl.update_sim_k(k=k, sim_k=add)
    frames[("grid", 1)] = concat("grid", 1, {"mpc_pb": sl})
    # ---- GridAgentMpc.grid_sim_dataframe: the grid itself, then its devices ordered by (type, id)
    dev_keys = sorted(k for k in frames if k[0] != "grid")
    df = pd.concat([frames[("grid", 1)]] + [frames[k] for k in dev_keys], axis=1)
    time_0 = "2018-12-10 00:00:00"
    df.index = pd.date_range(start=time_0, periods=steps, freq="15min")
    out = dict(values=df.to_numpy(dtype=float), columns=json.dumps([list(map(lambda x: x if isinstance(x, str) else int(x), c))
                                                                   for c in df.columns.tolist()]),
               column_names=json.dumps(list(df.columns.names)), index=np.array([str(t) for t in df.index]),
               time_0=time_0, dewh_ids=np.array(dewh_ids), P_h_Nom=P_nom, price=price)
    for cname, lgc in logs.items():
        for key, val in lgc.items():
            out["log_%s_%s" % (cname, key)] = val
    for name, src in (("pv", pv), ("resd", resd)):
        out[name + "_omega"], out[name + "_omega_hat"], out[name + "_gain"] = src["omega"], src["omega_hat"], src["gain"]
    for name, t in times.items():
        out["times_" + name] = t
    np.savez_compressed(os.path.join(HERE, "simlog_campaign.npz"), **out)
    print(df.shape, df.columns.names)
    print(df.iloc[:, :12])


if __name__ == "__main__":
    main()
