"""CPU: the result frame against a fixture produced by the unmodified reference's MldSimLog /
lsim_k / concat statements (tests/golden/make_golden_simlog.py): same columns in the same order, same index, same
values."""
import json
import os

import numpy as np
import pytest

from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import results

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simlog_campaign.npz")


def load_golden():
    g = np.load(GOLD, allow_pickle=False)
    logs = {}
    for key in g.files:
        if key.startswith("log_"):
            rest = key[4:]
            cname = "mpc_pb" if rest.startswith("mpc_pb_") else "thermo"
            logs.setdefault(cname, {})[rest[len(cname) + 1:]] = g[key]
    return g, logs


def dewh_params(g):
    from oracle import lsim as ol
    return [dict(ol.DEWH_PARAMS, P_h_Nom=float(p)) for p in g["P_h_Nom"]]


def sim_cons(log, params):
    """cons of the DEWH sim step (what closed_loop logs) from the golden's temperatures"""
    T = log["T"][:-1]
    x = np.where(T <= np.array([p["T_w"] for p in params]), np.array([p["T_w"] for p in params]) + 0.1, T)
    return np.stack([x - np.array([p["T_h_max"] for p in params]) <= 1e-6,
                     -x + np.array([p["T_h_min"] for p in params]) <= 1e-6], axis=2)


def grid_arrays(g, logs, params):
    """numpy twin of DewhFleet.grid_log (the GPU test compares the two)"""
    from oracle import lsim as ol, mld as omld
    lg = logs["mpc_pb"]
    order = np.argsort(g["dewh_ids"])
    steps = lg["u"].shape[0]
    P = np.array([p["P_h_Nom"] for p in params])
    out = {}
    for tag in ("", "_hat"):
        w = np.concatenate([(lg["u"] * P)[:, order], float(g["pv_gain"]) * g["pv_omega" + tag][:, None],
                            float(g["resd_gain"]) * g["resd_omega" + tag][:, None]], axis=1)
        full, d, _ = omld.complete({k: np.array(v, dtype=float) for k, v in ol.grid_mld(ol.GRID_PARAMS, w.shape[1]).items()})
        ys, ds, zs, cs = [], [], [], []
        for k in range(steps):
            y0 = float(np.ones(w.shape[1]) @ w[k])
            de, z = ol.grid_aux_closed_form(y0)
            _, y, cons = ol.lsim_k(full, np.zeros((0, 1)), np.zeros((0, 1)), np.array([[de]]), np.array([[z]]),
                                   np.zeros((0, 1)), w[k].reshape(-1, 1))
            ys.append(float(y.ravel()[0])); ds.append(de); zs.append(z); cs.append(np.asarray(cons).ravel())
        out.update({"omega" + tag: w, "y" + tag: np.array(ys), "delta" + tag: np.array(ds), "z" + tag: np.array(zs)})
        if not tag:
            out["cons"] = np.array(cs)
    out["price"] = g["price"]
    return out


def build_frame(g, logs, params, grid):
    ids = [int(i) for i in g["dewh_ids"]]
    steps = logs["mpc_pb"]["u"].shape[0]
    blocks = []
    for cname, lg in logs.items():
        lg = dict(lg, cons=sim_cons(lg, params))
        blocks.append(("dewh", ids, cname, results.dewh_log_blocks(lg, params, cname)))
    blocks.append(("resd", [1], "mpc_pb", results.source_log_blocks(g["resd_omega"], g["resd_omega_hat"], float(g["resd_gain"]),
                                                                  times=g["times_resd"])))
    blocks.append(("pv", [1], "mpc_pb", results.source_log_blocks(g["pv_omega"], g["pv_omega_hat"], float(g["pv_gain"]),
                                                                times=g["times_pv"])))
    blocks.append(("grid", [1], "mpc_pb", results.grid_log_blocks(grid, times=g["times_grid"])))
    return results.grid_sim_dataframe(blocks, steps, time_0=str(g["time_0"]))


def test_frame_matches_reference_layout_and_values():
    g, logs = load_golden()
    params = dewh_params(g)
    df = build_frame(g, logs, params, grid_arrays(g, logs, params))
    want_cols = [tuple(c) for c in json.loads(str(g["columns"]))]
    assert list(df.columns.names) == json.loads(str(g["column_names"]))
    assert [tuple(c) for c in df.columns.tolist()] == want_cols
    assert [str(t) for t in df.index] == [str(t) for t in g["index"]]
    np.testing.assert_allclose(df.to_numpy(dtype=float), g["values"], rtol=1e-12, atol=1e-12, equal_nan=True)
    # device order: grid, then (type, id) sorted; dewh 3 before dewh 7 although the fleet held 7 first
    firsts = []
    for c in want_cols:
        if not firsts or firsts[-1] != c[:2]:
            firsts.append(c[:2])
    assert firsts == [("grid", 1), ("dewh", 3), ("dewh", 7), ("pv", 1), ("resd", 1)]


def test_plain_k_index_and_sim_out_name(tmp_path):
    g, logs = load_golden()
    params = dewh_params(g)
    lg = dict(logs["mpc_pb"], cons=sim_cons(logs["mpc_pb"], params))
    df = results.grid_sim_dataframe([("dewh", [1, 2], "mpc_ce", results.dewh_log_blocks(lg, params, "mpc_ce"))], 4)
    assert df.index.name == "k" and list(df.index) == [0, 1, 2, 3]
    assert df[("dewh", 2, "mpc_ce", "mu_hat", 1)].shape == (4,)
    path = results.sim_out_path(str(tmp_path), 24, 10.0, 1.0, 20, 8, 50, 65.0, 50.0, "test")
    # micro_grid_control_simulation.py:257-259
    assert os.path.basename(path) == "sim_Np_24_st_10_sb_1_Ns_20_Nsr_8_Nh_50_Tmax_65_Tmin_50_test.sim_out"
    assert os.path.basename(os.path.dirname(path)) == "sim_out"
    results.save_sim_out(df, path)
    import pandas as pd
    back = pd.read_pickle(path)
    assert back.equals(df)


def test_shape_errors():
    with pytest.raises(ValueError):
        results.grid_sim_dataframe([("pv", [1], "mpc", results.source_log_blocks(np.zeros(3), np.zeros(3), 1.0))], 4)
