"""Generates tests/golden/profiles_inputs.npz from the UNMODIFIED reference: the profile / scenario / price window
methods of its agents (examples/.../modelling/micro_grid_agents.py:156-298, 563-606) and the three helper functions of
its simulation script (examples/.../micro_grid_control_simulation.py:56-83), run under oracle/ref_shim.

The agents are built with the reference's own constructors; their controllers are stand-ins that only carry
``N_tilde`` (a real MpcController needs cvxpy), which is all these methods read.  The script itself cannot be imported
(it loads data files that are not in the repository at import time), so the three helpers are taken from its source by
name (ast) and executed unchanged.

Run in the build container only:   python tests/golden/make_golden_profiles.py
"""
import ast
import importlib
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.simplefilter("ignore")

from oracle import ref_shim  # noqa: E402

ref_shim.load_symbolic()
# numpy 2 renamed unravel_index(dims=) to shape=; environment alias, like the others in the shim
_unravel = np.unravel_index
np.unravel_index = lambda indices, shape=None, order="C", dims=None: _unravel(
    indices, shape if shape is not None else dims, order=order)
import pandas as pd  # noqa: E402

ag = importlib.import_module("examples.residential_mg_with_pv_and_dewhs.modelling.micro_grid_agents")
from structdict import StructDict  # noqa: E402
from utils.matrix_utils import atleast_2d_col  # noqa: E402


def script_helpers():
    path = os.path.join(ref_shim.REFERENCE_ROOT, "examples", "residential_mg_with_pv_and_dewhs",
                        "micro_grid_control_simulation.py")
    tree = ast.parse(open(path).read())
    want = ("get_actual_omega_dewh_profiles", "get_dewh_random_initial_state", "get_min_max_dhw_scenario")
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = dict(np=np, pd=pd, StructDict=StructDict, atleast_2d_col=atleast_2d_col, steps_per_day=96)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return [ns[w] for w in want]


def main():
    rng = np.random.default_rng(424242)
    ts, ipd = 900.0, 96
    data = dict(ts=np.array(ts))
    for dev in (ag.DewhAgentMpc, ag.GridAgentMpc, ag.PvAgentMpc, ag.ResDemandAgentMpc):
        dev.delete_all_devices()

    # ---- disturbance profile windows of three water heaters with different horizons per controller
    n_rows, B = ipd * 4 + 17, 3
    profiles = rng.uniform(0, 0.02, size=(B, n_rows, 1)) * (rng.random((B, n_rows, 1)) < 0.3)
    data["profiles"] = profiles
    ks = np.array([0, 1, 5, 95, 96, 150, 250])
    horizons = dict(mpc_pb=25, mpc_ce=49, mpc_long=97)
    data["ks"], data["horizon_names"] = ks, np.array(list(horizons))
    data["horizon_values"] = np.array(list(horizons.values()))
    dewhs = []
    for b in range(B):
        d = ag.DewhAgentMpc(device_id=b + 1)
        d.set_omega_profile(profiles[b])
        for cname, nt in horizons.items():
            d._controllers[cname] = types.SimpleNamespace(N_tilde=nt)
        dewhs.append(d)
    det = dict(mpc_pb=True, mpc_ce=False, mpc_long=False)
    for cname, nt in horizons.items():
        act = np.full((len(ks), B, nt), np.nan)
        hat = np.full((len(ks), B, nt), np.nan)
        for i, k in enumerate(ks):
            for b, d in enumerate(dewhs):
                a = d.get_omega_tilde_k_act(int(k))[cname]
                h = d.get_omega_tilde_k_hat(int(k), deterministic_or_struct=det)[cname]
                act[i, b, :a.shape[0]] = a[:, 0]
                hat[i, b, :h.shape[0]] = h[:, 0]
        data["act_" + cname], data["hat_" + cname] = act, hat
    data["forecast_lag"] = np.array(dewhs[0].forecast_lag)

    # ---- scenario table and random scenario draws (global numpy.random state, as in the reference)
    scen_days = rng.uniform(0, 0.03, size=(ipd, 40)) * (rng.random((ipd, 40)) < 0.3)
    data["scenario_days"] = scen_days
    d = dewhs[0]
    d.set_omega_scenarios(omega_scenarios_profile=scen_days.flatten(order="f"))
    data["scenario_table"] = np.asarray(d.omega_scenarios.values)
    data["intervals_per_day"], data["num_scenarios"] = np.array(d.intervals_per_day), np.array(d.num_scenarios)
    draws = []
    cases = [(0, 49, 20, 1), (7, 49, 20, 2), (95, 25, 4, 3), (200, 97, 8, 4), (96 * 3 + 5, 13, 32, 5)]
    for k, nt, ns, seed in cases:
        np.random.seed(seed)
        draws.append(d.get_omega_tilde_scenario(k, N_tilde=nt, num_scenarios=ns))
    data["draw_cases"] = np.array(cases)
    for i, dr in enumerate(draws):
        data["draw_%d" % i] = dr
    np.random.seed(11)                                   # a fleet of 3 devices drawing one after the other
    data["fleet_draw"] = np.stack([d.get_omega_tilde_scenario(30, N_tilde=49, num_scenarios=6) for _ in range(3)])
    try:
        d.get_omega_tilde_scenario(0, N_tilde=97, num_scenarios=40)
        data["insufficient_raises"] = np.array(False)
    except ValueError:
        data["insufficient_raises"] = np.array(True)

    # a two-disturbance device: windows and scenario table with nomega = 2
    class TwoOmega(object):
        mld_info = types.SimpleNamespace(nomega=2, ts=ts)
        profile_t0 = dewhs[0].profile_t0
        scenarios_t0 = dewhs[0].scenarios_t0
        forecast_lag = "1D"
        omega_profile = None
        omega_scenarios = None
        controllers = _controllers = dict(c=types.SimpleNamespace(N_tilde=10))
        N_tilde = dict(c=10)
        OmegaTildeKActStruct = ag.MicroGridAgentBase.OmegaTildeKActStruct
        OmegaTildeKHatStruct = ag.MicroGridAgentBase.OmegaTildeKHatStruct
    two = TwoOmega()
    prof2 = rng.uniform(0, 1, size=(ipd * 2 + 30, 2))
    data["profile_two"] = prof2
    ag.MicroGridAgentBase.set_omega_profile(two, prof2)
    two.get_omega_tilde_k_act = types.MethodType(ag.MicroGridAgentBase.get_omega_tilde_k_act, two)
    data["act_two"] = np.stack([ag.MicroGridAgentBase.get_omega_tilde_k_act(two, k)["c"][:, 0] for k in (0, 3, 20)])
    data["hat_two"] = np.stack([ag.MicroGridAgentBase.get_omega_tilde_k_hat(two, k)["c"][:, 0] for k in (0, 3, 20)])
    scen2 = rng.uniform(0, 1, size=(ipd * 6, 2))
    data["scenario_profile_two"] = scen2
    ag.MicroGridAgentBase.set_omega_scenarios(two, scen2)
    data["scenario_table_two"] = np.asarray(two.omega_scenarios.values)
    np.random.seed(3)
    data["draw_two"] = ag.MicroGridAgentBase.get_omega_tilde_scenario(two, 50, N_tilde=10, num_scenarios=3)

    # ---- price windows of the grid agent
    grid = ag.GridAgentMpc(device_id=1)
    price = rng.uniform(0.1, 3.0, size=ipd * 4)
    data["price"] = price
    grid.set_price_profile(price_profile=price)
    for cname, nt in horizons.items():
        grid._controllers[cname] = types.SimpleNamespace(N_tilde=nt)
    for cname, nt in horizons.items():
        data["price_" + cname] = np.stack([grid.get_price_tilde_k(int(k))[cname][:, 0] for k in (0, 4, 100)])

    # ---- helpers of the simulation script
    get_actual, get_init, get_minmax = script_helpers()
    actual_scen = rng.uniform(0, 0.02, size=(ipd, 50))
    data["actual_scenarios"] = actual_scen
    prof = get_actual(actual_scenarios=actual_scen, N_h=5, size=12)
    data["actual_profiles"] = np.stack([prof[i][:, 0] for i in range(1, 6)])
    data["initial_states"] = np.array([get_init(i) for i in range(1, 41)])
    mn, mx = scen_days.min(axis=1), scen_days.max(axis=1)
    mm_cases = [(0, 49), (5, 49), (95, 97), (100, 25), (191, 193)]
    data["minmax_cases"] = np.array(mm_cases)
    for i, (k, nt) in enumerate(mm_cases):
        lo, hi = get_minmax(k=k, N_tilde=nt, min_dhw_day=mn, max_dhw_day=mx)
        data["minmax_%d" % i] = np.hstack([lo, hi])
    # ---- time-of-use tariff (examples/.../tariff_generator.py)
    from datetime import datetime as DateTime
    tg = importlib.import_module("examples.residential_mg_with_pv_and_dewhs.tariff_generator")
    gen = tg.TariffGenerator(low_off_peak=48.40, low_stnd=76.28, low_peak=110.84, high_off_peak=55.90,
                             high_stnd=102.95, high_peak=339.77)
    starts = [(2018, 12, 10, 0, 0), (2019, 5, 30, 17, 45), (2019, 8, 29, 3, 15), (2020, 2, 27, 23, 0)]
    data["tariff_starts"] = np.array(starts)
    for i, st in enumerate(starts):
        data["tariff_%d" % i] = np.asarray(gen.get_price_vector(DateTime(*st), 96 * 9, 900), dtype=float)
    data["tariff_hourly"] = np.asarray(gen.get_price_vector(DateTime(2019, 6, 1), 24 * 8, 3600.0), dtype=float)
    np.savez_compressed(os.path.join(HERE, "profiles_inputs.npz"), **data)
    print("profiles_inputs.npz", len(data), "arrays")


if __name__ == "__main__":
    main()
