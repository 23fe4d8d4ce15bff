"""CPU: the input side of the example's loop (examples/.../profiles.py) against vectors produced by the UNMODIFIED
reference's agents and script helpers (tests/golden/make_golden_profiles.py): forecast / actual windows, scenario
table and random scenario draws (same numpy.random seed -> same draws), price windows, min / max windows, device
profiles and initial states.  Index arithmetic only, so everything must be bit-identical."""
import os

import numpy as np
import pytest

from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import profiles as P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "profiles_inputs.npz"))
TS = float(G["ts"])
HORIZONS = dict(zip([str(n) for n in G["horizon_names"]], [int(v) for v in G["horizon_values"]]))


def _cmp_padded(got, ref):
    """reference windows were stored NaN-padded where the profile ends"""
    n = got.shape[-1]
    assert np.array_equal(got, ref[..., :n]) and np.isnan(ref[..., n:]).all()


def test_lag_steps():
    assert P.lag_steps(str(G["forecast_lag"]), TS) == 96
    assert P.lag_steps("1D", 900) == 96 and P.lag_steps("6H", 900) == 24 and P.lag_steps("30min", 900) == 2
    assert P.lag_steps(1800, 900) == 2 and P.lag_steps("0D", 900) == 0
    import datetime
    assert P.lag_steps(datetime.timedelta(days=2), 900) == 192
    with pytest.raises(ValueError):
        P.lag_steps("tomorrow", 900)


def test_forecast_and_actual_windows():
    prof = P.OmegaProfiles(G["profiles"], TS, forecast_lag=str(G["forecast_lag"]))
    assert prof.B == 3 and prof.lag == 96
    for cname, nt in HORIZONS.items():
        for i, k in enumerate(G["ks"]):
            _cmp_padded(prof.omega_tilde_k_act(k, nt), G["act_" + cname][i])
            _cmp_padded(prof.omega_tilde_k_hat(k, nt, deterministic=(cname == "mpc_pb")), G["hat_" + cname][i])
    assert np.array_equal(prof.omega_k_act(5)[:, 0], G["act_mpc_ce"][2][:, 0])
    with pytest.raises(ValueError, match="nomega"):
        P.OmegaProfiles(np.zeros((10, 2)), TS)


def test_closed_loop_arrays_reproduce_the_windows():
    prof = P.OmegaProfiles(G["profiles"], TS)
    sim_steps, nt = 200, 49
    arr = prof.closed_loop_arrays(sim_steps, nt)
    assert arr["demand"].shape == arr["demand_actual"].shape == (3, sim_steps + nt)
    for k in (0, 1, 95, 150, 199):
        assert np.array_equal(arr["demand"][:, k:k + nt], prof.omega_tilde_k_hat(k, nt))             # what closed_loop cuts
        assert np.array_equal(arr["demand_actual"][:, k:k + nt], prof.omega_tilde_k_act(k, nt))      # mpc_pb forecast
        assert np.array_equal(arr["demand_actual"][:, k], prof.omega_k_act(k)[:, 0])                 # applied draw
    with pytest.raises(ValueError, match="too short"):
        prof.closed_loop_arrays(400, nt)


def test_two_disturbances_per_step():
    prof = P.OmegaProfiles(G["profile_two"], TS, nomega=2)
    for i, k in enumerate((0, 3, 20)):
        assert np.array_equal(prof.omega_tilde_k_act(k, 10)[0], G["act_two"][i])
        assert np.array_equal(prof.omega_tilde_k_hat(k, 10)[0], G["hat_two"][i])
    sc = P.OmegaScenarios(G["scenario_profile_two"], TS, nomega=2)
    assert np.array_equal(sc.table, G["scenario_table_two"])
    np.random.seed(3)
    assert np.array_equal(sc.omega_tilde_scenario(50, 10, 3), G["draw_two"])
    with pytest.raises(ValueError):
        prof.closed_loop_arrays(5, 5)


def test_scenario_table_and_draws():
    sc = P.OmegaScenarios(G["scenario_days"].flatten(order="F"), TS)
    assert (sc.intervals_per_day, sc.num_scenarios) == (int(G["intervals_per_day"]), int(G["num_scenarios"]))
    assert np.array_equal(sc.table, G["scenario_table"]) and sc.table.flags.f_contiguous
    for i, (k, nt, ns, seed) in enumerate(G["draw_cases"]):
        np.random.seed(int(seed))                        # the reference draws from the global generator
        assert np.array_equal(sc.omega_tilde_scenario(k, nt, ns), G["draw_%d" % i]), i
        rs = np.random.RandomState(int(seed))            # ... an explicit RandomState gives the same columns
        assert np.array_equal(sc.omega_tilde_scenario(k, nt, ns, random_state=rs), G["draw_%d" % i]), i
    np.random.seed(11)
    fleet = sc.fleet_scenarios(30, 49, 6, B=3)
    assert fleet.shape == (3, 49, 6) and np.array_equal(fleet, G["fleet_draw"])
    assert bool(G["insufficient_raises"])
    with pytest.raises(ValueError, match="Insufficient number of scenarios"):
        sc.omega_tilde_scenario(0, 97, 40)
    with pytest.raises(ValueError):
        P.OmegaScenarios(np.zeros(100), TS)              # not whole days
    lo, hi = sc.min_max_day()
    assert np.array_equal(lo, G["scenario_days"].min(axis=1)) and np.array_equal(hi, G["scenario_days"].max(axis=1))


def test_price_windows():
    pr = P.PriceProfile(G["price"], TS)
    for cname, nt in HORIZONS.items():
        for i, k in enumerate((0, 4, 100)):
            got = pr.price_tilde_k(k, nt)
            assert np.array_equal(got, G["price_" + cname][i][:got.shape[0]])
    arr = pr.closed_loop_array(100, 49)
    for k in (0, 4, 99):
        assert np.array_equal(arr[k:k + 49], pr.price_tilde_k(k, 49))
    with pytest.raises(ValueError, match="too short"):
        pr.closed_loop_array(300, 49)


def test_script_helpers():
    prof = P.get_actual_omega_dewh_profiles(G["actual_scenarios"], N_h=5, size=12)
    assert sorted(prof) == [1, 2, 3, 4, 5]
    for i in range(1, 6):
        assert prof[i].shape == (96 * 12, 1) and np.array_equal(prof[i][:, 0], G["actual_profiles"][i - 1])
    assert [P.get_dewh_random_initial_state(i) for i in range(1, 41)] == G["initial_states"].tolist()
    lo_day, hi_day = G["scenario_days"].min(axis=1), G["scenario_days"].max(axis=1)
    for i, (k, nt) in enumerate(G["minmax_cases"]):
        lo, hi = P.get_min_max_dhw_scenario(k, nt, lo_day, hi_day)
        assert lo.shape == hi.shape == (nt, 1)
        assert np.array_equal(np.hstack([lo, hi]), G["minmax_%d" % i])
    with pytest.raises(ValueError, match="min_dhw_day"):
        P.get_min_max_dhw_scenario(0, 10, lo_day[:-1], hi_day)
    with pytest.raises(ValueError, match="max_dhw_day"):
        P.get_min_max_dhw_scenario(0, 10, lo_day, hi_day[:-1])
    mn, mx = P.min_max_closed_loop_arrays(300, 49, lo_day, hi_day)
    for k in (0, 5, 95, 100, 299):
        lo, hi = P.get_min_max_dhw_scenario(k, 49, lo_day, hi_day)
        assert np.array_equal(mn[k:k + 49], lo[:, 0]) and np.array_equal(mx[k:k + 49], hi[:, 0])


def test_time_of_use_tariff():
    import datetime
    rates = dict(low_off_peak=48.40, low_stnd=76.28, low_peak=110.84, high_off_peak=55.90, high_stnd=102.95,
                 high_peak=339.77)
    for i, st in enumerate(G["tariff_starts"]):
        got = P.tou_price_vector(datetime.datetime(*[int(v) for v in st]), 96 * 9, 900, **rates)
        assert got.shape == (96 * 9, 1) and np.array_equal(got, G["tariff_%d" % i]), i
    got = P.tou_price_vector(datetime.datetime(2019, 6, 1), 24 * 8, datetime.timedelta(hours=1), **rates)
    assert np.array_equal(got, G["tariff_hourly"])
    assert set(np.unique(got)) <= set(rates.values())
    with pytest.raises(TypeError):
        P.tou_price_vector("2019-06-01", 4, 900)
