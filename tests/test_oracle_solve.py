"""CPU: the solve oracle.  The reference pins nothing here (parity unpinned); HiGHS (scipy.optimize.milp), the
exhaustive enumeration and the QP branch-and-bound are checked against each other and against the committed
HiGHS goldens."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn


def _problem(wl, b, atoms=None, **kw):
    Nt = wl["Nt"]
    full, d, vt = omld.complete({k: v[b] for k, v in wl["mats"].items()}, nu_l=1)
    evo = oc.condense(full, d, Nt)
    atoms = atoms or dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b])
    return oa.build_problem(evo, d, vt, Nt, wl["x0"][b], wl["omega"][b], atoms=atoms, **kw), evo, d


def test_variable_layout_and_bounds():
    wl = syn.dewh_batch(1, 3, seed=0)
    prob, evo, d = _problem(wl, 0)
    assert prob.n == 12 and prob.H.shape == (8, 12)
    assert prob.is_bin.tolist() == [True, False, False] * 4           # v(k) = [u, mu0, mu1], binary u
    assert prob.lb.tolist() == [0.0] * 12 and prob.ub[::3].tolist() == [1.0] * 4 and np.isinf(prob.ub[1])
    np.testing.assert_allclose(prob.c[::3], wl["q_u"][0])
    np.testing.assert_allclose(prob.c[1::3], wl["q_mu"][0, 0])


def test_atom_grammar():
    assert oa.parse_atom_key("q_mu") == ("vector", "Linear", "mu", False, "")
    assert oa.parse_atom_key("Q_x") == ("matrix", "Quadratic", "x", False, "")
    assert oa.parse_atom_key("q_L1_du_N_p") == ("vector", "L1", "u", True, "N_p")
    assert oa.parse_atom_key("q_delta") == ("vector", "Linear", "delta", False, "")
    assert oa.parse_atom_key("Q_y_f") == ("matrix", "Quadratic", "y", False, "f")
    with pytest.raises(ValueError):
        oa.parse_atom_key("q_foo")


@pytest.mark.parametrize("N_p", [6, 9])
def test_highs_equals_enumeration(N_p):
    wl = syn.dewh_batch(4, N_p, seed=3)
    for b in range(4):
        prob, _, _ = _problem(wl, b)
        st, obj, v = osv.solve_milp(prob)
        st2, obj2, v2, second = osv.solve_enumerate(prob)
        assert st == st2 == osv.OPTIMAL
        assert abs(obj - obj2) <= 1e-7 * max(1, abs(obj2))
        if second - obj2 > 1e-6:  # unique optimum -> identical decisions
            assert np.array_equal(np.round(v[prob.is_bin]), np.round(v2[prob.is_bin]))


def test_miqp_bnb_equals_enumeration():
    wl = syn.dewh_batch(2, 6, seed=4)
    for b in range(2):
        prob, _, _ = _problem(wl, b, atoms=dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b], Q_x=np.array([[0.01]])))
        assert prob.P is not None
        st, obj, v, nodes = osv.solve_miqp(prob)
        st2, obj2, v2, _ = osv.solve_enumerate(prob)
        assert st == st2 == osv.OPTIMAL and abs(obj - obj2) <= 1e-6 * max(1, abs(obj2))


def test_scenario_rhs_is_rowwise_min():
    wl = syn.dewh_batch(1, 5, seed=5)
    prob, evo, d = _problem(wl, 0)
    rng = np.random.default_rng(0)
    W = np.abs(rng.standard_normal((wl["Nt"], 7))) * 0.01
    H, rhs = oa.evo_rhs(evo, d, wl["x0"][0], omega_scenarios=W)
    each = np.stack([oa.evo_rhs(evo, d, wl["x0"][0], W[:, s])[1] for s in range(7)], axis=1)
    np.testing.assert_allclose(rhs, each.min(axis=1))
    H2, rhs2 = oa.evo_rhs(evo, d, wl["x0"][0], W[:, 0], N_tilde=3)
    assert H2.shape[0] == 6 and np.allclose(rhs2, each[:6, 0])


@pytest.mark.parametrize("N_p", [24, 48])
def test_milp_goldens_reproduce(N_p):
    g = load_golden("milp", "dewh_N%d" % N_p)
    wl = syn.dewh_batch(int(g["B"]), N_p, seed=int(g["seed"]))
    for b in range(0, int(g["B"]), 4):
        prob, _, _ = _problem(wl, b)
        st, obj, v = osv.solve_milp(prob)
        assert abs(obj - g["obj"][b]) <= 1e-9 * max(1, abs(obj))
