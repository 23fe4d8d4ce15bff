"""GPU: the CUDA path against numbers produced by the UNMODIFIED reference's MpcController (build / solve / feedback /
sim_step_k under oracle/ref_shim.load_controllers; fixtures tests/golden/assembly_*.npz, generator
tests/golden/make_golden_assembly.py).  The reference's MILP backend there is HiGHS, not Gurobi: objectives are
compared at BASELINE.json's 1e-6, decisions where the optimum is unique.

(File name sorts last on purpose: these checks were added after the round's GPU budget was spent and have only been
exercised against the oracle on the CPU -- tests/test_oracle_assembly_pinned.py -- so they run after everything else.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEWH_MATS = ("A", "B1", "B4", "b5", "E", "F1", "Psi", "f5")


def _fixture(case):
    z = np.load(os.path.join(GOLDEN, "assembly_%s.npz" % case))
    return {k: z[k] for k in z.files}


# the general branch-and-cut kernel is exercised on the small instance only: the N_p = 48 fixture is a cold start
# (slack unavoidable), the kind of instance on which that kernel's node count explodes
@pytest.mark.parametrize("case,solver", [("dewh_N8_linear", "stage_dp"), ("dewh_N8_linear", "bnc"),
                                         ("dewh_N48_linear", "stage_dp"),
                                         ("dewh_N12_scenarios_minmax", "stage_dp")])
def test_control_instant_vs_reference(case, solver, cuda_device):
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.batch import BatchMpc
    g = _fixture(case)
    N_p, Nt = int(g["N_p"]), int(g["Nt"])
    atoms = {str(k): g["atom_%d" % i] for i, k in enumerate(g["atom_keys"])}
    assert set(atoms) == {"q_u", "q_mu"}
    mats = {k: g["in_" + k][None] for k in DEWH_MATS}
    bm = BatchMpc(mats, N_p, nu_l=1, device=cuda_device, solver=solver, opts=cabi.default_opts(mip_rel_gap=0.0),
                  dp_opts=cabi.stage_dp_default_opts(mip_rel_gap=0.0))
    bm.build()
    cost = np.zeros((1, Nt, 3))
    cost[0, :, 0] = atoms["q_u"]
    cost[0, :, 1:] = atoms["q_mu"][None, :]
    extra = []
    for j in range(int(g["n_extra"])):
        ec = {}
        if "extra_%d_omega_t" % j in g:
            ec["omega_tilde_k"] = g["extra_%d_omega_t" % j].reshape(1, Nt)
        if "extra_%d_omega_scenarios" % j in g:
            ec["omega_scenarios_k"] = g["extra_%d_omega_scenarios" % j].reshape(1, Nt, -1)
        if "extra_%d_N_tilde" % j in g:
            ec["N_tilde"] = int(g["extra_%d_N_tilde" % j])
        extra.append(ec)
    res = bm.solve(g["x_k"].reshape(1, 1), g["omega_tilde"].reshape(1, Nt), cost_v=cost.reshape(1, -1),
                   extra_constraints=extra)
    assert res["solver"] == solver and int(res["status"][0]) == 0
    obj, ref = float(res["obj"][0]), float(g["sol_obj"])
    assert abs(obj - ref) <= 1e-6 * max(1.0, abs(ref)), (obj, ref)
    v = res["v"][0].cpu().numpy()
    assert abs(float(cost.reshape(-1) @ v) - obj) <= 1e-9 * max(1.0, abs(obj))
    # the returned point satisfies every row the reference built (its rows, mapped to v~ by its own layout)
    G_v, h = g["G"] @ g["v_of_x"].T, g["h"]
    assert np.all(G_v @ v <= h + 1e-6 * np.maximum(1.0, np.abs(h)))
    isb = bm.is_bin_v.astype(bool)
    assert np.all((v[isb] == 0) | (v[isb] == 1)) and np.all(v[~isb] >= -1e-9)
    # decisions: the reference's solution has the same cost, so they must agree unless the optimum is not unique;
    # uniqueness is certified by enumeration where that is possible (2^9 assignments)
    if Nt <= 9:
        from oracle import assemble as oa, condense as oc, mld as omld, solve as osv
        full, dims, vt = omld.complete({k: g["in_" + k] for k in DEWH_MATS}, nu_l=1)
        prob = oa.build_problem(oc.condense(full, dims, Nt), dims, vt, Nt, g["x_k"], g["omega_tilde"], atoms=atoms)
        st, o2, v2, second = osv.solve_enumerate(prob)
        if second - o2 > 1e-6 * max(1.0, abs(o2)):
            v_ref = g["v_of_x"] @ g["sol_x"]
            assert np.array_equal(v[isb], np.round(v_ref[isb]))
            assert v[0] == round(float(g["fb_u"][0]))


def test_closed_loop_vs_reference(cuda_device):
    """the reference's own 14-instant loop (feedback -> sim_step_k on the re-parametrised simulation model) against
    DewhFleet.closed_loop with one heater: first controls, temperatures, objectives, constraint flags."""
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.fleet import DewhFleet
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs.parameters import dewh_param_struct
    z = _fixture("dewh_closed_loop")
    keys = ("C_w", "A_h", "U_h", "m_h", "T_w", "T_inf", "P_h_Nom", "T_h_min", "T_h_max", "T_h_Nom", "ts")
    p = dict(dewh_param_struct)
    p.update({k: float(v) for k, v in zip(keys, z["params"])})
    N_p, steps = int(z["N_p"]), int(z["steps"])
    fleet = DewhFleet([p], N_p, device=cuda_device)
    price = z["price"] / p["P_h_Nom"]                     # the fixture's q_u already carries P_h_Nom
    log = fleet.closed_loop(np.array([z["x"][0]]), z["demand"][None, :], price, steps, controller="mpc_ce")
    log = {k: v.cpu().numpy() for k, v in log.items()}
    assert (log["status"] == 0).all()
    np.testing.assert_allclose(log["obj"][:, 0], z["obj"], rtol=1e-6, atol=1e-9)
    assert np.array_equal(log["u"][:, 0], np.round(z["u"]))
    np.testing.assert_allclose(log["T"][:, 0], z["x"], rtol=1e-9)
    assert np.array_equal(log["cons"][:, 0, :].astype(bool), z["cons"].astype(bool))
