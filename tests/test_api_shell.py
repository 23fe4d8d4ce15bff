"""The reference-facing Python API (MldModel / MpcController / agents): host logic on the CPU, solves on the GPU."""
import numpy as np
import pytest

from pyhybridcontrol_b200.models.mld_model import MldModel, MldSystemModel
from pyhybridcontrol_b200.controllers.components.objective_atoms import ObjectiveAtoms, parse_atom_key
from pyhybridcontrol_b200.utils.structs import StructDict, ParNotSet

DEWH = dict(A=[[0.9970371127900564]], B1=[[4.298192277481107]], B4=[[-179.73320827515]], b5=[[0.07407218024859108]],
            E=[[1.0], [-1.0]], F1=[[0.0], [0.0]], Psi=[[-1.0, 0.0], [0.0, -1.0]], f5=[[65.0], [-50.0]])


def test_mld_model_dims_types_and_defaults():
    m = MldModel(nu_l=1, **DEWH)
    i = m.mld_info
    assert (i.nx, i.nu, i.ndelta, i.nz, i.nmu, i.nomega, i.ny, i.n_constraints, i.nv) == (1, 1, 0, 0, 2, 1, 1, 2, 3)
    assert i.var_type_v.ravel().tolist() == ["b", "c", "c"] and i.nu_l == 1
    assert np.array_equal(m.C, np.eye(1)) and m.D1.shape == (1, 1) and m["G"].shape == (2, 1)       # attr == item
    assert not m.A.flags.writeable
    grid = MldModel(D4=np.ones((1, 3)), F2=np.ones((6, 1)), F3=np.ones((6, 1)), f5=np.ones((6, 1)), G=np.ones((6, 1)))
    gi = grid.mld_info
    assert (gi.nx, gi.ndelta, gi.nz, gi.nomega, gi.ny, gi.n_constraints) == (0, 1, 1, 3, 1, 6)
    assert gi.var_type_v.ravel().tolist() == ["b", "c"]


def test_mld_model_errors_and_versioning():
    with pytest.raises(ValueError):
        MldModel(A=np.ones((2, 3)))
    with pytest.raises(ValueError):
        MldModel(foo=[[1.0]])
    with pytest.raises(ValueError):
        MldModel(E=[[1.0]], A=[[1.0]])           # constraint rows without f5
    with pytest.raises(ValueError):
        MldModel(A=[[1.0]], B1=[[1.0]], nu_l=2)
    with pytest.raises(TypeError):
        MldModel(A="not a matrix")
    assert MldModel(A=lambda: 1.0).mld_type == "callable"      # functions are traced (tests/test_callable_front_end.py)
    m = MldModel(nu_l=1, **DEWH)
    v0 = m.version
    m.update(f5=[[80.0], [-50.0]])
    assert m.version != v0 and m.f5[0, 0] == 80.0


def test_objective_atoms_grammar_and_weights():
    info = MldModel(nu_l=1, **DEWH).mld_info
    assert parse_atom_key("q_L22_x_N_p") == ("vector", "L22", "x", False, "N_p")
    assert parse_atom_key("q_ddelta") == ("vector", "Linear", "delta", True, "")
    atoms = ObjectiveAtoms(info, 3, 4, q_mu=[10.0, 1.0], q_u=np.arange(4.0), Q_x_f=[[2.0]])
    assert np.array_equal(atoms["mu"]["Linear_vector"].weight_N_tilde.ravel(), np.tile([10.0, 1.0], 4))
    assert np.array_equal(atoms["u"]["Linear_vector"].weight_N_tilde.ravel(), np.arange(4.0))
    W = atoms["x"]["Quadratic_matrix"].weight_N_tilde
    assert W.shape == (4, 4) and W[3, 3] == 2.0 and W.sum() == 2.0
    atoms.update_atoms(q_u_N_p=[5.0])                       # first N_p steps only
    assert atoms["u"]["Linear_vector"].weight_N_tilde.ravel().tolist() == [5.0, 5.0, 5.0, 3.0]
    atoms.update_atoms(q_u=np.zeros(4))                     # all-zero weight deletes the atom
    assert "Linear_vector" not in atoms["u"]
    with pytest.raises(ValueError):
        ObjectiveAtoms(info, 3, 4, q_mu=[1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        ObjectiveAtoms(info, 3, 4, q_nope=[1.0])


def test_structdict():
    s = StructDict(a=1)
    s.b = 2
    assert s["b"] == 2 and s.a == 1 and not ParNotSet
    with pytest.raises(AttributeError):
        s.c


@pytest.mark.gpu
def test_mpc_controller_lifecycle_and_parity(cuda_device):
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.controllers.mpc_controller import (MpcController, ControllerBuildRequiredError,
                                                                ControllerSolverError)
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    wl = syn.dewh_batch(1, 24, seed=21)
    mats = {k: v[0] for k, v in wl["mats"].items()}
    mld = MldModel(nu_l=1, **mats)
    ctrl = MpcController(model=MldSystemModel(mld_numeric=mld), N_p=24)
    assert ctrl.N_tilde == 25 and ctrl.build_required
    with pytest.raises(ControllerBuildRequiredError):
        ctrl.solve(k=0, x_k=wl["x0"][0], omega_tilde_k=wl["omega"][0])
    ctrl.set_std_obj_atoms(q_u=wl["q_u"][0], q_mu=wl["q_mu"][0])
    ctrl.build()
    assert not ctrl.build_required
    # condensed matrices are exposed under the reference's names
    full, d, vt = omld.complete(mats, nu_l=1)
    ref = oc.condense(full, d, 25)
    np.testing.assert_allclose(ctrl.mld_evo_matrices.constraint["H_v_N_tilde"], ref["H_v"], atol=1e-10)
    assert ctrl.mld_evo_matrices.constraint["H_v_N_p"].shape == (48, 75)
    obj = ctrl.solve(k=0, x_k=wl["x0"][0], omega_tilde_k=wl["omega"][0])
    prob = oa.build_problem(ref, d, vt, 25, wl["x0"][0], wl["omega"][0], atoms=dict(q_u=wl["q_u"][0], q_mu=wl["q_mu"][0]))
    st, oref, vref = osv.solve_milp(prob)
    assert abs(obj - oref) <= 1e-6 * max(1.0, abs(oref))
    fb = ctrl.feedback(k=0)
    full_s, first = oa.split_solution(vref, ref, d, 25, wl["x0"][0], wl["omega"][0])
    assert fb.u.shape == (1, 1) and fb.u[0, 0] == round(first["u"][0])
    np.testing.assert_allclose(fb.y.ravel(), first["y"], atol=1e-9)
    # simulation step + log
    l = ctrl.sim_step_k(k=0, omega_k=wl["omega"][0, :1], mu_k=None) if False else ctrl.sim_step_k(k=0, omega_k=wl["omega"][0, :1])
    assert 0 in ctrl.sim_log and np.allclose(ctrl.x_k, l.x_k1)
    df = ctrl.sim_log.get_concat_log()
    assert ("u", 0) in df.columns and df.index.name == "k"
    # changing the cost or the model requires a rebuild
    ctrl.set_std_obj_atoms(q_u=2 * wl["q_u"][0], q_mu=wl["q_mu"][0])
    assert ctrl.build_required
    ctrl.build()
    mld.update(f5=[[200.0], [-199.0]])          # impossible band, but soft constraints keep it feasible
    assert ctrl.build_required
    ctrl.build()
    assert np.isfinite(ctrl.solve(k=1))
    ctrl.build(disable_soft_constraints=True)   # now hard -> infeasible -> ControllerSolverError
    with pytest.raises(ControllerSolverError):
        ctrl.solve(k=1)


@pytest.mark.gpu
def test_lsim_k_aux_and_agents(cuda_device):
    from pyhybridcontrol_b200.controllers.mpc_controller import MpcController
    from pyhybridcontrol_b200.models.agents import Agent, MpcAgent
    mld = MldModel(nu_l=1, **DEWH)
    r = mld.lsim_k(x_k=58, u_k=1, omega_k=0.002, mu_k=[0, 0])
    assert abs(r.x_k1[0, 0] - 61.84095) < 1e-4 and r.y[0, 0] == 58 and r.cons.ravel().tolist() == [True, True]
    r2 = mld.lsim_k(x_k=70, u_k=0, omega_k=0.0)             # mu missing -> minimal slack from the GPU solver
    np.testing.assert_allclose(r2.mu.ravel(), [5.0, 0.0], atol=1e-9)
    Agent.delete_all_devices()
    ag = MpcAgent(device_type="dewh", device_id=1, sim_model=MldSystemModel(mld_numeric=mld), N_p=6)
    assert isinstance(ag.mpc_controller, MpcController) and ag.N_tilde == 7
    with pytest.raises(ValueError):
        MpcAgent(device_type="dewh", device_id=1, sim_model=MldSystemModel(mld_numeric=mld), N_p=6)
    ag.mpc_controller.set_std_obj_atoms(q_u=1.0, q_mu=[100.0, 10.0])
    ag.mpc_controller.build()
    fb = ag.mpc_controller.feedback(k=0, x_k=50.5, omega_tilde_k=np.full(7, 0.004))
    assert fb.u[0, 0] in (0.0, 1.0)


@pytest.mark.gpu
def test_mpc_controller_miqp_atoms(cuda_device):
    """Quadratic / L1 atoms through the reference's cost-atom grammar (objective_atoms.py:453-496): set-point tracking
    Q_x with a linear q_x, a squared slack penalty and an L1 output term -- an MIQP, checked against enumeration."""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.controllers.mpc_controller import MpcController
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    N_p = 8
    wl = syn.dewh_batch(1, N_p, seed=31)
    mats = {k: v[0] for k, v in wl["mats"].items()}
    scale = float(wl["q_u"].mean())
    atoms = dict(q_u=wl["q_u"][0], q_mu=wl["q_mu"][0], Q_x=0.5 * scale, q_x=-2 * 0.5 * scale * 62.0,
                 q_L22_mu=[1.2, 0.4], q_L1_y=0.05 * scale)
    ctrl = MpcController(model=MldSystemModel(mld_numeric=MldModel(nu_l=1, **mats)), N_p=N_p)
    ctrl.set_std_obj_atoms(**atoms)
    ctrl.build()
    obj = ctrl.solve(k=0, x_k=wl["x0"][0], omega_tilde_k=wl["omega"][0])
    full, d, vt = omld.complete(mats, nu_l=1)
    prob = oa.build_problem(oc.condense(full, d, N_p + 1), d, vt, N_p + 1, wl["x0"][0], wl["omega"][0], atoms=atoms)
    st, oref, vref, second = osv.solve_enumerate(prob)
    assert abs(obj - oref) <= 1e-6 * max(1.0, abs(oref)), (obj, oref)
    fb = ctrl.feedback(k=0)
    if second - oref > 1e-6 * max(1.0, abs(oref)):
        assert fb.u[0, 0] == round(vref[0])
    with pytest.raises(NotImplementedError):             # maximising a convex atom is not a convex problem
        ctrl.build(sense="maximize")
