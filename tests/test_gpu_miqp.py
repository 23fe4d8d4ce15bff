"""GPU parity tests of the general mixed-integer QP path (csrc/miqp_admm.cu behind pyhybridcontrol_b200/miqp.py):
every cost atom of the reference's grammar that the exact fast paths do not carry -- dense matrix weights, Linf (the
reference's norm1-with-weight quirk included), non-linear rate atoms, non-linear atoms on a vector-state MLD
(objective_atoms.py:185-206, 297-305, 320-363) -- against the oracle: numpy assembly pinned to the reference's own
expressions + HiGHS QP relaxations under depth-first branch and bound, or exhaustive enumeration of the binaries.
Tolerance: objectives 1e-6 relative (the ADMM relaxation is solved to 1e-9 of the scaled problem), binaries exact
where the runner-up assignment is worse by more than that."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def test_kernel_vs_enumeration_random_miqp(cuda_device):
    """the kernel alone, on the canonical form: random convex MIQPs with inequality rows, free / boxed / binary columns"""
    import torch
    from pyhybridcontrol_b200 import cabi
    from oracle import solve as osv
    from oracle.assemble import Problem
    rng = np.random.default_rng(5)
    B, n, m, nbin = 12, 14, 9, 6
    probs = []
    P = np.zeros((B, n, n)); c = np.zeros((B, n)); H = np.zeros((B, m, n)); rhs = np.zeros((B, m))
    lb = np.full(n, -np.inf); ub = np.full(n, np.inf)
    is_bin = np.zeros(n, dtype=np.uint8); is_bin[:nbin] = 1
    lb[:nbin], ub[:nbin] = 0.0, 1.0
    lb[nbin:nbin + 3], ub[nbin:nbin + 3] = -1.5, 2.0          # boxed continuous columns; the rest are free
    lb[-2:] = 0.0                                               # two non-negative columns (slack-like)
    for b in range(B):
        F = rng.normal(size=(n - 2, n)) * (rng.random((n - 2, n)) < 0.6)
        P[b] = F.T @ F + 1e-3 * np.eye(n)
        P[b][-2:, :] = 0.0; P[b][:, -2:] = 0.0                  # no curvature on the slack-like columns
        c[b] = rng.normal(size=n); c[b][-2:] = rng.uniform(5.0, 20.0, 2)
        H[b] = rng.normal(size=(m, n)) * (rng.random((m, n)) < 0.5)
        H[b][:2, -2:] = -np.eye(2)                              # two softened rows
        rhs[b] = rng.uniform(0.5, 2.0, m)
        pr = Problem(n)
        pr.P, pr.c, pr.H, pr.rhs, pr.lb, pr.ub, pr.is_bin = P[b], c[b], H[b], rhs[b], lb, ub, is_bin.astype(bool)
        probs.append(pr)
    dev = cuda_device
    t = lambda a: torch.as_tensor(a, dtype=torch.float64).to(dev)  # noqa: E731
    v, obj, status, stats = cabi.miqp_solve(t(c), t(H), t(rhs), t(lb), t(ub), torch.as_tensor(is_bin).to(dev), P=t(P))
    torch.cuda.synchronize()
    v, obj, status = v.cpu().numpy(), obj.cpu().numpy(), status.cpu().numpy()
    for b in range(B):
        st, oref, vref, second = osv.solve_enumerate(probs[b])
        if st != osv.OPTIMAL:
            assert status[b] == 1, (b, status[b])
            continue
        assert status[b] == 0, (b, status[b])
        assert _rel(obj[b], oref) <= 1e-6, (b, obj[b], oref)
        assert np.max(H[b] @ v[b] - rhs[b]) <= 1e-6
        if second - oref > 1e-5 * max(1.0, abs(oref)):
            assert np.array_equal(np.round(v[b][:nbin]), np.round(vref[:nbin])), b


def _vector_state_mld(rng):
    """nx = 2, one continuous + one binary input, one binary delta, slack-softened rows, an output"""
    nx, nu, nd, nw, ny, nc = 2, 2, 1, 1, 1, 4
    m = dict(A=np.array([[0.85, 0.1], [-0.05, 0.7]]) + rng.uniform(-0.05, 0.05, (nx, nx)),
             B1=rng.uniform(-1, 1, (nx, nu)), B2=rng.uniform(-1, 1, (nx, nd)), B4=rng.uniform(-1, 1, (nx, nw)),
             b5=rng.uniform(-0.2, 0.2, (nx, 1)), C=rng.uniform(-1, 1, (ny, nx)), D1=rng.uniform(-0.3, 0.3, (ny, nu)),
             E=rng.uniform(-1, 1, (nc, nx)), F1=rng.uniform(-0.5, 0.5, (nc, nu)), F2=rng.uniform(-0.5, 0.5, (nc, nd)),
             F4=rng.uniform(-0.3, 0.3, (nc, nw)), f5=rng.uniform(1.0, 3.0, (nc, 1)), G=rng.uniform(-0.3, 0.3, (nc, ny)),
             Psi=-np.eye(nc))
    return m


GENERAL_ATOMS = [
    dict(q_mu=[8.0, 9.0, 7.0, 10.0], Q_x=[[2.0, 0.6], [0.6, 1.5]], q_x=[-1.0, 0.5], Q_u=0.3 * np.eye(2)),      # dense Q_x
    dict(q_mu=[8.0, 9.0, 7.0, 10.0], q_L22_y=1.3, Q_x_f=[[4.0, -1.0], [-1.0, 3.0]], q_u=[0.2, 0.4]),          # terminal weight
    dict(q_mu=[8.0, 9.0, 7.0, 10.0], q_Linf_x=[0.7, 1.1], q_u=[0.1, 0.3], Q_u=0.2 * np.eye(2)),                # Linf quirk
    dict(q_mu=[8.0, 9.0, 7.0, 10.0], q_L1_du=[0.5, 0.8], Q_x=np.eye(2)),                                       # L1 rate atom
    dict(q_mu=[8.0, 9.0, 7.0, 10.0], Q_du=[[1.0, 0.2], [0.2, 2.0]], q_L1_x=[0.4, 0.25], q_delta=0.3),                  # quadratic rate
    dict(q_mu=[8.0, 9.0, 7.0, 10.0], Q_L1_y=[[1.5]], q_L22_dx=[0.6, 0.9], Q_v=0.05 * np.eye(7)),                      # matrix L1, Q_v
]


@pytest.mark.parametrize("case", range(len(GENERAL_ATOMS)))
def test_controller_general_atoms_on_vector_state_mld(case, cuda_device):
    """MpcController with atoms that need the general path, on an nx = 2 MLD with continuous and binary inputs:
    objective and first-step decisions against the oracle's assembly + QP branch and bound (cross-checked by
    enumeration: 10 binaries)."""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.controllers.mpc_controller import MpcController
    from pyhybridcontrol_b200.models.mld_model import MldModel, MldSystemModel
    rng = np.random.default_rng(100 + case)
    N_p = 4
    Nt = N_p + 1
    mats = _vector_state_mld(rng)
    atoms = GENERAL_ATOMS[case]
    x0 = rng.uniform(-1, 1, 2)
    om = rng.uniform(-1, 1, Nt)
    ctrl = MpcController(model=MldSystemModel(mld_numeric=MldModel(nu_l=1, **mats)), N_p=N_p)
    ctrl.set_std_obj_atoms(**atoms)
    ctrl.build()
    assert ctrl._general_path
    obj = ctrl.solve(k=0, x_k=x0, omega_tilde_k=om)
    full, d, vt = omld.complete(mats, nu_l=1)
    prob = oa.build_problem(oc.condense(full, d, Nt), d, vt, Nt, x0, om, atoms=atoms)
    st, oref, vref, _ = osv.solve_miqp(prob)
    assert st == osv.OPTIMAL
    st2, oenum, venum, second = osv.solve_enumerate(prob)
    assert _rel(oref, oenum) <= 1e-7, (oref, oenum)
    assert _rel(obj, oref) <= 1e-6, (case, obj, oref)
    fb = ctrl.feedback(k=0)
    if second - oenum > 1e-5 * max(1.0, abs(oenum)):
        nv = d["nv"]
        first = venum[:nv]
        assert fb.u[1, 0] == round(first[1])                    # the binary input
        assert fb.delta[0, 0] == round(first[2])
        assert abs(fb.u[0, 0] - first[0]) <= 1e-4 * max(1.0, abs(first[0]))


def test_dewh_with_rate_and_linf_atoms(cuda_device):
    """the reference example's water heater with atoms that leave the stage-DP class: a switching penalty (L1 on
    the rate of u), Linf on the slack, a dense terminal weight -- against enumeration (9 binaries)"""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.controllers.mpc_controller import MpcController
    from pyhybridcontrol_b200.models.mld_model import MldModel, MldSystemModel
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    N_p = 8
    wl = syn.dewh_batch(3, N_p, seed=41)
    for b in range(3):
        mats = {k: v[b] for k, v in wl["mats"].items()}
        scale = float(wl["q_u"].mean())
        atoms = dict(q_u=wl["q_u"][b], q_mu=wl["q_mu"][b], q_L1_du=0.4 * scale, q_Linf_mu=[0.3 * scale, 0.1 * scale],
                     Q_x_f=[[0.02 * scale]], q_x_f=[-2 * 0.02 * scale * 60.0])
        ctrl = MpcController(model=MldSystemModel(mld_numeric=MldModel(nu_l=1, **mats)), N_p=N_p)
        ctrl.set_std_obj_atoms(**atoms)
        ctrl.build()
        assert ctrl._general_path
        obj = ctrl.solve(k=0, x_k=wl["x0"][b], omega_tilde_k=wl["omega"][b])
        full, d, vt = omld.complete(mats, nu_l=1)
        prob = oa.build_problem(oc.condense(full, d, N_p + 1), d, vt, N_p + 1, wl["x0"][b], wl["omega"][b], atoms=atoms)
        st, oref, vref, second = osv.solve_enumerate(prob)
        assert _rel(obj, oref) <= 1e-6, (b, obj, oref)
        if second - oref > 1e-5 * max(1.0, abs(oref)):
            assert ctrl.feedback(k=0).u[0, 0] == round(vref[0])


def test_constraint_set_with_its_own_state(cuda_device):
    """gen_evo_constraints(x_k=...) -- a set evaluated at another state than the controller's
    (controller_base.py:411-456) -- on the exact path: same objective as the oracle's stacked problem"""
    from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
    from pyhybridcontrol_b200.controllers.mpc_controller import MpcController
    from pyhybridcontrol_b200.models.mld_model import MldModel, MldSystemModel
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn
    N_p = 10
    Nt = N_p + 1
    wl = syn.dewh_batch(1, N_p, seed=43)
    mats = {k: v[0] for k, v in wl["mats"].items()}
    atoms = dict(q_u=wl["q_u"][0], q_mu=wl["q_mu"][0])
    ctrl = MpcController(model=MldSystemModel(mld_numeric=MldModel(nu_l=1, **mats)), N_p=N_p)
    ctrl.set_std_obj_atoms(**atoms)
    x_alt = wl["x0"][0] - 3.0
    other = ctrl.gen_evo_constraints(x_k=x_alt, omega_tilde_k=1.3 * wl["omega"][0])
    ctrl.set_constraints(other_constraints=[other])
    ctrl.build()
    obj = ctrl.solve(k=0, x_k=wl["x0"][0], omega_tilde_k=wl["omega"][0])
    full, d, vt = omld.complete(mats, nu_l=1)
    evo = oc.condense(full, d, Nt)
    prob = oa.build_problem(evo, d, vt, Nt, wl["x0"][0], wl["omega"][0], atoms=atoms)
    H2, r2 = oa.evo_rhs(evo, d, x_alt, 1.3 * wl["omega"][0])
    prob.add_rows(H2, r2)
    st, oref, vref = osv.solve_milp(prob, polish=True)
    assert _rel(obj, oref) <= 1e-6, (obj, oref)
