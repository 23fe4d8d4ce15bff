"""CPU: the PRODUCT's host-side problem assembly -- the cost-atom grammar and weights
(controllers/components/objective_atoms.py), MpcController._cost_terms (Linear atoms on v / x / y / omega, rate form,
and the per-step diagonal terms handed to the stage-DP kernels for Quadratic / L22 / L1 atoms) and BatchMpc's variable
layout, bounds and integrality mask -- against what the UNMODIFIED reference assembles (tests/golden/assembly_*.npz,
tests/golden/make_golden_assembly.py).  No kernel runs: the controller method is called on a stand-in object with a
BatchMpc that is never built; the affine maps x~(v), y~(v) come from the (reference-pinned) oracle."""
import types

import numpy as np
import pytest

from oracle import assemble as oa
from pyhybridcontrol_b200.batch import BatchMpc
from pyhybridcontrol_b200.controllers.components.objective_atoms import ObjectiveAtoms
from pyhybridcontrol_b200.controllers.mpc_controller import MpcController
from pyhybridcontrol_b200.models.mld_model import MldModel
from test_oracle_assembly_pinned import CASES, MAT_NAMES, _load

LINEAR = [c for c in CASES if c not in ("dewh_N6_all_atoms", "rand_N5_quadratic", "dewh_N6_stage_dp_atoms")]


def _host_terms(g, Nt):
    mats = {k: g["in_" + k] for k in MAT_NAMES if g["in_" + k].size}
    mld = MldModel(nu_l=int(g["nu_l"]), **mats)
    atoms = {str(k): g["atom_%d" % i] for i, k in enumerate(g["atom_keys"])}
    batch = BatchMpc({k: np.array(v) for k, v in mld.items() if v.size}, int(g["N_p"]), Nt, nu_l=int(g["nu_l"]), B=1,
                     device="cpu")
    k_neg1 = {k[len("k_neg1_"):]: v.reshape(-1, 1) for k, v in g.items() if k.startswith("k_neg1_")}
    ctrl = types.SimpleNamespace(mld_info_k=mld.mld_info, N_tilde=Nt, _mld_evo_matrices=types.SimpleNamespace(batch=batch),
                                 _sense="minimize", _with_std_objective=True,
                                 _std_obj_atoms=ObjectiveAtoms(mld.mld_info, int(g["N_p"]), Nt, None, **atoms),
                                 variables_k_neg1=k_neg1, _omega_tilde_k=g["omega_tilde"].reshape(-1, 1))
    return MpcController._cost_terms(ctrl, 0), batch, mld


def _affine_cost(ct, maps):
    c, c0 = ct["cost_v"][0].copy(), ct["const"]
    for key, name in (("w_x", "x"), ("w_y", "y")):
        if ct[key] is not None:
            M, m0 = maps[name]
            c += ct[key][0] @ M
            c0 += float(ct[key][0] @ m0)
    return c, c0


@pytest.mark.parametrize("case", LINEAR)
def test_linear_cost_layout_and_bounds(case):
    g, prob, evo, dims, Nt = _load(case)
    ct, batch, mld = _host_terms(g, Nt)
    assert ct["quad"] is None
    c, c0 = _affine_cost(ct, oa.affine_maps(evo, dims, Nt, g["x_k"], g["omega_tilde"]))
    x_to_v = g["v_of_x"].argmax(axis=0)                    # where the reference's variables sit in v~
    scale = max(1.0, float(np.abs(g["c"]).max()))
    np.testing.assert_allclose(c[x_to_v], g["c"], rtol=0, atol=1e-12 * scale)
    assert abs(c0 - float(g["c0"])) <= 1e-10 * max(1.0, abs(float(g["c0"])))
    assert np.array_equal(batch.is_bin_v[x_to_v].astype(bool), g["integrality"].astype(bool))
    assert np.array_equal(batch.lb_v[x_to_v], g["lb"]) and np.array_equal(batch.ub_v[x_to_v], g["ub"])
    info = mld.mld_info
    for name in ("u", "delta", "z", "mu"):                 # var_index == the reference's stacking of v~
        assert np.array_equal(batch.var_index(name), oa.var_layout(dims, Nt)[name]), name
    assert (info.nv, info.nv_l) == (dims["nv"], int(g["integrality"].sum()) // Nt)


def test_stage_dp_terms_reproduce_the_reference_objective():
    """Quadratic / L22 / L1 atoms: cost_v + w_x'x~ + w_y'y~ + const + sum of the per-step diagonal terms equals the
    reference's objective at random points (vector weights enter squared, matrix weights by their diagonal, |mu| = mu
    folded into the linear cost, N_p / terminal suffixes)."""
    g, prob, evo, dims, Nt = _load("dewh_N6_stage_dp_atoms")
    ct, batch, mld = _host_terms(g, Nt)
    assert set(ct["quad"]) <= {"x2", "x1", "y2", "y1", "u2", "u1", "mu2"}
    maps = oa.affine_maps(evo, dims, Nt, g["x_k"], g["omega_tilde"])
    c, c0 = _affine_cost(ct, maps)
    for x, f in zip(g["points_x"], g["points_f"]):
        v = g["v_of_x"] @ x
        assert np.all(v[batch.var_index("mu")] >= 0)       # the fold |mu| = mu holds inside the bounds
        val = float(c @ v) + c0
        for key, w in ct["quad"].items():
            M, m0 = maps[key[:-1]]
            e = (M @ v + m0).reshape(1, Nt, -1)
            val += float(np.sum(w * (e ** 2 if key.endswith("2") else np.abs(e))))
        assert abs(val - f) <= 1e-10 * max(1.0, abs(f)), (val, f)


def test_atoms_outside_the_stage_dp_class_take_the_general_path():
    """which atoms of the reference's all-atom fixture leave the exact fast paths for the general MIQP assembly"""
    g, prob, evo, dims, Nt = _load("dewh_N6_all_atoms")
    mats = {k: g["in_" + k] for k in MAT_NAMES if g["in_" + k].size}
    mld = MldModel(nu_l=1, **mats)
    atoms = {str(k): g["atom_%d" % i] for i, k in enumerate(g["atom_keys"])}
    oas = ObjectiveAtoms(mld.mld_info, int(g["N_p"]), Nt, None, **atoms)
    batch = types.SimpleNamespace(stage_dp_ok=True)
    ctrl = types.SimpleNamespace(_mld_evo_matrices=types.SimpleNamespace(batch=batch), _sense="minimize")
    general = [(atom.atom_type, atom.var_name, atom.is_rate_atom) for atom in oas.iter_atoms()
               if MpcController._needs_general_path(ctrl, atom)]
    assert ("L1", "u", True) in general and any(a[0] == "Linf" for a in general)
    assert not any(a[0] == "Linear" for a in general)
    # maximising a convex atom is refused outright
    ctrl._sense = "maximize"
    with pytest.raises(NotImplementedError):
        for atom in oas.iter_atoms():
            MpcController._needs_general_path(ctrl, atom)


def test_update_std_obj_atoms_sequence():
    """set / update calls one after the other (a terminal weight merged next to a horizon weight, an N_p weight over
    it, an all-zero weight deleting the atom, an atom on x added later, a final set that replaces everything): the
    cost vector and constant after every call equal the reference's"""
    import os
    from oracle import condense as oc, mld as omld
    from test_oracle_assembly_pinned import GOLDEN
    z = np.load(os.path.join(GOLDEN, "assembly_update_sequence.npz"))
    N_p = int(z["N_p"])
    Nt = N_p + 1
    mats = {k: z["in_" + k] for k in MAT_NAMES if z["in_" + k].size}
    mld = MldModel(nu_l=1, **mats)
    full, dims, vt = omld.complete(mats, nu_l=1)
    maps = oa.affine_maps(oc.condense(full, dims, Nt), dims, Nt, z["x_k"], z["omega_tilde"])
    batch = BatchMpc({k: np.array(v) for k, v in mld.items() if v.size}, N_p, Nt, nu_l=1, B=1, device="cpu")
    atoms = None
    for i in range(int(z["n_steps"])):
        kw = {str(k): z["val_%d_%d" % (i, j)] for j, k in enumerate(z["keys_%d" % i])}
        if str(z["how_%d" % i]) == "set":
            if atoms is None:
                atoms = ObjectiveAtoms(mld.mld_info, N_p, Nt, None, **kw)
            else:
                atoms.set(None, **kw)
        else:
            atoms.update_atoms(None, **kw)
        ctrl = types.SimpleNamespace(mld_info_k=mld.mld_info, N_tilde=Nt, _mld_evo_matrices=types.SimpleNamespace(batch=batch),
                                     _sense="minimize", _with_std_objective=True, _std_obj_atoms=atoms,
                                     variables_k_neg1={}, _omega_tilde_k=z["omega_tilde"].reshape(-1, 1))
        c, c0 = _affine_cost(MpcController._cost_terms(ctrl, 0), maps)
        np.testing.assert_allclose(c, z["c_v_%d" % i], rtol=0, atol=1e-12 * max(1.0, np.abs(z["c_v_%d" % i]).max()),
                                   err_msg="call %d" % i)
        assert abs(c0 - float(z["c0_%d" % i])) <= 1e-10 * max(1.0, abs(float(z["c0_%d" % i]))), i
