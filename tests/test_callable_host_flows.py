"""CPU: host-side flows of the symbolic / callable front-end (MldModel.to_numeric, MldSystemModel.update_mld /
update_param_struct / get_mld_numeric / param_struct setter, the example's device models, get_mld_numeric_tilde) with
the kernel call ``cabi.param_eval`` replaced -- IN THIS TEST ONLY -- by the numpy twin of its interpreter
(tests/expr_vm_twin.py), so the bookkeeping around the kernel is exercised without a GPU.  The kernel itself is checked
in tests/test_gpu_callable.py; numbers are compared with golden vectors of the unmodified reference."""
import numpy as np
import pytest
import sympy as sp
import torch

import expr_vm_twin as twin
from test_callable_front_end import MAT_NAMES, load_fixture


@pytest.fixture()
def twin_kernel(monkeypatch):
    from pyhybridcontrol_b200 import cabi
    from pyhybridcontrol_b200.utils import matrix_utils as mu
    calls = []

    def fake_param_eval(program, n_regs, mat_sizes, params, out=None):
        calls.append(tuple(params.shape))
        return torch.from_numpy(twin.run(program.cpu().numpy(), int(n_regs), list(mat_sizes), params.cpu().numpy()))

    orig = mu.ExprProgram.param_table
    monkeypatch.setenv("HMPC_PARAM_EVAL", "v1")        # the CPU twin interprets the first kernel's instruction stream
    monkeypatch.setattr(cabi, "param_eval", fake_param_eval)
    monkeypatch.setattr(mu.ExprProgram, "param_table",
                        lambda self, ps, overrides=None, B=None, device="cuda": orig(self, ps, overrides, B, "cpu"))
    return calls


def _close(got, ref, rtol=1e-10):
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    assert got.shape == ref.shape or (got.size == 0 and ref.size == 0), (got.shape, ref.shape)
    if ref.size:
        np.testing.assert_allclose(got, ref, rtol=rtol, atol=0)


def test_device_models_and_parameter_updates(twin_kernel):
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import models as M
    for make, fixture in ((lambda: M.DewhModel(const_heat=True), "callable_dewh_control.npz"),
                          (lambda: M.DewhModel(const_heat=False), "callable_dewh_sim.npz"),
                          (lambda: M.GridModel(num_devices=3), "callable_grid_3dev.npz"),
                          (lambda: M.PvModel(), "callable_pv.npz"), (lambda: M.ResDemandModel(), "callable_resd.npz")):
        z, names, _, pnames = load_fixture(fixture)
        n0 = len(twin_kernel)
        model = make()
        assert len(twin_kernel) == n0 + 1 and twin_kernel[-1][0] == 1          # ONE launch for the whole model, B = 1
        num = model.mld_numeric
        assert num.mld_type == "numeric" and model.mld_symbolic.mld_type == "symbolic"
        for k in MAT_NAMES:
            _close(num[k], z["num_" + k])
            assert not np.asarray(num[k]).flags.writeable
        for key in ("nx", "nu", "ndelta", "nz", "nomega", "ny", "nmu", "nv", "n_constraints", "nu_l", "ndelta_l"):
            assert int(num.mld_info[key]) == int(z["info_" + key]), (fixture, key)
        other = dict(zip(pnames, z["params"][1]))
        got = model.get_mld_numeric(param_struct_subset=other)
        assert got is not model.mld_numeric and model.get_mld_numeric() is model.mld_numeric
        for k in names:
            _close(got[k], z["other_" + k])
        v0 = model.version
        model.update_param_struct(param_struct_subset=other)
        assert model.version != v0 and all(model.param_struct[k] == v for k, v in other.items())
        for k in names:
            _close(model.mld_numeric[k], z["other_" + k])
        model.param_struct = dict(model.param_struct, **dict(zip(pnames, z["params"][2])))      # setter = update
        third = load_fixture(fixture)[0]["out_" + names[0]][2]
        _close(model.mld_numeric[names[0]], third)
        with pytest.raises(ValueError, match="Invalid keys"):
            model.get_mld_numeric(param_struct_subset=dict(not_a_parameter=1.0))
        with pytest.raises(ValueError, match="missing from param_struct"):
            model.update_param_struct(param_struct={"ts": 900.0})


def test_grid_model_changes_its_number_of_devices(twin_kernel):
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import models as M
    grid = M.GridModel(num_devices=2)
    assert grid.mld_numeric.mld_info.nomega == 2 and grid.num_devices == 2
    grid.num_devices = 5                                     # micro_grid_models.py:127-132
    assert grid.mld_numeric.mld_info.nomega == 5 and np.array_equal(grid.mld_numeric.D4, np.ones((1, 5)))
    assert grid.mld_callable.mld_info.required_params == ["P_g_max", "P_g_min", "eps"]
    z = load_fixture("callable_grid_3dev.npz")[0]
    _close(grid.mld_numeric.F2, z["num_F2"])                 # the rows do not depend on the number of devices


def test_parameter_schedule_along_the_horizon(twin_kernel):
    """PvMldSystemModel.get_mld_numeric_tilde (models/mld_model.py:1208-1226): one numeric model per step"""
    from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import models as M
    model = M.DewhModel(const_heat=False)
    same = model.get_mld_numeric_tilde(4)
    assert len(same) == 4 and all(m is same[0] for m in same)
    sched = [dict(T_h=50.0 + 3 * k, D_h=0.001 * k) for k in range(4)]
    tilde = model.get_mld_numeric_tilde(4, schedule_params_tilde=sched)
    A = [float(m.A[0, 0]) for m in tilde]
    assert len(set(A)) == 4 and A[0] > A[1] > A[2] > A[3]    # a larger draw cools the tank faster
    for k, m in enumerate(tilde):
        one = model.get_mld_numeric(param_struct_subset=sched[k])
        _close(m.A, one.A, rtol=0)
    with pytest.raises(ValueError, match="N_tilde"):
        model.get_mld_numeric_tilde(3, schedule_params_tilde=sched)


def test_to_numeric_of_a_hand_made_symbolic_model(twin_kernel):
    from pyhybridcontrol_b200.models.mld_model import MldModel, MldSystemModel
    a, ts = sp.symbols("a ts")
    sym = MldModel(dict(A=sp.Matrix([[sp.exp(-a * ts), 0], [a, 1]]), B1=np.array([[1.0], [0.0]]),
                        E=np.array([[1.0, 0.0]]), f5=sp.Matrix([[10 * a]])), ts=0)
    num = sym.to_numeric(param_struct=dict(a=0.5, ts=2.0))
    assert num.mld_type == "numeric" and num.mld_info.ts == 2.0 and num.mld_info.param_struct["a"] == 0.5
    _close(num.A, [[np.exp(-1.0), 0.0], [0.5, 1.0]], rtol=1e-15)
    _close(num.f5, [[5.0]], rtol=0)
    _close(num.B1, [[1.0], [0.0]], rtol=0)
    assert (num.mld_info.nx, num.mld_info.nu, num.mld_info.n_constraints) == (2, 1, 1)
    call = sym.to_callable()
    _close(call.A(a=0.5, ts=2.0), num.A, rtol=0)             # a single CallableMatrix evaluates through the same path
    _close(call.A(0.5, 2.0), num.A, rtol=0)
    _close(call.A(param_struct=dict(a=0.5, ts=2.0, unused=1.0)), num.A, rtol=0)
    assert not call.A(a=0.5, ts=2.0).flags.writeable
    model = MldSystemModel(mld_symbolic=sym, param_struct=dict(a=0.5, ts=2.0))
    _close(model.mld_numeric.A, num.A, rtol=0)
    model.update_mld(mld_numeric=num)                        # switching to a numeric model drops the other forms
    assert model.mld_callable is None and model.get_required_params() == set()
    batch = sym.to_numeric_batch(param_struct=dict(a=0.5, ts=2.0), overrides=dict(a=np.array([0.5, 1.0, 2.0])),
                                 device="cpu")
    assert tuple(batch["A"].shape) == (3, 2, 2) and tuple(batch["B1"].shape) == (1, 2, 1)
    _close(batch["A"][:, 0, 0].numpy(), np.exp(-2.0 * np.array([0.5, 1.0, 2.0])), rtol=1e-15)
    _close(batch["f5"][:, 0, 0].numpy(), [5.0, 10.0, 20.0], rtol=0)
