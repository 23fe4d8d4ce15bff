"""CPU: symbolic / callable model front-end -- the oracle and the sympy -> register-program
compiler against golden vectors produced by the UNMODIFIED reference's CallableMatrix / MldSystemModel
(tests/golden/make_golden_callable.py), plus the host-side API rules.  No GPU: programs are executed by the numpy
twin of the kernel's interpreter (tests/expr_vm_twin.py); the kernel itself is checked in test_gpu_callable.py."""
import ctypes
import glob
import os

import numpy as np
import pytest
import sympy as sp

import expr_vm_twin as twin
from oracle import callable as ocall
from pyhybridcontrol_b200.models.mld_model import MldModel, MldSystemModel
from pyhybridcontrol_b200.utils import matrix_utils as mu
from pyhybridcontrol_b200.utils.matrix_utils import CallableMatrix, CallableMatrixConstant, ExprProgram
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import models as ex_models
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import parameters as ex_par

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "callable_*.npz")))
MAT_NAMES = ("A", "B1", "B2", "B3", "B4", "b5", "C", "D1", "D2", "D3", "D4", "d5",
             "E", "F1", "F2", "F3", "F4", "f5", "G", "Psi")


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name))
    names = [str(n) for n in z["names"]]
    mats = {k: sp.sympify(str(z["srepr_" + k])) for k in names}
    return z, names, mats, [str(n) for n in z["param_names"]]


def rel_err(got, ref):
    scale = np.maximum(np.abs(ref), 1e-300)
    return float(np.max(np.abs(got - ref) / scale)) if ref.size else 0.0


def test_fixtures_present():
    assert set(FIXTURES) >= {"callable_dewh_control.npz", "callable_dewh_sim.npz", "callable_grid_3dev.npz",
                             "callable_pv.npz", "callable_resd.npz", "callable_all_ops.npz"}


@pytest.mark.parametrize("fixture", FIXTURES)
def test_oracle_matches_reference(fixture):
    """oracle/callable.py (lambdify restatement) == the reference's own CallableMatrix, to the last bit."""
    z, names, mats, pnames = load_fixture(fixture)
    out = ocall.evaluate_batch(mats, pnames, z["params"])
    for k in names:
        assert out[k].shape == z["out_" + k].shape
        assert np.array_equal(out[k], z["out_" + k], equal_nan=True), (fixture, k, rel_err(out[k], z["out_" + k]))


@pytest.mark.parametrize("fixture", FIXTURES)
def test_compiled_program_matches_reference(fixture):
    """The register program (as the kernel executes it) reproduces the reference: + - * / in the printer's order are
    bit-exact; pow / powi chains may differ from numpy's pow by a few ulp -> 1e-13."""
    z, names, mats, pnames = load_fixture(fixture)
    prog = ExprProgram(mats, param_names=pnames)
    assert twin.valid(prog.instructions, prog.n_regs, len(pnames), prog.n_out)
    out = twin.run_program(prog, z["params"])
    for k in names:
        assert rel_err(out[k], z["out_" + k]) <= 1e-13, (fixture, k)
    if fixture in ("callable_dewh_control.npz", "callable_dewh_sim.npz", "callable_grid_3dev.npz"):
        for k in names:        # no pow in these models: every operation rounds like the reference's
            assert np.array_equal(out[k], z["out_" + k]), (fixture, k)


def test_program_shares_subexpressions_and_reuses_registers():
    z, names, mats, pnames = load_fixture("callable_dewh_control.npz")
    prog = ExprProgram(mats)
    ops = prog.instructions[:, 0].tolist()
    assert ops.count(mu.OP_EXP) == 1                      # e^{A_c ts} appears in A, B1, B4 and b5: evaluated once
    assert ops.count(mu.OP_OUT) == prog.n_out == 6
    assert ops.count(mu.OP_PARAM) == len(prog.param_names) == 11
    assert prog.n_regs <= 8                               # linear scan, not one register per value
    assert prog.param_names == tuple(sorted(prog.param_names))
    assert prog.bytes_per_agent() == 8 * (11 + 6)
    big = ExprProgram(load_fixture("callable_dewh_sim.npz")[2])
    assert big.n_ins < 260 and big.n_regs <= 24          # the reference's pinv() expression: ~1900 operations as text
    assert "EXP" in big.disassemble()


def test_own_dewh_model_matches_reference_numbers():
    """The package's DewhModel writes the input gain in closed form ((A - 1)/A_c instead of the reference's symbolic
    pinv); the numbers agree with the reference's to 1e-10 (BASELINE.json: matrices within 1e-10 relative)."""
    for const_heat, fixture in ((True, "callable_dewh_control.npz"), (False, "callable_dewh_sim.npz")):
        z, names, _, pnames = load_fixture(fixture)
        model = ex_models.DewhModel.get_dewh_mld_symbolic(const_heat=const_heat).to_callable()
        assert model.mld_type == "callable"
        assert list(model.mld_info.required_params) == sorted(str(n) for n in z["required_params"])
        prog = model.program
        assert sorted(prog.mat_names) == sorted(names)
        cols = [pnames.index(n) for n in prog.param_names]
        out = twin.run_program(prog, z["params"][:, cols])
        for k in names:
            assert rel_err(out[k], z["out_" + k]) <= 1e-10, (fixture, k, rel_err(out[k], z["out_" + k]))
        info = model.mld_info
        for key in ("nx", "nu", "ndelta", "nz", "nomega", "ny", "nmu", "nv", "n_constraints", "nu_l"):
            assert int(info[key]) == int(z["info_" + key]), key
        for k in MAT_NAMES:                                # constant blocks: identical to the reference's
            if model[k].is_constant:
                ref = z["num_" + k]
                got = np.asarray(model[k](), dtype=float)
                assert got.shape == ref.shape or (got.size == 0 and ref.size == 0), k
                if ref.size:
                    assert np.array_equal(got, ref), k


def test_own_grid_pv_resd_models_match_reference():
    for sym, fixture in ((ex_models.GridModel.get_grid_mld_symbolic(3), "callable_grid_3dev.npz"),
                         (ex_models.PvModel.get_pv_mld_symbolic(), "callable_pv.npz"),
                         (ex_models.ResDemandModel.get_res_demand_mld_symbolic(), "callable_resd.npz")):
        z, names, _, pnames = load_fixture(fixture)
        model = sym.to_callable()
        prog = model.program
        cols = [pnames.index(n) for n in prog.param_names]
        out = twin.run_program(prog, z["params"][:, cols])
        for k in names:
            assert np.array_equal(out[k], z["out_" + k]), (fixture, k)
        for key in ("nx", "nu", "ndelta", "nz", "nomega", "ny", "nmu", "n_constraints"):
            assert int(model.mld_info[key]) == int(z["info_" + key]), (fixture, key)


def test_callable_matrix_surface():
    a, b = sp.symbols("a b")
    cm = CallableMatrix(sp.Matrix([[a * b, 1], [0, sp.exp(a)]]), "M")
    assert type(cm) is CallableMatrix and not cm.is_constant and not cm.is_all_zero and not cm.is_empty
    assert (cm.shape, cm.size, cm.ndim, cm.matrix_name, cm.__name__) == ((2, 2), 4, 2, "M", "M")
    assert cm.required_params == ["a", "b"]
    assert "M(a, b, *, param_struct=None)" in repr(cm)
    const = CallableMatrix(np.array([[1.0, 2.0]]), "K")
    assert isinstance(const, CallableMatrixConstant) and const.is_constant and const.required_params == []
    val = const(param_struct=dict(a=1.0))
    assert np.array_equal(val, [[1.0, 2.0]]) and not val.flags.writeable          # no GPU involved
    assert CallableMatrix(np.zeros((2, 1))).is_all_zero
    assert CallableMatrix(np.zeros((0, 0))).is_empty
    assert isinstance(CallableMatrix(sp.Matrix([[2, 3]])), CallableMatrixConstant)
    assert CallableMatrix(5.0)().shape == (1, 1) and CallableMatrix([1.0, 2.0])().shape == (2, 1)
    cp = cm.copy()
    assert cp is not cm and cp.required_params == cm.required_params and cp.matrix_name == "M"
    with pytest.raises(TypeError):
        CallableMatrixConstant(sp.Matrix([[a]]))
    # argument binding errors are raised before anything is evaluated (utils/matrix_utils.py:441-462)
    with pytest.raises(TypeError, match="missing"):
        cm(param_struct=dict(a=1.0))
    with pytest.raises(TypeError, match="multiple values"):
        cm(a=1.0, param_struct=dict(a=2.0, b=1.0))
    with pytest.raises(TypeError, match="unexpected keyword"):
        cm(a=1.0, b=2.0, c=3.0)
    with pytest.raises(TypeError, match="dictionary like"):
        cm(param_struct=3)


def test_python_functions_are_traced():
    def A_fun(alpha, ts):
        return [[sp.exp(-alpha * ts), 0.0], [alpha ** 2 / (1 + ts), 1.0]]

    cm = CallableMatrix(A_fun, "A")
    assert cm.required_params == ["alpha", "ts"] and cm.shape == (2, 2)
    params = np.array([[0.3, 2.0], [1.5, 0.25]])
    out = twin.run_program(cm.program, params)["A"]
    for i, (al, ts) in enumerate(params):
        assert np.allclose(out[i], [[np.exp(-al * ts), 0.0], [al ** 2 / (1 + ts), 1.0]], rtol=1e-15, atol=0)

    assert isinstance(CallableMatrix(lambda: np.eye(2)), CallableMatrixConstant)

    def A_numpy(alpha, ts):                                # written the way users of the reference write them
        a = np.exp(-alpha * ts)
        row = np.array([a, np.sqrt(1 + alpha ** 2)])
        return np.array([[a, np.sqrt(1 + alpha ** 2), np.maximum(alpha, 1.0)],
                         [(a - 1) / (-alpha), np.tanh(ts) * np.abs(alpha - 2),
                          np.arctan2(alpha, ts) + np.minimum(ts, alpha) ** 2],
                         [np.power(alpha, 1.5), 2.0 * row[1] - row[0], np.sum(row * np.array([1.0, 2.0]))]])

    cm = CallableMatrix(A_numpy, "A")
    assert cm.shape == (3, 3) and cm.required_params == ["alpha", "ts"]
    params = np.array([[0.3, 2.0], [1.5, 0.25], [2.5, 3.0]])
    out = twin.run_program(cm.program, params)["A"]
    for i, (al, ts) in enumerate(params):
        np.testing.assert_allclose(out[i], A_numpy(al, ts), rtol=1e-14, atol=0)

    def keyword_only(alpha, *, ts, param_struct=None):
        return np.exp(-alpha * ts)                         # a scalar -> (1, 1)

    cm = CallableMatrix(keyword_only)
    assert cm.shape == (1, 1) and cm.required_params == ["alpha", "ts"] and cm.matrix_name == "keyword_only"
    with pytest.raises(TypeError, match="heaviside has no GPU instruction"):
        CallableMatrix(lambda x: [[np.heaviside(x, 0.5)]])

    def branches_on_value(x):
        return [[1.0 if x > 0 else 2.0]]

    with pytest.raises(TypeError, match="cannot be traced"):
        CallableMatrix(branches_on_value)
    with pytest.raises(TypeError):
        CallableMatrix(lambda *args: 1.0)


def test_unsupported_nodes_fail_loudly():
    x = sp.Symbol("x")
    with pytest.raises(NotImplementedError, match="Piecewise"):
        ExprProgram({"A": sp.Matrix([[sp.Piecewise((x, x > 0), (0, True))]])})
    with pytest.raises(NotImplementedError):
        ExprProgram({"A": sp.Matrix([[sp.I * x]])})
    with pytest.raises(ValueError):
        ExprProgram({"A": sp.Matrix([[x]])}, param_names=["y"])
    with pytest.raises(ValueError):
        ExprProgram({})
    # conj / re of a real parameter are the parameter (sympy's pinv leaves conjugates behind)
    prog = ExprProgram({"A": sp.Matrix([[sp.conjugate(x) * sp.re(x) + sp.im(x)]])})
    assert np.array_equal(twin.run_program(prog, [[3.0]])["A"], [[[9.0]]])


def test_mld_model_types_and_conversions():
    a, ts = sp.symbols("a ts")
    sym = MldModel(dict(A=sp.Matrix([[sp.exp(-a * ts)]]), B1=[[1.0]], E=np.array([[1], [-1]]),
                        f5=sp.Matrix([[a], [-a]]), Psi=-np.eye(2)), nu_l=1, ts=0)
    assert sym.mld_type == "symbolic" and sym.mld_info.required_params == ["a", "ts"]
    assert (sym.mld_info.nx, sym.mld_info.nu, sym.mld_info.nmu, sym.mld_info.n_constraints) == (1, 1, 2, 2)
    assert isinstance(sym.B1, np.ndarray) and isinstance(sym.A, sp.MatrixBase)    # mixed storage, as the reference's
    call = sym.to_callable()
    assert call.mld_type == "callable" and all(isinstance(m, CallableMatrix) for m in call.values())
    assert call.mld_info.nu_l == 1 and call.mld_info.required_params == ["a", "ts"]
    assert sorted(call.program.mat_names) == ["A", "f5"]
    assert MldModel(A=lambda: 1.0).mld_type == "callable"
    num = MldModel(A=[[0.5]])
    assert num.mld_type == "numeric" and num.mld_info.required_params is None and num.to_numeric() is num
    with pytest.raises(TypeError):
        num.program
    with pytest.raises(TypeError, match="param_struct"):
        call.to_numeric()                                  # parameters needed, none stored
    with pytest.raises(ValueError):
        MldModel(dict(A=sp.Matrix([[a, 1]])), ts=0)        # A not square
    with pytest.raises(TypeError):
        sym.lsim_k(x_k=[1.0], u_k=[0.0])                   # needs numbers first


def test_system_model_argument_rules():
    a = sp.Symbol("a")
    sym = MldModel(dict(A=sp.Matrix([[a]])), ts=0)
    with pytest.raises(ValueError, match="missing from param_struct"):
        MldSystemModel(mld_symbolic=sym, param_struct=dict(b=1.0))
    with pytest.raises(ValueError, match="Only one of"):
        MldSystemModel(mld_symbolic=sym, mld_numeric=MldModel(A=[[1.0]]), param_struct=dict(a=1.0))
    with pytest.raises(TypeError, match="mld_type"):
        MldSystemModel(mld_callable=sym, param_struct=dict(a=1.0))
    with pytest.raises(TypeError):
        MldSystemModel(mld_numeric=3)
    plain = MldSystemModel(mld_numeric=MldModel(A=[[1.0]]), param_struct=dict(k=2.0))
    assert plain.get_required_params() == set() and plain.mld_callable is None
    assert plain.get_mld_numeric() is plain.mld_numeric
    with pytest.raises(ValueError, match="Invalid keys"):
        plain.get_mld_numeric(param_struct_subset=dict(zzz=1.0))
    with pytest.raises(TypeError, match="mld_callable"):
        plain.get_mld_numeric(param_struct_subset=dict(k=3.0))
    with pytest.raises(ValueError):
        ex_models.GridModel(num_devices=2.5)
    assert ex_par.dewh_param_struct["ts"] == 900.0


def test_c_abi_argument_validation_without_gpu():
    from pyhybridcontrol_b200 import cabi
    lib = cabi.lib()
    sizes = (ctypes.c_int32 * 2)(1, 2)
    prog = (ctypes.c_int32 * 8)()
    f = lib.hmpc_param_eval_f64
    assert f(4, 1, 1, 2, None, 2, sizes, None, None, None) == -1              # no program
    assert f(4, 1, 0, 2, prog, 2, sizes, None, None, None) == -1              # no registers
    assert f(4, 1, 1, 2, prog, 21, sizes, None, None, None) == -1             # too many matrices
    assert f(-1, 0, 1, 2, prog, 2, sizes, None, None, None) == -1
    assert f(4, 1, 1, 2, prog, 2, sizes, None, None, None) == -1              # parameters expected, none given
    assert f(4, 0, 1, 2, prog, 2, sizes, None, None, None) == -1              # outputs expected, no buffer
    assert f(0, 0, 1, 2, prog, 2, sizes, None, None, None) == 0               # empty batch: nothing to do
    assert cabi.param_eval_bytes_per_agent(13, [1, 1, 1, 1, 2]) == 8 * 19


@pytest.mark.parametrize("fixture", FIXTURES)
def test_register_preloaded_program_form_is_equivalent(fixture):
    """the program form of the experimental kernel (hmpc_param_eval_v2_f64: parameters preloaded as registers 0..P-1,
    no PARAM instructions) performs the same operations in the same order: identical results to the last bit"""
    z, names, mats, pnames = load_fixture(fixture)
    prog = ExprProgram(mats, param_names=pnames)
    ins2, R2 = prog.instructions_v2
    P = len(pnames)
    assert twin.valid(ins2, R2, P, prog.n_out, preload=True)
    assert not np.any(ins2[:, 0] == mu.OP_PARAM) and ins2.shape[0] < prog.n_ins
    writes = ins2[ins2[:, 0] != mu.OP_OUT][:, 1]
    assert writes.size == 0 or writes.min() >= P                       # parameters are never overwritten
    a = twin.run(prog.instructions, prog.n_regs, prog.mat_sizes, z["params"])
    b = twin.run(ins2, R2, prog.mat_sizes, z["params"], preload=True)
    assert np.array_equal(a, b, equal_nan=True)
    # the v1 form is not a valid v2 program and vice versa where parameters exist
    if P:
        assert not twin.valid(prog.instructions, prog.n_regs, P, prog.n_out, preload=True)


def test_c_abi_v2_argument_validation_without_gpu():
    from pyhybridcontrol_b200 import cabi
    f = cabi.lib().hmpc_param_eval_v2_f64
    sizes = (ctypes.c_int32 * 1)(1)
    prog = (ctypes.c_int32 * 8)()
    assert f(4, 2, 2, 2, prog, 1, sizes, None, None, None) == -1               # n_regs must exceed n_params
    assert f(4, 0, 1, 2, None, 1, sizes, None, None, None) == -1               # no program
    assert f(0, 2, 3, 2, prog, 1, sizes, prog, None, None) == 0                # empty batch
