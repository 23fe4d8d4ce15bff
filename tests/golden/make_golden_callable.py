"""Generates tests/golden/callable_*.npz from the UNMODIFIED reference's symbolic / callable front-end
(utils/matrix_utils.py CallableMatrix, models/mld_model.py MldModel.to_callable / to_numeric, MldSystemModel,
examples/.../modelling/micro_grid_models.py) run under oracle/ref_shim.load_symbolic().

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_callable.py

Every fixture holds
  names            matrix names that depend on parameters
  srepr_<name>     sympy.srepr of the reference's symbolic matrix (so the test re-creates the SAME expression tree
                   without the reference)
  param_names      column order of params
  params [B, P]    parameter sets
  out_<name>       [B, rows, cols] = reference CallableMatrix(**params) for every row
and, for the device models, the complete numeric model at the default parameters:
  num_<name>       the 20 matrices of <Model>().mld_numeric ;  info_* its dimensions.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.simplefilter("ignore")

from oracle import ref_shim  # noqa: E402

MldModel, MldSystemModel, CallableMatrix, ref_models, ref_params = ref_shim.load_symbolic()
import sympy as sp  # noqa: E402

MAT_NAMES = ("A", "B1", "B2", "B3", "B4", "b5", "C", "D1", "D2", "D3", "D4", "d5",
             "E", "F1", "F2", "F3", "F4", "f5", "G", "Psi")
INFO = ("nx", "nu", "ndelta", "nz", "nomega", "ny", "nmu", "nv", "n_constraints", "nu_l", "ndelta_l", "nmu_l")


def fixture_from_model(sys_model, param_sets, fname):
    sym, call = sys_model.mld_symbolic, sys_model.mld_callable
    names = [k for k in MAT_NAMES if isinstance(sym[k], (sp.Expr, sp.MatrixBase)) and sym[k].free_symbols]
    param_names = sorted({str(s) for k in names for s in sym[k].free_symbols})
    data = dict(names=np.array(names), param_names=np.array(param_names),
                params=np.array([[float(p[n]) for n in param_names] for p in param_sets]),
                required_params=np.array(sorted(sys_model.get_required_params())))
    for k in names:
        data["srepr_" + k] = np.array(sp.srepr(sp.Matrix(sym[k])))
        data["out_" + k] = np.stack([np.asarray(call[k](param_struct=p), dtype=float) for p in param_sets])
    num = sys_model.mld_numeric
    for k in MAT_NAMES:
        data["num_" + k] = np.asarray(num[k], dtype=float)
    for k in INFO:
        data["info_" + k] = np.array(int(num.mld_info[k]))
    # MldSystemModel.get_mld_numeric for the first non-default parameter set (models/mld_model.py:1128-1149)
    other = sys_model.get_mld_numeric(param_struct=param_sets[1], invalid_param_check=False)
    for k in names:
        data["other_" + k] = np.asarray(other[k], dtype=float)
    np.savez_compressed(os.path.join(HERE, fname), **data)
    print(fname, names, data["params"].shape)


def jitter(base, rng, keys, rel=0.1, extra=None):
    p = dict(base)
    for k in keys:
        p[k] = float(base[k]) * (1.0 + rel * rng.uniform(-1, 1))
    for k, (lo, hi) in (extra or {}).items():
        p[k] = float(rng.uniform(lo, hi))
    return p


def main():
    rng = np.random.default_rng(20261018)
    dewh_keys = ("C_w", "A_h", "U_h", "m_h", "T_w", "T_inf", "P_h_Nom", "T_h_min", "T_h_max", "T_h_Nom")
    base = dict(ref_params.dewh_param_struct)
    sets = [dict(base)] + [jitter(base, rng, dewh_keys) for _ in range(63)]
    fixture_from_model(ref_models.DewhModel(const_heat=True), sets, "callable_dewh_control.npz")
    sets = [dict(base)] + [jitter(base, rng, dewh_keys, extra=dict(T_h=(20.0, 85.0), D_h=(0.0, 0.08)))
                           for _ in range(63)]
    fixture_from_model(ref_models.DewhModel(const_heat=False), sets, "callable_dewh_sim.npz")
    base = dict(ref_params.grid_param_struct)
    sets = [dict(base)] + [jitter(base, rng, ("P_g_min", "P_g_max")) for _ in range(7)]
    fixture_from_model(ref_models.GridModel(num_devices=3), sets, "callable_grid_3dev.npz")
    base = dict(ref_params.pv_param_struct)
    sets = [dict(base)] + [jitter(base, rng, ("P_pv_max",), extra=dict(P_pv_units=(1, 50))) for _ in range(7)]
    fixture_from_model(ref_models.PvModel(), sets, "callable_pv.npz")
    base = dict(ref_params.res_demand_param_struct)
    sets = [dict(base)] + [jitter(base, rng, ("P_res_ave",), extra=dict(P_res_units=(1, 50))) for _ in range(7)]
    fixture_from_model(ref_models.ResDemandModel(), sets, "callable_resd.npz")

    # a made-up 2-state model that exercises every instruction of the evaluator, through the reference's MldModel
    a, b, c, w = sp.symbols("a b c w")
    A = sp.Matrix([[sp.exp(-a * w), sp.sin(b) * sp.cos(c) / (1 + a ** 2)],
                   [sp.sqrt(a + b ** 2) - sp.log(a + 2), sp.tanh(c - b) * a ** -3]])
    B1 = sp.Matrix([[sp.Abs(b - c) ** sp.Rational(3, 2) + sp.Min(a, b, c)],
                    [sp.Max(a, b * c) - sp.atan2(b, a) + sp.floor(10 * c) / 7]])
    b5 = sp.Matrix([[sp.sinh(b / 4) + sp.cosh(c / 4) - sp.tan(a / 8)],
                    [sp.asin(b / 4) * sp.acos(c / 4) + sp.atan(a) - sp.sign(b - 1) * sp.ceiling(a) + a ** c
                     + sp.Rational(2, 3) * w - 1 / sp.sqrt(w)]])
    mld_sym = MldModel(dict(A=A, B1=B1, b5=b5), ts=0)
    base = dict(a=0.7, b=1.3, c=0.4, w=2.5, ts=0)
    sets = [dict(base)] + [dict(a=rng.uniform(0.1, 3), b=rng.uniform(0.1, 3.5), c=rng.uniform(0.05, 3),
                                w=rng.uniform(0.5, 4), ts=0) for _ in range(95)]
    fixture_from_model(MldSystemModel(mld_symbolic=mld_sym, param_struct=base), sets, "callable_all_ops.npz")


if __name__ == "__main__":
    main()
