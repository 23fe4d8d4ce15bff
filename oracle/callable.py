"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's symbolic / callable model evaluation.

Reference path (all citations relative to /root/reference):
  * ``CallableMatrix._process_matrix_func`` (utils/matrix_utils.py:334-347): a sympy matrix becomes
    ``sympy.lambdify(param_sym_tup, Matrix, modules="numpy", dummify=False)``;
  * ``_get_param_sym_tup`` (:372-380): the arguments are the free symbols sorted by name;
  * ``CallableMatrix._matrix_wrapper`` (:441-470): called with the entries of ``param_struct`` the function names,
    result made 2-D and read-only;
  * ``MldModel.to_numeric`` (models/mld_model.py:791-793): ``{mat_id: mat_callable(param_struct=param_struct)}``;
  * ``MldSystemModel.get_mld_numeric`` (:1128-1149): the numeric model for a parameter set.

Pinning: **pinned** -- the unmodified reference's own CallableMatrix / DewhModel / GridModel / PvModel /
ResDemandModel run in the build container under ``oracle/ref_shim.py`` (``load_symbolic``);
``tests/golden/make_golden_callable.py`` generated ``tests/golden/callable_*.npz`` from them, and
``tests/test_callable_front_end.py`` checks this restatement against those vectors.
"""
import numpy as np


def lambdify_matrix(matrix):
    """sympy matrix -> (function, argument names), as utils/matrix_utils.py:339-343 + :372-380."""
    import sympy as sp
    system_matrix = sp.Matrix(matrix)
    sym_dict = {str(sym): sym for sym in system_matrix.free_symbols}
    names = tuple(sorted(sym_dict))
    func = sp.lambdify(tuple(sym_dict[n] for n in names), system_matrix, modules="numpy", dummify=False)
    return func, names


def evaluate(matrix, param_struct):
    """One matrix for one parameter set (utils/matrix_utils.py:441-470)."""
    func, names = lambdify_matrix(matrix)
    ret = np.asarray(func(**{n: param_struct[n] for n in names}), dtype=np.float64)
    if ret.ndim < 2:
        ret = ret.reshape(-1, 1)
    return ret


def evaluate_batch(matrices, param_names, params):
    """dict name -> sympy matrix, params [B, P] in ``param_names`` order -> dict name -> [B, rows, cols];
    the per-agent loop of the reference (models/mld_model.py:791-793 for every agent)."""
    params = np.asarray(params, dtype=np.float64)
    funcs = {k: lambdify_matrix(m) for k, m in matrices.items()}
    out = {}
    for k, (func, names) in funcs.items():
        cols = [param_names.index(n) for n in names]
        rows = []
        for b in range(params.shape[0]):
            with np.errstate(all="ignore"):
                val = np.asarray(func(*[params[b, c] for c in cols]), dtype=np.float64)
            rows.append(val if val.ndim == 2 else val.reshape(-1, 1))
        out[k] = np.stack(rows) if rows else np.zeros((0,) + tuple(int(s) for s in matrices[k].shape))
    return out
