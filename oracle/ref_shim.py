"""TEST INFRASTRUCTURE ONLY -- compat shim that lets the UNMODIFIED reference run on this image.

The reference (michchr/pyhybridcontrol, mounted read-only at /root/reference) targets Python 3.6 /
numpy<1.20 / wrapt 1.x / cvxpy 1.0.x.  This module installs the aliases that those versions had and a
stub ``cvxpy`` so that the reference's own ``MldModel`` (models/mld_model.py:391), ``MldEvoMatrices``
(controllers/components/mld_evolution_matrices.py:19) and ``MldModel.lsim_k`` (models/mld_model.py:647)
import and run *unchanged* for numeric MLDs.  Nothing from the reference is copied; it is imported from
where it lies.

It is used ONLY by ``tests/golden/make_golden.py`` (to generate the committed golden fixtures) and by
the optional ``not gpu`` test that re-validates the numpy restatement when /root/reference is present.
/root/reference does not exist on the GPU box, so nothing in the gpu tests, smoke() or bench.py
imports this file.
"""
import collections
import collections.abc
import inspect
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HMPC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "controllers"))


def _formatargspec(args, varargs=None, varkw=None, defaults=None, kwonlyargs=(), kwonlydefaults=None,
                   annotations=None, formatannotation=None, **_ignored):
    """Replacement for inspect.formatargspec (removed in Python 3.11); the reference exec()s the result
    (utils/func_utils.py:72-83) so annotations have to be emitted as well."""
    annotations = annotations or {}
    kwonlydefaults = kwonlydefaults or {}
    fa = formatannotation or (lambda a: repr(a))

    def fmt(name):
        if name in annotations:
            return "%s: %s" % (name, fa(annotations[name]))
        return name

    specs = []
    first_default = len(args) - len(defaults) if defaults else None
    for i, a in enumerate(args):
        s = fmt(a)
        if defaults and i >= first_default:
            s += "=" + repr(defaults[i - first_default])
        specs.append(s)
    if varargs is not None:
        specs.append("*" + fmt(varargs))
    elif kwonlyargs:
        specs.append("*")
    for a in kwonlyargs:
        s = fmt(a)
        if a in kwonlydefaults:
            s += "=" + repr(kwonlydefaults[a])
        specs.append(s)
    if varkw is not None:
        specs.append("**" + fmt(varkw))
    out = "(" + ", ".join(specs) + ")"
    if "return" in annotations:
        out += " -> " + fa(annotations["return"])
    return out


def _install_cvxpy_stub():
    if "cvxpy" in sys.modules:
        return
    cvx = types.ModuleType("cvxpy")

    class Expression(object):
        pass

    class Parameter(Expression):
        def __init__(self, shape=(), name=None, value=None, **kw):
            self.shape, self.name, self.value = shape, name, value

    class Variable(Expression):
        def __init__(self, shape=(), **kw):
            raise NotImplementedError("cvxpy is not installed; shimmed reference is numeric-only")

    class Problem(object):
        def __init__(self, *a, **k):
            raise NotImplementedError("cvxpy is not installed; shimmed reference is numeric-only")

    class SolverError(Exception):
        pass

    err = types.ModuleType("cvxpy.error")
    err.SolverError = SolverError
    exprs = types.ModuleType("cvxpy.expressions")
    expr = types.ModuleType("cvxpy.expressions.expression")
    expr.Expression = Expression
    exprs.expression = expr
    cvx.Expression, cvx.Parameter, cvx.Variable, cvx.Problem = Expression, Parameter, Variable, Problem
    cvx.error, cvx.expressions = err, exprs
    cvx.GUROBI, cvx.CPLEX = "GUROBI", "CPLEX"
    sys.modules["cvxpy"] = cvx
    sys.modules["cvxpy.error"] = err
    sys.modules["cvxpy.expressions"] = exprs
    sys.modules["cvxpy.expressions.expression"] = expr


_installed = False


def install():
    """Idempotently install the aliases + stubs and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import warnings
    import numpy as np
    import wrapt
    import wrapt.decorators

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, val in (("int", int), ("str", str), ("bool", bool), ("float", float), ("object", object)):
            if name not in np.__dict__:
                setattr(np, name, val)
        if "NaN" not in np.__dict__:
            np.NaN = np.nan
        if "asscalar" not in np.__dict__:
            np.asscalar = lambda a: np.asarray(a).item()
        if "issubsctype" not in np.__dict__:
            np.issubsctype = lambda a, t: np.issubdtype(np.asarray(a).dtype, t)
    for name in ("Container", "Sequence", "Mapping", "MutableMapping", "Iterable", "Callable", "Hashable"):
        if not hasattr(collections, name):
            setattr(collections, name, getattr(collections.abc, name))
    if not hasattr(inspect, "formatargspec"):
        inspect.formatargspec = _formatargspec
    if not hasattr(wrapt.decorators, "AdapterWrapper"):
        wrapt.decorators.AdapterWrapper = wrapt.decorators._AdapterFunctionWrapper
    _install_cvxpy_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load():
    """Returns (MldModel, MldEvoMatrices) classes of the unmodified reference."""
    install()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.mld_model import MldModel  # noqa: reference module
        from controllers.components.mld_evolution_matrices import MldEvoMatrices  # noqa: reference module
        import models.mld_model as _mm
    if not getattr(_mm, "_hmpc_numeric_only", False):
        # wrapt 2.x proxies break isinstance() against the reference's CallableMatrix; the shim only ever
        # feeds numeric MLDs, so an empty placeholder class keeps `isinstance(mat, CallableMatrix)` False.
        _mm.CallableMatrix = type("CallableMatrix", (), {})
        _mm._hmpc_numeric_only = True
    return MldModel, MldEvoMatrices


def load_symbolic():
    """The reference's symbolic / callable front-end, unmodified: returns (MldModel, MldSystemModel, CallableMatrix,
    example micro_grid_models module, example parameters module).

    wrapt 2.x proxies make ``isinstance(x, CallableMatrix)`` raise inside abc (the class mixes ABCMeta with wrapt's
    ObjectProxy); the instance / subclass checks of the reference's metaclass are replaced by plain MRO look-ups --
    an environment alias like the others here, no arithmetic is touched.  Must be called in a process that has NOT
    called ``load()`` (which swaps CallableMatrix for a placeholder to keep the numeric-only path simple)."""
    install()
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import utils.matrix_utils as rmu  # noqa: reference module
        rmu.CallableMatrixMeta.__instancecheck__ = lambda cls, inst: cls in type(inst).__mro__
        rmu.CallableMatrixMeta.__subclasscheck__ = lambda cls, sub: isinstance(sub, type) and cls in sub.__mro__
        import models.mld_model as mm  # noqa: reference module
        if getattr(mm, "_hmpc_numeric_only", False):
            raise RuntimeError("load() already replaced the reference's CallableMatrix in this process")
        models = importlib.import_module("examples.residential_mg_with_pv_and_dewhs.modelling.micro_grid_models")
        params = importlib.import_module("examples.residential_mg_with_pv_and_dewhs.modelling.parameters")
    return mm.MldModel, mm.MldSystemModel, rmu.CallableMatrix, models, params


def load_controllers():
    """The reference's whole per-step path, unmodified, with ``oracle/mini_cvxpy.py`` standing in for cvxpy's
    modelling layer: returns a namespace with MldModel, MldSystemModel, MpcController, the controller_base module,
    the example's models / agents / parameters modules and the cvxpy stand-in.  Must be the FIRST loader called in
    the process (the stand-in has to be registered before the reference imports cvxpy)."""
    import importlib
    import types as _types
    import warnings
    import numpy as np
    from oracle import mini_cvxpy
    cvx = mini_cvxpy.install()
    MldModel, MldSystemModel, CallableMatrix, models, params = load_symbolic()
    if "dims" not in getattr(np.unravel_index, "__doc__", "") and not getattr(np, "_hmpc_unravel_alias", False):
        _unravel = np.unravel_index          # numpy 2 renamed unravel_index(dims=) to shape=
        np.unravel_index = lambda indices, shape=None, order="C", dims=None: _unravel(
            indices, shape if shape is not None else dims, order=order)
        np._hmpc_unravel_alias = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cb = importlib.import_module("controllers.controller_base")
        mpc = importlib.import_module("controllers.mpc_controller")
        agents = importlib.import_module("examples.residential_mg_with_pv_and_dewhs.modelling.micro_grid_agents")
    return _types.SimpleNamespace(MldModel=MldModel, MldSystemModel=MldSystemModel, MpcController=mpc.MpcController,
                                  controller_base=cb, models=models, params=params, agents=agents, cvx=cvx)


EVO_NAMES = ("Phi_x", "Gamma_v", "Gamma_omega", "Gamma_5",
             "L_x", "L_v", "L_omega", "L_5",
             "H_x", "H_v", "H_omega", "H_5")


def reference_condense(mats, N_p, N_tilde=None, bin_dims=None):
    """Run the reference's own condensing on a dict of numeric matrices -> dict of the 12 *_N_tilde arrays."""
    import numpy as np
    MldModel, MldEvoMatrices = load()
    N_tilde = N_p + 1 if N_tilde is None else N_tilde
    mld = MldModel(**{k: np.array(v, dtype=float) for k, v in mats.items()}, **(bin_dims or {}))
    evo = MldEvoMatrices(N_p=N_p, N_tilde=N_tilde, mld_numeric_k=mld, mld_numeric_tilde=None)
    out = {}
    for grp, names in (("state_input", EVO_NAMES[0:4]), ("output", EVO_NAMES[4:8]), ("constraint", EVO_NAMES[8:12])):
        for nm in names:
            out[nm] = np.array(evo[grp][nm + "_N_tilde"], dtype=float)
    info = mld.mld_info
    dims = {k: int(info[k]) for k in ("nx", "nu", "ndelta", "nz", "nmu", "nomega", "ny", "n_constraints", "nv",
                                       "nu_l", "ndelta_l", "nz_l", "nmu_l")}
    return out, dims, mld
