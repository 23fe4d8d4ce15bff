"""TEST INFRASTRUCTURE ONLY -- a stand-in for the part of cvxpy's MODELLING layer the reference touches, so that the
UNMODIFIED reference's problem assembly (EvoVariables, ObjectiveAtoms, gen_evo_constraints, MpcController.build /
solve / feedback) can run in the build container, where cvxpy is not installed.

What it is: lazily evaluated numeric expressions with cvxpy's documented shape and operator semantics --
2-D expressions (scalars have shape ``()``), ``@`` = matrix product, ``*`` = scaling when one side has a single entry
and matrix product otherwise (cvxpy 1.0), ``reshape`` in column-major (Fortran) order, ``vstack`` / ``hstack``,
slicing, ``.T``, ``multiply`` (elementwise), ``sum`` / ``sum_squares`` / ``quad_form`` / ``norm1`` / ``norm_inf`` with
``axis``, ``<=`` / ``>=`` / ``==`` constraints, ``Variable(shape, boolean=[index tuples], nonneg=...)``,
``Parameter(shape, value=...)``, ``Problem(Minimize(expr), constraints)``.

What it is not: a convex-optimisation compiler.  ``canonical_form`` recovers  min c'x + c0  s.t.  G x <= h, A x = b,
bounds, integrality  by PROBING the expression trees at 0 and at the unit vectors (exact for affine expressions;
an objective that is not affine is reported as such and can still be evaluated pointwise), and ``Problem.solve`` hands
that MILP to HiGHS (scipy.optimize.milp) -- the oracle's offline backend, not the reference's Gurobi.

Third-party semantics restated here: cvxpy (unpinned by the reference; its API usage implies 1.0.x --
``Variable(boolean=[...], nonneg=...)`` controllers/components/variables.py:219-221, ``Problem.solve(parallel=...)``
controllers/controller_base.py:509-512).  Used only by oracle/ref_shim.load_controllers and the golden generators.
"""
import itertools
import time
import types

import numpy as np

_counter = itertools.count()


class SolverError(Exception):
    pass


class _Unset(Exception):
    pass


def _shape_of(x):
    if isinstance(x, Expression):
        return x.shape
    return np.shape(x)


def _wrap(x):
    return x if isinstance(x, Expression) else Constant(x)


def _binary_shape(a, b):
    sa, sb = _shape_of(a), _shape_of(b)
    if int(np.prod(sa, dtype=int)) == 1:
        return sb if len(sb) >= len(sa) else sa
    if int(np.prod(sb, dtype=int)) == 1:
        return sa
    if sa != sb:
        raise ValueError("Incompatible dimensions %s %s" % (sa, sb))
    return sa


class Expression(object):
    __array_ufunc__ = None        # numpy operators defer to the reflected methods below
    __hash__ = object.__hash__

    def __init__(self, shape, args=()):
        self.shape = tuple(int(s) for s in shape)
        self.args = tuple(args)

    # ---- numbers
    def _eval(self):
        raise NotImplementedError

    @property
    def value(self):
        try:
            return self._eval()
        except _Unset:
            return None

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=int))

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def T(self):
        return _Lazy((self.shape[1], self.shape[0]) if self.ndim == 2 else self.shape, (self,),
                     lambda a: a.T)

    def is_scalar(self):
        return self.size == 1

    # ---- arithmetic
    def __add__(self, other):
        other = _wrap(other)
        return _Lazy(_binary_shape(self, other), (self, other), lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, other):
        other = _wrap(other)
        return _Lazy(_binary_shape(self, other), (self, other), lambda a, b: a - b)

    def __rsub__(self, other):
        return _wrap(other).__sub__(self)

    def __neg__(self):
        return _Lazy(self.shape, (self,), lambda a: -a)

    def __matmul__(self, other):
        other = _wrap(other)
        if self.ndim != 2 or other.ndim != 2 or self.shape[1] != other.shape[0]:
            raise ValueError("Incompatible dimensions %s %s" % (self.shape, other.shape))
        return _Lazy((self.shape[0], other.shape[1]), (self, other), lambda a, b: a @ b)

    def __rmatmul__(self, other):
        return _wrap(other).__matmul__(self)

    def __mul__(self, other):
        other = _wrap(other)
        if self.size == 1 or other.size == 1:            # scaling
            return _Lazy(_binary_shape(self, other), (self, other), lambda a, b: a * b)
        return self.__matmul__(other)                    # cvxpy 1.0: '*' of two matrices is the matrix product

    def __rmul__(self, other):
        return _wrap(other).__mul__(self)

    def __truediv__(self, other):
        other = _wrap(other)
        if other.size != 1:
            raise ValueError("Can only divide by a scalar constant.")
        return _Lazy(self.shape, (self, other), lambda a, b: a / b)

    def __getitem__(self, key):
        probe = np.empty(self.shape, dtype=np.int8)[key]
        return _Lazy(probe.shape, (self,), lambda a: a[key])

    # ---- constraints
    def __le__(self, other):
        return Constraint(self, _wrap(other), "<=")

    def __ge__(self, other):
        return Constraint(_wrap(other), self, "<=")

    def __eq__(self, other):
        return Constraint(self, _wrap(other), "==")

    def __repr__(self):
        return "<mini_cvxpy.%s %s>" % (type(self).__name__, self.shape)


class _Lazy(Expression):
    def __init__(self, shape, args, fn):
        super(_Lazy, self).__init__(shape, args)
        self._fn = fn

    def _eval(self):
        return self._fn(*[a._eval() for a in self.args])


class Constant(Expression):
    def __init__(self, value):
        self._value = np.asarray(value, dtype=np.float64)
        super(Constant, self).__init__(self._value.shape)

    def _eval(self):
        return self._value


class _Leaf(Expression):
    def __init__(self, shape=(), name=None, value=None):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        super(_Leaf, self).__init__(shape)
        self.id = next(_counter)
        self._name = name
        self._value = None
        if value is not None:
            self.value = value

    def name(self):
        return self._name if self._name is not None else "%s%d" % (type(self).__name__.lower(), self.id)

    def _eval(self):
        if self._value is None:
            raise _Unset(self.name())
        return self._value

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, val):
        if val is None:
            self._value = None
            return
        val = np.asarray(val, dtype=np.float64)
        if val.shape != self.shape:
            if val.size == self.size and self.size == 1:
                val = val.reshape(self.shape)
            else:
                raise ValueError("Invalid dimensions %s for %s value." % (val.shape, type(self).__name__))
        self._value = val


class Variable(_Leaf):
    def __init__(self, shape=(), name=None, boolean=None, nonneg=None, **attributes):
        super(Variable, self).__init__(shape, name=name)
        mask = np.zeros(self.shape, dtype=bool)
        if boolean is True:
            mask[...] = True
        elif boolean:
            for idx in boolean:
                mask[tuple(idx)] = True
        self.boolean_mask = mask
        self.attributes = dict(attributes, boolean=boolean, nonneg=bool(nonneg))


class Parameter(_Leaf):
    def __init__(self, shape=(), name=None, value=None, **attributes):
        super(Parameter, self).__init__(shape, name=name, value=value)


# ---- atoms ----------------------------------------------------------------------------------------------------
def _reduce(shape, axis):
    if axis is None:
        return ()
    return tuple(s for i, s in enumerate(shape) if i != axis)


def sum(expr, axis=None, keepdims=False):          # noqa: A001 (cvxpy's name)
    if isinstance(expr, (list, tuple)):
        total = Constant(0.0)
        for item in expr:
            total = total + item
        return total
    expr = _wrap(expr)
    return _Lazy(_reduce(expr.shape, axis), (expr,), lambda a: np.sum(a, axis=axis))


def sum_squares(expr):
    expr = _wrap(expr)
    return _Lazy((), (expr,), lambda a: np.sum(np.square(a)))


def quad_form(x, P):
    x, P = _wrap(x), _wrap(P)
    return _Lazy((), (x, P), lambda a, p: float(a.reshape(-1) @ p @ a.reshape(-1)))


def norm1(expr, axis=None):
    expr = _wrap(expr)
    return _Lazy(_reduce(expr.shape, axis), (expr,), lambda a: np.sum(np.abs(a), axis=axis))


def norm_inf(expr, axis=None):
    expr = _wrap(expr)
    return _Lazy(_reduce(expr.shape, axis), (expr,), lambda a: np.max(np.abs(a), axis=axis))


def multiply(a, b):
    a, b = _wrap(a), _wrap(b)
    return _Lazy(_binary_shape(a, b), (a, b), lambda x, y: x * y)


def reshape(expr, shape):
    expr = _wrap(expr)
    shape = tuple(int(s) for s in shape)
    if int(np.prod(shape, dtype=int)) != expr.size:
        raise ValueError("Invalid reshape dimensions %s." % (shape,))
    return _Lazy(shape, (expr,), lambda a: np.reshape(a, shape, order="F"))      # cvxpy reshapes column-major


def _stack(items, axis):
    items = [_wrap(i) for i in items]
    shapes = [i.shape if i.ndim == 2 else ((1, i.size) if axis == 0 else (i.size, 1)) for i in items]
    other = 1 - axis
    if len({s[other] for s in shapes}) != 1:
        raise ValueError("All the input dimensions except for axis %d must match exactly." % axis)
    out = list(shapes[0])
    out[axis] = int(np.sum([s[axis] for s in shapes]))
    return _Lazy(tuple(out), items,
                 lambda *vals: np.concatenate([np.reshape(v, s) for v, s in zip(vals, shapes)], axis=axis))


def vstack(items):
    return _stack(items, 0)


def hstack(items):
    return _stack(items, 1)


# ---- problems --------------------------------------------------------------------------------------------------
class Constraint(object):
    def __init__(self, lhs, rhs, kind):
        _binary_shape(lhs, rhs)
        self.lhs, self.rhs, self.kind = lhs, rhs, kind
        self.args = (lhs, rhs)

    def residual(self):
        """lhs - rhs, column-major vector: <= 0 or == 0"""
        diff = np.asarray(self.lhs._eval() - self.rhs._eval(), dtype=np.float64)
        return diff.reshape(-1, order="F")


class Minimize(object):
    sign = 1.0

    def __init__(self, expr):
        self.expr = _wrap(expr)
        if self.expr.size != 1:
            raise ValueError("The '%s' objective must resolve to a scalar." % type(self).__name__.lower())
        self.args = (self.expr,)


class Maximize(Minimize):
    sign = -1.0


def _variables_of(nodes):
    seen, out, stack = set(), [], list(nodes)
    while stack:
        node = stack.pop()
        if id(node) in seen:
            continue
        seen.add(id(node))
        if isinstance(node, Variable):
            out.append(node)
        stack.extend(getattr(node, "args", ()))
    return sorted(out, key=lambda v: v.id)


class Problem(object):
    def __init__(self, objective, constraints=None):
        self.objective = objective
        self.constraints = list(constraints or [])
        self.status = None
        self.value = None
        self.solver_stats = types.SimpleNamespace(solve_time=None)

    def variables(self):
        return _variables_of([self.objective] + self.constraints)

    # -- x = all variables, column-major, in creation order
    def _set_x(self, variables, x):
        off = 0
        for var in variables:
            var.value = np.reshape(x[off:off + var.size], var.shape, order="F")
            off += var.size

    def objective_at(self, x, variables=None):
        variables = variables if variables is not None else self.variables()
        self._set_x(variables, np.asarray(x, dtype=np.float64))
        return float(np.asarray(self.objective.expr._eval()).reshape(-1)[0])

    def canonical_form(self, check=True):
        """dict(variables, n, c, c0, objective_is_affine, G, h, A, b, lb, ub, integrality): the problem as
        min sign*(c'x + c0) s.t. G x <= h, A x = b, found by probing (exact when the expressions are affine)."""
        variables = self.variables()
        saved = [v.value for v in variables]
        n = int(np.sum([v.size for v in variables], dtype=int))
        try:
            zero = np.zeros(n)
            c0 = self.objective_at(zero, variables)
            res0 = [con.residual() for con in self.constraints]
            c = np.zeros(n)
            cols = [np.zeros((r.size, n)) for r in res0]
            for j in range(n):
                e = zero.copy()
                e[j] = 1.0
                c[j] = self.objective_at(e, variables) - c0
                for ci, con in enumerate(self.constraints):
                    cols[ci][:, j] = con.residual() - res0[ci]
            affine = True
            if check and n:
                rng = np.random.default_rng(0)
                for _ in range(3):
                    x = rng.uniform(-1.5, 1.5, n)
                    fx = self.objective_at(x, variables)
                    if abs(fx - (c @ x + c0)) > 1e-8 * (1.0 + abs(fx)):
                        affine = False
                    for ci, con in enumerate(self.constraints):
                        if not np.allclose(con.residual(), cols[ci] @ x + res0[ci], rtol=1e-9, atol=1e-9):
                            raise SolverError("constraint %d is not affine" % ci)
        finally:
            for v, val in zip(variables, saved):
                v.value = val
        ineq = [i for i, con in enumerate(self.constraints) if con.kind == "<="]
        eq = [i for i, con in enumerate(self.constraints) if con.kind == "=="]
        G = np.vstack([cols[i] for i in ineq]) if ineq else np.zeros((0, n))
        h = -np.concatenate([res0[i] for i in ineq]) if ineq else np.zeros(0)
        A = np.vstack([cols[i] for i in eq]) if eq else np.zeros((0, n))
        b = -np.concatenate([res0[i] for i in eq]) if eq else np.zeros(0)
        integrality = np.concatenate([v.boolean_mask.reshape(-1, order="F") for v in variables]) if variables \
            else np.zeros(0, dtype=bool)
        lb = np.full(n, -np.inf)
        ub = np.full(n, np.inf)
        nonneg = np.concatenate([np.full(v.size, v.attributes["nonneg"]) for v in variables]) if variables \
            else np.zeros(0, dtype=bool)
        lb[nonneg] = 0.0
        lb[integrality] = 0.0
        ub[integrality] = 1.0
        return dict(variables=variables, n=n, c=c, c0=c0, objective_is_affine=affine, sign=self.objective.sign,
                    G=G, h=h, A=A, b=b, lb=lb, ub=ub, integrality=integrality)

    def solve(self, solver=None, verbose=False, warm_start=True, parallel=False, method=None, **kwargs):
        from scipy.optimize import Bounds, LinearConstraint, milp
        cf = self.canonical_form()
        if not cf["objective_is_affine"]:
            raise SolverError("mini_cvxpy solves problems with an affine objective only")
        t0 = time.perf_counter()
        cons = []
        if cf["G"].shape[0]:
            cons.append(LinearConstraint(cf["G"], -np.inf, cf["h"]))
        if cf["A"].shape[0]:
            cons.append(LinearConstraint(cf["A"], cf["b"], cf["b"]))
        if cf["n"] == 0:
            feasible = bool(np.all(cf["h"] >= -1e-9) and np.all(np.abs(cf["b"]) <= 1e-9))
            self.status = "optimal" if feasible else "infeasible"
            self.value = cf["sign"] * cf["c0"] if feasible else cf["sign"] * np.inf
            self.solver_stats.solve_time = time.perf_counter() - t0
            return self.value
        res = milp(cf["sign"] * cf["c"], constraints=cons, integrality=cf["integrality"].astype(int),
                   bounds=Bounds(cf["lb"], cf["ub"]),
                   options=dict(mip_rel_gap=float(kwargs.get("MIPGap", 0.0)), disp=bool(verbose)))
        self.solver_stats.solve_time = time.perf_counter() - t0
        if res.status == 0:
            x = np.asarray(res.x)
            x[cf["integrality"]] = np.round(x[cf["integrality"]])
            self._set_x(cf["variables"], x)
            self.status = "optimal"
            self.value = float(cf["c"] @ x + cf["c0"])
        elif res.status == 2:
            self.status, self.value = "infeasible", cf["sign"] * np.inf
        elif res.status == 3:
            self.status, self.value = "unbounded", -cf["sign"] * np.inf
        else:
            raise SolverError("HiGHS status %d: %s" % (res.status, res.message))
        return self.value


def install():
    """Register this module as ``cvxpy`` (+ the sub-modules the reference imports).  Must run before any reference
    module is imported."""
    import sys
    me = sys.modules[__name__]
    if sys.modules.get("cvxpy") is me:
        return me
    if "cvxpy" in sys.modules:
        raise RuntimeError("another cvxpy (or the numeric-only stub of oracle/ref_shim) is already imported")
    err = types.ModuleType("cvxpy.error")
    err.SolverError = SolverError
    exprs = types.ModuleType("cvxpy.expressions")
    expr = types.ModuleType("cvxpy.expressions.expression")
    expr.Expression = Expression
    exprs.expression = expr
    me.error, me.expressions = err, exprs
    me.GUROBI, me.CPLEX = "GUROBI", "CPLEX"
    sys.modules["cvxpy"] = me
    sys.modules["cvxpy.error"] = err
    sys.modules["cvxpy.expressions"] = exprs
    sys.modules["cvxpy.expressions.expression"] = expr
    return me
