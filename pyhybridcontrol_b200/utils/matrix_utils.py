"""Symbolic / callable system matrices, compiled for the GPU.

Reference: ``CallableMatrix`` (utils/matrix_utils.py:279-562) wraps a matrix-valued function of the model
parameters -- for a sympy matrix, ``sympy.lambdify(sorted free symbols, matrix, modules="numpy")`` (:339-343,
:372-378) -- and ``MldModel.to_numeric`` calls every matrix with the ``param_struct`` (models/mld_model.py:791-793),
once per agent, on the host.

Here the expressions are compiled ONCE into a straight-line register program (``ExprProgram``) that
``hmpc_param_eval_f64`` (csrc/param_eval.cu) interprets for a whole batch of parameter sets, one thread per agent.
The compiler walks the sympy tree the way sympy's own code printer writes it for ``lambdify`` -- terms and factors
in printer order, ``a*b/(c*d)`` as one division of two products, ``x**-1`` as ``1/x``, ``x**(1/2)`` as ``sqrt`` -- so
that + - * / round exactly as in the reference and only exp / log / pow / trig differ (by the CUDA library's
<= 2 ulp).  Identical sub-expressions are evaluated once (they have the same value anyway).

Nothing here computes matrix values on the host: evaluation always goes through the CUDA library.
"""
import inspect
import os
import struct

import numpy as np

from .structs import atleast_2d_col

# opcodes of include/hmpc.h (hmpc_expr_ins)
OP_CONST, OP_PARAM, OP_OUT = 0, 1, 2
(OP_MOV, OP_NEG, OP_ABS, OP_SIGN, OP_SQRT, OP_EXP, OP_LOG, OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN,
 OP_SINH, OP_COSH, OP_TANH, OP_FLOOR, OP_CEIL, OP_POWI) = range(10, 29)
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_MIN, OP_MAX, OP_ATAN2 = range(40, 48)

OP_NAMES = {v: k[3:] for k, v in list(globals().items()) if k.startswith("OP_") and isinstance(v, int)}

_UNARY_FUNCS = {"exp": OP_EXP, "log": OP_LOG, "sin": OP_SIN, "cos": OP_COS, "tan": OP_TAN, "asin": OP_ASIN,
                "acos": OP_ACOS, "atan": OP_ATAN, "sinh": OP_SINH, "cosh": OP_COSH, "tanh": OP_TANH, "Abs": OP_ABS,
                "sign": OP_SIGN, "floor": OP_FLOOR, "ceiling": OP_CEIL}
MAX_POWI = 64            # |integer exponent| up to which x**n is a product chain; beyond it, pow(x, n)
MAX_OUT_MATS = 20


def _sympy():
    import sympy
    return sympy


def is_symbolic(obj):
    """True for sympy expressions and matrices (reference test: isinstance(obj, (sp.Expr, sp.Matrix)),
    models/mld_model.py:836-837)."""
    mod = type(obj).__module__
    if not mod.startswith("sympy"):
        return False
    sp = _sympy()
    return isinstance(obj, (sp.Expr, sp.MatrixBase))


def _const_bits(value):
    lo, hi = struct.unpack("<ii", struct.pack("<d", float(value)))
    return lo, hi


class _Emitter(object):
    """Expression trees -> instructions over virtual registers, with structural sharing of sub-expressions."""

    def __init__(self, param_index, preload_params=False):
        self.param_index = param_index
        self.preload = bool(preload_params)   # parameters ARE virtual registers 0..P-1 (no PARAM instructions)
        self.ins = []              # [op, dst, a, b] over virtual registers
        self.memo = {}
        self.n_vregs = len(param_index) if self.preload else 0

    def _new(self, op, a=0, b=0):
        dst = self.n_vregs
        self.n_vregs += 1
        self.ins.append([op, dst, a, b])
        return dst

    def const(self, value):
        key = ("const", float(value).hex())
        if key not in self.memo:
            lo, hi = _const_bits(value)
            self.memo[key] = self._new(OP_CONST, lo, hi)
        return self.memo[key]

    def _product(self, regs):
        acc = regs[0]
        for r in regs[1:]:
            acc = self._new(OP_MUL, acc, r)
        return acc

    def emit(self, e):
        sp = _sympy()
        e = sp.sympify(e)
        if e in self.memo:
            return self.memo[e]
        r = self._emit(e, sp)
        self.memo[e] = r
        return r

    def _emit(self, e, sp):
        if e.is_Symbol:
            name = str(e)
            if name not in self.param_index:
                raise KeyError("symbol %r is not a parameter of this program" % name)
            if self.preload:
                return self.param_index[name]
            return self._new(OP_PARAM, self.param_index[name])
        if e.is_number:
            if e is sp.nan:
                return self.const(float("nan"))
            if e is sp.oo or e is -sp.oo:
                return self.const(float("inf") if e is sp.oo else float("-inf"))
            if e.is_real is False or e is sp.zoo:
                raise NotImplementedError("non-real constant %s in a system matrix" % (e,))
            return self.const(float(e))
        if isinstance(e, (sp.conjugate, sp.re)):
            # parameters are real: conj(x) = re(x) = x (sympy's pinv() leaves conjugates in symbolic models,
            # e.g. the reference's const_heat=False DEWH model, micro_grid_models.py:52-57)
            return self.emit(e.args[0])
        if isinstance(e, sp.im):
            return self.const(0.0)
        if e.is_Add:
            terms = e.as_ordered_terms()
            acc = self.emit(terms[0])
            for term in terms[1:]:
                # the printer writes "+ t" or "- |t|"; both round the same, so the negative coefficient is split off
                # and subtracted, which saves the negation
                c, rest = term.as_coeff_Mul()
                if c.is_negative:
                    acc = self._new(OP_SUB, acc, self.emit(-term))
                else:
                    acc = self._new(OP_ADD, acc, self.emit(term))
            return acc
        if e.is_Mul:
            return self._emit_mul(e, sp)
        if e.is_Pow:
            return self._emit_pow(e, sp)
        if isinstance(e, (sp.Min, sp.Max)):
            op = OP_MIN if isinstance(e, sp.Min) else OP_MAX
            regs = [self.emit(a) for a in e.args]
            acc = regs[0]
            for r in regs[1:]:
                acc = self._new(op, acc, r)
            return acc
        if isinstance(e, sp.atan2):
            return self._new(OP_ATAN2, self.emit(e.args[0]), self.emit(e.args[1]))
        fname = type(e).__name__
        if e.is_Function and fname in _UNARY_FUNCS and len(e.args) == 1:
            return self._new(_UNARY_FUNCS[fname], self.emit(e.args[0]))
        raise NotImplementedError("no GPU instruction for %s (in %s); supported: + - * / ** exp log sqrt sin cos tan "
                                  "asin acos atan atan2 sinh cosh tanh Abs sign floor ceiling Min Max"
                                  % (type(e).__name__, e))

    def _emit_mul(self, e, sp):
        # numerator / denominator split of sympy's StrPrinter._print_Mul (what lambdify evaluates)
        c, rest = e.as_coeff_Mul()
        negate = bool(c.is_negative)
        if negate:
            e = -e
        num, den = [], []
        for item in e.as_ordered_factors():
            if item.is_commutative and item.is_Pow and item.exp.is_Rational and item.exp.is_negative:
                den.append(item.base if item.exp == -1 else sp.Pow(item.base, -item.exp, evaluate=False))
            elif item.is_Rational and item is not sp.S.Infinity and not item.is_Integer:
                if item.p != 1:
                    num.append(sp.Integer(item.p))
                den.append(sp.Integer(item.q))
            else:
                num.append(item)
        acc = self._product([self.emit(f) for f in num]) if num else self.const(1.0)
        if den:
            acc = self._new(OP_DIV, acc, self._product([self.emit(f) for f in den]))
        if negate:
            acc = self._new(OP_NEG, acc)
        return acc

    def _emit_pow(self, e, sp):
        base, ex = e.base, e.exp
        if base is sp.E:
            return self._new(OP_EXP, self.emit(ex))
        if ex.is_Integer:
            n = int(ex)
            if n == -1:
                return self._new(OP_DIV, self.const(1.0), self.emit(base))
            if abs(n) <= MAX_POWI:
                return self._new(OP_POWI, self.emit(base), n)
            return self._new(OP_POW, self.emit(base), self.const(float(n)))
        if ex == sp.S.Half:
            return self._new(OP_SQRT, self.emit(base))
        if ex == -sp.S.Half:
            return self._new(OP_DIV, self.const(1.0), self._new(OP_SQRT, self.emit(base)))
        return self._new(OP_POW, self.emit(base), self.emit(ex))


def _allocate_registers(ins, n_pinned=0):
    """``n_pinned`` leading virtual registers (preloaded parameters) keep their numbers and are never reused.
    Dead values dropped, then a linear scan virtual -> physical registers: a register is free after the last read
    of its value, and the destination of an instruction may reuse a source that dies there (the kernel reads both
    operands before it writes)."""
    live, kept = set(), []
    for op, dst, a, b in reversed(ins):
        if op != OP_OUT and dst not in live:
            continue
        live.update(_reads(op, dst, a, b))
        kept.append((op, dst, a, b))
    kept.reverse()
    last_use = {}
    for i, (op, dst, a, b) in enumerate(kept):
        for r in _reads(op, dst, a, b):
            last_use[r] = i
    free, phys, n_phys, out = [], {r: r for r in range(n_pinned)}, n_pinned, []
    for i, (op, dst, a, b) in enumerate(kept):
        reads = set(_reads(op, dst, a, b))
        pa = phys[a] if a in reads else a
        pb = phys[b] if (op >= OP_ADD and b in reads) else b
        for r in reads:
            if last_use[r] == i and r >= n_pinned:
                free.append(phys.pop(r))
        if op == OP_OUT:
            out.append((op, dst, pa, 0))
            continue
        if free:
            p = free.pop()
        else:
            p = n_phys
            n_phys += 1
        phys[dst] = p
        out.append((op, p, pa, pb))
    return out, max(n_phys, 1)


def _reads(op, dst, a, b):
    if op in (OP_CONST, OP_PARAM):
        return ()
    if op == OP_OUT or op < OP_ADD:
        return (a,)
    return (a, b)


class ExprProgram(object):
    """The compiled register program of a set of symbolic matrices.

    ``matrices``: ordered mapping name -> sympy Matrix (at most 20).  ``param_names``: order of the columns of the
    parameter table; default = sorted names of all free symbols (the reference sorts the lambdify arguments the same
    way, utils/matrix_utils.py:372-378)."""

    def __init__(self, matrices, param_names=None):
        sp = _sympy()
        mats = [(name, sp.Matrix(m) if not isinstance(m, sp.MatrixBase) else m) for name, m in matrices.items()]
        if not mats or len(mats) > MAX_OUT_MATS:
            raise ValueError("a program evaluates 1..%d matrices, got %d" % (MAX_OUT_MATS, len(mats)))
        free = set()
        for _, m in mats:
            free |= {str(s) for s in m.free_symbols}
        if param_names is None:
            param_names = sorted(free)
        else:
            param_names = list(param_names)
            missing = free.difference(param_names)
            if missing:
                raise ValueError("param_names lacks the symbols %s" % sorted(missing))
        self.param_names = tuple(param_names)
        self.required_params = tuple(sorted(free))
        self.mat_names = tuple(n for n, _ in mats)
        self.mat_shapes = tuple(tuple(int(s) for s in m.shape) for _, m in mats)
        self.mat_sizes = tuple(r * c for r, c in self.mat_shapes)
        self.n_out = int(sum(self.mat_sizes))
        self._mats = mats
        ins, self.n_regs = self._compile(preload_params=False)
        if not ins:
            raise ValueError("all matrices of the program are empty")
        self.instructions = np.array(ins, dtype=np.int32).reshape(-1, 4)
        self.n_ins = int(self.instructions.shape[0])
        self._v2 = None
        self._dev = {}

    def _compile(self, preload_params):
        em = _Emitter({n: i for i, n in enumerate(self.param_names)}, preload_params=preload_params)
        slot = 0
        for _, m in self._mats:
            rows, cols = m.shape
            for i in range(rows):
                for j in range(cols):
                    reg = em.emit(m[i, j])
                    em.ins.append([OP_OUT, slot, reg, 0])
                    slot += 1
        return _allocate_registers(em.ins, n_pinned=len(self.param_names) if preload_params else 0)

    @property
    def instructions_v2(self):
        """(instructions, n_regs) for the kernel hmpc_param_eval_v2_f64: registers 0..P-1 are the
        parameters (preloaded, never written), no PARAM instructions; same operations in the same order otherwise."""
        if self._v2 is None:
            ins, n_regs = self._compile(preload_params=True)
            self._v2 = (np.array(ins, dtype=np.int32).reshape(-1, 4), max(n_regs, len(self.param_names) + 1))
        return self._v2

    # ---- device side -----------------------------------------------------------------------------------------
    def _program_on(self, device, version=1):
        import torch
        key = (str(device), int(version))
        if key not in self._dev:
            ins = self.instructions if version == 1 else self.instructions_v2[0]
            self._dev[key] = torch.from_numpy(ins.copy()).to(device)
        return self._dev[key]

    def param_table(self, param_struct, overrides=None, B=None, device="cuda"):
        """[B, P] CUDA tensor: every column is the scalar of ``param_struct`` broadcast, or the per-agent vector in
        ``overrides`` (name -> array of length B)."""
        import torch
        overrides = overrides or {}
        lens = {int(np.asarray(v).reshape(-1).shape[0]) for v in overrides.values() if np.ndim(v) > 0}
        if B is None:
            B = lens.pop() if len(lens) == 1 else (1 if not lens else None)
        if B is None or any(n != B for n in lens):
            raise ValueError("per-agent parameter vectors must all have length B")
        tab = np.empty((B, len(self.param_names)), dtype=np.float64)
        for j, name in enumerate(self.param_names):
            if name in overrides:
                tab[:, j] = np.asarray(overrides[name], dtype=np.float64).reshape(-1)
            else:
                try:
                    tab[:, j] = float(param_struct[name])
                except (KeyError, TypeError):
                    if name in self.required_params:
                        raise TypeError("missing required parameter %r" % name)
                    tab[:, j] = 0.0
        return torch.from_numpy(tab).to(device)

    def evaluate(self, params, version=None):
        """params: CUDA float64 tensor [B, P] -> dict name -> CUDA tensor [B, rows, cols] (views of one buffer).
        version: 2 = hmpc_param_eval_v2_f64 (default); 1 = hmpc_param_eval_f64 (also HMPC_PARAM_EVAL=v1)."""
        from .. import cabi
        if params.dim() != 2 or params.shape[1] != len(self.param_names):
            raise ValueError("params must be [B, %d]" % len(self.param_names))
        if version is None:
            # v2 (parameters preloaded as registers, two agents per thread) is the default since round 2: bit-identical
            # results on the golden fixtures and the fuzz programs, 11-15 % faster at 2 M agents on a B200
            # (profiles/r2_notes.md); HMPC_PARAM_EVAL=v1 selects the first kernel
            version = 1 if os.environ.get("HMPC_PARAM_EVAL", "v2").lower() in ("v1", "1") else 2
        if int(version) == 1:
            flat = cabi.param_eval(self._program_on(params.device), self.n_regs, self.mat_sizes, params)
        else:
            flat = cabi.param_eval(self._program_on(params.device, 2), self.instructions_v2[1], self.mat_sizes, params,
                                   version=2)
        B = params.shape[0]
        out, off = {}, 0
        for name, (r, c), sz in zip(self.mat_names, self.mat_shapes, self.mat_sizes):
            out[name] = flat[B * off:B * (off + sz)].view(B, r, c)
            off += sz
        return out

    def bytes_per_agent(self):
        return 8 * (len(self.param_names) + self.n_out)

    def disassemble(self):
        lines = []
        for op, dst, a, b in self.instructions.tolist():
            if op == OP_CONST:
                val = struct.unpack("<d", struct.pack("<ii", a, b))[0]
                lines.append("r%d = %r" % (dst, val))
            elif op == OP_PARAM:
                lines.append("r%d = %s" % (dst, self.param_names[a]))
            elif op == OP_OUT:
                lines.append("out[%d] = r%d" % (dst, a))
            elif op == OP_POWI:
                lines.append("r%d = r%d ** %d" % (dst, a, b))
            elif op >= OP_ADD:
                lines.append("r%d = %s(r%d, r%d)" % (dst, OP_NAMES[op], a, b))
            else:
                lines.append("r%d = %s(r%d)" % (dst, OP_NAMES[op], a))
        return "\n".join(lines)


class _Tracer(object):
    """A sympy expression that also answers numpy's object-dtype protocol: ``np.exp(x)`` on an object looks for
    ``x.exp()``, ``np.sqrt(x)`` for ``x.sqrt()`` and so on, so matrix functions written with numpy ufuncs -- the way
    the reference's users write them, since it calls them with floats -- can be traced as well.  Anything that needs
    the VALUE (``if x > 0``, ``float(x)``, ``int(x)``) raises: such a function cannot run on the GPU."""
    __array_priority__ = 1000.0
    _UFUNCS = dict(exp="exp", log="log", sqrt="sqrt", sin="sin", cos="cos", tan="tan", arcsin="asin", arccos="acos",
                   arctan="atan", sinh="sinh", cosh="cosh", tanh="tanh", fabs="Abs", absolute="Abs", sign="sign",
                   floor="floor", ceil="ceiling", conjugate="conjugate", conj="conjugate")

    _BINARY = dict(add=lambda sp, a, b: a + b, subtract=lambda sp, a, b: a - b, multiply=lambda sp, a, b: a * b,
                   true_divide=lambda sp, a, b: a / b, divide=lambda sp, a, b: a / b, power=lambda sp, a, b: a ** b,
                   float_power=lambda sp, a, b: a ** b, maximum=lambda sp, a, b: sp.Max(a, b),
                   minimum=lambda sp, a, b: sp.Min(a, b), fmax=lambda sp, a, b: sp.Max(a, b),
                   fmin=lambda sp, a, b: sp.Min(a, b), arctan2=lambda sp, a, b: sp.atan2(a, b))

    def __init__(self, expr):
        self.e = _sympy().sympify(expr)

    @staticmethod
    def _un(x):
        return x.e if isinstance(x, _Tracer) else x

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """``np.exp(x)``, ``np.maximum(x, 0.0)``, ... called directly on a tracer"""
        if method != "__call__" or kwargs:
            return NotImplemented
        if any(isinstance(i, np.ndarray) for i in inputs):
            # tracer (op) array: let numpy's object loop apply the operator element by element
            arrs = [np.asarray(i, dtype=object) if isinstance(i, np.ndarray) else i for i in inputs]
            return getattr(ufunc, method)(*[np.array(a, dtype=object) if isinstance(a, _Tracer) else a for a in arrs])
        sp = _sympy()
        name = ufunc.__name__
        args = [self._un(i) for i in inputs]
        if len(args) == 1:
            if name in _Tracer._UFUNCS:
                return _Tracer(getattr(sp, _Tracer._UFUNCS[name])(args[0]))
            if name == "negative":
                return _Tracer(-args[0])
            if name == "positive":
                return _Tracer(args[0])
            if name == "square":
                return _Tracer(args[0] ** 2)
            if name == "reciprocal":
                return _Tracer(1 / args[0])
        elif len(args) == 2 and name in _Tracer._BINARY:
            return _Tracer(_Tracer._BINARY[name](sp, args[0], args[1]))
        raise TypeError("numpy.%s has no GPU instruction" % name)

    def __getattr__(self, name):
        fname = _Tracer._UFUNCS.get(name)
        if fname is None:
            raise AttributeError(name)
        sp = _sympy()
        return lambda: _Tracer(getattr(sp, fname)(self.e))

    def square(self):
        return _Tracer(self.e ** 2)

    def reciprocal(self):
        return _Tracer(1 / self.e)

    def __add__(self, o): return _Tracer(self.e + self._un(o))                     # noqa: E704
    def __radd__(self, o): return _Tracer(self._un(o) + self.e)                    # noqa: E704
    def __sub__(self, o): return _Tracer(self.e - self._un(o))                     # noqa: E704
    def __rsub__(self, o): return _Tracer(self._un(o) - self.e)                    # noqa: E704
    def __mul__(self, o): return _Tracer(self.e * self._un(o))                     # noqa: E704
    def __rmul__(self, o): return _Tracer(self._un(o) * self.e)                    # noqa: E704
    def __truediv__(self, o): return _Tracer(self.e / self._un(o))                 # noqa: E704
    def __rtruediv__(self, o): return _Tracer(self._un(o) / self.e)                # noqa: E704
    def __pow__(self, o): return _Tracer(self.e ** self._un(o))                    # noqa: E704
    def __rpow__(self, o): return _Tracer(self._un(o) ** self.e)                   # noqa: E704
    def __neg__(self): return _Tracer(-self.e)                                     # noqa: E704
    def __pos__(self): return self                                                 # noqa: E704
    def __abs__(self): return _Tracer(_sympy().Abs(self.e))                        # noqa: E704

    def _needs_value(self, *a, **k):
        raise TypeError("the function needs the VALUE of a parameter (a comparison, float() or int())")

    __bool__ = __float__ = __int__ = __index__ = _needs_value
    __lt__ = __le__ = __gt__ = __ge__ = _needs_value

    def _sympy_(self):
        return self.e


def _unwrap_matrix(ret):
    """whatever the function returned -> sympy Matrix (2-D; scalars (1,1), vectors columns)"""
    sp = _sympy()
    if isinstance(ret, _Tracer):
        ret = ret.e
    if is_symbolic(ret):
        return sp.Matrix(ret) if isinstance(ret, sp.MatrixBase) else sp.Matrix([[ret]])
    arr = np.asarray(atleast_2d_col(ret), dtype=object)
    if arr.ndim != 2:
        raise TypeError("expected a 2-D matrix, got %d dimensions" % arr.ndim)
    return sp.Matrix(arr.shape[0], arr.shape[1],
                     lambda i, j: arr[i, j].e if isinstance(arr[i, j], _Tracer) else sp.sympify(arr[i, j]))


def _trace_function(func):
    """Python matrix function -> sympy Matrix, by calling it with one symbolic tracer per argument.  The reference
    calls such functions with floats on the host (utils/matrix_utils.py:334-337, 441-470); on the GPU path the function
    has to be expressible as arithmetic and elementary functions of its arguments -- written with Python operators,
    numpy ufuncs or sympy functions -- and must not branch on their values."""
    sp = _sympy()
    spec = inspect.getfullargspec(func)
    if spec.varargs or spec.varkw:
        raise TypeError("matrix function %s(): *args / **kwargs cannot be traced" % func.__name__)
    names = [n for n in list(spec.args) + list(spec.kwonlyargs) if n != "param_struct"]
    if inspect.ismethod(func):
        names = names[1:]
    last = None
    for make in (lambda n: sp.Symbol(n), lambda n: _Tracer(sp.Symbol(n))):
        try:
            ret = func(**{n: make(n) for n in names})
            mat = _unwrap_matrix(ret)
            break
        except Exception as exc:  # numpy ufuncs on plain symbols, branches on values, ragged returns, ...
            last = exc
    else:
        raise TypeError("matrix function %s() cannot be traced symbolically (%s: %s); write it with arithmetic, numpy "
                        "ufuncs or sympy functions of its arguments, or pass the sympy matrix itself"
                        % (func.__name__, type(last).__name__, last))
    return mat, tuple(names)


class CallableMatrix(object):
    """A system matrix as a function of the model parameters (reference: utils/matrix_utils.py:279-562).

    ``CallableMatrix(matrix, matrix_name)`` accepts a sympy expression / matrix, a Python function of the parameters
    (traced with symbols), a numeric array, or another CallableMatrix.  Calling it --
    ``cm(param_struct=...)``, ``cm(**params)`` or positionally in ``required_params`` order -- evaluates on the GPU
    and returns a read-only 2-D numpy array; constant matrices return their stored array
    (``CallableMatrixConstant.__call__``, :558-562)."""

    def __new__(cls, matrix=None, matrix_name=None):
        if cls is CallableMatrix:
            probe = object.__new__(CallableMatrix)
            probe._setup(matrix, matrix_name)
            if probe.is_constant:
                probe.__class__ = CallableMatrixConstant
            return probe
        return object.__new__(cls)

    def __init__(self, matrix=None, matrix_name=None):
        if not hasattr(self, "_expr"):
            self._setup(matrix, matrix_name)
            if type(self) is CallableMatrixConstant and not self.is_constant:
                raise TypeError("Cannot initialize CallableMatrixConstant object with non-constant matrix.")

    def _setup(self, matrix, matrix_name):
        sp = _sympy()
        if isinstance(matrix, CallableMatrix):
            self._expr, self._arg_names = matrix._expr, matrix._arg_names
            self._wrapped_name = matrix._wrapped_name
            self._matrix_name = matrix_name if matrix_name is not None else matrix._matrix_name
        elif inspect.isfunction(matrix) or inspect.ismethod(matrix):
            self._expr, self._arg_names = _trace_function(matrix)
            self._wrapped_name = matrix.__name__
            self._matrix_name = matrix_name if matrix_name is not None else matrix.__name__
        elif is_symbolic(matrix):
            self._expr = sp.Matrix(matrix) if not isinstance(matrix, sp.MatrixBase) else sp.Matrix(matrix)
            self._arg_names = tuple(sorted(str(s) for s in self._expr.free_symbols))
            self._wrapped_name = "_lambdifygenerated"
            self._matrix_name = matrix_name if matrix_name is not None else self._wrapped_name
        elif callable(matrix):
            raise TypeError("matrix must be a function, a sympy expression or numeric, not %s" % type(matrix).__name__)
        else:
            arr = np.array(atleast_2d_col(matrix))
            if not np.issubdtype(arr.dtype, np.number) and arr.dtype != bool:
                raise TypeError("System matrices must be numeric, callable, or symbolic.")
            self._expr = sp.Matrix(arr.tolist()) if arr.size else sp.zeros(*arr.shape)
            self._arg_names = ()
            self._wrapped_name = "constant_matrix_func"
            self._matrix_name = matrix_name if matrix_name is not None else self._wrapped_name
            self._constant = np.array(arr)
            self._constant.setflags(write=False)
        self._required = tuple(n for n in self._arg_names)
        free = {str(s) for s in self._expr.free_symbols}
        self._is_constant = not free
        if self._is_constant and not hasattr(self, "_constant"):
            vals = np.array(self._expr.tolist(), dtype=np.float64).reshape(self._expr.shape) if self._expr.shape[0] * \
                self._expr.shape[1] else np.zeros(self._expr.shape)
            vals.setflags(write=False)
            self._constant = vals
        self._program = None

    # ---- reference attribute surface (utils/matrix_utils.py:487-545) ---------------------------------------
    @property
    def __name__(self):
        return self._matrix_name

    @property
    def matrix_name(self):
        return self._matrix_name

    @property
    def required_params(self):
        return list(self._required)

    @property
    def expr(self):
        return self._expr

    @property
    def shape(self):
        return tuple(int(s) for s in self._expr.shape)

    @property
    def size(self):
        return self.shape[0] * self.shape[1]

    @property
    def ndim(self):
        return 2

    @property
    def dtype(self):
        return self._constant.dtype if self._is_constant else np.dtype(np.float64)

    @property
    def itemsize(self):
        return self.dtype.itemsize

    @property
    def nbytes(self):
        return self.size * self.itemsize

    @property
    def is_empty(self):
        return self.size == 0

    @property
    def is_all_zero(self):
        return bool(self._is_constant and np.all(self._constant == 0)) if self.size else True

    @property
    def is_constant(self):
        return self._is_constant

    @property
    def program(self):
        if self._program is None:
            self._program = ExprProgram({self._matrix_name: self._expr}, param_names=self._arg_names)
        return self._program

    def _bind(self, args, kwargs):
        param_struct = kwargs.pop("param_struct", None)
        if len(args) > len(self._arg_names):
            raise TypeError("%s() takes %d positional arguments but %d were given"
                            % (self._matrix_name, len(self._arg_names), len(args)))
        bound = dict(zip(self._arg_names, args))
        for k, v in kwargs.items():
            if k not in self._arg_names:
                raise TypeError("%s() got an unexpected keyword argument '%s'" % (self._matrix_name, k))
            if k in bound:
                raise TypeError("%s() got multiple values for argument '%s'" % (self._matrix_name, k))
            bound[k] = v
        if param_struct:
            try:
                common = set(self._arg_names).intersection(param_struct)
            except TypeError as te:
                raise TypeError("'param_struct' must be dictionary like or None: %s" % te.args[0])
            dup = common.intersection(bound)
            if dup:
                raise TypeError("%s() got multiple values for argument '%s' - values in kwargs are duplicated in "
                                "param_struct." % (self._matrix_name, sorted(dup)[0]))
            bound.update({k: param_struct[k] for k in common})
        missing = [n for n in self._arg_names if n not in bound]
        if missing:
            raise TypeError("%s() missing %d required argument(s): %s"
                            % (self._matrix_name, len(missing), ", ".join(repr(m) for m in missing)))
        return bound

    def __call__(self, *args, **kwargs):
        bound = self._bind(args, kwargs)
        if self._is_constant:
            return self._constant
        prog = self.program
        out = prog.evaluate(prog.param_table(bound, B=1))[self._matrix_name]
        ret = out[0].cpu().numpy()
        ret.setflags(write=False)
        return ret

    def copy(self):
        return type(self)(self, self._matrix_name)

    __copy__ = copy

    def __deepcopy__(self, memo=None):
        return self.copy()

    def __reduce__(self):
        return (CallableMatrix, (self._expr if not hasattr(self, "_constant") else self._constant,
                                 self._matrix_name))

    def __repr__(self):
        sig = ", ".join(self._arg_names)
        sig = (sig + ", " if sig else "") + "*, param_struct=None"
        empty = ", shape=%s" % (self.shape,) if not self.size else ""
        return "<%s %s(%s)%s>" % (type(self).__name__, self._matrix_name, sig, empty)


class CallableMatrixConstant(CallableMatrix):
    """A CallableMatrix without parameters (reference: utils/matrix_utils.py:551-562)."""

    def __call__(self, *args, **kwargs):
        kwargs.pop("param_struct", None)
        if args or kwargs:
            self._bind(args, kwargs)
        return self._constant
