"""MLD data model + one-step simulation; numeric, symbolic (sympy) and callable models.

Mirrors the reference's ``MldInfo`` / ``MldModel`` / ``MldSystemModel`` (models/mld_model.py:109, 391, 1001):

    x(k+1) = A x + B1 u + B2 delta + B3 z + B4 omega + b5
    y(k)   = C x + D1 u + D2 delta + D3 z + D4 omega + d5
    E x + F1 u + F2 delta + F3 z + F4 omega + G y + Psi mu <= f5 ,   mu >= 0

The arithmetic (``lsim_k``, auxiliary-variable computation) runs on the GPU through the C ABI
(hmpc_lsim_step_f64, hmpc_milp_solve_f64); this module is host-side bookkeeping: dimension rules, defaults,
variable types, change tracking.  Symbolic / callable models (reference: utils/matrix_utils.py:279-562,
MldModel.to_callable / to_numeric models/mld_model.py:768-833, MldSystemModel :1001-1164) keep their matrices as
sympy expressions / ``CallableMatrix`` objects; parameter -> matrix evaluation is ONE launch of
hmpc_param_eval_f64 over the compiled program of the whole model, for one parameter set (``to_numeric``) or a
batch of agents (``to_numeric_batch``, ``MldSystemModel.get_mld_numeric_batch``).
"""
import itertools

import numpy as np

from ..utils.structs import StructDict, ParNotSet, atleast_2d_col
from ..utils.matrix_utils import CallableMatrix, ExprProgram, is_symbolic

_version_counter = itertools.count(1)


def _next_version():
    return next(_version_counter)


class MldInfo(StructDict):
    """Dimensions and variable types of an MLD (reference: models/mld_model.py:109-387)."""
    _var_names = ["x", "u", "delta", "z", "omega", "y", "mu", "v"]
    _controllable_var_names = ["u", "delta", "z", "mu"]
    _slack_var_names = ["mu"]
    _sys_dim_names = ["n_states", "n_outputs", "n_constraints"]

    def get_var_dim(self, var_name):
        return self["n" + var_name]

    def get_var_type(self, var_name):
        return self["var_type_" + var_name]

    def get_var_bin_dim(self, var_name):
        return self["n" + var_name + "_l"]

    @property
    def var_dims_struct(self):
        return StructDict((v, self.get_var_dim(v)) for v in self._var_names)


class MldModel(StructDict):
    """Container of the 20 named system matrices (2-D float64 numpy arrays, read-only) and their MldInfo."""
    _state_input_mat_names = ["A", "B1", "B2", "B3", "B4", "b5"]
    _output_mat_names = ["C", "D1", "D2", "D3", "D4", "d5"]
    _constraint_mat_names = ["E", "F1", "F2", "F3", "F4", "f5", "G", "Psi"]
    _sys_mat_names = _state_input_mat_names + _output_mat_names + _constraint_mat_names
    _bin_dim_names = ("nu_l", "ndelta_l", "nz_l", "nmu_l")
    MldModelTypes = StructDict(numeric="numeric", callable="callable", symbolic="symbolic")

    def __init__(self, system_matrices=None, ts=ParNotSet, param_struct=None, bin_dims_struct=None,
                 var_types_struct=None, **kwargs):
        super(MldModel, self).__init__()
        object.__setattr__(self, "_mld_info", MldInfo())
        object.__setattr__(self, "_version", _next_version())
        object.__setattr__(self, "_given", {})
        object.__setattr__(self, "_bin_dims", {})
        object.__setattr__(self, "_meta", dict(ts=None, param_struct=None))
        object.__setattr__(self, "_mld_type", self.MldModelTypes.numeric)
        object.__setattr__(self, "_program", None)
        self.update(system_matrices=system_matrices, ts=ts, param_struct=param_struct,
                    bin_dims_struct=bin_dims_struct, var_types_struct=var_types_struct, _from_init=True, **kwargs)

    # ---- reference API -------------------------------------------------------------------------------
    @property
    def mld_info(self):
        return self._mld_info

    @property
    def mld_type(self):
        return self._mld_type

    @property
    def version(self):
        return self._version

    def update(self, system_matrices=None, ts=ParNotSet, param_struct=None, bin_dims_struct=None,
               var_types_struct=None, _from_init=False, **kwargs):
        bin_dims = dict(bin_dims_struct or {})
        for key in list(kwargs):
            if key in self._bin_dim_names:
                bin_dims[key] = kwargs.pop(key)
        if system_matrices and kwargs:
            raise ValueError("Individual matrix arguments cannot be set if 'system_matrices' argument is set")
        creation = system_matrices if system_matrices else kwargs
        if not isinstance(creation, dict):
            try:
                creation = dict(creation)
            except TypeError as te:
                raise TypeError("Argument:'system_matrices' must be dictionary like: %s" % te.args[0])
        for name, mat in creation.items():
            if name not in self._sys_mat_names:
                raise ValueError("Invalid matrix name in %s: %s" % ("kwargs" if name in kwargs else "system_matrices",
                                                                      name))
            if mat is None:
                continue
            if is_symbolic(mat):
                import sympy as sp
                self._given[name] = sp.Matrix(mat) if not isinstance(mat, sp.MatrixBase) else mat
                continue
            if isinstance(mat, CallableMatrix):
                self._given[name] = mat
                continue
            if callable(mat):
                self._given[name] = CallableMatrix(mat, name)
                continue
            try:
                mat = np.array(atleast_2d_col(mat), dtype=np.float64)
            except (TypeError, ValueError):
                raise TypeError("System matrices must be numeric, callable, or symbolic.")
            self._given[name] = mat
        if bin_dims.get("ndelta_l") or bin_dims.get("nz_l"):
            raise ValueError("Cannot manually set ndelta_l or nz_l - these are fixed by the MLD specification")
        self._bin_dims.update({k: int(v) for k, v in bin_dims.items() if v is not None})
        if var_types_struct:
            for key, vt in var_types_struct.items():
                if vt is None:
                    continue
                name = key.replace("var_type_", "")
                if name in ("delta", "z", "v"):
                    raise ValueError("Cannot manually set var types of delta, z or v")
                vt = [str(t) for t in np.asarray(vt).ravel()]
                if any(t not in ("c", "b") for t in vt):
                    raise ValueError("All elements of var_type vectors must be in {'c', 'b'}")
                nb = sum(t == "b" for t in vt)
                if vt != ["c"] * (len(vt) - nb) + ["b"] * nb:
                    raise NotImplementedError("binary entries must be the trailing entries of a variable")
                self._bin_dims["n%s_l" % name] = nb
        if ts is not ParNotSet:
            self._meta["ts"] = ts
        if param_struct is not None:
            if not isinstance(param_struct, dict):
                raise TypeError("'param_struct' must be dictionary like or None.")
            self._meta["param_struct"] = param_struct
        self._rebuild()
        object.__setattr__(self, "_version", _next_version())

    def _rebuild(self):
        g = self._given
        T = self.MldModelTypes
        # model type (reference: _set_mld_type, models/mld_model.py:835-849): symbolic if any matrix is a sympy
        # object, else callable if any matrix is a function -- then EVERY matrix is wrapped as a CallableMatrix
        if any(is_symbolic(m) for m in g.values()):
            mld_type = T.symbolic
        elif any(isinstance(m, CallableMatrix) for m in g.values()):
            mld_type = T.callable
        else:
            mld_type = T.numeric
        object.__setattr__(self, "_mld_type", mld_type)
        object.__setattr__(self, "_program", None)

        def shape_of(k):
            if k not in g:
                return (0, 0)
            sh = tuple(int(v) for v in g[k].shape)
            return sh if 0 not in sh else (0, 0)
        shp = {k: shape_of(k) for k in self._sys_mat_names}
        if "C" not in g:  # C defaults to eye(*A.shape) (reference :515-520)
            n = shp["A"][0]
            g["C"] = np.eye(n)
            shp["C"] = (n, n) if n else (0, 0)
        A_shape = shp["A"]
        if A_shape[0] != A_shape[1]:
            raise ValueError("Invalid shape for state matrix A:'%s', must be a square matrix or scalar" % (A_shape,))

        def rows(names):
            return max(shp[n][0] for n in names)

        def cols(names):
            return max(shp[n][1] for n in names)
        d = dict(nx=rows(self._state_input_mat_names), ny=rows(self._output_mat_names),
                 nc=rows(self._constraint_mat_names), nu=cols(("B1", "D1", "F1")), ndelta=cols(("B2", "D2", "F2")),
                 nz=cols(("B3", "D3", "F3")), nomega=cols(("B4", "D4", "F4")), nmu=shp["Psi"][1])
        coldim = dict(A="nx", B1="nu", B2="ndelta", B3="nz", B4="nomega", b5=None, C="nx", D1="nu", D2="ndelta",
                      D3="nz", D4="nomega", d5=None, E="nx", F1="nu", F2="ndelta", F3="nz", F4="nomega", f5=None,
                      G="ny", Psi="nmu")
        rowdim = {}
        rowdim.update({k: "nx" for k in self._state_input_mat_names})
        rowdim.update({k: "ny" for k in self._output_mat_names})
        rowdim.update({k: "nc" for k in self._constraint_mat_names})
        for k in self._sys_mat_names:
            r = d[rowdim[k]]
            c = 1 if coldim[k] is None else d[coldim[k]]
            if shp[k] == (0, 0):
                if k == "f5" and d["nc"]:
                    raise ValueError("Constraint vector 'f5' can only be null if all constraint matrices are null.")
                mat = np.zeros((r, c))
            else:
                mat = g[k]
                if shp[k][0] != r:
                    raise ValueError("Invalid shape for matrix/vector '%s':%s, row dimension must be equal to system "
                                     "dimension %d" % (k, shp[k], r))
                if shp[k][1] != c:
                    raise ValueError("Invalid shape for matrix/vector '%s':%s, column dimension must be %d"
                                     % (k, shp[k], c))
            if isinstance(mat, np.ndarray):
                mat = np.ascontiguousarray(mat, dtype=np.float64)
                mat.setflags(write=False)
            if mld_type == T.callable and not isinstance(mat, CallableMatrix):
                mat = CallableMatrix(mat, k)
            dict.__setitem__(self, k, mat)
        nu_l = self._bin_dims.get("nu_l", 0)
        nmu_l = self._bin_dims.get("nmu_l", 0)
        if nu_l > d["nu"] or nmu_l > d["nmu"]:
            raise ValueError("number of binary entries exceeds the variable dimension")
        info = self._mld_info
        info.clear()
        info.update(nx=d["nx"], nu=d["nu"], ndelta=d["ndelta"], nz=d["nz"], nomega=d["nomega"], ny=d["ny"],
                    nmu=d["nmu"], nv=d["nu"] + d["ndelta"] + d["nz"] + d["nmu"],
                    n_states=d["nx"], n_outputs=d["ny"], n_constraints=d["nc"],
                    nx_l=0, nu_l=nu_l, ndelta_l=d["ndelta"], nz_l=0, nomega_l=0, ny_l=0, nmu_l=nmu_l,
                    ts=self._meta["ts"], param_struct=self._meta["param_struct"],
                    required_params=self._get_required_params() if mld_type != T.numeric else None)
        for name, nb in (("x", 0), ("u", nu_l), ("delta", d["ndelta"]), ("z", 0), ("omega", 0), ("y", 0),
                         ("mu", nmu_l)):
            dim = info["n" + name]
            info["var_type_" + name] = atleast_2d_col(list("c" * (dim - nb) + "b" * nb)) if dim else np.empty((0, 1), str)
        info["var_type_v"] = np.vstack([info["var_type_" + n] for n in MldInfo._controllable_var_names])
        info["nv_l"] = nu_l + d["ndelta"] + nmu_l

    def _get_required_params(self):
        """Sorted names of every parameter some matrix depends on (reference: models/mld_model.py:954-963)."""
        names = set()
        for mat in self.values():
            if is_symbolic(mat):
                names.update(str(sym) for sym in mat.free_symbols)
            elif isinstance(mat, CallableMatrix):
                names.update(mat.required_params)
        return sorted(names)

    def __setitem__(self, key, value):
        if key in self._sys_mat_names:
            self.update(**{key: value})
        else:
            raise KeyError("key: '%s' is not a system matrix name." % key)

    def __setattr__(self, key, value):
        if key.startswith("_"):
            object.__setattr__(self, key, value)
        else:
            self[key] = value

    # ---- symbolic -> callable -> numeric (reference: models/mld_model.py:768-833) ----------------------------
    def _resolve_params(self, param_struct, ts):
        param_struct = param_struct if param_struct is not None else self._meta["param_struct"]
        if ts is ParNotSet:
            if param_struct and param_struct.get("ts", ParNotSet) is not ParNotSet:
                ts = param_struct["ts"]
            else:
                ts = self._meta["ts"]
        if param_struct is not None:
            param_struct = type(param_struct)(param_struct) if isinstance(param_struct, dict) else dict(param_struct)
            if param_struct.get("ts", ParNotSet) is not ParNotSet:
                param_struct["ts"] = ts
        if self._meta["ts"] is None and ts is not None:
            raise NotImplementedError("Discretization required")
        return param_struct, ts

    def _sibling(self, matrices, param_struct, ts):
        return MldModel(system_matrices=matrices, ts=ts, param_struct=param_struct,
                        bin_dims_struct={k: v for k, v in self._bin_dims.items() if k in ("nu_l", "nmu_l")})

    def to_callable(self, param_struct=None, ts=ParNotSet, copy=False):
        """Every matrix as a CallableMatrix (reference: models/mld_model.py:807-833)."""
        param_struct, ts = self._resolve_params(param_struct, ts)
        mats = {k: (m.copy() if copy else m) if isinstance(m, CallableMatrix) else CallableMatrix(m, k)
                for k, m in self.items()}
        return self._sibling(mats, param_struct, ts)

    @property
    def program(self):
        """The compiled register program of all non-constant matrices of a callable model (None if there are none)."""
        if self._mld_type != self.MldModelTypes.callable:
            raise TypeError("only callable models have a program; use to_callable() first")
        if self._program is None:
            varying = {k: m.expr for k, m in self.items() if not m.is_constant and m.size}
            object.__setattr__(self, "_program", ExprProgram(varying) if varying else False)
        return self._program or None

    def to_numeric(self, param_struct=None, ts=ParNotSet, copy=False):
        """Numeric model for one parameter set (reference: models/mld_model.py:768-805): the non-constant matrices
        come from one launch of hmpc_param_eval_f64 with B = 1."""
        T = self.MldModelTypes
        if self._mld_type == T.numeric:
            return self
        if self._mld_type == T.symbolic:
            # the reference prints a performance warning here and converts (:796-799)
            return self.to_callable(param_struct=param_struct, ts=ts).to_numeric(param_struct=param_struct, ts=ts)
        param_struct, ts = self._resolve_params(param_struct, ts)
        mats = {k: np.array(m(), dtype=np.float64) for k, m in self.items() if m.is_constant}
        prog = self.program
        if prog is not None:
            if param_struct is None:
                raise TypeError("to_numeric() of a callable model needs a param_struct with %s"
                                % list(prog.required_params))
            vals = prog.evaluate(prog.param_table(param_struct, B=1))
            mats.update({k: v[0].cpu().numpy() for k, v in vals.items()})
        return self._sibling(mats, param_struct, ts)

    def to_numeric_batch(self, param_struct=None, overrides=None, B=None, device="cuda"):
        """Batched counterpart of ``to_numeric``: name -> CUDA float64 tensor, ``[B, rows, cols]`` for the matrices
        that depend on parameters and ``[1, rows, cols]`` for the constant ones -- the ``mats`` argument of
        ``BatchMpc`` / the C ABI's mats[] with strides.  ``overrides``: name -> per-agent vector of length B; every
        other parameter is the scalar of ``param_struct``.  Empty matrices are left out."""
        import torch
        model = self if self._mld_type == self.MldModelTypes.callable else self.to_callable(param_struct=param_struct)
        if self._mld_type == self.MldModelTypes.numeric:
            return {k: torch.as_tensor(np.array(m), dtype=torch.float64).unsqueeze(0).to(device)
                    for k, m in self.items() if m.size}
        param_struct = param_struct if param_struct is not None else self._meta["param_struct"]
        out = {k: torch.as_tensor(np.array(m(), dtype=np.float64)).unsqueeze(0).to(device)
               for k, m in model.items() if m.is_constant and m.size}
        prog = model.program
        if prog is not None:
            out.update(prog.evaluate(prog.param_table(param_struct or {}, overrides=overrides, B=B, device=device)))
        return out

    # ---- simulation ------------------------------------------------------------------------------------
    LSimStruct_k = StructDict

    def lsim_k(self, x_k=ParNotSet, u_k=ParNotSet, delta_k=ParNotSet, z_k=ParNotSet, mu_k=ParNotSet, v_k=ParNotSet,
               omega_k=ParNotSet, solver=None, cons_tol=1e-6):
        """One simulation step on the GPU (reference: models/mld_model.py:647-699)."""
        if self._mld_type != self.MldModelTypes.numeric:
            raise TypeError("lsim_k needs a numeric model; call to_numeric(param_struct) first")
        from .. import sim
        return sim.lsim_k_single(self, x_k, u_k, delta_k, z_k, mu_k, v_k, omega_k, cons_tol)


class MldSystemModel(object):
    """A model in its three forms + the current parameter set (reference: models/mld_model.py:1001-1164).

    Exactly one of ``mld_numeric`` / ``mld_callable`` / ``mld_symbolic`` is given; a symbolic model is compiled to
    its callable form once and ``mld_numeric`` is re-evaluated (on the GPU) whenever the parameters change."""
    MldNames = StructDict(numeric="mld_numeric", callable="mld_callable", symbolic="mld_symbolic")

    def __init__(self, mld_numeric=None, mld_callable=None, mld_symbolic=None, param_struct=None, copy=False):
        self._param_struct = None
        self._mld_numeric = None
        self._mld_callable = None
        self._mld_symbolic = None
        self._version = _next_version()
        self.update_mld(mld_numeric=mld_numeric, mld_callable=mld_callable, mld_symbolic=mld_symbolic,
                        param_struct=param_struct, copy=copy, missing_param_check=True)

    def update_mld(self, mld_numeric=None, mld_callable=None, mld_symbolic=None, param_struct=None,
                   param_struct_subset=None, copy=False, missing_param_check=True, invalid_param_check=False,
                   **kwargs):
        mlds = (mld_numeric, mld_callable, mld_symbolic)
        types = (MldModel.MldModelTypes.numeric, MldModel.MldModelTypes.callable, MldModel.MldModelTypes.symbolic)
        if sum(m is not None for m in mlds) > 1:
            raise ValueError("Only one of {'mld_numeric', 'mld_callable', 'mld_symbolic'} can be used to "
                             "construct/update an %s" % type(self).__name__)
        if not all(m is None or isinstance(m, MldModel) for m in mlds):
            raise TypeError("Each of {'mld_numeric', 'mld_callable', 'mld_symbolic'} is required to be an instance of "
                            "MldModel or None.")
        for name, m, t in zip(("mld_numeric", "mld_callable", "mld_symbolic"), mlds, types):
            if m is not None and m.mld_type != t:
                raise TypeError("'%s' is required to be an instance of MldModel with mld_type:%s, not mld_type:'%s'"
                                % (name, t, m.mld_type))
        if any(m is not None for m in mlds):
            self._mld_numeric, self._mld_callable, self._mld_symbolic = mlds
        param_struct = param_struct if param_struct is not None else (self._param_struct or {})
        try:
            self._param_struct = self._resolve_params(param_struct, param_struct_subset, kwargs, missing_param_check,
                                                      invalid_param_check)
        except ValueError as ve:
            raise ValueError("A valid 'param_struct' is required, the argument was not provided or is invalid. %s"
                             % ve.args[0])
        if self._mld_callable is not None or self._mld_symbolic is not None:
            if self._mld_callable is None:
                self._mld_callable = self._mld_symbolic.to_callable(copy=copy)
            self._mld_numeric = self._mld_callable.to_numeric(param_struct=self._param_struct, copy=copy)
        self._version = _next_version()

    @property
    def mld_numeric(self):
        return self._mld_numeric

    @property
    def mld_callable(self):
        return self._mld_callable

    @property
    def mld_symbolic(self):
        return self._mld_symbolic

    @property
    def param_struct(self):
        return self._param_struct

    @param_struct.setter
    def param_struct(self, param_struct):
        self.update_param_struct(param_struct=param_struct)

    @property
    def version(self):
        return (self._version, self._mld_numeric.version if self._mld_numeric is not None else 0)

    # ---- parameters.  The stored parameter set is a table (name -> value); every call that takes parameters resolves
    # them against that table with the same three inputs the reference accepts (models/mld_model.py:1072-1149): a full
    # replacement (`param_struct`), a partial override (`param_struct_subset` and keyword arguments), and two checks.
    def _resolve_params(self, replacement, overrides, keywords, require_all, known_only):
        """-> the parameter table a call works with; the STORED table itself when nothing was given (callers test
        identity to reuse the stored numeric model)."""
        stored = self._param_struct
        if overrides is None:
            overrides = {}
        if not hasattr(overrides, "update") or not hasattr(overrides, "keys"):
            raise TypeError("Invalid type for 'param_struct_subset', must be dictionary like or None.")
        overrides.update(keywords)
        replacing = replacement is not None and replacement is not stored
        if not replacing and not overrides:
            return stored
        try:
            table = StructDict(replacement if replacing else stored)
            table.update(overrides)
        except (AttributeError, TypeError, ValueError):
            raise TypeError("Invalid type for 'param_struct', must be dictionary like or None.")
        if require_all:
            absent = set(self.get_required_params()) - set(table.keys())
            if absent:
                raise ValueError("The following keys are missing from param_struct: '%s'" % absent)
        if known_only:
            named = table if replacing else overrides          # what the caller spelled out
            unknown = set(named.keys()) - set((stored or {}).keys())
            if unknown:
                raise ValueError("Invalid keys:'%s' in kwargs/param_struct - keys must all exist in "
                                 "self.param_struct. Hint: either disable 'invalid_param_check' or update "
                                 "self.param_struct." % unknown)
        return table

    def update_param_struct(self, param_struct=None, param_struct_subset=None, missing_param_check=True,
                            invalid_param_check=False, **kwargs):
        """Store a new parameter table and re-evaluate the numeric model from it (one param-eval launch)."""
        table = self._resolve_params(param_struct, param_struct_subset, kwargs, missing_param_check, invalid_param_check)
        self._mld_numeric = self._evaluate(table)
        self._param_struct = table
        self._version = _next_version()

    def _evaluate(self, table):
        if table is self._param_struct:
            return self._mld_numeric
        if self._mld_callable is None:
            raise TypeError("AgentModel does not contain valid mld_callable.")
        return self._mld_callable.to_numeric(table)

    def get_mld_numeric(self, param_struct=None, param_struct_subset=None, missing_param_check=False,
                        invalid_param_check=True, copy=False, **kwargs):
        """The stored numeric model, or a fresh evaluation when other parameters are given."""
        table = self._resolve_params(param_struct, param_struct_subset, kwargs, missing_param_check, invalid_param_check)
        if table is self._param_struct:
            return self._mld_numeric
        if self._mld_callable is None:
            raise TypeError("AgentModel does not contain valid mld_callable.")
        return self._mld_callable.to_numeric(table, copy=copy)

    def get_mld_numeric_batch(self, overrides=None, B=None, param_struct=None, device="cuda"):
        """The model of B agents at once: name -> CUDA tensor ``[B|1, rows, cols]`` (one kernel launch).  ``overrides``
        maps parameter names to per-agent vectors; the remaining parameters come from ``param_struct`` (default: the
        stored one).  This is the batched form of calling ``get_mld_numeric(param_struct_i)`` for every agent i, which
        is what the reference's agents do (examples/.../micro_grid_agents.py:389-408)."""
        unknown = set(overrides or {}).difference(self._param_struct or {}) if param_struct is None else set()
        if unknown:
            raise ValueError("Invalid keys:'%s' in overrides - keys must all exist in self.param_struct." % unknown)
        src = self._mld_callable if self._mld_callable is not None else self._mld_numeric
        if src is None:
            raise TypeError("the model is empty")
        return src.to_numeric_batch(param_struct=param_struct if param_struct is not None else self._param_struct,
                                    overrides=overrides, B=B, device=device)

    def get_required_params(self):
        if self._mld_symbolic is not None:
            return set(self._mld_symbolic.mld_info.required_params)
        elif self._mld_callable is not None:
            return set(self._mld_callable.mld_info.required_params)
        return set()

    def __repr__(self):
        return "%s(mld_numeric=%s, mld_callable=%s, mld_symbolic=%s)" % (
            type(self).__name__, *("None" if m is None else "<MldModel %s>" % m.mld_type
                                   for m in (self._mld_numeric, self._mld_callable, self._mld_symbolic)))
