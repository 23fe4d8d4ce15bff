"""MLD data model + one-step simulation, numeric models only.

Mirrors the reference's ``MldInfo`` / ``MldModel`` / ``MldSystemModel`` (models/mld_model.py:109, 391, 1001):

    x(k+1) = A x + B1 u + B2 delta + B3 z + B4 omega + b5
    y(k)   = C x + D1 u + D2 delta + D3 z + D4 omega + d5
    E x + F1 u + F2 delta + F3 z + F4 omega + G y + Psi mu <= f5 ,   mu >= 0

The arithmetic (``lsim_k``, auxiliary-variable computation) runs on the GPU through the C ABI
(hmpc_lsim_step_f64, hmpc_milp_solve_f64); this module is host-side bookkeeping: dimension rules, defaults,
variable types, change tracking.  Symbolic / callable models (reference: utils/matrix_utils.py:279-562) are
outside the hot path and are rejected with ``NotImplementedError``.
"""
import itertools

import numpy as np

from ..utils.structs import StructDict, ParNotSet, atleast_2d_col

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
        self.update(system_matrices=system_matrices, ts=ts, param_struct=param_struct,
                    bin_dims_struct=bin_dims_struct, var_types_struct=var_types_struct, _from_init=True, **kwargs)

    # ---- reference API -------------------------------------------------------------------------------
    @property
    def mld_info(self):
        return self._mld_info

    @property
    def mld_type(self):
        return self.MldModelTypes.numeric

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
            if callable(mat) or type(mat).__module__.startswith("sympy"):
                raise NotImplementedError("callable / symbolic MLD matrices are outside the GPU hot path; pass the "
                                          "numeric matrices (reference: MldModel.to_numeric, mld_model.py:768-805)")
            mat = np.array(atleast_2d_col(mat), dtype=np.float64)
            if not np.issubdtype(mat.dtype, np.number):
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
        shp = {k: (g[k].shape if k in g and 0 not in g[k].shape else (0, 0)) for k in self._sys_mat_names}
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
                if mat.shape[0] != r:
                    raise ValueError("Invalid shape for matrix/vector '%s':%s, row dimension must be equal to system "
                                     "dimension %d" % (k, mat.shape, r))
                if mat.shape[1] != c:
                    raise ValueError("Invalid shape for matrix/vector '%s':%s, column dimension must be %d"
                                     % (k, mat.shape, c))
            mat = np.ascontiguousarray(mat, dtype=np.float64)
            mat.setflags(write=False)
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
                    ts=self._meta["ts"], param_struct=self._meta["param_struct"])
        for name, nb in (("x", 0), ("u", nu_l), ("delta", d["ndelta"]), ("z", 0), ("omega", 0), ("y", 0),
                         ("mu", nmu_l)):
            dim = info["n" + name]
            info["var_type_" + name] = atleast_2d_col(list("c" * (dim - nb) + "b" * nb)) if dim else np.empty((0, 1), str)
        info["var_type_v"] = np.vstack([info["var_type_" + n] for n in MldInfo._controllable_var_names])
        info["nv_l"] = nu_l + d["ndelta"] + nmu_l

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

    def to_numeric(self, param_struct=None, ts=ParNotSet, copy=False):
        return self

    # ---- simulation ------------------------------------------------------------------------------------
    LSimStruct_k = StructDict

    def lsim_k(self, x_k=ParNotSet, u_k=ParNotSet, delta_k=ParNotSet, z_k=ParNotSet, mu_k=ParNotSet, v_k=ParNotSet,
               omega_k=ParNotSet, solver=None, cons_tol=1e-6):
        """One simulation step on the GPU (reference: models/mld_model.py:647-699)."""
        from .. import sim
        return sim.lsim_k_single(self, x_k, u_k, delta_k, z_k, mu_k, v_k, omega_k, cons_tol)


class MldSystemModel(object):
    """Numeric-only counterpart of the reference's MldSystemModel (models/mld_model.py:1001-1164)."""

    def __init__(self, mld_numeric=None, mld_callable=None, mld_symbolic=None, param_struct=None, copy=False):
        if mld_callable is not None or mld_symbolic is not None:
            raise NotImplementedError("callable / symbolic models are outside the GPU hot path (SURVEY.md 8 f4)")
        if mld_numeric is not None and not isinstance(mld_numeric, MldModel):
            raise TypeError("'mld_numeric' is required to be an instance of MldModel or None.")
        self._mld_numeric = mld_numeric
        self._param_struct = StructDict(param_struct or {})
        self._version = _next_version()

    def update_mld(self, mld_numeric=None, param_struct=None, **kwargs):
        if mld_numeric is not None:
            self._mld_numeric = mld_numeric
        if param_struct is not None:
            self._param_struct = StructDict(param_struct)
        self._version = _next_version()

    @property
    def mld_numeric(self):
        return self._mld_numeric

    @property
    def param_struct(self):
        return self._param_struct

    @property
    def version(self):
        return (self._version, self._mld_numeric.version if self._mld_numeric is not None else 0)

    def get_mld_numeric(self, param_struct=None, **kwargs):
        return self._mld_numeric
