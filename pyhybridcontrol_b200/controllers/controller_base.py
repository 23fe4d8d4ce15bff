"""Controller lifecycle: build / solve / feedback / sim_step_k, constraint sets, simulation log.

Host-side mirror of the reference's ``ControllerBase`` / ``ConstraintSolvedController``
(controllers/controller_base.py:150-548) with the same method names, argument meaning and error behaviour.
The arithmetic is delegated to the GPU: condensing in ``MldEvoMatrices`` (K1), right-hand sides (K2) and the
mixed-integer solve (K3/K4) through ``BatchMpc`` with a batch of one.
"""
import time

import numpy as np
import torch

from ..models.mld_model import MldModel, MldSystemModel
from ..utils.structs import StructDict, ParNotSet, atleast_2d_col
from .components.mld_evolution_matrices import MldEvoMatrices
from .components.objective_atoms import ObjectiveAtoms


class ControllerBuildRequiredError(RuntimeError):
    pass


class ControllerSolverError(RuntimeError):
    pass


class MldSimLog(object):
    """Columnar simulation log: one ``[rows, dim]`` array per logged variable plus the list of instants k, so that a
    closed loop appends rows and the result frame (``get_concat_log``; reference layout: controllers/
    controller_base.py:116-146, columns ``(var_names, var_index)``, index k) is a concatenation of whole arrays.  An
    entry that lacks a variable reads as NaN (None for non-numeric variables).  ``log[k]`` gives the entry of instant k
    as a struct of column vectors, the form ``ControllerBase.sim_step_k`` and the rate atoms consume."""

    def __init__(self):
        self._row_of = {}            # k -> row
        self._ks = []                # row -> k
        self._cols = {}              # name -> [capacity, dim] array (float or object)
        self._has = {}               # name -> [capacity] bool: the entry has this variable

    # -- mapping surface
    def __contains__(self, k):
        return k in self._row_of

    def __len__(self):
        return len(self._ks)

    def __iter__(self):
        return iter(list(self._ks))

    def keys(self):
        return list(self._ks)

    def get(self, k, default=None):
        return self[k] if k in self._row_of else default

    def __getitem__(self, k):
        r = self._row_of[k]
        out = StructDict()
        for name, col in self._cols.items():
            out[name] = col[r].reshape(-1, 1) if self._has[name][r] else self._blank(col)[0].reshape(-1, 1)
        return out

    def pop(self, k, default=None):
        if k not in self._row_of:
            return default
        entry = self[k]
        r = self._row_of.pop(k)
        self._ks.pop(r)
        for name in self._cols:
            self._cols[name] = np.delete(self._cols[name], r, axis=0)
            self._has[name] = np.delete(self._has[name], r, axis=0)
        self._row_of = {kk: i for i, kk in enumerate(self._ks)}
        return entry

    @staticmethod
    def _blank(col, rows=1):
        if col.dtype == object:
            return np.full((rows, col.shape[1]), None, dtype=object)
        return np.full((rows, col.shape[1]), np.nan)

    def _row(self, k):
        r = self._row_of.get(k)
        if r is None:
            r = len(self._ks)
            self._row_of[k] = r
            self._ks.append(k)
            for name, col in self._cols.items():
                if r >= col.shape[0]:
                    grow = max(16, col.shape[0])
                    self._cols[name] = np.concatenate([col, self._blank(col, grow)], axis=0)
                    self._has[name] = np.concatenate([self._has[name], np.zeros(grow, dtype=bool)])
        return r

    # -- writers (same call forms as the reference's log)
    def set_sim_k(self, k, sim_k=None, **kwargs):
        self.pop(k, None)
        self.update_sim_k(k=k, sim_k=sim_k, **kwargs)

    def update_sim_k(self, k, sim_k=None, **kwargs):
        items = dict(sim_k) if sim_k is not None else {}
        items.update(kwargs)
        r = self._row(k)
        for name, var in items.items():
            if var is None:
                continue
            var = np.asarray(atleast_2d_col(var))
            numeric = np.issubdtype(var.dtype, np.number) or np.issubdtype(var.dtype, np.bool_)
            col = self._cols.get(name)
            if col is None:
                cap = max(16, len(self._ks))
                col = (np.full((cap, var.shape[0]), np.nan) if numeric else np.full((cap, var.shape[0]), None, dtype=object))
                self._cols[name] = col
                self._has[name] = np.zeros(cap, dtype=bool)
            if var.shape != (col.shape[1], 1):
                raise ValueError("shape of var_k must match previous inserts")
            if numeric and col.dtype != object:
                col[r] = var[:, 0]
            else:
                if col.dtype != object:
                    col = col.astype(object)
                    self._cols[name] = col
                col[r] = list(var[:, 0])
            self._has[name][r] = True

    def get_concat_log(self, add_column_levels=None):
        import pandas as pd
        order = np.argsort(np.array(self._ks, dtype=object)) if self._ks else np.zeros(0, dtype=int)
        index = [self._ks[i] for i in order]
        frames = {}
        for name, col in self._cols.items():
            if col.shape[1]:
                block = col[:len(self._ks)][order]
                if col.dtype == object:
                    block = np.array(block.tolist(), dtype=object)
                    try:
                        block = block.astype(float)
                    except (TypeError, ValueError):
                        pass
                frames[name] = pd.DataFrame(block)
        df = pd.concat(frames, keys=list(frames), axis=1)
        df.columns.names = ["var_names", "var_index"]
        df.index = index
        df.index.name = "k"
        if add_column_levels:
            df = pd.concat([df], keys=[add_column_levels], axis=1)
        return df


class EvoConstraint(object):
    """What ``gen_evo_constraints`` returns: one set of rows  H_v v <= H_x x0 + H_omega(.) + H_5."""

    def __init__(self, x_k=None, omega_tilde_k=None, omega_scenarios_k=None, N_tilde=None):
        self.x_k, self.omega_tilde_k, self.omega_scenarios_k, self.N_tilde = x_k, omega_tilde_k, omega_scenarios_k, N_tilde


class ControllerBase(object):
    def __init__(self, model=None, x_k=None, omega_tilde_k=None, N_p=None, N_tilde=None, agent=None,
                 mld_numeric=None, mld_callable=None, mld_symbolic=None, param_struct=None, device="cuda"):
        self._N_p = N_p if N_p is not None else 0
        self._N_tilde = N_tilde if N_tilde is not None else self._N_p + 1
        self._build_required = True
        self._device = device
        if agent is not None and model is not None:
            raise ValueError("agent and model cannot both be set")
        elif agent is not None:
            self._agent, self._model = agent, None
        else:
            self._model = model if model is not None else MldSystemModel(mld_numeric=mld_numeric,
                                                                         mld_symbolic=mld_symbolic,
                                                                         mld_callable=mld_callable,
                                                                         param_struct=param_struct)
            self._agent = None
        self._sim_log = MldSimLog()
        self._solve_time_overall = 0
        self._solve_time_solver = 0
        self._built_version = None
        self.reset_components(x_k=x_k, omega_tilde_k=omega_tilde_k)

    # ---- models
    @property
    def N_p(self):
        return self._N_p

    @property
    def N_tilde(self):
        return self._N_tilde

    @property
    def sim_model(self):
        return self._agent.sim_model if self._agent else self._model

    @property
    def control_model(self):
        return self._agent.control_model if self._agent else self._model

    @property
    def mld_numeric_k(self) -> MldModel:
        return self.control_model.mld_numeric

    @property
    def mld_numeric_tilde(self):
        return None

    @property
    def mld_info_k(self):
        return self.mld_numeric_k.mld_info

    @property
    def sim_log(self):
        return self._sim_log

    def _model_version(self):
        return (id(self.mld_numeric_k), self.mld_numeric_k.version)

    @property
    def build_required(self):
        return self._build_required or self._built_version != self._model_version()

    def set_build_required(self):
        self._build_required = True

    def reset_components(self, x_k=None, omega_tilde_k=None):
        info = self.mld_info_k
        self._x_k = self._as_col(x_k, (info.nx, 1), "x_k") if x_k is not None else np.zeros((info.nx, 1))
        shape = (info.nomega * self.N_tilde, 1)
        self._omega_tilde_k = (self._as_col(omega_tilde_k, shape, "omega_tilde_k") if omega_tilde_k is not None
                               else np.zeros(shape))
        self._build_required = True

    @staticmethod
    def _as_col(value, shape, name):
        value = np.asarray(atleast_2d_col(value), dtype=np.float64)
        if value.dtype == np.object_:
            raise TypeError("'new_value' must be a numeric array like object or None.")
        if value.shape != shape:
            raise ValueError("Incorrect shape:%s for %s, a shape of %s is required." % (value.shape, name, shape))
        return value

    def update_horizons(self, N_p=ParNotSet, N_tilde=ParNotSet):
        old = (self._N_p, self._N_tilde)
        self._N_p = N_p if N_p is not ParNotSet else self._N_p or 0
        self._N_tilde = N_tilde if N_tilde is not ParNotSet else self._N_p + 1
        if old != (self._N_p, self._N_tilde):
            self.reset_components()

    def get_sim_k(self, k, default=None, x_k=None, u_k=None, omega_k=None):
        if k in self.sim_log:
            return self.sim_log[k]
        elif default is not None:
            return default
        return None

    def sim_step_k(self, k, x_k=None, u_k=None, omega_k=None, mld_numeric_k=None, solver=None, step_state=True):
        """reference: controllers/controller_base.py:229-253."""
        var_k = self.variables_k
        omega_k = omega_k if omega_k is not None else var_k.omega
        x_k = x_k if x_k is not None else var_k.x
        u_k = u_k if u_k is not None else var_k.u
        sim_model = mld_numeric_k if mld_numeric_k is not None else self.sim_model.mld_numeric
        lsim_k = sim_model.lsim_k(x_k=x_k, u_k=u_k, omega_k=omega_k, solver=solver)
        if step_state:
            lsim_k.update({name + "_hat": var for name, var in var_k.items()})
            self.sim_log.set_sim_k(k=k, sim_k=lsim_k)
            self.sim_log.update_sim_k(k=k, time_solve_overall=self._solve_time_overall,
                                      time_in_solver=self._solve_time_solver)
            self.x_k = lsim_k.x_k1
        else:
            del lsim_k["x_k1"]
        return lsim_k


class ConstraintSolvedController(ControllerBase):
    def reset_components(self, x_k=None, omega_tilde_k=None):
        super(ConstraintSolvedController, self).reset_components(x_k=x_k, omega_tilde_k=omega_tilde_k)
        if self.control_model.mld_numeric is not None:
            self._mld_evo_matrices = MldEvoMatrices(self, device=self._device)
            self._std_obj_atoms = ObjectiveAtoms(self.mld_info_k, self.N_p, self.N_tilde)
        else:
            self._mld_evo_matrices = None
            self._std_obj_atoms = None
        self._std_evo_constraints = []
        self._other_constraints = []
        self._constraints = []
        self._disable_soft = False
        self._problem = None
        self._solution = None
        self._variables_k_neg1 = None
        self._build_required = True

    @property
    def mld_evo_matrices(self):
        return self._mld_evo_matrices

    @property
    def constraints(self):
        return self._constraints

    @property
    def problem(self):
        return self._problem

    @property
    def x_k(self):
        return self._x_k

    @x_k.setter
    def x_k(self, value):
        self._x_k = self._as_col(value, (self.mld_info_k.nx, 1), "x_k")

    @property
    def omega_tilde_k(self):
        return self._omega_tilde_k

    @omega_tilde_k.setter
    def omega_tilde_k(self, value):
        self._omega_tilde_k = self._as_col(value, (self.mld_info_k.nomega * self.N_tilde, 1), "omega_tilde_k")

    @property
    def variables_k_neg1(self):
        return self._variables_k_neg1

    @variables_k_neg1.setter
    def variables_k_neg1(self, value):
        self._variables_k_neg1 = value

    # ---- constraints (reference :411-474)
    def gen_evo_constraints(self, x_k=None, omega_tilde_k=None, omega_scenarios_k=ParNotSet, N_p=ParNotSet,
                            N_tilde=ParNotSet, mld_numeric_k=ParNotSet, mld_numeric_tilde=ParNotSet,
                            mld_evo_matrices=ParNotSet):
        N_tilde = N_tilde if N_tilde is not ParNotSet else self.N_tilde
        N_p = N_p if N_p is not ParNotSet else self.N_p
        if not N_tilde <= self.N_tilde:
            raise ValueError("N_tilde: %s must be less or equal to self.N_tilde: %s" % (N_tilde, self.N_tilde))
        if not N_p <= self.N_tilde:
            raise ValueError("N_p: %s must be less or equal to self.N_tilde: %s" % (N_p, self.N_tilde))
        if mld_numeric_k is not ParNotSet or mld_numeric_tilde is not ParNotSet or mld_evo_matrices is not ParNotSet:
            raise NotImplementedError("constraints from a different model than the controller's are not supported")
        if omega_scenarios_k is None:
            return None
        sc = None
        if omega_scenarios_k is not ParNotSet:
            sc = np.asarray(atleast_2d_col(omega_scenarios_k), dtype=np.float64)
            if sc.shape[0] != self.mld_info_k.nomega * N_tilde and sc.shape[0] != self.mld_info_k.nomega * self.N_tilde:
                raise ValueError("omega_scenarios_k must have nomega*N_tilde rows")
        w = None if omega_tilde_k is None else np.asarray(atleast_2d_col(omega_tilde_k), dtype=np.float64)
        x = None if x_k is None else self._as_col(x_k, (self.mld_info_k.nx, 1), "x_k")
        return EvoConstraint(x_k=x, omega_tilde_k=w, omega_scenarios_k=sc,
                             N_tilde=None if N_tilde == self.N_tilde else N_tilde)

    def set_constraints(self, std_evo_constaints=ParNotSet, other_constraints=ParNotSet, disable_soft_constraints=False):
        if std_evo_constaints is not ParNotSet:
            self._std_evo_constraints = (std_evo_constaints if std_evo_constaints is not None
                                         else [EvoConstraint()])
        if other_constraints is not ParNotSet:
            self._other_constraints = other_constraints if other_constraints is not None else []
        for con in list(self._std_evo_constraints) + list(self._other_constraints):
            if not isinstance(con, EvoConstraint):
                raise TypeError("constraints must be generated with gen_evo_constraints()")
        self._disable_soft = bool(disable_soft_constraints and self.mld_info_k.nmu)
        self._constraints = list(self._std_evo_constraints) + list(self._other_constraints)
        self._build_required = True

    def build(self, with_std_constraints=True, disable_soft_constraints=True):
        self._mld_evo_matrices.update()
        self.set_constraints(std_evo_constaints=None if with_std_constraints else [],
                             disable_soft_constraints=disable_soft_constraints)
        self._cost_atoms = None
        self._finish_build()

    def _finish_build(self):
        self._problem = StructDict(constraints=self._constraints, sense=getattr(self, "_sense", "minimize"))
        self._build_required = False
        self._built_version = self._model_version()

    # ---- solve (reference :491-548)
    def _cost_terms(self, k):
        """-> dict(cost_v, w_x, w_y, const) from the Linear atoms; overridden by MpcController."""
        return dict(cost_v=None, w_x=None, w_y=None, const=0.0)

    def solve(self, k, x_k=None, omega_tilde_k=None, external_solve=None, solver=None, verbose=False, warm_start=True,
              parallel=False, *args, method=None, **kwargs):
        t_start = time.time()
        try:
            if x_k is not None:
                self.x_k = x_k
            if omega_tilde_k is not None:
                self.omega_tilde_k = omega_tilde_k
            k_neg1 = k - 1 if k is not None else k
            self.variables_k_neg1 = self.get_sim_k(k=k_neg1)
            if self.build_required:
                raise ControllerBuildRequiredError(
                    "%s problem has not been built or needs to be rebuilt." % self.__class__.__name__)
            if external_solve is None:
                solution = self._solve_on_gpu(k, **kwargs)
                if not np.isfinite(solution):
                    raise ControllerSolverError("solve() failed with objective: '%s', and status: %s"
                                                % (solution, self._status_name))
            else:
                self._solve_time_solver = 0
                solution = external_solve
            return solution
        finally:
            self._solve_time_overall = time.time() - t_start

    def _solve_on_gpu(self, k, **solver_kwargs):
        from .. import cabi
        batch = self._mld_evo_matrices.batch
        info = self.mld_info_k
        opts = cabi.default_opts()
        if "MIPGap" in solver_kwargs:
            opts.mip_rel_gap = float(solver_kwargs["MIPGap"])
        if "mip_rel_gap" in solver_kwargs:
            opts.mip_rel_gap = float(solver_kwargs["mip_rel_gap"])
        if "max_nodes" in solver_kwargs:
            opts.max_nodes = int(solver_kwargs["max_nodes"])
        batch.opts = opts
        batch.dp_opts.mip_rel_gap = opts.mip_rel_gap
        if "max_nodes" in solver_kwargs:
            batch.dp_opts.max_nodes = int(solver_kwargs["max_nodes"])
        batch.disable_soft_constraints = self._disable_soft
        terms = self._cost_terms(k)
        x0 = self._x_k.reshape(1, -1)
        w = self._omega_tilde_k.reshape(1, -1)
        extra = []
        with_std = False
        for con in self._constraints:
            if (con.x_k is None and con.omega_tilde_k is None and con.omega_scenarios_k is None and con.N_tilde is None
                    and not with_std):
                with_std = True
                continue
            sc = con.omega_scenarios_k
            if sc is not None:
                sc = sc[:info.nomega * self.N_tilde]
                full = np.zeros((info.nomega * self.N_tilde, sc.shape[1]))
                full[:sc.shape[0]] = sc
                sc = full[None]
            ww = con.omega_tilde_k
            if ww is not None:
                full = np.zeros((1, info.nomega * self.N_tilde))
                full[0, :ww.size] = ww.ravel()
                ww = full
            xk = None if con.x_k is None else np.asarray(con.x_k, dtype=np.float64).reshape(1, -1)
            extra.append(dict(omega_tilde_k=ww, omega_scenarios_k=sc, N_tilde=con.N_tilde, x_k=xk))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if getattr(self, "_general_path", False):
            # a general MIQP: assembled on the device from the condensed matrices, solved by the ADMM branch and bound
            from .. import miqp
            sign = -1.0 if getattr(self, "_sense", "minimize").lower().startswith("max") else 1.0
            atoms = list(self._std_obj_atoms.iter_atoms()) if (self._with_std_objective and self._std_obj_atoms) else []
            prev = {name: np.asarray(val, dtype=np.float64).ravel() for name, val in (self.variables_k_neg1 or {}).items()
                    if val is not None and name in ("x", "u", "delta", "z", "omega", "y", "mu", "v")}
            prob = miqp.assemble(batch, x0, w, atoms, prev=prev, constraint_sets=extra, with_std_constraints=with_std,
                                 sign=sign)
            qopts = cabi.miqp_default_opts(mip_rel_gap=opts.mip_rel_gap)
            if "max_nodes" in solver_kwargs:
                qopts.max_nodes = int(solver_kwargs["max_nodes"])
            res = miqp.solve(prob, qopts)
            terms = dict(const=0.0)
            res["obj"] = sign * res["obj"]           # (assemble() already applied the sense to the linear atoms)
        else:
            res = batch.solve(x0, w, cost_v=terms["cost_v"], w_x=terms["w_x"], w_y=terms["w_y"], extra_constraints=extra,
                              with_std_constraints=with_std, quad=terms.get("quad"))
        status = int(res["status"].cpu()[0])
        if status in (2, 5) and res.get("solver") == "stage_dp" and not terms.get("quad"):
            # search budget exhausted (or the agent fell outside the class): the general kernel takes over
            keep = batch.solver
            batch.solver = "bnc"
            try:
                res = batch.solve(x0, w, cost_v=terms["cost_v"], w_x=terms["w_x"], w_y=terms["w_y"],
                                  extra_constraints=extra, with_std_constraints=with_std)
            finally:
                batch.solver = keep
            status = int(res["status"].cpu()[0])
        ev1.record()
        torch.cuda.synchronize()
        self._solve_time_solver = ev0.elapsed_time(ev1) * 1e-3
        self._status_name = cabi.SOLVE_STATUS.get(status, str(status))
        self._stats = dict(zip(cabi.STAT_NAMES, res["stats"].cpu().numpy()[0].tolist()))
        obj_dev = float(res["obj"].cpu()[0])
        # a search that ran out of budget still returns its incumbent, like a solver stopped by a node or time limit
        # (the reference runs Gurobi with TimeLimit / MIPGap, micro_grid_control_simulation.py:232); the status and the
        # certified gap stay readable in status_name / solver_stats
        if status not in (0, 2, 3) or not np.isfinite(obj_dev):
            self._solution = None
            return float("inf") if status == 1 else float("nan")
        sign = -1.0 if getattr(self, "_sense", "minimize").lower().startswith("max") else 1.0
        v = res["v"]
        xt, yt = batch.predictions(v, x0, w)
        self._solution = StructDict(v=v.cpu().numpy().reshape(-1, 1),
                                    x=(xt.cpu().numpy().reshape(-1, 1) if xt is not None else np.empty((0, 1))),
                                    y=(yt.cpu().numpy().reshape(-1, 1) if yt is not None else np.empty((0, 1))))
        if getattr(self, "_general_path", False):
            return obj_dev                            # (sense and constants are part of the assembled problem)
        return sign * (obj_dev + terms["const"])

    def feedback(self, k, x_k=None, omega_tilde_k=None, external_solve=None, solver=None, verbose=False,
                 warm_start=True, parallel=False, *args, method=None, **kwargs):
        self.solve(k=k, x_k=x_k, omega_tilde_k=omega_tilde_k, external_solve=external_solve, solver=solver,
                   warm_start=warm_start, verbose=verbose, parallel=parallel, method=method, **kwargs)
        return self.variables_k

    # ---- results (reference: controllers/components/variables.py:75-85, 288-317)
    def variables_N_tilde(self):
        info = self.mld_info_k
        batch = self._mld_evo_matrices.batch
        out = StructDict()
        sol = self._solution
        v = sol.v if sol is not None else np.full((info.nv * self.N_tilde, 1), np.nan)
        for name in ("u", "delta", "z", "mu"):
            out[name] = v[batch.var_index(name)] if info["n" + name] else np.empty((0, 1))
        out["v"] = v
        out["x"] = sol.x if sol is not None else np.full((info.nx * self.N_tilde, 1), np.nan)
        out["y"] = sol.y if sol is not None else np.full((info.ny * self.N_tilde, 1), np.nan)
        out["omega"] = self._omega_tilde_k
        return out

    @property
    def variables_k(self):
        info = self.mld_info_k
        full = self.variables_N_tilde()
        out = StructDict()
        for name in ("x", "u", "delta", "z", "omega", "y", "mu", "v"):
            dim = info.nv if name == "v" else info["n" + name]
            out[name] = full[name][:dim]
        if info.nx:
            out["x"] = self._x_k
        return out


class PredictiveController(ConstraintSolvedController):
    pass


class NonPredictiveController(ControllerBase):
    pass
