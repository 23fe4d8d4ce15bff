"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of how the reference turns (condensed matrices, x_k, omega_tilde_k, cost atoms) into the
optimisation problem it hands to cvxpy:

* decision vector layout, integrality, mu >= 0         : controllers/components/variables.py:189-243
  (``v`` = per step ``[u; delta; z; mu]``; col-major reshape :298-307)
* state / output predictions as affine maps of v       : variables.py:245-286
* constraint  H_v v <= H_x x0 + H_w w + H_5, scenario row-min, reduced horizon
                                                       : controllers/controller_base.py:411-456
* ``mu == 0`` when soft constraints are disabled       : controller_base.py:467-472
* cost-atom grammar and semantics                      : controllers/components/objective_atoms.py:453-496, 308-363,
  weights :79-206, rate variables :297-305

Parity status: PINNED against the unmodified reference's own assembly code (EvoVariables, ObjectiveAtoms,
gen_evo_constraints, set_constraints, MpcController.build), run in the build container under
``oracle/ref_shim.load_controllers`` with ``oracle/mini_cvxpy.py`` standing in for cvxpy's modelling layer (cvxpy is a
third-party dependency that is not installed and not pinned by the reference; the stand-in restates its documented
shape / operator semantics).  Vectors: ``tests/golden/assembly_*.npz`` (``tests/golden/make_golden_assembly.py``);
checker: ``tests/test_oracle_assembly_pinned.py`` -- variable layout, every row and right-hand side, bounds,
integrality, cost vector, and the objective value at random points for every atom type.

The result is the canonical form used by every solver in this repo:

    minimise   0.5 v' P v + c' v + c0        (P may be None)
    subject to H v <= rhs ,  lb <= v <= ub ,  v[j] in {0,1} for is_bin[j]
"""
import re

import numpy as np

VAR_NAMES = ("x", "u", "delta", "z", "omega", "y", "mu", "v")
_ATOM_RE = re.compile(r"(Linear)|(Quadratic)|([L](1|(22)|(inf)))")
_RATE_RE = re.compile(r"[dD][^e]")


class Problem(object):
    def __init__(self, n):
        self.n = n
        self.P = None
        self.c = np.zeros(n)
        self.c0 = 0.0
        self.H = np.zeros((0, n))
        self.rhs = np.zeros(0)
        self.lb = np.full(n, -np.inf)
        self.ub = np.full(n, np.inf)
        self.is_bin = np.zeros(n, dtype=bool)
        self.n_v = n  # leading entries that are the MLD's v~ (aux epigraph variables come after)

    def add_rows(self, H, rhs):
        H = np.asarray(H, dtype=float)
        if H.shape[1] < self.n:
            H = np.hstack([H, np.zeros((H.shape[0], self.n - H.shape[1]))])
        self.H = np.vstack([self.H, H])
        self.rhs = np.concatenate([self.rhs, np.asarray(rhs, dtype=float).ravel()])

    def add_cols(self, k, lb=-np.inf, ub=np.inf):
        self.c = np.concatenate([self.c, np.zeros(k)])
        self.lb = np.concatenate([self.lb, np.full(k, lb)])
        self.ub = np.concatenate([self.ub, np.full(k, ub)])
        self.is_bin = np.concatenate([self.is_bin, np.zeros(k, dtype=bool)])
        self.H = np.hstack([self.H, np.zeros((self.H.shape[0], k))])
        if self.P is not None:
            P = np.zeros((self.n + k, self.n + k))
            P[:self.n, :self.n] = self.P
            self.P = P
        self.n += k
        return np.arange(self.n - k, self.n)

    def objective(self, v):
        v = np.asarray(v, dtype=float).ravel()
        val = self.c @ v + self.c0
        if self.P is not None:
            val += 0.5 * v @ self.P @ v
        return float(val)


def var_layout(dims, Nt):
    """Index arrays into v~ for u/delta/z/mu (each (dim*Nt,), step-major) -- variables.py:233-241."""
    nv = dims["nv"]
    off = 0
    idx = {}
    for name in ("u", "delta", "z", "mu"):
        d = dims["n" + name]
        idx[name] = (np.arange(Nt)[:, None] * nv + off + np.arange(d)[None, :]).ravel()
        off += d
    idx["v"] = np.arange(nv * Nt)
    return idx


def affine_maps(evo, dims, Nt, x0, omega_t):
    """var_N_tilde = M v + m0 for every variable name -- variables.py:245-286, 288-317."""
    n = dims["nv"] * Nt
    idx = var_layout(dims, Nt)
    maps = {}
    for name in ("u", "delta", "z", "mu", "v"):
        M = np.zeros((idx[name].size, n))
        M[np.arange(idx[name].size), idx[name]] = 1.0
        maps[name] = (M, np.zeros(idx[name].size))
    x0 = np.asarray(x0, dtype=float).reshape(-1)
    w = np.asarray(omega_t, dtype=float).reshape(-1)
    if dims["nx"]:
        maps["x"] = (evo["Gamma_v"], evo["Phi_x"] @ x0 + evo["Gamma_omega"] @ w + evo["Gamma_5"].ravel())
    else:
        maps["x"] = (np.zeros((0, n)), np.zeros(0))
    if dims["ny"]:
        lx = evo["L_x"] @ x0 if dims["nx"] else 0.0
        maps["y"] = (evo["L_v"], lx + evo["L_omega"] @ w + evo["L_5"].ravel())
    else:
        maps["y"] = (np.zeros((0, n)), np.zeros(0))
    maps["omega"] = (np.zeros((w.size, n)), w)
    return maps


def evo_rhs(evo, dims, x0, omega_t=None, omega_scenarios=None, N_tilde=None):
    """RHS of the evolution constraint; scenario form takes the row-wise min over scenario columns
    (controller_base.py:440-452); reduced horizon keeps the first N_tilde*nc rows
    (mld_evolution_matrices.py:89-105)."""
    x0 = np.asarray(x0, dtype=float).reshape(-1)
    rows = evo["H_v"].shape[0] if N_tilde is None else N_tilde * dims["nc"]
    Hw = evo["H_omega"][:rows]
    if omega_scenarios is not None:
        how = np.min(Hw @ np.asarray(omega_scenarios, dtype=float), axis=1)
    else:
        how = Hw @ np.asarray(omega_t, dtype=float).reshape(-1)
    hx = evo["H_x"][:rows] @ x0 if dims["nx"] else 0.0
    return evo["H_v"][:rows], hx + how + evo["H_5"][:rows].ravel()


def parse_atom_key(key):
    """'q_mu', 'Q_x', 'q_L1_du_N_p', ... -> (weight_type, atom_type, var_name, is_rate, post_fix)
    -- objective_atoms.py:453-471."""
    info = key.split("_")
    wtype = "vector" if "".join(info[0:1]).islower() else "matrix"
    atom = "".join(info[1:2]).capitalize()
    if not _ATOM_RE.search(atom):
        atom = "Linear" if wtype == "vector" else "Quadratic"
        var = "".join(info[1:2]).lower()
        post = "_".join(info[2:])
    else:
        var = "".join(info[2:3]).lower()
        post = "_".join(info[3:])
    rate = False
    if _RATE_RE.search(var):
        var = var[1:]
        rate = True
    if var not in VAR_NAMES or post not in ("N_p", "N_tilde", "f", ""):
        raise ValueError("weight_name: '%s' is not valid" % key)
    return wtype, atom, var, rate, post


def expand_weight(wtype, value, dim, Nt, N_p, post, prev=None):
    """Tile / block-repeat a weight to the full horizon -- objective_atoms.py:79-206, 480-485."""
    value = np.asarray(value, dtype=float)
    if value.ndim == 0:
        value = value.reshape(1, 1)
    elif value.ndim == 1:
        value = value[:, None]
    if not post:
        post = "N_tilde" if value.shape[0] in (dim, dim * Nt) else "N_p"
    length = dict(N_tilde=Nt, N_p=N_p, f=1)[post]
    if wtype == "vector":
        W = np.zeros((dim * Nt, 1)) if prev is None else prev.copy()
        if value.shape[1] != 1:
            raise ValueError("vector weight must have one column")
        if post == "f":
            if value.shape[0] != dim:
                raise ValueError("bad terminal weight")
            W[-dim:] = value
        else:
            if value.shape[0] == dim:
                value = np.tile(value, (length, 1))
            elif value.shape[0] != dim * length:
                raise ValueError("bad weight rows")
            W[:dim * length] = value
        return W
    W = np.zeros((dim * Nt, dim * Nt)) if prev is None else prev.copy()
    if value.shape[0] != value.shape[1]:
        raise ValueError("matrix weight must be square")
    if post == "f":
        if value.shape[0] != dim:
            raise ValueError("bad terminal weight")
        W[-dim:, -dim:] = value
    else:
        if value.shape[0] == dim:
            full = np.zeros((dim * length, dim * length))
            for k in range(length):
                full[k * dim:(k + 1) * dim, k * dim:(k + 1) * dim] = value
            value = full
        elif value.shape[0] != dim * length:
            raise ValueError("bad weight rows")
        W[:dim * length, :dim * length] = value
    return W


def build_problem(evo, dims, var_types, Nt, x0, omega_t, atoms=None, N_p=None, disable_soft_constraints=False,
                  omega_scenarios=None, extra_constraints=(), var_k_neg1=None, with_std_constraints=True):
    """-> Problem.  ``atoms`` is a dict like {'q_u': ..., 'Q_x': ...}; ``extra_constraints`` is a list of
    dicts with keys among omega_t / omega_scenarios / N_tilde (the ``other_constraints`` the example adds,
    examples/.../micro_grid_control_simulation.py:200-227)."""
    N_p = Nt - 1 if N_p is None else N_p
    n = dims["nv"] * Nt
    prob = Problem(n)
    idx = var_layout(dims, Nt)
    bin_step = np.array([t == "b" for t in var_types], dtype=bool)
    prob.is_bin = np.tile(bin_step, Nt)
    prob.lb[prob.is_bin] = 0.0
    prob.ub[prob.is_bin] = 1.0
    prob.lb[idx["mu"]] = np.maximum(prob.lb[idx["mu"]], 0.0)         # nonneg slack -- variables.py:221
    if disable_soft_constraints and dims["nmu"]:
        prob.lb[idx["mu"]] = 0.0
        prob.ub[idx["mu"]] = 0.0
    if dims["nc"]:
        if with_std_constraints:
            H, rhs = evo_rhs(evo, dims, x0, omega_t, omega_scenarios=omega_scenarios)
            prob.add_rows(H, rhs)
        for ec in extra_constraints:
            H, rhs = evo_rhs(evo, dims, x0, ec.get("omega_t", omega_t), omega_scenarios=ec.get("omega_scenarios"),
                             N_tilde=ec.get("N_tilde"))
            prob.add_rows(H, rhs)
    maps = affine_maps(evo, dims, Nt, x0, omega_t)
    var_k_neg1 = var_k_neg1 or {}
    merged = {}
    for key, value in (atoms or {}).items():
        if value is None:
            continue
        wtype, atom, var, rate, post = parse_atom_key(key)
        dim = dims["nv"] if var == "v" else dims["n" + var]
        if dim == 0:
            continue
        slot = (wtype, atom, var, rate)
        merged[slot] = expand_weight(wtype, value, dim, Nt, N_p, post, prev=merged.get(slot))
    for (wtype, atom, var, rate), W in merged.items():
        if np.allclose(W, 0.0):
            continue
        dim = dims["nv"] if var == "v" else dims["n" + var]
        M, m0 = maps[var]
        M, m0 = M[:, :n], m0
        if rate:  # v(k) - v(k-1), first step uses var_k_neg1 -- objective_atoms.py:297-305
            prev = np.asarray(var_k_neg1.get(var, np.zeros(dim)), dtype=float).reshape(-1)
            M = M - np.vstack([np.zeros((dim, n)), M[:-dim]])
            m0 = m0 - np.concatenate([prev, m0[:-dim]])
        _apply_atom(prob, wtype, atom, W, M, m0, dim, n)
    return prob


def _pad(M, ntot):
    return np.hstack([M, np.zeros((M.shape[0], ntot - M.shape[1]))]) if M.shape[1] < ntot else M


def _apply_atom(prob, wtype, atom, W, M, m0, dim, n):
    if atom == "Linear":
        g = W.ravel() if wtype == "vector" else W.sum(axis=0)        # w'e  |  sum(W e)
        prob.c[:n] += g @ M
        prob.c0 += float(g @ m0)
    elif atom in ("Quadratic", "L22"):
        Q = np.diag(W.ravel() ** 2) if wtype == "vector" else W     # ||w.e||^2 (weights enter squared) | e'We
        Qs = 0.5 * (Q + Q.T)
        if prob.P is None:
            prob.P = np.zeros((prob.n, prob.n))
        prob.P[:n, :n] += 2.0 * (M.T @ Qs @ M)
        prob.c[:n] += 2.0 * (M.T @ (Qs @ m0))
        prob.c0 += float(m0 @ Q @ m0)
    else:  # L1, and Linf-with-weight which the reference evaluates as a per-step norm1 (:357-363)
        if atom == "Linf" and wtype is None:
            raise NotImplementedError
        A = np.diag(W.ravel()) @ M if wtype == "vector" else W @ M
        a0 = W.ravel() * m0 if wtype == "vector" else W @ m0
        t = prob.add_cols(A.shape[0], lb=0.0)
        ntot = prob.n
        Ap = _pad(A, ntot)
        sel = np.zeros((A.shape[0], ntot))
        sel[np.arange(A.shape[0]), t] = 1.0
        prob.add_rows(Ap - sel, -a0)      #  (A v + a0) <= t
        prob.add_rows(-Ap - sel, a0)      # -(A v + a0) <= t
        prob.c[t] += 1.0


def split_solution(v, evo, dims, Nt, x0, omega_t):
    """Per-variable horizon stacks + first-step values (``variables_k``, variables.py:75-85, 291)."""
    maps = affine_maps(evo, dims, Nt, x0, omega_t)
    v = np.asarray(v, dtype=float).ravel()[:dims["nv"] * Nt]
    full, first = {}, {}
    for name in VAR_NAMES:
        M, m0 = maps[name]
        val = M @ v + m0
        dim = dims["nv"] if name == "v" else dims["n" + name]
        full[name] = val
        first[name] = val[:dim]
    return full, first
