"""Assembly of the GENERAL mixed-integer quadratic program of a controller -- every cost atom of the reference's
grammar on any MLD -- in the canonical form of ``hmpc_miqp_solve_f64``:

    minimise 0.5 v'P v + c'v + c0    s.t.    H v <= rhs,  lb <= v <= ub,  v_j in {0, 1}

Reference: the atoms' cvxpy expressions, controllers/components/objective_atoms.py:308-363 (Linear ``w'e`` /
``sum(W e)``, Quadratic and L22 ``||w.e||^2`` / ``e'W e``, L1 ``||w.e||_1`` / ``||W e||_1``, Linf -- which the reference
evaluates as a per-step norm1 whenever a weight is given, :357-363), the rate form ``e(k) - e(k-1)`` with the logged
previous value (:297-305), and the affine maps of the predicted states / outputs (variables.py:245-286).

Everything is device-side tensor algebra on the condensed matrices K1 produced (``P += 2 M'QM`` are GEMMs on
``Gamma_v`` / ``L_v``); L1 / Linf atoms become epigraph columns ``t >= |A v + a0|`` appended after ``v~``.  The fast paths
(linear atoms; separable atoms on scalar-state MLDs) never come here: see MpcController._general_path.
"""
import numpy as np
import torch

from . import cabi


def _t(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)


class GeneralProblem(object):
    """Device tensors of the canonical form for a batch of B agents that share the atom structure."""

    def __init__(self, B, n, dev):
        self.B, self.n, self.n_v, self.dev = B, n, n, dev
        self.P = None
        self.c = torch.zeros((B, n), dtype=torch.float64, device=dev)
        self.c0 = torch.zeros((B,), dtype=torch.float64, device=dev)
        self.H = torch.zeros((B, 0, n), dtype=torch.float64, device=dev)
        self.rhs = torch.zeros((B, 0), dtype=torch.float64, device=dev)
        self.lb = np.full(n, -np.inf)
        self.ub = np.full(n, np.inf)
        self.is_bin = np.zeros(n, dtype=np.uint8)

    def add_rows(self, H, rhs):
        if H.shape[0] == 1 and self.B > 1:
            H = H.expand(self.B, H.shape[1], H.shape[2])
        if H.shape[2] < self.n:
            H = torch.cat([H, torch.zeros((H.shape[0], H.shape[1], self.n - H.shape[2]), dtype=torch.float64, device=self.dev)], dim=2)
        self.H = torch.cat([self.H, H], dim=1)
        self.rhs = torch.cat([self.rhs, rhs.reshape(self.B, -1)], dim=1)

    def add_cols(self, k, lb=-np.inf, ub=np.inf):
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=self.dev)  # noqa: E731
        self.c = torch.cat([self.c, z(self.B, k)], dim=1)
        self.H = torch.cat([self.H, z(self.B, self.H.shape[1], k)], dim=2)
        if self.P is not None:
            P = z(self.P.shape[0], self.n + k, self.n + k)
            P[:, :self.n, :self.n] = self.P
            self.P = P
        self.lb = np.concatenate([self.lb, np.full(k, lb)])
        self.ub = np.concatenate([self.ub, np.full(k, ub)])
        self.is_bin = np.concatenate([self.is_bin, np.zeros(k, dtype=np.uint8)])
        self.n += k
        return np.arange(self.n - k, self.n)


def affine_maps(batch, x0, omega):
    """var name -> (M [B|1, rows, n], m0 [B, rows]) with var~ = M v~ + m0 (variables.py:245-286, 288-317)."""
    d, B, Nt, dev = batch.dims, batch.B, batch.Nt, batch.device
    n = batch.nvt
    maps = {}
    eye = torch.eye(n, dtype=torch.float64, device=dev)
    for name in ("u", "delta", "z", "mu"):
        idx = torch.as_tensor(batch.var_index(name), dtype=torch.long, device=dev)
        maps[name] = (eye[idx].unsqueeze(0), torch.zeros((B, idx.numel()), dtype=torch.float64, device=dev))
    maps["v"] = (eye.unsqueeze(0), torch.zeros((B, n), dtype=torch.float64, device=dev))
    evo = batch.evo
    if d.nx:
        x_free = cabi.predict(evo["Phi_x"], None, evo["Gamma_omega"], evo["Gamma_5"].reshape(B, -1), x0, None, omega)
        maps["x"] = (evo["Gamma_v"], x_free.reshape(B, -1))
    else:
        maps["x"] = (torch.zeros((1, 0, n), dtype=torch.float64, device=dev), torch.zeros((B, 0), dtype=torch.float64, device=dev))
    if d.ny:
        y_free = cabi.predict(evo["L_x"], None, evo["L_omega"], evo["L_5"].reshape(B, -1), x0, None, omega)
        maps["y"] = (evo["L_v"], y_free.reshape(B, -1))
    else:
        maps["y"] = (torch.zeros((1, 0, n), dtype=torch.float64, device=dev), torch.zeros((B, 0), dtype=torch.float64, device=dev))
    nw = d.nomega * Nt
    w = omega if omega is not None else torch.zeros((B, nw), dtype=torch.float64, device=dev)
    maps["omega"] = (torch.zeros((1, nw, n), dtype=torch.float64, device=dev), w.reshape(B, nw))
    return maps


def assemble(batch, x0, omega, atoms, prev=None, constraint_sets=(), with_std_constraints=True, sign=1.0):
    """atoms: iterable of ObjectiveAtom (weights are host arrays shared by the batch); prev: name -> previous-step value
    (dim,) for rate atoms; constraint_sets: dicts as BatchMpc.solve's ``extra_constraints`` (+ optional ``x_k``)."""
    d, B, Nt, dev = batch.dims, batch.B, batch.Nt, batch.device
    n = batch.nvt
    prob = GeneralProblem(B, n, dev)
    lb, ub = batch.lb_v.copy(), batch.ub_v.copy()
    if batch.disable_soft_constraints and d.nmu:
        idx = batch.var_index("mu")
        lb[idx] = 0.0
        ub[idx] = 0.0
    prob.lb, prob.ub, prob.is_bin = lb, ub, batch.is_bin_v.astype(np.uint8).copy()
    x0 = None if x0 is None or not d.nx else _t(x0, dev).reshape(B, d.nx)
    omega = None if omega is None or not batch.nwt else _t(omega, dev).reshape(B, batch.nwt)
    if d.nc:
        if with_std_constraints:
            H, r = batch.constraint_rows(x0, omega)
            prob.add_rows(H, r)
        for ec in constraint_sets:
            w2 = ec.get("omega_tilde_k")
            sc = ec.get("omega_scenarios_k")
            xk = ec.get("x_k")
            w2 = omega if w2 is None else _t(w2, dev).reshape(B, batch.nwt)
            sc = None if sc is None else _t(sc, dev).reshape(B, batch.nwt, -1)
            xk = x0 if xk is None else _t(xk, dev).reshape(B, d.nx)
            H, r = batch.constraint_rows(xk, w2, scenarios=sc, N_tilde=ec.get("N_tilde"))
            prob.add_rows(H, r)
    maps = affine_maps(batch, x0, omega)
    prev = prev or {}
    for atom in atoms:
        W = np.asarray(atom.weight_N_tilde, dtype=np.float64)
        dim, name = atom.dim, atom.var_name
        if dim == 0 or np.allclose(W, 0.0):
            continue
        M, m0 = maps[name]
        if atom.is_rate_atom:        # e(k) - e(k-1); the first step uses the logged previous value (:297-305)
            e_prev = np.asarray(prev.get(name, np.zeros(dim)), dtype=np.float64).reshape(-1)
            if e_prev.size != dim or not np.all(np.isfinite(e_prev)):
                e_prev = np.zeros(dim)
            Msh = torch.zeros_like(M)
            Msh[:, dim:, :] = M[:, :-dim, :]
            M = M - Msh
            m0 = m0 - torch.cat([_t(e_prev, dev).expand(B, dim), m0[:, :-dim]], dim=1)
        vec = atom.weight_type == "vector"
        if atom.atom_type == "Linear":
            g = _t(W.ravel() if vec else W.sum(axis=0), dev)                    # w'e | sum(W e)
            prob.c[:, :n] += sign * torch.matmul(g, M).reshape(-1, n)
            prob.c0 += sign * (m0 @ g)
        elif atom.atom_type in ("Quadratic", "L22"):
            if sign < 0:
                raise NotImplementedError("maximising a convex cost atom is not a convex problem")
            Q = np.diag(W.ravel() ** 2) if vec else W                           # ||w.e||^2 | e'W e
            Qs = _t(0.5 * (Q + Q.T), dev)
            QM = torch.matmul(Qs, M)                                            # [B|1, rows, n]
            if prob.P is None:
                prob.P = torch.zeros((M.shape[0] if M.shape[0] > 1 else 1, prob.n, prob.n), dtype=torch.float64, device=dev)
            P_add = 2.0 * torch.matmul(M.transpose(1, 2), QM)
            if P_add.shape[0] > prob.P.shape[0]:
                prob.P = prob.P.expand(P_add.shape[0], prob.n, prob.n).clone()
            prob.P[:, :n, :n] += P_add
            prob.c[:, :n] += 2.0 * torch.matmul(m0.unsqueeze(1), QM.expand(B, -1, -1)).reshape(B, n)
            prob.c0 += ((m0 @ _t(Q, dev)) * m0).sum(dim=1)
        else:                                                                   # L1, and Linf with a weight (:357-363)
            if sign < 0:
                raise NotImplementedError("maximising a convex cost atom is not a convex problem")
            Wd = _t(np.diag(W.ravel()) if vec else W, dev)
            Am = torch.matmul(Wd, M)                                            # [B|1, rows, n]
            a0 = m0 @ Wd.T
            rows = Am.shape[1]
            t = prob.add_cols(rows, lb=0.0)
            sel = torch.zeros((1, rows, prob.n), dtype=torch.float64, device=dev)
            sel[0, torch.arange(rows, device=dev), torch.as_tensor(t, device=dev)] = 1.0
            Ap = torch.cat([Am, torch.zeros((Am.shape[0], rows, prob.n - n), dtype=torch.float64, device=dev)], dim=2)
            prob.add_rows(Ap - sel, -a0)                                        #  (A v + a0) <= t
            prob.add_rows(-Ap - sel, a0)                                        # -(A v + a0) <= t
            prob.c[:, torch.as_tensor(t, device=dev)] += 1.0
    return prob


def solve(prob, opts=None):
    """-> dict(v [B, n_v] (the MLD's v~ only), obj [B] (with c0), status, stats)."""
    dev = prob.dev
    lb, ub = _t(prob.lb, dev), _t(prob.ub, dev)
    isb = torch.as_tensor(prob.is_bin, dtype=torch.uint8).to(dev)
    v, obj, status, stats = cabi.miqp_solve(prob.c, prob.H, prob.rhs, lb, ub, isb, P=prob.P, opts=opts)
    return dict(v=v[:, :prob.n_v].contiguous(), v_full=v, obj=obj + prob.c0, status=status, stats=stats, c0=prob.c0,
                solver="miqp")
