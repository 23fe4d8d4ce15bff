"""Simulation step on the GPU (K5) and the auxiliary-variable computation of ``MldModel.lsim_k``.

reference: models/mld_model.py:647-766.  When delta / z / mu are not supplied the reference solves a cvxpy
feasibility MIP with Gurobi (``_compute_aux``); here the same feasibility set is handed to the GPU mixed-integer
solver with the (admissible) objective ``min sum(mu)`` so that the answer is the minimal slack.
"""
import numpy as np
import torch

from . import cabi
from .utils.structs import StructDict, ParNotSet, atleast_2d_col


def _col(var, dim, name, required=False):
    if var is None or dim == 0:
        return np.zeros((dim, 1))
    if var is ParNotSet:
        if required:
            raise ValueError("variable %s cannot be set to ParNotSet" % name)
        return ParNotSet
    return np.asarray(atleast_2d_col(var), dtype=np.float64).reshape(dim, 1)


def compute_aux(mld, x, u, delta, z, mu, omega, device="cuda"):
    """Fill in the missing ones of delta / z / mu (reference: models/mld_model.py:701-766)."""
    info = mld.mld_info
    free = [(name, val, info["n" + name]) for name, val in (("delta", delta), ("z", z), ("mu", mu))]
    unknown = [(n, d) for n, v, d in free if v is ParNotSet and d]
    vals = dict(delta=delta, z=z, mu=mu)
    for n, v, d in free:
        if v is ParNotSet and d == 0:
            vals[n] = np.zeros((0, 1))
    if not unknown:
        return vals["delta"], vals["z"], vals["mu"]
    nc = info.n_constraints
    # eliminate y:  (F2 + G D2) delta + (F3 + G D3) z + Psi mu <= f5 - E x - F1 u - F4 w - G (C x + D1 u + D4 w + d5)
    yc = mld.C @ x + mld.D1 @ u + mld.D4 @ omega + mld.d5
    rhs = mld.f5 - mld.E @ x - mld.F1 @ u - mld.F4 @ omega - mld.G @ yc
    blocks = dict(delta=mld.F2 + mld.G @ mld.D2, z=mld.F3 + mld.G @ mld.D3, mu=mld.Psi)
    cols, cost, lb, ub, isb = [], [], [], [], []
    for n, v, d in free:
        if not d:
            continue
        if v is ParNotSet:
            cols.append(blocks[n])
            cost += [1.0 if n == "mu" else 0.0] * d
            lb += [0.0 if n != "z" else -np.inf] * d
            ub += [1.0 if n == "delta" else np.inf] * d
            isb += [1 if n == "delta" else 0] * d
        else:
            rhs = rhs - blocks[n] @ v
            yc = yc  # known auxiliaries only shift the right-hand side
    H = np.hstack(cols)
    dev = torch.device(device)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    v, obj, status, stats = cabi.milp_solve(t(np.array(cost)[None, :]), t(H[None]), t(rhs.reshape(1, nc)),
                                            t(np.array(lb)), t(np.array(ub)), t(np.array(isb, dtype=np.uint8), torch.uint8))
    sol = v.cpu().numpy().ravel()
    ok = int(status.cpu()[0]) == 0
    o = 0
    for n, val, d in free:
        if d and val is ParNotSet:
            vals[n] = sol[o:o + d].reshape(d, 1) if ok else np.full((d, 1), np.nan)
            o += d
    return vals["delta"], vals["z"], vals["mu"]


def lsim_k_single(mld, x_k=ParNotSet, u_k=ParNotSet, delta_k=ParNotSet, z_k=ParNotSet, mu_k=ParNotSet, v_k=ParNotSet,
                  omega_k=ParNotSet, cons_tol=1e-6, device="cuda"):
    info = mld.mld_info
    x = _col(None if x_k is ParNotSet else x_k, info.nx, "x_k")
    omega = _col(omega_k, info.nomega, "omega_k", required=True)
    if v_k is not ParNotSet:
        if not all(a is ParNotSet for a in (u_k, delta_k, z_k, mu_k)):
            raise ValueError("Either supply concatenated input in 'v_k' or supply individual inputs 'u_k', 'delta_k', "
                             "'z_k' and 'mu_k', but not both.")
        v = np.asarray(atleast_2d_col(v_k), dtype=np.float64)
        o1, o2, o3 = info.nu, info.nu + info.ndelta, info.nu + info.ndelta + info.nz
        u, delta, z, mu = v[:o1], v[o1:o2], v[o2:o3], v[o3:]
    else:
        u = _col(u_k, info.nu, "u_k", required=True)
        delta = _col(delta_k, info.ndelta, "delta_k")
        z = _col(z_k, info.nz, "z_k")
        mu = _col(mu_k, info.nmu, "mu_k")
        if any(a is ParNotSet for a in (delta, z, mu)):
            delta, z, mu = compute_aux(mld, x, u, delta, z, mu, omega, device=device)
        v = np.vstack((u, delta, z, mu))
    dev = torch.device(device)
    d = cabi.make_dims(1, 1, nx=info.nx, nu=info.nu, ndelta=info.ndelta, nz=info.nz, nmu=info.nmu, nomega=info.nomega,
                       ny=info.ny, nc=info.n_constraints)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a.reshape(1, -1)), dtype=torch.float64).to(dev)
    mats = {k: torch.as_tensor(np.ascontiguousarray(mld[k]), dtype=torch.float64).to(dev).unsqueeze(0)
            for k in cabi.MAT_NAMES if mld[k].size}
    x1, y, cons = cabi.lsim_step(d, mats, t(x), t(u), t(delta), t(z), t(omega), cons_tol)
    return StructDict(x_k1=x1.cpu().numpy().reshape(-1, 1), x=x, u=u, delta=delta, z=z, mu=mu, v=v,
                      y=y.cpu().numpy().reshape(-1, 1), omega=omega,
                      cons=cons.cpu().numpy().astype(bool).reshape(-1, 1))
