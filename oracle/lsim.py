"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

One-step MLD simulation and the DEWH model closed forms.

* ``lsim_k``           : models/mld_model.py:647-699 (mu is ignored in ``cons`` -- :694; tolerance 1e-6 -- :648)
* ``aux_closed_form``  : what ``_compute_aux`` (models/mld_model.py:701-766) needs for the example's models:
                         DEWH: any feasible slack is admissible, the minimal one mu = max(0, violation) is used;
                         grid: delta = [y >= 0], z = delta*y are the unique solution of the grid rows
                         (examples/.../micro_grid_models.py:145-168).
* ``dewh_mld``         : examples/.../micro_grid_models.py:27-100 with parameters.py:10-27
* sim-model clamp      : examples/.../micro_grid_agents.py:395-399

Pinned against the unmodified reference for ``lsim_k`` (tests/golden/lsim_*.npz).  The DEWH closed form
replaces sympy-lambdified expressions that cannot run here (CallableMatrix needs wrapt 1.x): unpinned,
checked against the survey's published numeric values (SURVEY.md section 8 header).
"""
import numpy as np

DEWH_PARAMS = dict(C_w=4.1816e3, A_h=2.35, U_h=0.88, m_h=150.0, T_w=15.0, T_inf=25.0, P_h_Nom=3000.0,
                   T_h_min=50.0, T_h_max=65.0, T_h_Nom=45.0, T_h=45.0, D_h=0.0, ts=900.0)
GRID_PARAMS = dict(P_g_min=-2e4, P_g_max=2e4, eps=float(np.finfo(float).eps), ts=900.0)


def lsim_k(full, x, u, delta, z, mu, omega, cons_tol=1e-6):
    """All arguments column vectors (n,1).  Returns (x_k1, y, cons[bool])."""
    x1 = full["A"] @ x + full["B1"] @ u + full["B2"] @ delta + full["B3"] @ z + full["B4"] @ omega + full["b5"]
    y = full["C"] @ x + full["D1"] @ u + full["D2"] @ delta + full["D3"] @ z + full["D4"] @ omega + full["d5"]
    cons = (full["E"] @ x + full["F1"] @ u + full["F2"] @ delta + full["F3"] @ z + full["F4"] @ omega
            + full["G"] @ y + full["Psi"] @ (mu * 0) - full["f5"] <= cons_tol)
    return x1, y, cons


def dewh_scalars(p, const_heat=True, T_h=None, D_h=None):
    """-> (A, B1, B4, b5) scalars of the discretised DEWH model."""
    p1 = p["U_h"] * p["A_h"]
    p2 = p["m_h"] * p["C_w"]
    ts = p["ts"]
    if const_heat:
        A_c = -p1 / p2
        B4_c = p["C_w"] * (p["T_w"] - p["T_h_Nom"]) / p2
    else:
        T_h = p["T_h"] if T_h is None else T_h
        D_h = p["D_h"] if D_h is None else D_h
        r = (p["T_h_Nom"] - p["T_w"]) / (T_h - p["T_w"])
        A_c = -((D_h * p["C_w"] * r) + p1) / p2
        B4_c = p["C_w"] * p["T_w"] * r / p2
    B1_c = p["P_h_Nom"] / p2
    b5_c = p1 * p["T_inf"] / p2
    A = np.exp(A_c * ts)
    em = (A - 1.0) / A_c
    return A, em * B1_c, em * B4_c, em * b5_c


def dewh_mld(p, const_heat=True, T_h=None, D_h=None):
    """Numeric DEWH MLD matrices (binary input form): nx=1, nu=1 (binary), nmu=2, nomega=1, nc=2."""
    A, B1, B4, b5 = dewh_scalars(p, const_heat, T_h, D_h)
    return dict(A=[[A]], B1=[[B1]], B4=[[B4]], b5=[[b5]], E=[[1.0], [-1.0]], F1=[[0.0], [0.0]],
                Psi=[[-1.0, 0.0], [0.0, -1.0]], f5=[[p["T_h_max"]], [-p["T_h_min"]]])


def dewh_sim_step(p, x, u, omega):
    """micro_grid_agents.py:389-408: clamp T_h, re-parametrise the sim model with (D_h=omega, T_h=x), step it."""
    x = float(x)
    if x <= p["T_w"]:
        x = p["T_w"] + 0.1
    A, B1, B4, b5 = dewh_scalars(p, const_heat=False, T_h=x, D_h=float(omega))
    x1 = A * x + B1 * float(u) + B4 * float(omega) + b5
    cons = np.array([x - p["T_h_max"] <= 1e-6, -x + p["T_h_min"] <= 1e-6])
    return x1, x, cons


def dewh_thermostat(p, T_h, u_prev):
    """theromstat_control.py:50-62: hysteresis between T_h_max - T_h_max_sub_T_h_on and T_h_max - T_h_max_sub_T_h_off;
    inside the band the previous input is kept only if it equals 1."""
    if T_h <= p["T_h_max"] - p["T_h_max_sub_T_h_on"]:
        return 1
    if T_h >= p["T_h_max"] - p["T_h_max_sub_T_h_off"]:
        return 0
    return 1 if u_prev == 1 else 0


def grid_mld(p, num_devices):
    """micro_grid_models.py:137-172: nx=0, ndelta=1, nz=1, nomega=num_devices, ny=1, nc=6."""
    lo, hi, eps = p["P_g_min"], p["P_g_max"], p["eps"]
    return dict(D4=np.ones((1, num_devices)),
                F2=[[-lo], [-(hi + eps)], [-hi], [lo], [-lo], [hi]],
                F3=[[0.0], [0.0], [1.0], [-1.0], [1.0], [-1.0]],
                f5=[[-lo], [-eps], [0.0], [0.0], [-lo], [hi]],
                G=[[-1.0], [1.0], [0.0], [0.0], [-1.0], [1.0]])


def grid_aux_closed_form(y):
    """delta = [y >= 0], z = delta * y (import power)."""
    d = 1.0 if y >= 0 else 0.0
    return d, d * y
